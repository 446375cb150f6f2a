"""GPU tests of the reference-facing plugin API (step classes + datasets) and of BASELINE-size runs."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def pb():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pandora_b200

    pandora_b200.get_engine("cuda:0")
    return pandora_b200


def cones():
    from PIL import Image

    left = np.array(Image.open(os.path.join(GOLD, "cones", "left.png"))).astype(np.float32)
    right = np.array(Image.open(os.path.join(GOLD, "cones", "right.png"))).astype(np.float32)
    gt = np.array(Image.open(os.path.join(GOLD, "cones", "disp_left.tif"))).astype(np.float32)
    return left, right, gt


def bad_pixel_ratio(disp, gt, thr=1.0):
    """tests/functional_tests/test_basic.py:45-68: ground truth is positive, Pandora's disparities negative."""
    valid = gt > 0
    return float(np.mean(np.abs(disp[valid] + gt[valid]) > thr))


def test_c0_cones_sad_wta_matches_cpu_path(pb, oracle):
    """BASELINE configs[0]: cones, SAD w5, disp [-63, 0], WTA -- whole plugin flow vs the CPU restatement."""
    left, right, gt = cones()
    cfg = {"pipeline": {"matching_cost": {"matching_cost_method": "sad", "window_size": 5, "subpix": 1},
                        "disparity": {"disparity_method": "wta", "invalid_disparity": -9999}}}
    dl = pb.create_image_dataset(left, disparity=[-63, 0])
    dr = pb.create_image_dataset(right)
    disp, cv = pb.run(dl, dr, cfg)
    ref_cv, attrs = oracle.sad_ssd_cost_volume(left, right, 5, -63, 0, "sad")
    vm = oracle.validity_mask(375, 450, -63, 0, 2)
    oracle.cv_masked(ref_cv, vm, 2)
    ref_disp, inv = oracle.wta(ref_cv, np.arange(-63, 1))
    np.testing.assert_array_equal(cv["cost_volume"].data, ref_cv)
    np.testing.assert_array_equal(disp["disparity_map"].data, ref_disp)
    np.testing.assert_array_equal(disp["validity_mask"].data, oracle.wta_validity_mask(vm, inv))
    np.testing.assert_array_equal(cv["disp_indices"].data, ref_disp)
    assert cv.attrs["cmax"] == attrs["cmax"] and cv.attrs["type_measure"] == "min"
    assert 0.25 < bad_pixel_ratio(ref_disp, gt) < 0.40          # 0.316 measured on the CPU path


def test_cones_census_cbca_sgm_plugin_flow(pb, oracle):
    """Census -> CBCA -> SGM -> WTA through the step classes, volume resident in HBM between steps."""
    left, right, gt = cones()
    cfg = {"pipeline": {"matching_cost": {"matching_cost_method": "census", "window_size": 5},
                        "aggregation": {"aggregation_method": "cbca"},
                        "optimization": {"optimization_method": "sgm", "penalty": {"P1": 8, "P2": 32}},
                        "disparity": {"disparity_method": "wta"}}}
    dl = pb.create_image_dataset(left, disparity=[-60, 0])
    dr = pb.create_image_dataset(right)
    disp, cv = pb.run(dl, dr, cfg)
    assert cv["cost_volume"].device_tensor() is not None        # never copied to the host so far
    ref, attrs = oracle.census_cost_volume(left, right, 5, -60, 0)
    ref, cmax = oracle.cbca_cost_volume(left, right, ref, 2, -60, 5, 30.0, attrs["cmax"])
    ref = oracle.sgm_cost_volume(ref, 8, 32, cmax=cmax)
    ref_disp, _ = oracle.wta(ref, np.arange(-60, 1))
    np.testing.assert_array_equal(disp["disparity_map"].data, ref_disp)
    np.testing.assert_allclose(cv["cost_volume"].data, ref, rtol=1e-5)
    assert cv.attrs["aggregation"] == "cbca" and cv.attrs["optimization"] == "sgm" and cv.attrs["cmax"] == 25 * 81


def test_cones_census_sgm_quality_gate(pb):
    """The only reference-derived SGM check (tests/functional_tests/test_basic.py:135-166): bad pixels @1px <= 0.20
    for Census w5 + SGM P1=8 P2=32 (the reference adds vfit + median filtering, which only lower the ratio)."""
    left, right, gt = cones()
    pipe = pb.StereoPipeline(375, 450, -60, 0, "census", 5, sgm=(8, 32))
    disp = pipe.run_host(left, right).copy()
    assert bad_pixel_ratio(disp, gt) <= 0.20


def test_error_behaviour(pb):
    with pytest.raises(KeyError, match="No optimization method named foo supported"):
        pb.AbstractOptimization(None, optimization_method="foo")
    with pytest.raises(pb.Pb200Error):
        eng = pb.get_engine()
        import torch
        z = torch.zeros((8, 8), device=eng.device)
        eng.census(z, z, 4, -1, 1)                                # illegal window reaches the C-ABI -> error code


@pytest.mark.parametrize("cfg", ["C1", "C2", "C3"])
def test_baseline_sizes_properties(pb, oracle, cfg):
    """BASELINE.json configs at full size: size-independent properties + oracle on a row band."""
    import torch

    eng = pb.get_engine()
    H, W, D, cbca, sgm = {"C1": (1024, 1024, 128, None, None), "C2": (2048, 2048, 192, (5, 30.0), None),
                          "C3": (4096, 4096, 256, None, (8, 32))}[cfg]
    left, right, g = oracle.synthetic_pair(H, W, D)
    dmin = -(D - 1)
    pipe = pb.StereoPipeline(H, W, dmin, 0, "census", 5, cbca=cbca, sgm=sgm)
    disp = pipe.run_host(left, right).copy()
    cv = pipe.final_cv
    # (1) fused WTA == stand-alone WTA kernel on the final volume (checksum of the whole map)
    d2, _ = eng.wta(cv, dmin)
    assert torch.equal(d2, pipe.disp)
    # (2) the volume is finite exactly where the census geometry says (NaN bookkeeping), checked by counts
    n_nan = int(torch.isnan(cv).sum().item())
    xs = np.arange(W)
    valid_x = (xs >= 2) & (xs < W - 2)
    per_x = np.array([np.sum((x + dmin + np.arange(D) >= 2) & (x + dmin + np.arange(D) < W - 2)) for x in xs]) * valid_x
    assert n_nan == H * W * D - int(per_x.sum()) * (H - 4)
    # (3) matching is right: the recovered disparity equals the synthetic ground truth on most block interiors
    inner = np.zeros((H, W), bool)
    inner[8:-8, D + 8:-8] = True
    agree = np.mean(disp[inner] == g[inner])
    assert agree > (0.80 if cfg == "C1" else 0.85), agree
    # (4) oracle on a band of rows that is self-contained for the matching-cost stage
    if cfg == "C1":
        band = slice(100, 164)
        ref, _ = oracle.census_cost_volume(left[band], right[band], 5, dmin, 0)
        np.testing.assert_array_equal(cv[band][2:-2].cpu().numpy(), ref[2:-2])
        exp, _ = oracle.wta(ref, np.arange(dmin, 1))
        np.testing.assert_array_equal(disp[band][2:-2], exp[2:-2])
    # (5) determinism / idempotence: a second run gives the identical map
    assert np.array_equal(pipe.run_host(left, right), disp)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [dict(sgm=(8, 32)), dict(), dict(cbca=(5, 30.0))])
def test_run_host_banded_upload_equals_device_run(pb, oracle, cfg):
    """run_host uploads the images in row bands and starts the Census fill of a band as soon as it has arrived:
    same disparity map and volume as the single-shot device run, from pinned and from pageable host images."""
    import torch

    H, W, D = 1300, 257, 64
    left, right, _ = oracle.synthetic_pair(H, W, D)
    pipe = pb.StereoPipeline(H, W, -(D - 1), 0, "census", 5, **cfg)
    ref = pipe.run_device(pipe.eng.to_device(left), pipe.eng.to_device(right)).cpu().numpy()
    ref_cv = pipe.final_cv.clone()
    pipe.cv_a.fill_(-1.0)
    got = pipe.run_host(left, right).copy()                                # pageable numpy input
    np.testing.assert_array_equal(got, ref)
    assert torch.equal(torch.nan_to_num(pipe.final_cv, nan=-7.0), torch.nan_to_num(ref_cv, nan=-7.0))
    hl, hr = torch.from_numpy(left).pin_memory(), torch.from_numpy(right).pin_memory()
    pipe.cv_a.fill_(-1.0)
    got = pipe.run_host(hl, hr).copy()                                     # pinned tensors: no staging copy
    np.testing.assert_array_equal(got, ref)


def _oracle_median_filter(oracle, disp, mask):
    """MedianFilter.filter_disparity, filter/median.py:96-132."""
    masked = disp.copy()
    masked[(mask & oracle.MSK_INVALID) != 0] = np.nan
    valid = np.isfinite(masked)
    med = oracle.median_filter3(masked)
    out = disp.copy()
    out[valid] = med[valid]
    return out


def test_reference_sample_config_a_semi_global_matching(pb, oracle):
    """The reference's flagship sample pipeline (data_samples/json_conf_files/a_semi_global_matching.json: census 5x5 ->
    SGM P1=8 P2=32 -> WTA (invalid = NaN) -> vfit -> 3x3 median -> cross_checking_accurate (threshold 1) -> 3x3 median) on
    cones through run(), every step on the device for the left AND the right image, against the same chain of oracle
    functions -- bit-exact -- and against the reference's functional gate (tests/functional_tests/test_basic.py:135-166)."""
    left, right, gt = cones()
    H, W = left.shape
    cfg = {"pipeline": {
        "matching_cost": {"matching_cost_method": "census", "window_size": 5, "subpix": 1},
        "optimization": {"optimization_method": "sgm", "overcounting": False,
                         "penalty": {"penalty_method": "sgm_penalty", "P1": 8, "P2": 32, "p2_method": "constant"}},
        "disparity": {"disparity_method": "wta", "invalid_disparity": "NaN"},
        "refinement": {"refinement_method": "vfit"},
        "filter": {"filter_method": "median", "filter_size": 3},
        "validation": {"validation_method": "cross_checking_accurate", "cross_checking_threshold": 1},
        "filter.this_time_after_validation": {"filter_method": "median", "filter_size": 3}}}
    dl = pb.create_image_dataset(left, disparity=[-60, 0])
    dr = pb.create_image_dataset(right)
    disp, cv, rdisp = pb.run(dl, dr, cfg, return_right=True)

    def side(a, b, dmin, dmax):
        ccv, attrs = oracle.census_cost_volume(a, b, 5, dmin, dmax)
        vm = oracle.validity_mask(H, W, dmin, dmax, 2)
        oracle.cv_masked(ccv, vm, 2)
        scv = oracle.sgm_cost_volume(ccv, 8, 32, cmax=attrs["cmax"])
        d, inv = oracle.wta(scv, np.arange(dmin, dmax + 1), invalid_disparity=np.nan)
        m = oracle.wta_validity_mask(vm, inv)
        itp, d, m = oracle.refinement(scv, d, m, dmin, dmax, 1, "min", "vfit")
        return scv, _oracle_median_filter(oracle, d, m), m, itp

    scv_l, d_l, m_l, itp_l = side(left, right, -60, 0)
    _, d_r, m_r, _ = side(right, left, 0, 60)
    m_l2, conf_l = oracle.cross_checking(d_l, m_l, d_r, 1, -60, 0, 2)
    m_r2, _ = oracle.cross_checking(d_r, m_r, d_l, 1, 0, 60, 2)
    d_l2 = _oracle_median_filter(oracle, d_l, m_l2)
    d_r2 = _oracle_median_filter(oracle, d_r, m_r2)
    np.testing.assert_array_equal(cv["cost_volume"].data, scv_l)
    np.testing.assert_array_equal(disp["validity_mask"].data, m_l2)
    np.testing.assert_array_equal(disp["disparity_map"].data, d_l2)
    np.testing.assert_array_equal(disp["interpolated_coeff"].data, itp_l)
    np.testing.assert_array_equal(np.asarray(disp["confidence_measure"].data)[:, :, 0], conf_l)
    np.testing.assert_array_equal(rdisp["validity_mask"].data, m_r2)
    np.testing.assert_array_equal(rdisp["disparity_map"].data, d_r2)
    assert disp.attrs["filter"] == "median" and disp.attrs["validation"] == "cross_checking_accurate"
    # functional gate of the reference on the valid pixels of the final map
    ok = (disp["validity_mask"].data & oracle.MSK_INVALID) == 0
    sel = ok & (gt > 0)
    assert float(np.mean(np.abs(disp["disparity_map"].data[sel] + gt[sel]) > 1.0)) <= 0.20


@pytest.mark.gpu
def test_pandora_plugin_classes_run_the_pipeline(pb, oracle):
    """The classes `pandora_plugin_b200` registers with Pandora's factories (here: stand-in factories with the same decorator
    API), driven like the state machine drives them -- compute_cost_volume, cv_masked, optimize_cv, to_disp -- give the oracle's
    disparity map, with and without host copies between the steps."""
    import pandora_plugin_b200 as plug

    def factory(key):
        class Base:
            avail = {}

            def __new__(cls, *args, **cfg):
                return super().__new__(cls.avail[cfg[key]] if cls is Base else cls)

            @classmethod
            def register_subclass(cls, name, *aliases):
                def deco(sub):
                    cls.avail[name] = sub
                    return sub
                return deco
        return Base

    mc_f, agg_f, opt_f, disp_f = (factory(k) for k in ("matching_cost_method", "aggregation_method", "optimization_method", "disparity_method"))
    plug.build_classes(mc_f, agg_f, opt_f, disp_f)
    H, W, D = 40, 200, 64
    left, right, _ = oracle.synthetic_pair(H, W, D)
    dmin, dmax = -(D - 1), 0
    cv_ref, attrs = oracle.census_cost_volume(left, right, 5, dmin, dmax)
    exp, _ = oracle.wta(oracle.sgm_cost_volume(cv_ref, 8, 32, cmax=attrs["cmax"]), np.arange(dmin, dmax + 1))
    for keep in (True, False):
        plug.keep_host_copy = keep
        try:
            il, ir = pb.create_image_dataset(left, disparity=[dmin, dmax]), pb.create_image_dataset(right)
            mc = mc_f(matching_cost_method="census_b200", window_size=5, subpix=1)
            grids = (il["disparity"].data[0], il["disparity"].data[1])
            cv = pb.AbstractMatchingCost.allocate_cost_volume(mc._impl, il, grids)
            cv = pb.validity_mask(il, ir, cv)
            cv = mc.compute_cost_volume(il, ir, cv)
            mc.cv_masked(il, ir, cv, *grids)
            cv = opt_f(il, optimization_method="sgm_b200", penalty={"P1": 8, "P2": 32}).optimize_cv(cv, il, ir)
            assert cv.attrs["pb200_resident"]["cost_volume"].is_cuda
            disp = disp_f(disparity_method="wta_b200").to_disp(cv, il, ir)
            np.testing.assert_array_equal(np.asarray(disp["disparity_map"].data), exp)
        finally:
            plug.keep_host_copy = True


def test_staged_upload_of_large_host_arrays():
    """Engine.to_device cuts large pageable arrays into row chunks staged through page-locked memory by worker threads:
    same bytes on the device, also back to back (one shared staging buffer) and for non-float dtypes."""
    import torch

    import pandora_b200

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    eng = pandora_b200.get_engine("cuda:0")
    g = np.random.default_rng(3)
    a = g.random((2100, 4099), dtype=np.float32)                  # 34 MB, odd shape
    b = g.random((2100, 4099), dtype=np.float32)
    da, db = eng.to_device(a), eng.to_device(b)
    a_copy = a.copy()
    a[:] = -1.0                                                   # the caller may reuse its array as soon as to_device returns
    assert torch.equal(da.cpu(), torch.from_numpy(a_copy)) and torch.equal(db.cpu(), torch.from_numpy(b))
    m = g.integers(0, 3, (4096, 4096)).astype(np.int16)           # 33 MB mask
    dm = eng.to_device(m, dtype=torch.int16)
    assert dm.dtype == torch.int16 and torch.equal(dm.cpu(), torch.from_numpy(m))


test_staged_upload_of_large_host_arrays = pytest.mark.gpu(test_staged_upload_of_large_host_arrays)


@pytest.mark.gpu
def test_variable_left_grids_with_accurate_cross_checking(pb, oracle):
    """Per-pixel left disparity grids + cross_checking_accurate through run(): the right image's grids come from
    reverse_disp_range (state_machine.py:668-683, matching_cost.cpp:59-131), both volumes are masked to their grids, and the
    result equals the chain of oracle functions bit for bit (left and right maps, masks, left-right distances)."""
    from pandora_b200.dataset import DataArray

    H, W = 60, 140
    left, right, _ = oracle.synthetic_pair(H, W, 24, seed=77)
    lmin = np.full((H, W), -20, np.float32)
    lmax = np.full((H, W), -3, np.float32)
    lmin[:, 70:] = -12
    lmax[:, 70:] = 0
    lmin[20:40, 30:60] = -6                                       # a narrow window somewhere
    lmax[20:40, 30:60] = -5
    cfg = {"pipeline": {"matching_cost": {"matching_cost_method": "census", "window_size": 5, "subpix": 1},
                        "disparity": {"disparity_method": "wta", "invalid_disparity": "NaN"},
                        "validation": {"validation_method": "cross_checking_accurate", "cross_checking_threshold": 1}}}
    dl = pb.create_image_dataset(left)
    dl["disparity"] = (("band_disp", "row", "col"), np.stack([lmin, lmax]))
    dl.coords["band_disp"] = DataArray(np.array(["min", "max"]), ("band_disp",))
    dr = pb.create_image_dataset(right)
    disp, cv, rdisp = pb.run(dl, dr, cfg, return_right=True)

    def side(a, b, gmin, gmax):
        dmin, dmax = int(np.nanmin(gmin)), int(np.nanmax(gmax))
        ccv, _ = oracle.census_cost_volume(a, b, 5, dmin, dmax)
        vm = oracle.validity_mask(H, W, dmin, dmax, 2)
        oracle.cv_masked_full(ccv, vm, 2, 5, dmin, grid_min=gmin, grid_max=gmax)
        d, inv = oracle.wta(ccv, np.arange(dmin, dmax + 1), invalid_disparity=np.nan)
        return ccv, d, oracle.wta_validity_mask(vm, inv), dmin, dmax

    cv_l, d_l, m_l, dmin, dmax = side(left, right, lmin, lmax)
    rmin, rmax = oracle.reverse_disp_range(lmin, lmax)
    _, d_r, m_r, rdmin, rdmax = side(right, left, rmin, rmax)
    m_l2, conf_l = oracle.cross_checking(d_l, m_l, d_r, 1, dmin, dmax, 2)
    m_r2, _ = oracle.cross_checking(d_r, m_r, d_l, 1, rdmin, rdmax, 2)
    np.testing.assert_array_equal(cv["cost_volume"].data, cv_l)
    np.testing.assert_array_equal(disp["disparity_map"].data, d_l)
    np.testing.assert_array_equal(rdisp["disparity_map"].data, d_r)
    np.testing.assert_array_equal(disp["validity_mask"].data, m_l2)
    np.testing.assert_array_equal(rdisp["validity_mask"].data, m_r2)
    np.testing.assert_array_equal(np.asarray(disp["confidence_measure"].data)[:, :, 0], conf_l)


def _norm_percentile(amb):
    """Ambiguity.normalize_with_percentile, ambiguity.py:172-186 (percentile 1)."""
    a = np.copy(amb)
    lo, hi = np.percentile(a, 1.0), np.percentile(a, 99.0)
    np.clip(a, lo, hi, out=a)
    return (a - np.min(a)) / (np.max(a) - np.min(a))


@pytest.mark.gpu
def test_sgm_use_confidence_sample_pipeline(pb, oracle):
    """use_confidence (plugin_libsgm.rst:38-47: every cost of a pixel is multiplied by the pixel's ambiguity confidence before
    the recurrence) on the pipeline of data_samples/json_conf_files/a_semi_global_matching_with_confidence.json up to the
    disparity step: census -> ambiguity ".before" -> SGM on the weighted costs -> ambiguity ".after" -> WTA, against the same
    chain of oracle functions; and a missing confidence band means confidence 1 (same result as without the option)."""
    H, W, D = 40, 120, 32
    left, right, _ = oracle.synthetic_pair(H, W, D, seed=9)
    dmin, dmax = -(D - 1), 0
    base = {"matching_cost": {"matching_cost_method": "census", "window_size": 5, "subpix": 1}}
    sgm = {"optimization_method": "sgm", "overcounting": False, "penalty": {"penalty_method": "sgm_penalty", "P1": 8, "P2": 32, "p2_method": "constant"}}
    cfg = {"pipeline": {**base,
                        "cost_volume_confidence.before": {"confidence_method": "ambiguity", "eta_max": 0.7, "eta_step": 0.01},
                        "optimization": {**sgm, "use_confidence": "cost_volume_confidence.before"},
                        "cost_volume_confidence.after": {"confidence_method": "ambiguity", "eta_max": 0.7, "eta_step": 0.01},
                        "disparity": {"disparity_method": "wta", "invalid_disparity": "NaN"}}}
    dl, dr = pb.create_image_dataset(left, disparity=[dmin, dmax]), pb.create_image_dataset(right)
    disp, cv = pb.run(dl, dr, cfg)
    assert list(disp.coords["indicator"].data) == ["confidence_from_ambiguity.before", "confidence_from_ambiguity.after"]

    ccv, attrs = oracle.census_cost_volume(left, right, 5, dmin, dmax)
    vm = oracle.validity_mask(H, W, dmin, dmax, 2)
    oracle.cv_masked(ccv, vm, 2)
    etas = np.arange(0.0, 0.7, 0.01)
    grids = np.stack([np.full((H, W), dmin), np.full((H, W), dmax)]).astype(np.int64)
    disps = np.arange(dmin, dmax + 1).astype(np.float32)
    conf = (1 - _norm_percentile(oracle.ambiguity(ccv, etas, grids, disps))).astype(np.float32)
    np.testing.assert_array_equal(np.asarray(disp["confidence_measure"].data)[:, :, 0], conf)
    weighted = (ccv * conf[:, :, None]).astype(np.float32)
    scv = oracle.sgm_cost_volume(weighted, 8, 32, cmax=attrs["cmax"])
    np.testing.assert_array_equal(cv["cost_volume"].data, scv)
    d, _ = oracle.wta(scv, np.arange(dmin, dmax + 1), invalid_disparity=np.nan)
    np.testing.assert_array_equal(disp["disparity_map"].data, d)

    # a band that does not exist: confidence 1 everywhere == the plain SGM step (the fused stage may run again)
    cfg1 = {"pipeline": {**base, "optimization": {**sgm, "use_confidence": "cost_volume_confidence.before"},
                         "disparity": {"disparity_method": "wta", "invalid_disparity": "NaN"}}}
    cfg0 = {"pipeline": {**base, "optimization": sgm, "disparity": {"disparity_method": "wta", "invalid_disparity": "NaN"}}}
    d1, _ = pb.run(pb.create_image_dataset(left, disparity=[dmin, dmax]), pb.create_image_dataset(right), cfg1)
    d0, _ = pb.run(pb.create_image_dataset(left, disparity=[dmin, dmax]), pb.create_image_dataset(right), cfg0)
    np.testing.assert_array_equal(d1["disparity_map"].data, d0["disparity_map"].data)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,over", [((9, 21, 64), False), ((17, 40, 40), True), ((6, 33, 200), False), ((5, 12, 300), False), ((1, 1, 8), False)])
def test_sgm_min_cost_paths_kernels_vs_oracle(pb, oracle, shape, over):
    """pb200_sgm_min_cost_paths: optimised volume and nb_of_directions map against the oracle (NaN cells, whole pixels without
    a valid cost, every per-lane register count of the path kernel)."""
    eng = pb.get_engine("cuda:0")
    g = np.random.default_rng(shape[1] + shape[2])
    cv = g.integers(0, 26, shape).astype(np.float32)
    cv[g.random(shape) < 0.12] = np.nan
    if shape[0] > 2:
        cv[2, 1, :] = np.nan
    ref, nb_ref = oracle.sgm_min_cost_paths(cv, 8, 32, cmax=25, overcounting=over)
    out, nb = eng.sgm_min_cost_paths(eng.to_device(cv), 8, 32, oracle.sgm_invalid_value(25, 32), overcounting=over)
    np.testing.assert_array_equal(out.cpu().numpy(), ref)
    np.testing.assert_array_equal(nb.cpu().numpy(), nb_ref)


@pytest.mark.gpu
def test_sgm_min_cost_paths_through_run(pb, oracle):
    """The option through run(cfg): the band optimization_plugin_libsgm_nb_of_directions (docs/source/userguide/output.rst:22)
    joins the confidence measures of the cost volume and of the disparity dataset."""
    H, W, D = 30, 90, 32
    left, right, _ = oracle.synthetic_pair(H, W, D, seed=21)
    dmin, dmax = -(D - 1), 0
    cfg = {"pipeline": {"matching_cost": {"matching_cost_method": "census", "window_size": 5, "subpix": 1},
                        "optimization": {"optimization_method": "sgm", "min_cost_paths": True, "penalty": {"P1": 8, "P2": 32}},
                        "disparity": {"disparity_method": "wta", "invalid_disparity": "NaN"}}}
    disp, cv = pb.run(pb.create_image_dataset(left, disparity=[dmin, dmax]), pb.create_image_dataset(right), cfg)
    ccv, attrs = oracle.census_cost_volume(left, right, 5, dmin, dmax)
    vm = oracle.validity_mask(H, W, dmin, dmax, 2)
    oracle.cv_masked(ccv, vm, 2)
    scv, nb = oracle.sgm_min_cost_paths(ccv, 8, 32, cmax=attrs["cmax"])
    assert list(cv.coords["indicator"].data) == ["optimization_plugin_libsgm_nb_of_directions"]
    np.testing.assert_array_equal(cv["cost_volume"].data, scv)
    np.testing.assert_array_equal(np.asarray(cv["confidence_measure"].data)[:, :, 0], nb)
    np.testing.assert_array_equal(np.asarray(disp["confidence_measure"].data)[:, :, 0], nb)
    d, _ = oracle.wta(scv, np.arange(dmin, dmax + 1), invalid_disparity=np.nan)
    np.testing.assert_array_equal(disp["disparity_map"].data, d)
