"""GPU parity tests of the SURVEY.md 8(f) rows -- input masks / disparity grids, fast cross-checking (right disparity
map from the left volume + consistency check), vfit / quadratic refinement, ambiguity / risk -- through the C-ABI
against the CPU oracle and the reference's golden vectors.  Bar: bit-exact everywhere (integer / index work, and
float32 arithmetic in the reference's operation order)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REF = "test_refinement.py::TestRefinement."
VAL = "test_validation.py::TestValidation."
AMB = "test_confidence/test_ambiguity.py::"
RISK = "test_confidence/test_risk.py::"
CONF = "test_confidence/conftest.py::"
MC = "test_matching_cost/test_matching_cost.py::"


@pytest.fixture(scope="module")
def eng():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pandora_b200

    return pandora_b200.get_engine("cuda:0")


def dev(eng, a):
    return eng.to_device(np.ascontiguousarray(a, dtype=np.float32))


def dev_mask(eng, m):
    return eng.to_device(np.ascontiguousarray(m).astype(np.uint16).view(np.int16), dtype=None)


def host(t):
    return t.detach().cpu().numpy()


def host_mask(t):
    return host(t).view(np.uint16)


def random_volume(seed, H, W, D, nan_frac=0.08, integer=True, hi=40):
    gen = np.random.default_rng(seed)
    cv = gen.integers(0, hi, (H, W, D)).astype(np.float32)
    if not integer:
        cv += gen.random((H, W, D)).astype(np.float32)
    cv[gen.random(cv.shape) < nan_frac] = np.nan
    return cv


# ---- refinement ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("method,test", [("quadratic", "test_quadratic"), ("vfit", "test_vfit")])
def test_refinement_goldens(eng, goldens, method, test):
    cv = goldens[REF + "setUp::self.cv@0"].astype(np.float32)
    disp = dev(eng, goldens[REF + "setUp::self.disp@0"])
    mask = dev_mask(eng, goldens[REF + "setUp::self.disp@1"])
    itp = eng.refinement(dev(eng, cv), disp, mask, -2, 2, 1, False, method)
    np.testing.assert_allclose(host(disp), goldens[REF + test + "::gt_sub_disp"], rtol=1e-6)
    np.testing.assert_allclose(host(itp), goldens[REF + test + "::gt_sub_cost"], rtol=1e-6)
    np.testing.assert_array_equal(host_mask(mask), goldens[REF + test + "::gt_mask"])


def test_approximate_refinement_golden(eng, goldens):
    t = REF + "test_vfit_approximate_subpixel_refinement::"
    disp, mask = dev(eng, goldens[t + "disp_right@0"]), dev_mask(eng, goldens[t + "disp_right@1"])
    itp = eng.refinement(dev(eng, goldens[t + "cv_left@0"]), disp, mask, -3, 2, 1, False, "vfit", approximate=True)
    np.testing.assert_allclose(host(disp), goldens[t + "gt_sub_disp"], rtol=1e-6)
    np.testing.assert_allclose(host(itp), goldens[t + "gt_sub_costs"], rtol=1e-6)
    np.testing.assert_array_equal(host_mask(mask), goldens[t + "gt_mask"])


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("method", ["vfit", "quadratic"])
@pytest.mark.parametrize("measure", ["min", "max"])
def test_refinement_vs_oracle(eng, oracle, seed, method, measure):
    subpix = (1, 2, 4)[seed % 3]
    H, W, D = 61, 83, 4 * subpix * 3 + 1
    cv = random_volume(seed, H, W, D, integer=seed % 2 == 0)
    if measure == "max":
        cv = -cv
    d_min, d_max = -4.0, -4.0 + (D - 1) / subpix
    work = np.where(np.isnan(cv), np.inf if measure == "min" else -np.inf, cv)
    idx = np.argmin(work, axis=2) if measure == "min" else np.argmax(work, axis=2)
    disp = (d_min + idx / subpix).astype(np.float32)
    gen = np.random.default_rng(seed + 50)
    mask = np.zeros((H, W), dtype=np.uint16)
    mask[gen.random((H, W)) < 0.1] = 2
    mask[gen.random((H, W)) < 0.1] |= 4
    exp_itp, exp_disp, exp_mask = oracle.refinement(cv, disp, mask, d_min, d_max, subpix, measure, method)
    d_disp, d_mask = dev(eng, disp), dev_mask(eng, mask)
    itp = eng.refinement(dev(eng, cv), d_disp, d_mask, d_min, d_max, subpix, measure == "max", method)
    np.testing.assert_array_equal(host(itp), exp_itp)
    np.testing.assert_array_equal(host(d_disp), exp_disp)
    np.testing.assert_array_equal(host_mask(d_mask), exp_mask)


@pytest.mark.parametrize("seed", range(3))
def test_right_refinement_from_left_volume(eng, oracle, seed):
    """mode 2 == loop_refinement on reverse_cost_volume(left) without the right volume (state_machine.py:488-490)."""
    H, W, D = 37, 90, 12
    dmin, dmax = -9, 2
    cv = random_volume(seed, H, W, D)
    right_cv = oracle.reverse_cost_volume(cv, -dmax)
    rdisp, inv = oracle.wta(right_cv, np.arange(-dmax, -dmin + 1))
    mask = np.where(inv, 0x3C3, 0).astype(np.uint16)
    exp_itp, exp_disp, exp_mask = oracle.refinement(right_cv, rdisp, mask, -dmax, -dmin, 1, "min", "vfit")
    d_disp, d_mask = dev(eng, rdisp), dev_mask(eng, mask)
    itp = eng.refinement(dev(eng, cv), d_disp, d_mask, -dmax, -dmin, 1, False, "vfit", approximate=2)
    np.testing.assert_array_equal(host(itp), exp_itp)
    np.testing.assert_array_equal(host(d_disp), exp_disp)
    np.testing.assert_array_equal(host_mask(d_mask), exp_mask)


# ---- fast cross-checking ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,rng", [((9, 70, 8), (-5, 2)), ((5, 300, 64), (-63, 0)), ((4, 333, 128), (-100, 27)), ((3, 100, 256), (-255, 0)),
                                       ((6, 41, 12), (3, 14)), ((6, 50, 16), (-40, -25)), ((2, 20, 64), (-70, -7))])
@pytest.mark.parametrize("measure", ["min", "max"])
def test_wta_right_vs_oracle(eng, oracle, shape, rng, measure):
    H, W, D = shape
    dmin, dmax = rng
    assert dmax - dmin + 1 == D
    cv = random_volume(D + W, H, W, D, nan_frac=0.2, hi=6)                # few levels: many ties
    cv[0, : W // 2, :] = np.nan
    exp, inv = oracle.right_disparity_fast(cv, dmin, dmax, measure)
    disp, flags = eng.wta_right(dev(eng, cv), -dmax, measure == "max")
    np.testing.assert_array_equal(host(disp), exp)
    np.testing.assert_array_equal(host(flags).astype(bool), inv)


def test_wta_right_fallback_shape(eng, oracle):
    cv = random_volume(1, 5, 33, 7)                                          # D % 4 != 0 -> reverse + wta
    exp, _ = oracle.right_disparity_fast(cv, -3, 3)
    disp, _ = eng.wta_right(dev(eng, cv), -3)
    np.testing.assert_array_equal(host(disp), exp)


def test_wta_right_c3_row_band_property(eng):
    """Full-width property at the BASELINE size (4096 columns, D = 256): the right map of a volume whose minimum sits on
    a known diagonal is that diagonal."""
    import torch

    H, W, D = 8, 4096, 256
    dmin = -255
    g = torch.Generator(device="cuda").manual_seed(3)
    cv = torch.randint(5, 30, (H, W, D), device="cuda", generator=g).float()
    true_d = torch.randint(dmin, 1, (H, 1), device="cuda", generator=g)      # left disparity per row
    cols = torch.arange(W, device="cuda")[None, :]
    cv.scatter_(2, (true_d - dmin).expand(H, W)[:, :, None], 0.0)            # cost 0 at the true disparity
    disp, _ = eng.wta_right(cv, 0)                                           # right range [0, 255]
    exp = (-true_d).expand(H, W).float().clone()
    reach = (cols - true_d) < W                                              # the matching left column x = j - d exists
    np.testing.assert_array_equal(host(disp)[host(reach)], host(exp)[host(reach)])


@pytest.mark.parametrize("seed", range(5))
def test_cross_checking_vs_oracle(eng, oracle, seed):
    gen = np.random.default_rng(seed)
    H, W = 40, 97
    dmin, dmax = -7, 5
    left = gen.integers(dmin, dmax + 1, (H, W)).astype(np.float32)
    right = (-left + gen.integers(-1, 2, (H, W))).astype(np.float32)
    if seed % 2:
        left += gen.choice([0.0, 0.25, 0.5], (H, W)).astype(np.float32)
        right += gen.choice([0.0, -0.25, 0.5], (H, W)).astype(np.float32)
    right[gen.random((H, W)) < 0.05] = np.nan
    mask = np.zeros((H, W), dtype=np.uint16)
    mask[gen.random((H, W)) < 0.1] = 2
    mask[gen.random((H, W)) < 0.1] |= 4
    off = seed % 3
    exp_mask, exp_conf = oracle.cross_checking(left, mask, right, 1.0 if seed else 0.0, dmin, dmax, off)
    d_mask = dev_mask(eng, mask)
    conf = eng.cross_checking(dev(eng, left), d_mask, dev(eng, right), 1.0 if seed else 0.0, dmin, dmax, off)
    np.testing.assert_array_equal(host_mask(d_mask), exp_mask)
    np.testing.assert_array_equal(host(conf), exp_conf)


def test_cross_checking_goldens(eng, goldens):
    left, lmask = goldens[VAL + "setUp::self.left@0"], goldens[VAL + "setUp::self.left@2"]
    d_mask = dev_mask(eng, lmask)
    conf = eng.cross_checking(dev(eng, left), d_mask, dev(eng, goldens[VAL + "setUp::self.right@0"]), 0.0, -2, 2, 0)
    np.testing.assert_array_equal(host_mask(d_mask), goldens[VAL + "test_cross_checking::gt_mask"])
    np.testing.assert_array_equal(host(conf), goldens[VAL + "test_cross_checking::gt_dist"][:, :, 1])
    t = VAL + "test_cross_checking_float_disparity::"
    d_mask = dev_mask(eng, goldens[t + "left@2"])
    eng.cross_checking(dev(eng, goldens[t + "left@0"]), d_mask, dev(eng, goldens[t + "right@0"]), 0.0, -2, 2, 0)
    np.testing.assert_array_equal(host_mask(d_mask), goldens[t + "gt_mask"])


# ---- confidence ---------------------------------------------------------------------------------------------------
def test_confidence_goldens(eng, goldens):
    t = AMB + "test_compute_ambiguity_and_sampled_ambiguity::"
    out = eng.confidence(dev(eng, goldens[t + "cv_"]), goldens[t + "etas"], goldens[t + "grids"], goldens[t + "disparity_range"],
                         sampled_ambiguity=True)
    np.testing.assert_allclose(host(out["ambiguity"]), goldens[t + "gt_amb_int"], rtol=1e-6)
    np.testing.assert_allclose(host(out["sampled_ambiguity"]), goldens[t + "gt_sam_amb"], rtol=1e-6)
    t = RISK + "test_compute_risk_and_sampled_risk::"
    out = eng.confidence(dev(eng, goldens[t + "cv_"]), goldens[t + "etas"], goldens[t + "grids"], goldens[t + "disparity_range"],
                         ambiguity=False, risk=True, sampled_risk=True, sampled_ambiguity_in=goldens[t + "sampled_ambiguity"])
    for key, gt in [("risk_max", "gt_risk_max"), ("risk_min", "gt_risk_min"), ("disp_sup", "gt_disp_sup"), ("disp_inf", "gt_disp_inf"),
                    ("sampled_risk_max", "gt_sampled_risk_max"), ("sampled_risk_min", "gt_sampled_risk_min")]:
        np.testing.assert_allclose(host(out[key]), goldens[t + gt], rtol=1e-6)


@pytest.mark.parametrize("seed", range(6))
def test_confidence_vs_oracle(eng, oracle, seed):
    gen = np.random.default_rng(seed)
    H, W, D = 33, 47, (12, 64, 33)[seed % 3]
    cv = random_volume(seed, H, W, D, nan_frac=0.15, integer=seed % 2 == 0, hi=60)
    cv[0, :3, :] = np.nan
    is_max = seed == 5
    dr = np.arange(-5, -5 + D).astype(np.float32)
    gmin = gen.integers(-5, 0, (H, W))
    gmax = gen.integers(1, D - 6, (H, W))
    grids = np.array([gmin, gmax], dtype=np.int64)
    etas = np.arange(0.0, 0.7, 0.01) if seed % 2 else np.arange(0.0, 0.2, 0.1)
    src = -cv if is_max else cv
    exp_amb, exp_samp = oracle.ambiguity(src, etas, grids, dr, sampled=True)
    exp_risk = oracle.risk(src, exp_samp, etas, grids, dr, sampled=True)
    out = eng.confidence(dev(eng, cv), etas, grids, dr, is_max=is_max, sampled_ambiguity=True, risk=True, sampled_risk=True)
    np.testing.assert_array_equal(host(out["ambiguity"]), exp_amb)
    np.testing.assert_array_equal(host(out["sampled_ambiguity"]), exp_samp)
    for key, exp in zip(["risk_max", "risk_min", "disp_sup", "disp_inf", "sampled_risk_max", "sampled_risk_min"], exp_risk):
        np.testing.assert_array_equal(host(out[key]), exp)
    # ambiguity alone, whole range (grids = None == grids spanning every disparity)
    full = np.array([np.full((H, W), dr[0]), np.full((H, W), dr[-1])], dtype=np.int64)
    out2 = eng.confidence(dev(eng, cv), etas, None, dr, is_max=is_max)
    np.testing.assert_array_equal(host(out2["ambiguity"]), oracle.ambiguity(src, etas, full, dr))


# ---- masks / disparity grids --------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", range(6))
def test_validity_mask_with_masks_goldens(eng, oracle, goldens, case):
    """tests/test_criteria.py:723-1310 through the plugin-API mirror: validity_mask + SAD + cv_masked with image masks."""
    import pandora_b200 as pb

    t = f"test_criteria.py::TestCriteria.test_validity_mask[{case}]::"
    attrs = {"valid_pixels": int(goldens[t + "left_attrs.valid_pixels"]), "no_data_mask": int(goldens[t + "left_attrs.no_data_mask"])}
    dmin, dmax = (int(v) for v in goldens[t + "disparity"])
    left = pb.create_image_dataset(goldens[t + "left_data"], disparity=(dmin, dmax), msk=goldens[t + "left_msk"])
    right = pb.create_image_dataset(goldens[t + "right_data"], msk=goldens[t + "right_msk"])
    left.attrs.update(attrs)
    right.attrs.update(attrs)
    mc = pb.AbstractMatchingCost(matching_cost_method="sad", window_size=int(goldens[t + "window_size"]), subpix=1)
    grids = (left["disparity"].data[0], left["disparity"].data[1])
    cv = mc.allocate_cost_volume(left, grids)
    cv = pb.validity_mask(left, right, cv)
    cv = mc.compute_cost_volume(left, right, cv)
    mc.cv_masked(left, right, cv, *grids)
    np.testing.assert_array_equal(cv["validity_mask"].data, goldens[t + "gt_mask"])


@pytest.mark.parametrize("seed", range(5))
def test_cv_masked_vs_oracle(eng, oracle, seed):
    gen = np.random.default_rng(seed)
    H, W = 37, 71
    dmin, dmax = (-9, 6) if seed % 2 else (-20, -5)
    D = dmax - dmin + 1
    w = (1, 3, 5)[seed % 3]
    off = (w - 1) // 2
    left = gen.integers(0, 255, (H, W)).astype(np.float32)
    right = gen.integers(0, 255, (H, W)).astype(np.float32)
    lm = gen.choice([0, 0, 0, 0, 0, 0, 0, 1, 2, 5], (H, W)).astype(np.int16)
    rm = gen.choice([0, 0, 0, 0, 0, 0, 0, 1, 3], (H, W)).astype(np.int16)
    gmin = gen.integers(dmin, dmin + 4, (H, W)).astype(np.float32)
    gmax = gen.integers(dmax - 4, dmax + 1, (H, W)).astype(np.float32)
    gmin[0, 0], gmax[0, 1] = dmin, dmax
    cv, _ = oracle.sad_ssd_cost_volume(left, right, w, dmin, dmax, "sad")
    exp_cv = cv.copy()
    exp_vm = oracle.validity_mask_with_masks(H, W, dmin, dmax, off, w, lm, rm, 0, 1, gmin, gmax)
    oracle.cv_masked_full(exp_cv, exp_vm, off, w, dmin, lm, rm, 0, 1, gmin, gmax)
    fl, fr = eng.mask_flags(lm, 0, 1, w), eng.mask_flags(rm, 0, 1, w)
    mask = eng.validity_mask_init(H, W, dmin, dmax, off)
    eng.validity_mask_masks(mask, dmin, dmax, off, fl, fr, dev(eng, gmin), dev(eng, gmax))
    d_cv = dev(eng, cv)
    flags = eng.cv_masked(d_cv, dmin, fl, fr, dev(eng, gmin), dev(eng, gmax))
    mask = eng.validity_mask(H, W, dmin, dmax, off, flags, mask=mask)
    np.testing.assert_array_equal(host(d_cv), exp_cv)
    np.testing.assert_array_equal(host_mask(mask), exp_vm)


# ---- the extended run(): matching cost -> SGM -> WTA -> confidence -> refinement -> cross-checking ----------------
def test_run_with_next_rows(eng, oracle):
    import pandora_b200 as pb

    H, W, D = 48, 160, 32
    left_np, right_np, _ = oracle.synthetic_pair(H, W, D)
    dmin, dmax = -(D - 1), 0
    left = pb.create_image_dataset(left_np, disparity=(dmin, dmax))
    right = pb.create_image_dataset(right_np)
    cfg = {"pipeline": {"matching_cost": {"matching_cost_method": "census", "window_size": 5},
                        "optimization": {"optimization_method": "sgm"},
                        "cost_volume_confidence": {"confidence_method": "ambiguity", "normalization": False},
                        "disparity": {"disparity_method": "wta"},
                        "refinement": {"refinement_method": "vfit"},
                        "validation": {"validation_method": "cross_checking_fast", "cross_checking_threshold": 1.0}}}
    disp, cv = pb.run(left, right, cfg)
    # oracle chain
    ccv, attrs = oracle.census_cost_volume(left_np, right_np, 5, dmin, dmax)
    vm = oracle.validity_mask(H, W, dmin, dmax, 2)
    oracle.cv_masked(ccv, vm, 2)
    scv = oracle.sgm_cost_volume(ccv, 8, 32, cmax=attrs["cmax"])
    disps = np.arange(dmin, dmax + 1)
    ld, inv = oracle.wta(scv, disps)
    lmask = oracle.wta_validity_mask(vm, inv)
    rd, rinv = oracle.right_disparity_fast(scv, dmin, dmax)
    rmask = np.where(rinv, 0x3C3, 0).astype(np.uint16)
    grids = np.array([np.full((H, W), dmin), np.full((H, W), dmax)], dtype=np.int64)
    amb = oracle.ambiguity(scv, np.arange(0.0, 0.7, 0.01), grids, disps.astype(np.float32))
    itp, ld, lmask = oracle.refinement(scv, ld, lmask, dmin, dmax, 1, "min", "vfit")
    _, rd, rmask = oracle.refinement(oracle.reverse_cost_volume(scv, -dmax), rd, rmask, -dmax, -dmin, 1, "min", "vfit")
    exp_mask, exp_conf = oracle.cross_checking(ld, lmask, rd, 1.0, dmin, dmax, 2)
    np.testing.assert_array_equal(disp["disparity_map"].data, ld)
    np.testing.assert_array_equal(disp["validity_mask"].data, exp_mask)
    np.testing.assert_array_equal(disp["interpolated_coeff"].data, itp)
    conf = np.asarray(disp["confidence_measure"].data)
    assert list(disp.coords["indicator"].data) == ["confidence_from_ambiguity", "confidence_from_left_right_consistency"]
    np.testing.assert_array_equal(conf[:, :, 0], 1 - amb)
    np.testing.assert_array_equal(conf[:, :, 1], exp_conf)


def test_confidence_generic_inputs_sequential_sums(eng, oracle):
    """Non-dyadic disparity coordinates and an external sampled ambiguity take the eta-ordered float32 sums of
    risk.cpp:128-175 (no parallel reduction); unevenly spaced etas exercise the threshold search's correction loops."""
    gen = np.random.default_rng(11)
    H, W, D = 17, 29, 20
    cv = random_volume(11, H, W, D, nan_frac=0.1, integer=False, hi=30)
    dr = (np.arange(D) * 0.3 - 2.1).astype(np.float32)                      # not multiples of 1/16
    grids = np.array([np.full((H, W), -2), np.full((H, W), 3)], dtype=np.int64)
    etas = np.sort(gen.random(37) * 0.6)
    etas[0] = 0.0
    exp_amb, exp_samp = oracle.ambiguity(cv, etas, grids, dr, sampled=True)
    fake_samp = (exp_samp + gen.random(exp_samp.shape).astype(np.float32)).astype(np.float32)
    exp_risk = oracle.risk(cv, fake_samp, etas, grids, dr, sampled=True)
    out = eng.confidence(dev(eng, cv), etas, grids, dr, sampled_ambiguity=True, risk=True, sampled_risk=True, sampled_ambiguity_in=fake_samp)
    np.testing.assert_array_equal(host(out["ambiguity"]), exp_amb)
    np.testing.assert_array_equal(host(out["sampled_ambiguity"]), exp_samp)
    for key, exp in zip(["risk_max", "risk_min", "disp_sup", "disp_inf", "sampled_risk_max", "sampled_risk_min"], exp_risk):
        np.testing.assert_array_equal(host(out[key]), exp)
    # internal samples + non-dyadic disparities: still the sequential sums
    exp_risk2 = oracle.risk(cv, exp_samp, etas, grids, dr)
    out2 = eng.confidence(dev(eng, cv), etas, grids, dr, ambiguity=False, risk=True)
    for key, exp in zip(["risk_max", "risk_min", "disp_sup", "disp_inf"], exp_risk2):
        np.testing.assert_array_equal(host(out2[key]), exp)


# ---- reverse_disp_range (matching_cost.cpp:59-131) -------------------------------------------------------------------------
@pytest.mark.parametrize("seed,H,W,lo,hi", [(0, 9, 31, -7, 2), (1, 5, 64, -3, 3), (2, 7, 40, 2, 9), (3, 6, 50, -30, -20), (4, 12, 300, -120, 40),
                                            (5, 3, 1000, -260, 0)])
def test_reverse_disp_range_vs_oracle(eng, oracle, seed, H, W, lo, hi):
    gen = np.random.default_rng(seed)
    a = gen.integers(lo, hi + 1, (H, W)).astype(np.float32)
    b = a + gen.integers(0, 9, (H, W)).astype(np.float32)
    if seed % 2:
        a += 0.5
    a[gen.random((H, W)) < 0.1] = np.nan
    b[gen.random((H, W)) < 0.1] = np.nan
    if seed == 4:
        a[3], b[3] = np.nan, np.nan                                # a row nothing is seen from
    rmin, rmax = eng.reverse_disp_range(eng.to_device(a), eng.to_device(b))
    emin, emax = oracle.reverse_disp_range(a, b)
    np.testing.assert_array_equal(rmin.cpu().numpy(), emin)
    np.testing.assert_array_equal(rmax.cpu().numpy(), emax)


def test_reverse_disp_range_host_entry_and_mirror(eng, oracle):
    import ctypes

    import pandora_b200
    from pandora_b200.matching_cost import AbstractMatchingCost

    gen = np.random.default_rng(11)
    a = gen.integers(-9, 3, (10, 37)).astype(np.float32)
    b = a + gen.integers(0, 5, a.shape).astype(np.float32)
    emin, emax = oracle.reverse_disp_range(a, b)
    lib = pandora_b200.load()
    rmin, rmax = np.empty_like(a), np.empty_like(a)
    p = lambda x: x.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
    assert lib.pb200_reverse_disp_range_host(p(a), p(b), 10, 37, p(rmin), p(rmax)) == 0
    np.testing.assert_array_equal(rmin, emin)
    np.testing.assert_array_equal(rmax, emax)
    mmin, mmax = AbstractMatchingCost.reverse_disp_range(a, b)
    np.testing.assert_array_equal(mmin, emin)
    np.testing.assert_array_equal(mmax, emax)
