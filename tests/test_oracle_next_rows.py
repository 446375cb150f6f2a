"""Pins the oracle restatements of the SURVEY.md 8(f) rows (input masks + disparity grids, fast cross-checking,
sub-pixel refinement, cost-volume confidence) against the reference's own golden vectors
(tests/golden/reference_goldens.npz) and against the UNMODIFIED reference C++ compiled into oracle/_ref
(refinement_cpp, cost_volume_confidence_cpp, criteria_cpp).  CPU only."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REF = "test_refinement.py::TestRefinement."
VAL = "test_validation.py::TestValidation."
AMB = "test_confidence/test_ambiguity.py::"
RISK = "test_confidence/test_risk.py::"
CONF = "test_confidence/conftest.py::"
MC = "test_matching_cost/test_matching_cost.py::"


@pytest.fixture(scope="module")
def ref_next():
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import cost_volume_confidence_cpp  # noqa: PLC0415
        import criteria_cpp  # noqa: PLC0415
        import refinement_cpp  # noqa: PLC0415
    except ImportError:
        pytest.skip("oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")
    return refinement_cpp, cost_volume_confidence_cpp, criteria_cpp


# ---- refinement: tests/test_refinement.py ------------------------------------------------------------------
def _setup(goldens):
    return (goldens[REF + "setUp::self.cv@0"].astype(np.float32), goldens[REF + "setUp::self.disp@0"].astype(np.float32),
            goldens[REF + "setUp::self.disp@1"].astype(np.uint16))


@pytest.mark.parametrize("method,test", [("quadratic", "test_quadratic"), ("vfit", "test_vfit")])
def test_refinement_goldens(oracle, goldens, method, test):
    cv, disp, mask = _setup(goldens)                                     # disp coords [-2 .. 2], subpix 1
    itp, d, m = oracle.refinement(cv, disp, mask, -2, 2, 1, "min", method)
    np.testing.assert_allclose(d, goldens[REF + test + "::gt_sub_disp"], rtol=1e-6)
    np.testing.assert_allclose(itp, goldens[REF + test + "::gt_sub_cost"], rtol=1e-6)
    np.testing.assert_array_equal(m, goldens[REF + test + "::gt_mask"])


@pytest.mark.parametrize("method,test", [("quadratic", "test_quadratic_subpix"), ("vfit", "test_vfit_subpix"),
                                          ("quadratic", "test_quadratic_with_nan_and_subpix"), ("vfit", "test_vfit_with_nan_and_subpix")])
def test_refinement_subpix_goldens(oracle, goldens, method, test):
    cv = goldens[REF + test + "::subpix_cv@0"].astype(np.float32)        # disp coords [-4, -3.5, ..., -2] style: subpix 2
    disp = goldens[REF + test + "::subpix_disp@0"].astype(np.float32)
    mask = goldens[REF + test + "::subpix_disp@1"].astype(np.uint16)
    itp, d, m = oracle.refinement(cv, disp, mask, -1, 1, 2, "min", method)
    np.testing.assert_allclose(itp, goldens[REF + test + "::gt_sub_cost"], rtol=1e-6)
    np.testing.assert_array_equal(m, goldens[REF + test + "::gt_mask"])


def test_refinement_vfit_nan_golden(oracle, goldens):
    cv = goldens[REF + "test_vfit_with_nan::cv@0"].astype(np.float32)
    disp = goldens[REF + "test_vfit_with_nan::disp@0"].astype(np.float32)
    mask = goldens[REF + "test_vfit_with_nan::disp@1"].astype(np.uint16)
    itp, d, m = oracle.refinement(cv, disp, mask, -1, 1, 1, "min", "vfit")
    np.testing.assert_allclose(d, goldens[REF + "test_vfit_with_nan::gt_sub_disp"], rtol=1e-6)
    np.testing.assert_allclose(itp, goldens[REF + "test_vfit_with_nan::gt_sub_cost"], rtol=1e-6)
    np.testing.assert_array_equal(m, goldens[REF + "test_vfit_with_nan::gt_mask"])


def test_approximate_refinement_golden(oracle, goldens):
    t = REF + "test_vfit_approximate_subpixel_refinement::"
    cv = goldens[t + "cv_left@0"].astype(np.float32)                     # disp coords [-3 .. 2]
    disp = goldens[t + "disp_right@0"].astype(np.float32)
    mask = goldens[t + "disp_right@1"].astype(np.uint16)
    itp, d, m = oracle.refinement(cv, disp, mask, -3, 2, 1, "min", "vfit", approximate=True)
    np.testing.assert_allclose(d, goldens[t + "gt_sub_disp"], rtol=1e-6)
    np.testing.assert_allclose(itp, goldens[t + "gt_sub_costs"], rtol=1e-6)
    np.testing.assert_array_equal(m, goldens[t + "gt_mask"])


def _random_refinement_case(seed, H=23, W=37, D=11, subpix=1):
    gen = np.random.default_rng(seed)
    cv = gen.integers(0, 40, (H, W, D)).astype(np.float32)
    if seed % 2:
        cv += gen.random((H, W, D)).astype(np.float32)
    cv[gen.random(cv.shape) < 0.08] = np.nan
    d_min = -4.0
    d_max = d_min + (D - 1) / subpix
    idx = np.argmin(np.where(np.isnan(cv), np.inf, cv), axis=2)
    disp = (d_min + idx / subpix).astype(np.float32)
    mask = np.zeros((H, W), dtype=np.uint16)
    mask[gen.random((H, W)) < 0.1] = 2                                 # invalid pixels
    mask[gen.random((H, W)) < 0.1] |= 4                                # information bit only
    return cv, disp, mask, d_min, d_max


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("method", ["vfit", "quadratic"])
@pytest.mark.parametrize("measure", ["min", "max"])
def test_refinement_vs_reference_cpp(oracle, ref_next, seed, method, measure):
    rc = ref_next[0]
    subpix = (1, 2, 4)[seed % 3]
    cv, disp, mask, d_min, d_max = _random_refinement_case(seed, subpix=subpix)
    if measure == "max":
        cv = -cv
    fn = (lambda c, d, m: rc.vfit_refinement_method(c, d, m, 8)) if method == "vfit" else (lambda c, d, m: rc.quadratic_refinement_method(c, d, m, 8))
    ref_itp, ref_disp, ref_mask = rc.loop_refinement(cv, disp.copy(), mask.astype(np.int64), d_min, d_max, subpix, measure, fn, 0x3C3, 8)
    itp, d, m = oracle.refinement(cv, disp, mask, d_min, d_max, subpix, measure, method)
    np.testing.assert_array_equal(itp, ref_itp)
    np.testing.assert_array_equal(d, ref_disp)
    np.testing.assert_array_equal(m, ref_mask.astype(np.uint16))


@pytest.mark.parametrize("seed", range(4))
def test_approximate_refinement_vs_reference_cpp(oracle, ref_next, seed):
    rc = ref_next[0]
    gen = np.random.default_rng(100 + seed)
    H, W, D = 9, 41, 9
    d_min, d_max = -5.0, 3.0
    cv = gen.integers(0, 50, (H, W, D)).astype(np.float32)
    cv[gen.random(cv.shape) < 0.05] = np.nan
    # right disparities whose diagonal stays inside the image (col + raw in [0, W)) and whose neighbours dsp +- 1 stay
    # inside the disparity axis: the reference compares raw with the LEFT bounds (refinement.cpp:147-150), so raw =
    # -d_max / -d_min would make it read out of bounds (undefined there; the oracle answers NaN)
    raw = gen.integers(-2, 5, (H, W)).astype(np.float32)
    cols = np.arange(W)[None, :]
    mask = np.where((cols + raw < 0) | (cols + raw >= W), 2, 0).astype(np.uint16)
    fn = lambda c, d, m: rc.vfit_refinement_method(c, d, m, 8)  # noqa: E731
    ref_itp, ref_disp, ref_mask = rc.loop_approximate_refinement(cv, raw.copy(), mask.astype(np.int64), d_min, d_max, 1, "min", fn, 0x3C3, 8)
    itp, d, m = oracle.refinement(cv, raw, mask, d_min, d_max, 1, "min", "vfit", approximate=True)
    np.testing.assert_array_equal(itp, ref_itp)
    np.testing.assert_array_equal(d, ref_disp)
    np.testing.assert_array_equal(m, ref_mask.astype(np.uint16))


# ---- cross-checking: tests/test_validation.py ----------------------------------------------------------------
def test_cross_checking_golden(oracle, goldens):
    left, lmask = goldens[VAL + "setUp::self.left@0"], goldens[VAL + "setUp::self.left@2"]
    right = goldens[VAL + "setUp::self.right@0"]
    vm, conf = oracle.cross_checking(left, lmask, right, 0.0, -2, 2, 0)
    np.testing.assert_array_equal(vm, goldens[VAL + "test_cross_checking::gt_mask"])
    np.testing.assert_array_equal(conf, goldens[VAL + "test_cross_checking::gt_dist"][:, :, 1])


def test_cross_checking_float_golden(oracle, goldens):
    t = VAL + "test_cross_checking_float_disparity::"
    vm, _ = oracle.cross_checking(goldens[t + "left@0"], goldens[t + "left@2"], goldens[t + "right@0"], 0.0, -2, 2, 0)
    np.testing.assert_array_equal(vm, goldens[t + "gt_mask"])


def test_cross_checking_distance_golden(oracle, goldens):
    t = VAL + "test_distance_lr_rl::"
    _, conf = oracle.cross_checking(goldens[t + "left@0"], goldens[t + "left@2"], goldens[t + "right@0"], 0.0, -2, 2, 0)
    np.testing.assert_array_equal(conf, goldens[t + "gt_dist"][:, :, 1])


def test_right_disparity_fast_is_reverse_then_wta(oracle):
    gen = np.random.default_rng(3)
    cv = gen.integers(0, 9, (7, 19, 6)).astype(np.float32)
    cv[gen.random(cv.shape) < 0.1] = np.nan
    disp, inv = oracle.right_disparity_fast(cv, -4, 1)
    ref = oracle.reverse_cost_volume(cv, -1)
    exp, einv = oracle.wta(ref, np.arange(-1, 5))
    np.testing.assert_array_equal(disp, exp)
    np.testing.assert_array_equal(inv, einv)


# ---- confidence: tests/test_confidence ------------------------------------------------------------------------
def test_ambiguity_goldens(oracle, goldens):
    t = AMB + "test_compute_ambiguity_and_sampled_ambiguity::"
    amb, samp = oracle.ambiguity(goldens[t + "cv_"], goldens[t + "etas"], goldens[t + "grids"], goldens[t + "disparity_range"], sampled=True)
    np.testing.assert_allclose(amb, goldens[t + "gt_amb_int"], rtol=1e-6)
    np.testing.assert_allclose(samp, goldens[t + "gt_sam_amb"], rtol=1e-6)
    t = AMB + "test_compute_compute_ambiguity_and_sampled_ambiguity_with_variable_disparity::"
    cv = goldens[CONF + "create_cv_for_variable_disparities::cv_#1"]
    grids = goldens[CONF + "create_grids_and_disparity_range_with_variable_disparities::grids"]
    dr = goldens[CONF + "create_grids_and_disparity_range_with_variable_disparities::disparity_range"]
    amb, samp = oracle.ambiguity(cv, goldens[t + "etas"], grids, dr, sampled=True)
    np.testing.assert_allclose(amb, goldens[t + "gt_amb_int"], rtol=1e-6)
    np.testing.assert_allclose(samp, goldens[t + "gt_sampl_amb"], rtol=1e-6)


def test_risk_goldens(oracle, goldens):
    t = RISK + "test_compute_risk_and_sampled_risk::"
    out = oracle.risk(goldens[t + "cv_"], goldens[t + "sampled_ambiguity"], goldens[t + "etas"], goldens[t + "grids"],
                      goldens[t + "disparity_range"], sampled=True)
    for got, key in zip(out, ["gt_risk_max", "gt_risk_min", "gt_disp_sup", "gt_disp_inf", "gt_sampled_risk_max", "gt_sampled_risk_min"]):
        np.testing.assert_allclose(got, goldens[t + key], rtol=1e-6)
    t = RISK + "test_compute_risk_with_variable_disparity::"
    cv = goldens[CONF + "create_cv_for_variable_disparities::cv_#1"]
    grids = goldens[CONF + "create_grids_and_disparity_range_with_variable_disparities::grids"]
    dr = goldens[CONF + "create_grids_and_disparity_range_with_variable_disparities::disparity_range"]
    out = oracle.risk(cv, goldens[t + "amb_sampl"], goldens[t + "etas"], grids, dr)
    for got, key in zip(out, ["gt_risk_max", "gt_risk_min", "gt_disp_sup", "gt_disp_inf"]):
        np.testing.assert_allclose(got, goldens[t + key], rtol=1e-6)


@pytest.mark.parametrize("seed", range(5))
def test_confidence_vs_reference_cpp(oracle, ref_next, seed):
    cc = ref_next[1]
    gen = np.random.default_rng(seed)
    H, W, D = 13, 17, 12
    cv = gen.integers(0, 60, (H, W, D)).astype(np.float32)
    if seed % 2:
        cv += gen.random(cv.shape).astype(np.float32)
    cv[gen.random(cv.shape) < 0.15] = np.nan
    cv[0, :3, :] = np.nan
    dr = np.arange(-5, -5 + D).astype(np.float32)
    gmin = gen.integers(-5, 0, (H, W))
    gmax = gen.integers(1, 7, (H, W))
    grids = np.array([gmin, gmax], dtype=np.int64)
    etas = np.arange(0.0, 0.7, 0.01)
    ref_amb, ref_samp = cc.compute_ambiguity_and_sampled_ambiguity(cv, etas, len(etas), grids, dr, True)
    amb, samp = oracle.ambiguity(cv, etas, grids, dr, sampled=True)
    np.testing.assert_array_equal(amb, ref_amb)
    np.testing.assert_array_equal(samp, ref_samp)
    ref = cc.compute_risk_and_sampled_risk(cv, ref_samp, etas, len(etas), grids, dr, True)
    out = oracle.risk(cv, samp, etas, grids, dr, sampled=True)
    for got, exp in zip(out, ref):
        np.testing.assert_array_equal(got, exp)


# ---- masks / disparity grids: tests/test_matching_cost/test_matching_cost.py:699-1786, tests/test_criteria.py -----
def _image(goldens, name):
    return goldens[MC + f"{name}::data"].astype(np.float32)


@pytest.mark.parametrize("case", range(4))
def test_cv_masked_pixellic_goldens(oracle, goldens, case):
    t = MC + f"TestCvMasked.test_pixellic[{case}]::"
    left_name = bytes(goldens[t + "make_cv_masked_parameters.left_image"]).decode()
    right_name = bytes(goldens[t + "make_cv_masked_parameters.right_image"]).decode()
    left, right = _image(goldens, "TestCvMasked." + left_name), _image(goldens, "TestCvMasked." + right_name)
    w = int(goldens[t + "make_cv_masked_parameters.cfg.window_size"])
    lm, rm = goldens[t + "make_cv_masked_parameters.left_mask"], goldens[t + "make_cv_masked_parameters.right_mask"]
    H, W = left.shape
    cv, _ = oracle.census_cost_volume(left, right, w, -1, 1)             # the fixture uses census-free "ssd"/census alike NaN layout
    vm = oracle.validity_mask_with_masks(H, W, -1, 1, (w - 1) // 2, w, lm, rm)
    oracle.cv_masked_full(cv, vm, (w - 1) // 2, w, -1, lm, rm)
    np.testing.assert_array_equal(np.isnan(cv), goldens[t + "expected_nan_mask"])


def test_cv_masked_window1_golden(oracle, goldens):
    t = MC + "TestCvMaskedWithWindowSizeOf1.test_pixellic[0]::"
    left, right = _image(goldens, "TestCvMaskedWithWindowSizeOf1.left_2x5"), _image(goldens, "TestCvMaskedWithWindowSizeOf1.right_2x5")
    lm, rm = goldens[t + "make_cv_masked_parameters.left_mask"], goldens[t + "make_cv_masked_parameters.right_mask"]
    cv, _ = oracle.sad_ssd_cost_volume(left, right, 1, -1, 1, "sad")
    vm = oracle.validity_mask_with_masks(2, 5, -1, 1, 0, 1, lm, rm)
    oracle.cv_masked_full(cv, vm, 0, 1, -1, lm, rm)
    np.testing.assert_array_equal(np.isnan(cv), goldens[t + "expected_nan_mask"])


def test_cv_masked_grid_golden(oracle, goldens):
    t = MC + "TestCvMaskedWithGrid."
    left = goldens[t + "left_with_disparity_grid_4x11::data"].astype(np.float32)
    right = goldens[t + "right_without_disparity_4x11::data"].astype(np.float32)
    gmin, gmax = goldens[t + "disparity_grid_4x11::dmin_grid"], goldens[t + "disparity_grid_4x11::dmax_grid"]
    w = int(goldens[t + "test_pixellic[0]::make_cv_masked_parameters.cfg.window_size"])
    dmin, dmax = int(gmin.min()), int(gmax.max())
    cv, _ = oracle.census_cost_volume(left, right, w, dmin, dmax) if w in (3, 5) else oracle.sad_ssd_cost_volume(left, right, w, dmin, dmax)
    vm = oracle.validity_mask(4, 11, dmin, dmax, (w - 1) // 2)
    oracle.cv_masked_full(cv, vm, (w - 1) // 2, w, dmin, grid_min=gmin, grid_max=gmax)
    np.testing.assert_array_equal(np.isnan(cv), goldens[t + "test_pixellic[0]::expected_nan_mask"])


@pytest.mark.parametrize("seed", range(5))
def test_partially_missing_vs_reference_cpp(oracle, ref_next, seed):
    crit = ref_next[2]
    gen = np.random.default_rng(seed)
    H, W = 9, 40
    gmin = gen.integers(-6, 0, (H, W)).astype(np.float32)
    gmax = (gmin + gen.integers(0, 8, (H, W))).astype(np.float32)
    invalid = gen.random((H, W)) < 0.12
    ref = crit.partially_missing_variable_ranges(np.array([gmin, gmax]), invalid)
    np.testing.assert_array_equal(oracle.partially_missing_variable_ranges(gmin, gmax, invalid), ref)


def test_binary_dilation_matches_scipy(oracle):
    from scipy.ndimage import binary_dilation  # noqa: PLC0415

    gen = np.random.default_rng(0)
    for w in (1, 3, 5, 7):
        msk = (gen.random((17, 23)) < 0.06).astype(np.int16)
        ref = binary_dilation(msk == 1, structure=np.ones((w, w)), iterations=1)
        np.testing.assert_array_equal(oracle.binary_dilation_nodata(msk, 1, w), ref)


@pytest.mark.parametrize("case", range(6))
def test_validity_mask_with_masks_goldens(oracle, goldens, case):
    """tests/test_criteria.py:723-1310: validity_mask + SAD compute_cost_volume + cv_masked with left / right masks."""
    t = f"test_criteria.py::TestCriteria.test_validity_mask[{case}]::"
    left, right = goldens[t + "left_data"].astype(np.float32), goldens[t + "right_data"].astype(np.float32)
    lm, rm = goldens[t + "left_msk"], goldens[t + "right_msk"]
    valid, nodata = int(goldens[t + "left_attrs.valid_pixels"]), int(goldens[t + "left_attrs.no_data_mask"])
    dmin, dmax = (int(v) for v in goldens[t + "disparity"])
    w = int(goldens[t + "window_size"])
    H, W = left.shape
    off = (w - 1) // 2
    gmin, gmax = np.full((H, W), dmin), np.full((H, W), dmax)
    vm = oracle.validity_mask_with_masks(H, W, dmin, dmax, off, w, lm, rm, valid, nodata, gmin, gmax)
    cv, _ = oracle.sad_ssd_cost_volume(left, right, w, dmin, dmax, "sad")
    oracle.cv_masked_full(cv, vm, off, w, dmin, lm, rm, valid, nodata, gmin, gmax)
    np.testing.assert_array_equal(vm, goldens[t + "gt_mask"])


def test_sgm_min_cost_paths_known_answers(oracle):
    """min_cost_paths (plugin_libsgm.rst:411-413).  A single pixel: every path starts there, L_r = C for all eight, so all eight
    agree with the sum; a pixel without any valid cost counts 0; the optimised volume is the one of the plain call."""
    cv = np.array([[[5, 2, 7, 2]]], dtype=np.float32)
    out, nb = oracle.sgm_min_cost_paths(cv, 8, 32, cmax=25)
    np.testing.assert_array_equal(out, 8 * cv)
    assert nb[0, 0] == 8.0
    g = np.random.default_rng(4)
    cv = g.integers(0, 26, (6, 8, 5)).astype(np.float32)
    cv[1, 2, :] = np.nan
    cv[3, 3, 1] = np.nan
    out, nb = oracle.sgm_min_cost_paths(cv, 8, 32, cmax=25)
    np.testing.assert_array_equal(out, oracle.sgm_cost_volume(cv, 8, 32, cmax=25))
    assert nb[1, 2] == 0.0 and nb.max() <= 8.0 and nb.min() >= 0.0 and nb.dtype == np.float32
    # one row: E and W are the only paths longer than a pixel; at the first column W has seen the whole row, the other seven
    # directions start there (L_r = C)
    row = g.integers(0, 26, (1, 9, 4)).astype(np.float32)
    out, nb = oracle.sgm_min_cost_paths(row, 8, 32, cmax=25)
    k_c, k_s = int(np.argmin(row[0, 0])), int(np.argmin(out[0, 0]))
    assert nb[0, 0] >= (7.0 if k_c == k_s else 0.0)
