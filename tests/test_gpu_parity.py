"""GPU parity tests: every CUDA kernel, called through the C-ABI (ctypes -> libpandora_b200.so), against
the CPU oracle on the same seeded inputs, and against the reference's golden vectors.

Bar: bit-exact (np.testing.assert_array_equal, NaN == NaN) for integer / index work and for every
float path whose operation order is reproduced (SAD/SSD, SGM); stated tolerances otherwise (ZNCC,
CBCA on float costs).  Run on the B200 box with ``pytest -m gpu``.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CENSUS = "test_matching_cost/test_matching_cost_census.py"
SAD = "test_matching_cost/test_matching_cost_sad.py"
AGG = "test_aggregation.py"
DISP = "test_disparity.py"


@pytest.fixture(scope="module")
def eng():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pandora_b200

    return pandora_b200.get_engine("cuda:0")


def dev(eng, a):
    return eng.to_device(np.ascontiguousarray(a, dtype=np.float32))


def host(t):
    return t.detach().cpu().numpy()


def rand_pair(seed, H, W, levels=256, as_float=False):
    g = np.random.default_rng(seed)
    if as_float:
        return (g.random((H, W)) * 255).astype(np.float32), (g.random((H, W)) * 255).astype(np.float32)
    return g.integers(0, levels, (H, W)).astype(np.float32), g.integers(0, levels, (H, W)).astype(np.float32)


# ------------------------------------------------------------------------------------------------
# Census
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("w", [3, 5, 7, 9, 11, 13])
@pytest.mark.parametrize("shape,rng", [((37, 53), (-9, 6)), ((29, 64), (-40, -3)), ((40, 45), (2, 33)), ((16, 20), (-30, 30))])
def test_census_vs_oracle(eng, oracle, w, shape, rng):
    left, right = rand_pair(w * 1000 + shape[1], *shape, levels=7)       # few levels: ties exercise the strict '>'
    dmin, dmax = rng
    ref, _ = oracle.census_cost_volume(left, right, w, dmin, dmax)
    got = host(eng.census(dev(eng, left), dev(eng, right), w, dmin, dmax))
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("D", [1, 3, 61, 64, 100, 128, 256, 300])
def test_census_disparity_counts(eng, oracle, D):
    left, right = rand_pair(D, 21, 333)
    dmin = -(D - 1) + 2
    ref, _ = oracle.census_cost_volume(left, right, 5, dmin, dmin + D - 1)
    got = host(eng.census(dev(eng, left), dev(eng, right), 5, dmin, dmin + D - 1))
    np.testing.assert_array_equal(got, ref)


def test_census_float_images_and_tiny(eng, oracle):
    left, right = rand_pair(5, 33, 47, as_float=True)
    ref, _ = oracle.census_cost_volume(left, right, 5, -11, 4)
    np.testing.assert_array_equal(host(eng.census(dev(eng, left), dev(eng, right), 5, -11, 4)), ref)
    # image smaller than the window: everything NaN
    left, right = rand_pair(6, 4, 4)
    got = host(eng.census(dev(eng, left), dev(eng, right), 5, -1, 1))
    assert np.isnan(got).all()


def test_census_fused_wta(eng, oracle):
    for (H, W, dmin, dmax, w) in [(45, 130, -63, 0, 5), (33, 77, -5, 9, 3), (20, 50, 3, 40, 7)]:
        left, right = rand_pair(H * W, H, W, levels=5)
        ref, _ = oracle.census_cost_volume(left, right, w, dmin, dmax)
        exp_disp, exp_inv = oracle.wta(ref, oracle.disparity_range(dmin, dmax), "min", -9999)
        cv, disp, flags = eng.census(dev(eng, left), dev(eng, right), w, dmin, dmax, fuse_wta=True)
        np.testing.assert_array_equal(host(cv), ref)
        np.testing.assert_array_equal(host(disp), exp_disp)
        np.testing.assert_array_equal(host(flags).astype(bool), exp_inv)


@pytest.mark.parametrize("case", range(7))
def test_census_reference_goldens(eng, goldens, case):
    """tests/test_matching_cost/test_matching_cost_census.py:379-729 through the CUDA path."""
    k = f"{CENSUS}::test_census[{case}]::"
    w = int(goldens[k + "window_size"])
    dmin, dmax = (int(v) for v in goldens[k + "disp_interval"])
    got = host(eng.census(dev(eng, goldens[k + "left_data"]), dev(eng, goldens[k + "right_data"]), w, dmin, dmax))
    np.testing.assert_array_equal(got[:, :, int(goldens[k + "tested_layer"])], goldens[k + "ref_out"])


def test_census_cost_golden_w3(eng, goldens):
    k = f"{CENSUS}::test_census_cost::"
    got = host(eng.census(dev(eng, goldens[k + "data"]), dev(eng, goldens[k + "data#1"]), 3, -1, 1))
    for i, name in enumerate(["census_ground_truth_d1", "census_ground_truth_d2", "census_ground_truth_d3"]):
        np.testing.assert_array_equal(got[:, :, i], goldens[k + name])


# ------------------------------------------------------------------------------------------------
# SAD / SSD / ZNCC
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", ["sad", "ssd"])
@pytest.mark.parametrize("w", [1, 3, 5, 7])
@pytest.mark.parametrize("as_float", [False, True])
def test_sad_ssd_vs_oracle(eng, oracle, method, w, as_float):
    left, right = rand_pair(w + 10 * as_float, 23, 41, as_float=as_float)
    ref, _ = oracle.sad_ssd_cost_volume(left, right, w, -7, 5, method)
    got = host(eng.sad_ssd(dev(eng, left), dev(eng, right), w, -7, 5, squared=(method == "ssd")))
    np.testing.assert_array_equal(got, ref)      # bit-exact even for float images: same summation order as numpy


def test_sad_reference_goldens(eng, goldens):
    left = goldens["common.py::matching_cost_tests_setup::data"]
    right = goldens["common.py::matching_cost_tests_setup::data#1"]
    k = f"{SAD}::TestMatchingCostSAD.test_sad_cost::"
    np.testing.assert_array_equal(host(eng.sad_ssd(dev(eng, left), dev(eng, right), 1, -1, 1))[:, :, 1], goldens[k + "ad_ground_truth"])
    np.testing.assert_array_equal(host(eng.sad_ssd(dev(eng, left), dev(eng, right), 5, -1, 1))[:, :, 1], goldens[k + "sad_ground_truth"])
    k = f"{SAD}::TestMatchingCostSAD.test_cost_volume::"
    got = host(eng.sad_ssd(dev(eng, goldens[k + "data"]), dev(eng, goldens[k + "data#1"]), 3, -2, 1))
    np.testing.assert_array_equal(got, goldens[k + "ground_truth"])


@pytest.mark.parametrize("method", ["sad", "ssd"])
@pytest.mark.parametrize("H,W,dmin,dmax,w", [(90, 75, -40, 3, 5), (70, 50, -150, 20, 3), (41, 37, -5, 70, 7), (37, 33, -3, 3, 13),
                                             (20, 18, 2, 9, 1), (9, 70, -20, 0, 11)])
def test_sad_ssd_running_sums(eng, oracle, method, H, W, dmin, dmax, w):
    """Integer images take the separable running-sum kernel (several row blocks, bands, disparity chunks, ranges leaving
    the image on both sides): bit-exact vs the oracle and vs the tap-ordered kernel."""
    import pandora_b200

    left, right = rand_pair(H + w, H, W)
    ref, _ = oracle.sad_ssd_cost_volume(left, right, w, dmin, dmax, method)
    got = host(eng.sad_ssd(dev(eng, left), dev(eng, right), w, dmin, dmax, squared=(method == "ssd")))
    assert pandora_b200.last_path("sad")[0] == "sad_running"
    np.testing.assert_array_equal(got, ref)
    with pandora_b200.option("sad.taps", 1):
        taps = host(eng.sad_ssd(dev(eng, left), dev(eng, right), w, dmin, dmax, squared=(method == "ssd")))
        assert pandora_b200.last_path("sad")[0] == "sad_taps"
    np.testing.assert_array_equal(taps, ref)


@pytest.mark.parametrize("case", ["fraction", "nan", "large", "negative", "float16bit"])
def test_sad_running_sums_data_condition(eng, oracle, case):
    """A CTA that stages anything but a small integer falls back to the reference's tap order before using the value:
    one fractional / NaN / huge pixel in the middle of a band, negative integers (still exact), 16-bit data."""
    left, right = rand_pair(77, 120, 64)
    if case == "fraction":
        left[70, 30] = 17.25
        right[33, 11] = 0.5
    elif case == "nan":
        right[64, 20] = np.nan
    elif case == "large":
        left[50, 40] = 3.0e7
    elif case == "negative":
        left -= 100.0
        right -= 60.0
    else:
        left, right = rand_pair(78, 120, 64, levels=65536)
    import pandora_b200

    for method in ("sad", "ssd"):
        if case == "nan":                         # the oracle's cmax attribute cannot take a NaN image: tap-ordered kernel as the checker
            with pandora_b200.option("sad.taps", 1):
                ref = host(eng.sad_ssd(dev(eng, left), dev(eng, right), 5, -30, 4, squared=(method == "ssd")))
            assert np.isnan(ref[64, 50, 0]) and np.isfinite(ref[64, 56, 0])
        else:
            ref, _ = oracle.sad_ssd_cost_volume(left, right, 5, -30, 4, method)
        got = host(eng.sad_ssd(dev(eng, left), dev(eng, right), 5, -30, 4, squared=(method == "ssd")))
        np.testing.assert_array_equal(got, ref)


def test_zncc_running_sums(eng, oracle):
    left, right = rand_pair(5, 100, 90)
    for w, dmin, dmax in ((5, -70, 3), (3, -10, 140), (9, -4, 4)):
        ref, _ = oracle.zncc_cost_volume(left, right, w, dmin, dmax)
        got = host(eng.zncc(dev(eng, left), dev(eng, right), w, dmin, dmax))
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("w", [1, 3, 5, 9])
@pytest.mark.parametrize("as_float", [False, True])
def test_zncc_vs_oracle(eng, oracle, w, as_float):
    left, right = rand_pair(w + 50 * as_float, 25, 38, as_float=as_float)
    ref, _ = oracle.zncc_cost_volume(left, right, w, -6, 4)
    got = host(eng.zncc(dev(eng, left), dev(eng, right), w, -6, 4))
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)     # tolerance stated by north_star: 1e-5 relative


# ------------------------------------------------------------------------------------------------
# WTA, validity mask, reverse
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D", [1, 5, 61, 64, 256, 300])
@pytest.mark.parametrize("mode", ["min", "max"])
def test_wta_vs_oracle(eng, oracle, D, mode):
    g = np.random.default_rng(D)
    cv = g.integers(0, 6, (19, 23, D)).astype(np.float32)           # many ties
    cv[g.random(cv.shape) < 0.3] = np.nan
    cv[0, 0, :] = np.nan                                            # all-NaN pixel
    cv[0, 1, :] = np.inf if mode == "min" else -np.inf              # all-inf pixel -> index 0
    cv[0, 2, : max(1, D // 2)] = np.nan
    disps = np.arange(-3, -3 + D)
    exp, exp_inv = oracle.wta(cv, disps, mode, -9999)
    disp, flags = eng.wta(dev(eng, cv), -3, mode == "max", -9999)
    np.testing.assert_array_equal(host(disp), exp)
    np.testing.assert_array_equal(host(flags).astype(bool), exp_inv)


def test_validity_mask_vs_oracle(eng, oracle):
    for (dmin, dmax, off) in [(-4, 3, 2), (-9, -2, 1), (2, 8, 2), (-3, 3, 0)]:
        left, right = rand_pair(100 + off + dmax, 17, 31)
        w = 2 * off + 1
        cv = oracle.census_cost_volume(left, right, w, dmin, dmax)[0] if w >= 3 else oracle.sad_ssd_cost_volume(left, right, 1, dmin, dmax)[0]
        vm = oracle.validity_mask(17, 31, dmin, dmax, off)
        oracle.cv_masked(cv, vm, off)
        _, inv = oracle.wta(cv, np.arange(dmin, dmax + 1))
        exp = oracle.wta_validity_mask(vm, inv)
        _, flags = eng.wta(dev(eng, cv), dmin)
        m1 = eng.validity_mask(17, 31, dmin, dmax, off, flags)
        np.testing.assert_array_equal(host(m1).view(np.uint16), vm)
        m2 = eng.validity_mask(17, 31, dmin, dmax, 0, flags, wta_invalidate=True, mask=m1.clone())
        np.testing.assert_array_equal(host(m2).view(np.uint16), exp)


def test_reverse_cost_volume(eng, oracle):
    g = np.random.default_rng(3)
    cv = g.random((7, 29, 11)).astype(np.float32)
    cv[g.random(cv.shape) < 0.2] = np.nan
    for md in (-10, -4, 0, 5):
        np.testing.assert_array_equal(host(eng.reverse_cost_volume(dev(eng, cv), md)), oracle.reverse_cost_volume(cv, md))


@pytest.mark.parametrize("shape,md", [((5, 70, 64), -63), ((3, 100, 37), -20), ((4, 33, 256), -255), ((2, 500, 128), 6), ((6, 45, 8), -3),
                                      ((3, 64, 5), 0)])
def test_reverse_cost_volume_tiled(eng, oracle, shape, md):
    """The tiled kernel (D >= 8): disparity counts that are odd / not multiples of 4, partial tiles, ranges on either side;
    bit-identical to the oracle (matching_cost.cpp:26-57) and to the plain gather kernel."""
    g = np.random.default_rng(shape[1] + shape[2])
    cv = g.random(shape).astype(np.float32)
    cv[g.random(shape) < 0.2] = np.nan
    got = host(eng.reverse_cost_volume(dev(eng, cv), md))
    np.testing.assert_array_equal(got, oracle.reverse_cost_volume(cv, md))
    import pandora_b200 as pb

    assert pb.last_path("reverse")[0] == ("reverse_tiled" if shape[2] >= 8 else "reverse_gather")
    with pb.option("reverse.gather", 1):
        np.testing.assert_array_equal(host(eng.reverse_cost_volume(dev(eng, cv), md)), got)
        assert pb.last_path("reverse")[0] == "reverse_gather"


# ------------------------------------------------------------------------------------------------
# CBCA
# ------------------------------------------------------------------------------------------------
def test_median3_vs_oracle(eng, oracle):
    g = np.random.default_rng(4)
    img = (g.random((41, 37)) * 100).astype(np.float32)
    img[g.random(img.shape) < 0.25] = np.nan
    np.testing.assert_array_equal(host(eng.median3(dev(eng, img))), oracle.median_filter3(img))
    img = g.integers(0, 255, (20, 33)).astype(np.float32)
    np.testing.assert_array_equal(host(eng.median3(dev(eng, img))), oracle.median_filter3(img))


@pytest.mark.parametrize("arms,tau,off", [(3, 5.0, 0), (5, 30.0, 2), (9, 2.5, 1), (1, 4.0, 0), (17, 50.0, 3)])
def test_cross_support_vs_oracle(eng, oracle, arms, tau, off):
    g = np.random.default_rng(arms)
    img = g.integers(0, 60, (27, 35)).astype(np.float32)
    img[g.random(img.shape) < 0.05] = np.nan
    ref_in = np.nan_to_num(img.copy(), nan=np.inf)
    if off:
        ref_in = np.ascontiguousarray(ref_in[off:-off, off:-off])
    got = host(eng.cross_support(dev(eng, img), arms, tau, off, nan_as_inf=True))
    np.testing.assert_array_equal(got, oracle.cross_support(ref_in, arms, tau))


def test_cross_support_golden(eng, goldens):
    left = goldens[f"{AGG}::TestAggregation.setUp::data"]
    k = f"{AGG}::TestAggregation.test_cross_support_region::csr_ground_truth_"
    csr = host(eng.cross_support(dev(eng, left), 3, 5.0))
    for i, name in enumerate(["left_arm", "right_arm", "top_arm", "bottom_arm"]):
        np.testing.assert_array_equal(csr[:, :, i], goldens[k + name])


@pytest.mark.parametrize("cfg", [(31, 45, -9, 4, 5, 5, 30.0), (40, 33, -20, -1, 3, 3, 12.0), (25, 60, 0, 37, 5, 9, 40.0),
                                 (50, 41, -3, 3, 7, 17, 25.0), (12, 14, -2, 2, 3, 5, 1000.0)])
def test_cbca_census_vs_oracle(eng, oracle, cfg):
    H, W, dmin, dmax, w, dist, tau = cfg
    left, right = rand_pair(H + W, H, W)
    cv, attrs = oracle.census_cost_volume(left, right, w, dmin, dmax)
    ref, _ = oracle.cbca_cost_volume(left, right, cv, w // 2, dmin, dist, tau)
    got = host(eng.cbca(dev(eng, left), dev(eng, right), dev(eng, cv), w // 2, dmin, dist, tau))
    np.testing.assert_array_equal(got, ref)          # integer costs: sums and counts exact, one division


@pytest.mark.parametrize("cfg", [(97, 131, -40, 23, 5, 5, 30.0), (64, 258, -100, -5, 5, 4, 12.0), (33, 67, 3, 70, 3, 5, 8.0),
                                 (150, 41, -7, 8, 5, 2, 20.0), (260, 90, -95, 0, 5, 5, 25.0)])
def test_cbca_register_kernel_vs_oracle_and_staged_kernel(eng, oracle, cfg):
    """cbca_distance <= 5 runs the register kernel (no barriers, thread-private prefix rings): bit-exact against the
    oracle on integer costs, for disparity counts that are not multiples of 32, widths that are not multiples of 4,
    more rows than the ring is deep, ranges that leave the image on either side -- and identical to the staged kernel."""
    import torch

    H, W, dmin, dmax, w, dist, tau = cfg
    left, right = rand_pair(H * 3 + W, H, W)
    left[5:9, 10:20] = left[5, 10]                               # flat patches: long arms
    right[5:9, 12:22] = left[5, 10]
    cv, _ = oracle.census_cost_volume(left, right, w, dmin, dmax)
    ref, _ = oracle.cbca_cost_volume(left, right, cv, w // 2, dmin, dist, tau)
    dl, dr, dcv = dev(eng, left), dev(eng, right), dev(eng, cv)
    got = eng.cbca(dl, dr, dcv, w // 2, dmin, dist, tau)
    np.testing.assert_array_equal(host(got), ref)
    import pandora_b200 as pb

    assert pb.last_path("cbca")[0] == "cbca_reg"
    with pb.option("cbca.pipe", 1):
        staged = eng.cbca(dl, dr, dcv, w // 2, dmin, dist, tau)
        assert pb.last_path("cbca")[0] == "cbca_pipe"
    assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(staged, nan=-7.0))


def test_cbca_float_costs_tolerance(eng, oracle):
    left, right = rand_pair(9, 33, 47, as_float=True)
    cv, _ = oracle.zncc_cost_volume(left, right, 3, -6, 6)
    ref, _ = oracle.cbca_cost_volume(left, right, cv, 1, -6, 5, 30.0)
    got = host(eng.cbca(dev(eng, left), dev(eng, right), dev(eng, cv), 1, -6, 5, 30.0))
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    # the reference's float32 row-prefix differences carry ~1e-7 * |prefix| absolute error: compare
    # with an absolute tolerance scaled by the row prefix magnitude (|cost| <= 1, W = 47)
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=47 * 1.2e-7)


def test_cbca_reference_goldens(eng, goldens, oracle):
    left = goldens[f"{AGG}::TestAggregation.setUp::data"].astype(np.float32)
    right = goldens[f"{AGG}::TestAggregation.setUp::data#1"].astype(np.float32)
    cv = np.full((3, 5, 3), np.nan, dtype=np.float32)
    cv[:, 1:, 0] = abs(left[:, 1:] - right[:, :4])
    cv[:, :, 1] = abs(left - right)
    cv[:, :4, 2] = abs(left[:, :4] - right[:, 1:])
    got = host(eng.cbca(dev(eng, left), dev(eng, right), dev(eng, cv), 0, -1, 3, 5.0))
    np.testing.assert_allclose(got, goldens[f"{AGG}::TestAggregation.test_compute_cbca::aggregated_ground_truth"], rtol=1e-7)
    k = f"{AGG}::TestAggregation.test_compute_cbca_with_offset::"
    left, right = goldens[k + "data"], goldens[k + "data#1"]
    cv = host(eng.sad_ssd(dev(eng, left), dev(eng, right), 3, -1, 1))
    got = host(eng.cbca(dev(eng, left), dev(eng, right), dev(eng, cv), 1, -1, 3, 5.0))
    np.testing.assert_allclose(got, goldens[k + "aggregated_ground_truth"], rtol=1e-7)


# ------------------------------------------------------------------------------------------------
# SGM
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(9, 11, 5), (23, 31, 61), (17, 40, 64), (12, 19, 200), (10, 13, 300), (1, 30, 16), (30, 1, 16)])
@pytest.mark.parametrize("over", [False, True])
def test_sgm_integer_costs_vs_oracle(eng, oracle, shape, over):
    g = np.random.default_rng(sum(shape))
    cv = g.integers(0, 26, shape).astype(np.float32)
    cv[g.random(shape) < 0.15] = np.nan
    ref = oracle.sgm_cost_volume(cv, 8, 32, cmax=25, overcounting=over)
    got = host(eng.sgm(dev(eng, cv), 8, 32, oracle.sgm_invalid_value(25, 32), overcounting=over))
    np.testing.assert_array_equal(got, ref)


def test_sgm_float_costs_bit_exact(eng, oracle):
    g = np.random.default_rng(12)
    cv = g.random((21, 27, 48)).astype(np.float32)
    cv[g.random(cv.shape) < 0.1] = np.nan
    ref = oracle.sgm_cost_volume(cv, 0.3, 1.7, cmax=1.0)
    got = host(eng.sgm(dev(eng, cv), 0.3, 1.7, oracle.sgm_invalid_value(1.0, 1.7)))
    np.testing.assert_array_equal(got, ref)          # same operation order as the oracle -> identical rounding


@pytest.mark.parametrize("shape", [(20, 700, 64), (33, 300, 256), (7, 1200, 40), (64, 613, 130), (5, 4100, 8)])
def test_sgm_strip_sweep_many_strips(eng, oracle, shape):
    """Wide images: the vertical groups run as multi-strip cooperative sweeps (ring hand-over between CTAs);
    the result must equal both the oracle and the direction-by-direction path kernels."""
    import torch

    g = np.random.default_rng(shape[1])
    cv = g.integers(0, 26, shape).astype(np.float32)
    cv[g.random(shape) < 0.1] = np.nan
    ref = oracle.sgm_cost_volume(cv, 8, 32, cmax=25)
    d_cv = dev(eng, cv)
    got, disp, flags = eng.sgm(d_cv, 8, 32, 58.0, fuse_wta=True, dmin=-(shape[2] - 1))
    np.testing.assert_array_equal(host(got), ref)
    exp_disp, exp_inv = oracle.wta(ref, np.arange(-(shape[2] - 1), 1))
    np.testing.assert_array_equal(host(disp), exp_disp)
    np.testing.assert_array_equal(host(flags).astype(bool), exp_inv)
    one = torch.empty_like(d_cv)
    for r in range(8):                                   # single-direction calls never take the sweep path
        eng.sgm(d_cv, 8, 32, 58.0, out=one, dir_mask=1 << r, init_final=(1 if r == 0 else 0) | (2 if r == 7 else 0))
    np.testing.assert_array_equal(host(one), ref)


def test_sgm_strip_sweep_float_costs_repeatable(eng, oracle):
    g = np.random.default_rng(77)
    cv = g.random((19, 500, 96)).astype(np.float32)
    cv[g.random(cv.shape) < 0.1] = np.nan
    ref = oracle.sgm_cost_volume(cv, 0.3, 1.7, cmax=1.0)
    for _ in range(3):
        got = host(eng.sgm(dev(eng, cv), 0.3, 1.7, oracle.sgm_invalid_value(1.0, 1.7)))
        np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("shape", [(9, 21, 64), (31, 333, 128), (16, 600, 256), (3, 4, 64), (1, 50, 128), (40, 1, 64), (23, 410, 192), (5, 33, 192)])
@pytest.mark.parametrize("over,p1,p2", [(False, 8, 32), (True, 8, 32), (False, 10, 120), (True, 3, 200)])
def test_sgm_packed_integer_path(eng, oracle, shape, over, p1, p2):
    """D in {64, 128, 192, 256} with integer costs takes the packed 16-bit path (sgm_narrow.cu): bit-identical to the
    oracle, including NaN cells, all-NaN pixels, overcounting and the fused WTA."""
    g = np.random.default_rng(shape[0] * 1000 + shape[2] + p2)
    cv = g.integers(0, 26, shape).astype(np.float32)
    cv[g.random(shape) < 0.15] = np.nan
    cv[g.random(shape[:2]) < 0.1] = np.nan                   # whole pixels without any valid cost
    ref = oracle.sgm_cost_volume(cv, p1, p2, cmax=25, overcounting=over)   # P2 = 32: byte storage tier; larger: 16-bit tier
    dmin = -(shape[2] - 1)
    got, disp, flags = eng.sgm(dev(eng, cv), p1, p2, oracle.sgm_invalid_value(25, p2), overcounting=over, fuse_wta=True, dmin=dmin)
    np.testing.assert_array_equal(host(got), ref)
    exp_disp, exp_inv = oracle.wta(ref, np.arange(dmin, 1))
    np.testing.assert_array_equal(host(disp), exp_disp)
    np.testing.assert_array_equal(host(flags).astype(bool), exp_inv)


@pytest.mark.parametrize("shape", [(5, 4096, 256), (7, 3000, 128), (4, 4095, 64), (33, 1500, 256), (2, 4736, 64), (6, 4100, 192)])
def test_sgm_wavefront_wide_images(eng, oracle, shape):
    """Wide images: strips of many warps (K = 2 * warps columns), partial last strips, odd widths -- the 4-direction
    wavefront passes (sgm_wave_kernel) against the oracle, bit-exact, with the fused WTA."""
    g = np.random.default_rng(shape[1])
    cv = g.integers(0, 26, shape).astype(np.float32)
    cv[g.random(shape) < 0.1] = np.nan
    ref = oracle.sgm_cost_volume(cv, 8, 32, cmax=25)
    dmin = -(shape[2] - 1)
    got, disp, flags = eng.sgm(dev(eng, cv), 8, 32, oracle.sgm_invalid_value(25, 32), fuse_wta=True, dmin=dmin)
    np.testing.assert_array_equal(host(got), ref)
    exp_disp, exp_inv = oracle.wta(ref, np.arange(dmin, 1))
    np.testing.assert_array_equal(host(disp), exp_disp)
    np.testing.assert_array_equal(host(flags).astype(bool), exp_inv)


def test_sgm_wavefront_repeatable_under_load(eng, oracle):
    """The wavefront passes synchronise through mailboxes and progress counters, not barriers: the same input must give
    the same bits every time, also while another stream keeps the SMs busy (different warp interleavings)."""
    import torch

    shape = (48, 4096, 256)
    g = np.random.default_rng(5)
    cv = g.integers(0, 26, shape).astype(np.float32)
    cv[g.random(shape) < 0.05] = np.nan
    ref = oracle.sgm_cost_volume(cv, 8, 32, cmax=25)
    d_cv = dev(eng, cv)
    side = torch.cuda.Stream()
    junk = torch.empty(64 << 20, device="cuda")
    for rep in range(6):
        if rep >= 3:
            with torch.cuda.stream(side):                         # memory traffic next to the sweep
                for _ in range(20):
                    junk.add_(1.0)
        got = host(eng.sgm(d_cv, 8, 32, oracle.sgm_invalid_value(25, 32)))
        torch.cuda.synchronize()
        np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("kind", ["float", "one_fraction", "negative", "too_large", "big_penalty"])
def test_sgm_packed_path_falls_back_exactly(eng, oracle, kind):
    """Volumes that do not qualify for the packed path (checked on the device while it runs) are redone by the
    float kernels behind the gate: the result is still bit-identical to the oracle."""
    g = np.random.default_rng(3)
    shape = (13, 37, 64)
    p1, p2, cmax = 8.0, 32.0, 25.0
    cv = g.integers(0, 26, shape).astype(np.float32)
    if kind == "float":
        cv = (g.random(shape) * 25).astype(np.float32)
    elif kind == "one_fraction":
        cv[-1, -1, -1] = 3.5
    elif kind == "negative":
        cv[5, 7, 9] = -2.0
    elif kind == "too_large":
        cv[2, 3, 4] = 9000.0
        cmax = 9000.0
    elif kind == "big_penalty":
        p2 = 9000.0
    cv[g.random(shape) < 0.1] = np.nan
    ref = oracle.sgm_cost_volume(cv, p1, p2, cmax=cmax)
    got = host(eng.sgm(dev(eng, cv), p1, p2, oracle.sgm_invalid_value(cmax, p2)))
    np.testing.assert_array_equal(got, ref)


def test_sgm_fused_wta_and_census_pipeline(eng, oracle):
    left, right, _ = oracle.synthetic_pair(48, 160, 64)
    cv, attrs = oracle.census_cost_volume(left, right, 5, -63, 0)
    ref = oracle.sgm_cost_volume(cv, 8, 32, cmax=attrs["cmax"])
    exp_disp, exp_inv = oracle.wta(ref, np.arange(-63, 1))
    S, disp, flags = eng.sgm(eng.census(dev(eng, left), dev(eng, right), 5, -63, 0), 8, 32, 58.0, fuse_wta=True, dmin=-63)
    np.testing.assert_array_equal(host(S), ref)
    np.testing.assert_array_equal(host(disp), exp_disp)
    np.testing.assert_array_equal(host(flags).astype(bool), exp_inv)


def test_sgm_row_tiles_with_halo_equal_whole_image(eng, oracle):
    """Row-tiled execution with path-state hand-over (the multi-GPU scheme, run here on one GPU) is bit-identical."""
    import torch

    left, right, _ = oracle.synthetic_pair(40, 96, 32)
    cv = eng.census(dev(eng, left), dev(eng, right), 5, -31, 0)
    whole = host(eng.sgm(cv, 8, 32, 58.0))
    H, W, D = cv.shape
    cuts = [0, 13, 27, 40]
    tiles = [cv[a:b].contiguous() for a, b in zip(cuts[:-1], cuts[1:])]
    outs = [torch.empty_like(t) for t in tiles]
    n = len(tiles)
    for t, o in zip(tiles, outs):
        eng.sgm(t, 8, 32, 58.0, out=o, dir_mask=0x03, init_final=1)
    halo = None
    for i in range(n):                                   # downward sweep, top tile first
        nxt = torch.empty((3, W, D), device=cv.device)
        eng.sgm(tiles[i], 8, 32, 58.0, out=outs[i], dir_mask=0x1C, init_final=0, halo_in_top=halo, halo_out_bottom=nxt)
        halo = nxt
    halo = None
    for i in reversed(range(n)):                         # upward sweep, bottom tile first
        nxt = torch.empty((3, W, D), device=cv.device)
        eng.sgm(tiles[i], 8, 32, 58.0, out=outs[i], dir_mask=0xE0, init_final=2, halo_in_bottom=halo, halo_out_top=nxt)
        halo = nxt
    np.testing.assert_array_equal(np.concatenate([host(o) for o in outs]), whole)
    np.testing.assert_array_equal(whole, oracle.sgm_cost_volume(host(cv), 8, 32, cmax=25))


@pytest.mark.parametrize("kind", ["integer", "float"])
@pytest.mark.parametrize("shape", [(40, 96, 64), (37, 333, 128), (30, 200, 256)])
def test_sgm_row_tiles_packed_split_calls(eng, oracle, kind, shape):
    """The split-call sequence of a row-tiled run with packed intermediates (include/pandora_b200.h, init_final
    bits 2 and 3): bit-identical to the whole-image run, on the packed path (integer costs) and on the gated
    float fallback (float costs)."""
    import torch

    g = np.random.default_rng(shape[1])
    H, W, D = shape
    if kind == "integer":
        cvh = g.integers(0, 26, shape).astype(np.float32)
        p1, p2, inv = 8.0, 32.0, 58.0
    else:
        cvh = (g.random(shape) * 25).astype(np.float32)
        p1, p2, inv = 8.0, 32.0, 58.0
    cvh[g.random(shape) < 0.1] = np.nan
    cv = dev(eng, cvh)
    dmin = -(D - 1)
    whole, wdisp, wflags = eng.sgm(cv, p1, p2, inv, fuse_wta=True, dmin=dmin)
    whole, wdisp = host(whole), host(wdisp)
    np.testing.assert_array_equal(whole, oracle.sgm_cost_volume(cvh, p1, p2, cmax=25))
    cuts = [0, H // 3, 2 * H // 3 + 1, H]
    tiles = [cv[a:b].contiguous() for a, b in zip(cuts[:-1], cuts[1:])]
    outs = [torch.empty_like(t) for t in tiles]
    disps = [torch.empty(t.shape[:2], device=cv.device) for t in tiles]
    flags = [torch.empty(t.shape[:2], dtype=torch.uint8, device=cv.device) for t in tiles]
    n = len(tiles)
    for t, o in zip(tiles, outs):
        eng.sgm(t, p1, p2, inv, out=o, dir_mask=0x03, init_final=1, packed=True)
        eng.sgm(t, p1, p2, inv, out=o, dir_mask=0x03, init_final=1, packed=True, float_only=True)
    halo = None
    for i in range(n):                                   # downward wave, top tile first
        nxt = torch.empty((3, W, D), device=cv.device)
        eng.sgm(tiles[i], p1, p2, inv, out=outs[i], dir_mask=0x1C, init_final=0, halo_in_top=halo, halo_out_bottom=nxt, packed=True)
        halo = nxt
    halo = None
    for i in reversed(range(n)):                         # upward wave, bottom tile first, final + fused WTA
        nxt = torch.empty((3, W, D), device=cv.device)
        eng.sgm(tiles[i], p1, p2, inv, out=outs[i], dir_mask=0xE0, init_final=2, halo_in_bottom=halo, halo_out_top=nxt, packed=True,
                fuse_wta=True, dmin=dmin, disp=disps[i], flags=flags[i])
        halo = nxt
    np.testing.assert_array_equal(np.concatenate([host(o) for o in outs]), whole)
    np.testing.assert_array_equal(np.concatenate([host(d) for d in disps]), wdisp)


# ------------------------------------------------------------------------------------------------
# host-buffer C-ABI entry points (what a reference-side binding calls)
# ------------------------------------------------------------------------------------------------
def _p(a):
    return a.ctypes.data


def test_host_entry_points(eng, oracle):
    from pandora_b200 import _native

    lib = _native.load()
    left, right = rand_pair(77, 30, 44)
    disps = np.arange(-8, 3).astype(np.float32)
    cv = np.empty((30, 44, 11), dtype=np.float32)
    _native.check(lib.pb200_census_cost_volume_host(_p(left), _p(right), 30, 44, 5, _p(disps), 11, _p(cv)))
    ref, _ = oracle.census_cost_volume(left, right, 5, -8, 2)
    np.testing.assert_array_equal(cv, ref)

    rcv = np.empty_like(cv)
    _native.check(lib.pb200_reverse_cost_volume_host(_p(cv), 30, 44, 11, -2, _p(rcv)))
    np.testing.assert_array_equal(rcv, oracle.reverse_cost_volume(cv, -2))

    img = np.nan_to_num(left.copy())
    cross = np.empty((30, 44, 4), dtype=np.int16)
    _native.check(lib.pb200_cross_support_host(_p(img), 30, 44, 5, 30.0, _p(cross)))
    np.testing.assert_array_equal(cross, oracle.cross_support(img, 5, 30.0))

    cross_r = oracle.cross_support(right, 5, 30.0)
    for k, d in enumerate(range(-8, 3)):
        sl = np.ascontiguousarray(cv[:, :, k])
        cols = np.arange(44, dtype=np.int64)
        ok = (cols + d >= 0) & (cols + d < 44)
        rc, rcr = np.ascontiguousarray(cols[ok]), np.ascontiguousarray(cols[ok] + d)
        s4 = np.empty((30, 44), dtype=np.float32)
        n4 = np.empty((30, 44), dtype=np.float32)
        _native.check(lib.pb200_cbca_host(_p(sl), 30, 44, _p(cross), _p(cross_r), _p(rc), _p(rcr), len(rc), _p(s4), _p(n4)))
        e4, en = oracle.cbca_slice(sl, cross, cross_r, d)
        np.testing.assert_array_equal(s4, e4)
        np.testing.assert_array_equal(n4, en)


def test_disparity_host_pipelines(eng, oracle):
    from pandora_b200 import _native

    lib = _native.load()
    left, right, _ = oracle.synthetic_pair(40, 120, 48)
    H, W = left.shape
    dmin, dmax = -47, 0
    for (cbca, sgm) in [(0, 0.0), (5, 0.0), (0, 32.0), (5, 32.0)]:
        disp = np.empty((H, W), dtype=np.float32)
        vm = np.empty((H, W), dtype=np.uint16)
        out = np.empty((H, W, 48), dtype=np.float32)
        _native.check(lib.pb200_disparity_host(_p(left), _p(right), H, W, 0, 5, dmin, dmax, cbca, 30.0, 8.0, sgm, 0, -9999.0,
                                               _p(disp), _p(vm), _p(out)))
        cv, attrs = oracle.census_cost_volume(left, right, 5, dmin, dmax)
        mask = oracle.validity_mask(H, W, dmin, dmax, 2)
        oracle.cv_masked(cv, mask, 2)
        cmax = attrs["cmax"]
        if cbca:
            cv, cmax = oracle.cbca_cost_volume(left, right, cv, 2, dmin, cbca, 30.0, cmax)
        if sgm:
            cv = oracle.sgm_cost_volume(cv, 8, sgm, cmax=cmax)
        exp, inv = oracle.wta(cv, np.arange(dmin, dmax + 1))
        np.testing.assert_array_equal(out, cv)
        np.testing.assert_array_equal(disp, exp)
        np.testing.assert_array_equal(vm, oracle.wta_validity_mask(mask, inv))


# ------------------------------------------------------------------------------------------------
# sub-pixel Census (subpix 2 / 4): census.cpp:128-155 with the list of shifted right images
# ------------------------------------------------------------------------------------------------
def test_census_subpix_reference_golden(eng, oracle, goldens):
    """tests/test_matching_cost/test_matching_cost_census.py:637-683 (window 3, subpix 2): the reference's own golden volume."""
    k = f"{CENSUS}::test_census[7]::"
    assert int(goldens[k + "subpix"]) == 2 and int(goldens[k + "window_size"]) == 3
    left, right = goldens[k + "left_data"].astype(np.float32), goldens[k + "right_data"].astype(np.float32)
    dmin, dmax = (int(v) for v in goldens[k + "disp_interval"])
    rights = [np.ascontiguousarray(r, dtype=np.float32) for r in oracle.shift_right_img(right, 2, 1)]
    got = host(eng.census_subpix(dev(eng, left), [dev(eng, r) for r in rights], 3, dmin, (dmax - dmin) * 2 + 1))
    np.testing.assert_array_equal(got, goldens[k + "ref_out"])


@pytest.mark.parametrize("w", [3, 5, 7, 11])
@pytest.mark.parametrize("subpix", [2, 4])
@pytest.mark.parametrize("shape,rng", [((19, 37), (-7, 5)), ((12, 40), (-25, -2)), ((9, 21), (3, 14))])
def test_census_subpix_vs_oracle(eng, oracle, w, subpix, shape, rng):
    """Every branch of census.cpp:128-155: ranges leaving the image on either side, the shorter shifted images (their last
    usable centre is one column earlier), one- to four-word descriptors -- and the `compute_matching_costs`-style host entry."""
    import ctypes

    from pandora_b200 import _native

    g = np.random.default_rng(w * 100 + subpix + shape[1])
    left, right = (g.integers(0, 50, shape).astype(np.float32) for _ in range(2))
    dmin, dmax = rng
    ref, _ = oracle.census_cost_volume_subpix(left, right, w, dmin, dmax, subpix)
    rights = [np.ascontiguousarray(r, dtype=np.float32) for r in oracle.shift_right_img(right, subpix, 1)]
    n_disp = (dmax - dmin) * subpix + 1
    got = host(eng.census_subpix(dev(eng, left), [dev(eng, r) for r in rights], w, dmin, n_disp))
    np.testing.assert_array_equal(got, ref)
    import pandora_b200 as pb

    assert pb.last_path("census")[0] == "census_subpix"
    out = np.empty(shape + (n_disp,), dtype=np.float32)
    ptrs = (ctypes.c_void_p * subpix)(*[r.ctypes.data for r in rights])
    disps = np.append(np.arange(dmin, dmax, 1 / subpix), [dmax]).astype(np.float32)
    _native.check(_native.load().pb200_census_cost_volume_multi_host(left.ctypes.data, ptrs, subpix, shape[0], shape[1], w, disps.ctypes.data,
                                                                     n_disp, out.ctypes.data))
    np.testing.assert_array_equal(out, ref)


def test_census_subpix_through_the_step_classes(oracle):
    """matching_cost (census, subpix 4 -- the reference's a_local_block_matching.json uses it) -> disparity through run(cfg):
    fractional disparities disps[argmin] like disparity.py:434-455."""
    import pandora_b200 as pb

    H, W = 30, 64
    g = np.random.default_rng(5)
    left, right = (g.integers(0, 255, (H, W)).astype(np.float32) for _ in range(2))
    cfg = {"pipeline": {"matching_cost": {"matching_cost_method": "census", "window_size": 5, "subpix": 4},
                        "disparity": {"disparity_method": "wta", "invalid_disparity": -9999}}}
    disp, cv = pb.run(pb.create_image_dataset(left, disparity=[-6, 3]), pb.create_image_dataset(right), cfg)
    ref, _ = oracle.census_cost_volume_subpix(left, right, 5, -6, 3, 4)
    np.testing.assert_array_equal(np.asarray(cv["cost_volume"].data), ref)
    disps = np.asarray(cv.coords["disp"].data)
    assert len(disps) == 37 and disps[1] == -5.75
    filled = np.where(np.isnan(ref), np.inf, ref)
    exp = disps[np.argmin(filled, axis=2)].astype(np.float32)
    exp[np.all(np.isnan(ref), axis=2)] = -9999
    np.testing.assert_array_equal(np.asarray(disp["disparity_map"].data), exp)
