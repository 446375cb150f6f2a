"""Differential tests: oracle restatement vs the UNMODIFIED reference C++ compiled into oracle/_ref
(oracle/build_ref.sh).  Skipped where oracle/_ref was never built.  CPU only."""
import numpy as np
import pytest


@pytest.mark.parametrize("w", [3, 5, 7, 9, 11, 13])
@pytest.mark.parametrize("shape,rng", [((17, 31), (-6, 3)), ((20, 24), (-30, -2)), ((15, 40), (2, 9))])
def test_census_vs_reference(oracle, ref_modules, w, shape, rng):
    mc, _ = ref_modules
    gen = np.random.default_rng(w * 100 + shape[1])
    left = gen.integers(0, 6, shape).astype(np.float32)          # few levels -> many ties (strict > matters)
    right = gen.integers(0, 6, shape).astype(np.float32)
    dmin, dmax = rng
    disps = np.arange(dmin, dmax + 1).astype(np.float32)
    cv0 = np.full(shape + (len(disps),), np.nan, dtype=np.float32)
    ref = mc.compute_matching_costs(left, [right], cv0, disps, w, w)          # census.py:140-147
    got, _ = oracle.census_cost_volume(left, right, w, dmin, dmax)
    np.testing.assert_array_equal(got, ref)


def test_census_vs_reference_float_images(oracle, ref_modules):
    mc, _ = ref_modules
    gen = np.random.default_rng(7)
    left = gen.normal(size=(33, 47)).astype(np.float32)
    right = gen.normal(size=(33, 47)).astype(np.float32)
    disps = np.arange(-11, 5).astype(np.float32)
    ref = mc.compute_matching_costs(left, [right], np.full((33, 47, 16), np.nan, np.float32), disps, 5, 5)
    got, _ = oracle.census_cost_volume(left, right, 5, -11, 4)
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("cfg", [(9, 17, 3, -4, 3, 2), (11, 20, 5, -6, 2, 4), (8, 15, 5, 1, 5, 2), (9, 14, 3, -3, 0, 4), (7, 30, 7, -9, 9, 2)])
def test_census_subpix_vs_reference(oracle, ref_modules, cfg):
    """Sub-pixel Census (subpix 2 and 4): the oracle's restatement of census.py:109-153 + census.cpp:97-180 with the
    shifted right images of img_tools.shift_right_img against the compiled reference on the same images."""
    mc, _ = ref_modules
    H, W, w, dmin, dmax, subpix = cfg
    gen = np.random.default_rng(H * W + subpix)
    left = gen.integers(0, 9, (H, W)).astype(np.float32)
    right = gen.integers(0, 9, (H, W)).astype(np.float32)
    got, attrs = oracle.census_cost_volume_subpix(left, right, w, dmin, dmax, subpix)
    disps = np.arange(dmin, dmax + 1e-9, 1.0 / subpix).astype(np.float32)
    assert got.shape == (H, W, len(disps)) and attrs["cmax"] == w * w
    shifted = [x.astype(np.float32) for x in oracle.shift_right_img(right, subpix)]
    ref = mc.compute_matching_costs(left, shifted, np.full(got.shape, np.nan, np.float32), disps, w, w)
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("arms,tau", [(3, 5.0), (5, 30.0), (9, 2.5), (1, 4.0)])
def test_cross_support_vs_reference(oracle, ref_modules, arms, tau):
    _, agg = ref_modules
    gen = np.random.default_rng(arms)
    img = gen.integers(0, 40, (23, 29)).astype(np.float32)
    img[gen.random(img.shape) < 0.05] = np.inf                    # NaN->inf pixels of cbca.py:233
    np.testing.assert_array_equal(oracle.cross_support(img, arms, tau), agg.cross_support(img, arms, tau))


@pytest.mark.parametrize("seed", range(6))
def test_cbca_slice_vs_reference(oracle, ref_modules, seed):
    """one aggregation_cpp.cbca call (cbca.py:152-164) on a strided slice, integer and float costs."""
    _, agg = ref_modules
    gen = np.random.default_rng(seed)
    H, W, D = 19, 27, 7
    dmin = -4
    left = gen.integers(0, 60, (H, W)).astype(np.float32)
    right = gen.integers(0, 60, (H, W)).astype(np.float32)
    cl = oracle.cross_support(left, 5, 30.0)
    cr = oracle.cross_support(right, 5, 30.0)
    cv = gen.integers(0, 26, (H, W, D)).astype(np.float32) if seed % 2 == 0 else gen.random((H, W, D)).astype(np.float32)
    cv[gen.random(cv.shape) < 0.1] = np.nan
    for k in range(D):
        d = dmin + k
        cols = np.arange(W)
        colr = cols + d
        ok = np.where((colr >= 0) & (colr < W))
        ref4, refn = agg.cbca(cv[:, :, k], cl, cr, cols[ok], colr[ok].astype(int))
        got4, gotn = oracle.cbca_slice(cv[:, :, k], cl, cr, d)
        if seed % 2 == 0:
            np.testing.assert_array_equal(got4, ref4)
        else:
            # float costs: column 0 of row 0 reads one float before the step-1 buffer in the reference
            # (aggregation.cpp:113-114, index -1); everything else must be bit-identical
            np.testing.assert_array_equal(got4[1:], ref4[1:])
            np.testing.assert_allclose(got4[0], ref4[0], rtol=1e-6, atol=1e-6)
        np.testing.assert_array_equal(gotn, refn)


def test_cbca_volume_vs_reference_driver(oracle, ref_modules):
    """whole driver loop of cbca.py:127-177 restated with the reference's own C++ per disparity."""
    _, agg = ref_modules
    gen = np.random.default_rng(11)
    H, W, dmin, dmax = 21, 33, -9, 2
    left = gen.integers(0, 255, (H, W)).astype(np.float32)
    right = gen.integers(0, 255, (H, W)).astype(np.float32)
    cv, attrs = oracle.census_cost_volume(left, right, 5, dmin, dmax)
    got, cmax = oracle.cbca_cost_volume(left, right, cv, 2, dmin, 5, 30.0, attrs["cmax"])
    cl, cr = oracle.computes_cross_supports(left, right, 2, 5, 30.0)
    cv_data = cv[2:-2, 2:-2]
    n_col_, n_row_, nb_disp = cv_data.shape
    aggv = np.zeros((nb_disp, n_row_, n_col_), dtype=np.float32)
    aggv += np.swapaxes(cv_data, 0, 2)
    aggv *= 0
    cols = np.arange(n_row_)
    for k in range(nb_disp):
        colr = cols + (dmin + k)
        ok = np.where((colr >= 0) & (colr < n_row_))
        s4, n4 = agg.cbca(cv_data[:, :, k], cl, cr, cols[ok], colr[ok].astype(int))
        n4 += 1
        aggv[k] += np.swapaxes(s4, 0, 1)
        aggv[k] /= np.swapaxes(n4, 0, 1)
    exp = cv.copy()
    exp[2:-2, 2:-2] = np.swapaxes(aggv, 0, 2)
    np.testing.assert_array_equal(got, exp)
    assert cmax == 25 * 81


def test_reverse_cost_volume_vs_reference(oracle, ref_modules):
    mc, _ = ref_modules
    gen = np.random.default_rng(3)
    cv = gen.random((5, 13, 6)).astype(np.float32)
    cv[gen.random(cv.shape) < 0.2] = np.nan
    for min_disp in (-5, -2, 0, 3):
        np.testing.assert_array_equal(oracle.reverse_cost_volume(cv, min_disp), mc.reverse_cost_volume(cv, min_disp))


def _disp_grids(seed, H, W, lo, hi, nan_frac=0.1):
    gen = np.random.default_rng(seed)
    a = gen.integers(lo, hi + 1, (H, W)).astype(np.float32)
    b = a + gen.integers(0, 6, (H, W)).astype(np.float32)
    if seed % 2:
        a += 0.5                                                  # static_cast<int> truncates toward zero
    a[gen.random((H, W)) < nan_frac] = np.nan
    b[gen.random((H, W)) < nan_frac] = np.nan
    return a, b


@pytest.mark.parametrize("seed,lo,hi", [(0, -7, 2), (1, -3, 3), (2, 2, 9), (3, -30, -20), (4, -40, 40)])
def test_reverse_disp_range_vs_reference(oracle, ref_modules, seed, lo, hi):
    """matching_cost.cpp:59-131: the oracle against the compiled reference, NaN bounds, ranges leaving the row, right
    pixels that nothing reaches."""
    mc, _ = ref_modules
    a, b = _disp_grids(seed, 9, 31, lo, hi)
    rmin, rmax = oracle.reverse_disp_range(a, b)
    ref_min, ref_max = mc.reverse_disp_range(a, b)
    np.testing.assert_array_equal(rmin, ref_min)
    np.testing.assert_array_equal(rmax, ref_max)


def test_reverse_disp_range_constant_grid(oracle):
    """A constant left range [-3, 1]: right pixel rc sees col - rc for every col in [rc - 1, rc + 3] inside the row."""
    a, b = np.full((2, 8), -3, np.float32), np.full((2, 8), 1, np.float32)
    rmin, rmax = oracle.reverse_disp_range(a, b)
    np.testing.assert_array_equal(rmin[0], [-0.0, -1, -1, -1, -1, -1, -1, -1])
    np.testing.assert_array_equal(rmax[0], [3, 3, 3, 3, 3, 2, 1, 0])
