"""CPU oracle for Pandora's dense cost-volume hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``bench.py`` (``cpu_baseline`` leg and ``--impl reference``) and
``__graft_entry__.smoke()`` may import this module.  ``pandora_b200`` never does.

Two layers:

* ``libpandora_oracle.so`` (``oracle/pandora_oracle.c``, plain C): census, cross support, CBCA,
  reverse cost volume, SGM and WTA restated from the reference's C++ / the published algorithm.
* numpy restatements of the reference's *Python* glue that cannot be imported in this image
  (xarray, json_checker, rasterio and transitions are absent): SAD/SSD, ZNCC, the 3x3 NaN-median,
  the CBCA driver, cost-volume allocation, validity mask and ``cv_masked`` (no-mask branch).

All citations are relative to ``/root/reference/src/pandora``.

Pinning: every function below is checked in ``tests/test_oracle_goldens.py`` against the literal
golden arrays of the reference's own unit tests (``tests/golden/reference_goldens.npz``, extracted
by ``tests/golden/extract_reference_goldens.py``) and, for census / cross support / CBCA / reverse,
against the unmodified reference C++ compiled into ``oracle/_ref`` (``tests/test_oracle_vs_ref.py``).
**SGM: parity unpinned** -- libSGM is an un-vendored dependency (pyproject.toml:59-61) and the
reference holds no numeric SGM test; see the header of ``pbo_sgm`` in ``pandora_oracle.c``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import warnings
from math import ceil, floor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# validity-mask bits, constants.py:28-55
MSK_INVALID = 0b01111000011
MSK_LEFT_NODATA_OR_BORDER = 1 << 0
MSK_RIGHT_NODATA_OR_DISPARITY_RANGE_MISSING = 1 << 1
MSK_RIGHT_INCOMPLETE_DISPARITY_RANGE = 1 << 2
MSK_STOPPED_INTERPOLATION = 1 << 3
MSK_IN_VALIDITY_MASK_LEFT = 1 << 6
MSK_IN_VALIDITY_MASK_RIGHT = 1 << 7
MSK_OCCLUSION = 1 << 8
MSK_MISMATCH = 1 << 9
MSK_INCOMPLETE_VARIABLE_DISPARITY_RANGE = 1 << 12


def build(force: bool = False) -> str:
    """Compile ``pandora_oracle.c`` (gcc) if the shared object is missing or stale."""
    so = os.path.join(_HERE, "_build", "libpandora_oracle.so")
    src = os.path.join(_HERE, "pandora_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O3", "-std=c11", "-fPIC", "-shared", "-fvisibility=hidden", "-ffp-contract=off", src, "-o", so, "-lm"]
        )
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        f32p, i16p, u8p = (ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int16), ctypes.POINTER(ctypes.c_uint8))
        ci, cf, cl = ctypes.c_int, ctypes.c_float, ctypes.c_long
        _LIB.pbo_census_cost.argtypes = [f32p, f32p, ci, ci, ci, ci, ci, f32p]
        _LIB.pbo_census_cost.restype = ci
        _LIB.pbo_cross_support.argtypes = [f32p, ci, ci, ci, cf, i16p]
        _LIB.pbo_cross_support.restype = None
        _LIB.pbo_cbca_slice.argtypes = [f32p, cl, cl, ci, ci, ci, i16p, i16p, f32p, f32p]
        _LIB.pbo_cbca_slice.restype = ci
        _LIB.pbo_cbca_volume.argtypes = [f32p, ci, ci, ci, ci, i16p, i16p, f32p]
        _LIB.pbo_cbca_volume.restype = ci
        _LIB.pbo_reverse_cost_volume.argtypes = [f32p, ci, ci, ci, ci, f32p]
        _LIB.pbo_reverse_cost_volume.restype = None
        _LIB.pbo_reverse_disp_range.argtypes = [f32p, f32p, ci, ci, f32p, f32p]
        _LIB.pbo_reverse_disp_range.restype = None
        _LIB.pbo_sgm.argtypes = [f32p, ci, ci, ci, cf, cf, cf, ci, ci, f32p]
        _LIB.pbo_sgm.restype = ci
        _LIB.pbo_sgm_direction.argtypes = [f32p, f32p, ci, ci, ci, cf, cf, ci, ci, f32p, f32p]
        _LIB.pbo_sgm_direction.restype = ci
        _LIB.pbo_wta.argtypes = [f32p, ci, ci, ci, f32p, ci, cf, f32p, u8p]
        _LIB.pbo_wta.restype = None
        u16p, i32p, f64p, cd = ctypes.POINTER(ctypes.c_uint16), ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_double), ctypes.c_double
        _LIB.pbo_refinement.argtypes = [f32p, ci, ci, ci, cd, cd, ci, ci, ci, ci, f32p, u16p, f32p]
        _LIB.pbo_refinement.restype = None
        _LIB.pbo_ambiguity.argtypes = [f32p, ci, ci, ci, f32p, ci, i32p, f32p, f32p, f32p]
        _LIB.pbo_ambiguity.restype = None
        _LIB.pbo_risk.argtypes = [f32p, f32p, ci, ci, ci, f64p, ci, i32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p]
        _LIB.pbo_risk.restype = None
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


# --------------------------------------------------------------------------------------------
# cost volume container pieces: matching_cost/matching_cost.py:330-427
# --------------------------------------------------------------------------------------------
def disparity_range(dmin: int, dmax: int) -> np.ndarray:
    """``get_disparity_range`` for subpix == 1, matching_cost.py:409-427."""
    return np.arange(dmin, dmax + 1)


def allocate_cost_volume(H: int, W: int, dmin: int, dmax: int) -> np.ndarray:
    """NaN-filled float32 (row, col, disp) volume, matching_cost.py:377-407."""
    return np.full((H, W, dmax - dmin + 1), np.nan, dtype=np.float32)


# --------------------------------------------------------------------------------------------
# Census: matching_cost/census.py:109-153 + cpp/src/census.cpp:45-180
# --------------------------------------------------------------------------------------------
def census_cost_volume(left, right, window: int, dmin: int, dmax: int):
    """Returns (cv float32 (H, W, D), attrs) -- ``type_measure='min'``, ``cmax=w*w`` (census.py:116-122)."""
    left, right = _f32(left), _f32(right)
    H, W = left.shape
    D = dmax - dmin + 1
    cv = np.empty((H, W, D), dtype=np.float32)
    rc = lib().pbo_census_cost(_p(left, ctypes.c_float), _p(right, ctypes.c_float), H, W, window, dmin, D, _p(cv, ctypes.c_float))
    if rc:
        raise ValueError(f"pbo_census_cost failed: {rc}")
    return cv, {"type_measure": "min", "cmax": int(window**2)}


def shift_right_img(right, subpix: int, order: int = 1):
    """img_tools.shift_right_img (img_tools.py:713-752): the right image and its subpix - 1 resampled copies at column
    offsets k / subpix (scipy.ndimage.zoom along the columns, one column shorter than the image)."""
    from scipy.ndimage import zoom  # noqa: PLC0415

    right = np.asarray(right)
    nx = right.shape[1]
    out = [right]
    for ind in range(1, subpix):
        out.append(zoom(right, (1, (nx * subpix - (subpix - 1)) / float(nx)), order=order)[:, ind::subpix])
    return out


def _census_bits(img: np.ndarray, w: int) -> np.ndarray:
    """census_transform (census.cpp:45-95) as an (H, W, w*w) boolean array; False outside the valid interior."""
    H, W = img.shape
    half = w // 2
    bits = np.zeros((H, W, w * w), dtype=bool)
    if H > 2 * half and W > 2 * half:
        c = img[half:H - half, half:W - half]
        for wy in range(w):
            for wx in range(w):
                bits[half:H - half, half:W - half, wy * w + wx] = img[wy:wy + H - 2 * half, wx:wx + W - 2 * half] > c
    return bits


def census_cost_volume_subpix(left, right, window: int, dmin: int, dmax: int, subpix: int, order: int = 1):
    """Census volume with sub-pixel disparities (census.py:109-153 + census.cpp:97-180, every branch): disparity index
    k = subpix * (d - dmin) + id_right uses the id_right-th shifted right image; a shifted image is one column shorter,
    its last usable centre is therefore one column earlier (census.cpp:144-150).  Small sizes only (pure numpy)."""
    left = _f32(left)
    rights = [_f32(r) for r in shift_right_img(_f32(right), subpix, order)]
    H, W = left.shape
    half = window // 2
    n_disp = (dmax - dmin) * subpix + 1
    cv = np.full((H, W, n_disp), np.nan, dtype=np.float32)
    bl = _census_bits(left, window)
    brs = [_census_bits(r, window) for r in rights]
    for row in range(half, H - half):
        for col in range(half, W - half):
            for disp in range(0, n_disp, subpix):
                rx = col + disp // subpix + dmin
                if rx < half or rx >= W - half:
                    continue
                for idr in range(subpix):
                    if disp + idr >= n_disp:
                        break
                    if idr != 0 and rx >= W - half - 1:
                        break
                    cv[row, col, disp + idr] = np.count_nonzero(bl[row, col] != brs[idr][row, rx])
    return cv, {"type_measure": "min", "cmax": int(window**2)}


# --------------------------------------------------------------------------------------------
# SAD / SSD: matching_cost/sad_ssd.py:110-207, 209-224, 340-368 ; point_interval matching_cost.py:429-482
# --------------------------------------------------------------------------------------------
def point_interval(nx_left: int, nx_right: int, disp: float):
    """matching_cost.py:429-482."""
    if abs(disp) > nx_left:
        point_p = (nx_left, nx_left)
    else:
        point_p = (max(0 - disp, 0), min(nx_left - disp, nx_left))
    if abs(disp) > nx_right:
        point_q = (nx_right, nx_right)
    else:
        point_q = (max(0 + disp, 0), min(nx_right + disp, nx_right))
    rnd = ceil if disp < 0 else floor
    return (int(rnd(point_p[0])), int(rnd(point_p[1]))), (int(rnd(point_q[0])), int(rnd(point_q[1])))


def pixel_wise_aggregation(cost_volume: np.ndarray, window: int) -> np.ndarray:
    """Window sum through a strided view exactly as sad_ssd.py:340-368 (same numpy reduction)."""
    nb_disp, nx_, ny_ = cost_volume.shape
    str_disp, str_col, str_row = cost_volume.strides
    shape_windows = (window, window, nb_disp, nx_ - (window - 1), ny_ - (window - 1))
    strides_windows = (str_row, str_col, str_disp, str_col, str_row)
    view = np.lib.stride_tricks.as_strided(cost_volume, shape_windows, strides_windows, writeable=False)
    return np.sum(view, (0, 1))


def sad_ssd_cost_volume(left, right, window: int, dmin: int, dmax: int, method: str = "sad"):
    """sad_ssd.py:75-207 for subpix == 1, single band.  Images are used in their own dtype like the
    reference (tests feed float64, the pipeline float32)."""
    left = np.asarray(left)
    right = np.asarray(right)
    H, W = left.shape
    off = (window - 1) // 2
    disps = disparity_range(dmin, dmax)
    mx = max(abs(np.amax(left) - np.amin(right)), abs(np.amax(right) - np.amin(left)))
    cmax = int(mx * window**2) if method == "sad" else int(mx**2 * window**2)
    cv_enlarge = np.full((len(disps), W + 2 * off, H + 2 * off), np.nan, dtype=np.float32)
    cv = cv_enlarge[:, off : W + off, off : H + off] if off else cv_enlarge
    for k, d in enumerate(disps):
        p, q = point_interval(W, W, d)
        diff = left[:, p[0] : p[1]] - right[:, q[0] : q[1]]
        cost = abs(diff) if method == "sad" else diff**2
        cv[k, p[0] : p[1], :] = np.swapaxes(cost, 0, 1)
    out = pixel_wise_aggregation(cv_enlarge, window)
    out = np.swapaxes(out, 0, 2)
    if off:
        out[:off, :, :] = np.nan
        out[-off:, :, :] = np.nan
        out[:, :off, :] = np.nan
        out[:, -off:, :] = np.nan
    return np.ascontiguousarray(out), {"type_measure": "min", "cmax": cmax}


# --------------------------------------------------------------------------------------------
# ZNCC: matching_cost/zncc.py:73-241, 244-277 ; img_tools.py:834-879 (mean), 915-952 (std)
# --------------------------------------------------------------------------------------------
def compute_mean_raster(img: np.ndarray, win: int) -> np.ndarray:
    """img_tools.py:834-879 -- float64 integral images (np.r_ with np.zeros promotes)."""
    ny_, nx_ = img.shape
    r_mean = np.r_[np.zeros((1, nx_)), img]
    r_mean = np.nancumsum(r_mean, axis=0)
    r_mean = r_mean[win:, :] - r_mean[:-win, :]
    r_mean = np.c_[np.zeros(ny_ - (win - 1)), r_mean]
    r_mean = np.cumsum(r_mean, axis=1)
    r_mean = r_mean[:, win:] - r_mean[:, :-win]
    return r_mean / float(win * win)


def compute_std_raster(img: np.ndarray, win: int) -> np.ndarray:
    """img_tools.py:915-952."""
    mean_ = compute_mean_raster(img, win)
    mean_power_two = compute_mean_raster(img**2, win)
    var = mean_power_two - mean_**2
    var[np.where(var < (10 ** (-15) * abs(mean_power_two)))] = 0
    return np.sqrt(var)


def zncc_cost_volume(left, right, window: int, dmin: int, dmax: int):
    """zncc.py:114-241 for subpix == 1, single band."""
    left = np.asarray(left)
    right = np.asarray(right)
    H, W = left.shape
    off = (window - 1) // 2
    disps = disparity_range(dmin, dmax)
    l_std, r_std = compute_std_raster(left, window), compute_std_raster(right, window)
    l_mean, r_mean = compute_mean_raster(left, window), compute_mean_raster(right, window)
    cv = np.full((len(disps), W, H), np.nan, dtype=np.float32)
    cv_crop = cv[:, off : W - off, off : H - off] if off else cv
    for k, d in enumerate(disps):
        p, q = point_interval(W, W, d)
        if abs(d) > W - (int(window / 2) * 2):                       # zncc.py:108-110
            p, q = (W, W), (W, W)
        p_std = (p[0], p[1] - (int(window / 2) * 2))
        q_std = (q[0], q[1] - (int(window / 2) * 2))
        prod = left[:, p[0] : p[1]] * right[:, q[0] : q[1]]
        if prod.shape[1] < window:
            continue
        zncc_ = compute_mean_raster(prod, window)
        zncc_ -= l_mean[:, p_std[0] : p_std[1]] * r_mean[:, q_std[0] : q_std[1]]
        divide_standard = np.multiply(l_std[:, p_std[0] : p_std[1]], r_std[:, q_std[0] : q_std[1]])
        valid = np.where(divide_standard > 0)
        zncc_[valid] /= divide_standard[valid]
        zncc_[np.where(divide_standard <= 0)] = 0
        cv_crop[k, p[0] : p_std[1], :] = np.swapaxes(zncc_, 0, 1)
    return np.ascontiguousarray(np.swapaxes(cv, 0, 2)), {"type_measure": "max", "cmax": 1}


# --------------------------------------------------------------------------------------------
# validity mask (no-mask branch) and cv_masked: criteria.py:66-158, 291-353 ; matching_cost.py:770-872
# --------------------------------------------------------------------------------------------
def validity_mask(H: int, W: int, dmin: int, dmax: int, offset: int) -> np.ndarray:
    """criteria.py:106-147 without left/right ``msk``; columns are 0..W-1."""
    vm = np.zeros((H, W), dtype=np.uint16)
    col = np.arange(W)
    if dmax < 0:
        bit_1 = np.where((col + dmax) < (col[0] + offset))[0]
        vm[:, np.where(((col + dmax) >= (col[0] + offset)) & ((col + dmin) < (col[0] + offset)))[0]] += (
            MSK_RIGHT_INCOMPLETE_DISPARITY_RANGE
        )
    elif dmin > 0:
        bit_1 = np.where((col + dmin) > (col[-1] - offset))[0]
        vm[:, np.where(((col + dmin) <= (col[-1] - offset)) & ((col + dmax) > (col[-1] - offset)))[0]] += (
            MSK_RIGHT_INCOMPLETE_DISPARITY_RANGE
        )
    else:
        bit_1 = np.array([], dtype=int)
        vm[:, np.where(((col + dmin) < (col[0] + offset)) | (col + dmax > (col[-1]) - offset))[0]] += (
            MSK_RIGHT_INCOMPLETE_DISPARITY_RANGE
        )
    vm[:, bit_1] += MSK_RIGHT_NODATA_OR_DISPARITY_RANGE_MISSING
    return vm


def cv_masked(cv: np.ndarray, vm: np.ndarray, offset: int) -> None:
    """matching_cost.py:815-872 with no ``msk`` and a fixed disparity range: the mask additions and
    the per-pixel range masking are value no-ops; what remains is
    ``mask_invalid_variable_disparity_range`` (criteria.py:291-322) and ``mask_border`` (:325-353)."""
    missing = np.min(np.isnan(cv), axis=2)
    upd = missing & ((vm & MSK_RIGHT_NODATA_OR_DISPARITY_RANGE_MISSING) == 0)
    vm[upd] += MSK_RIGHT_NODATA_OR_DISPARITY_RANGE_MISSING
    if offset > 0:
        vm[:offset, :] = MSK_LEFT_NODATA_OR_BORDER
        vm[-offset:, :] = MSK_LEFT_NODATA_OR_BORDER
        vm[offset:-offset, :offset] = MSK_LEFT_NODATA_OR_BORDER
        vm[offset:-offset, -offset:] = MSK_LEFT_NODATA_OR_BORDER


# --------------------------------------------------------------------------------------------
# 3x3 NaN-aware median (CBCA pre-filter): filter/median.py:134-179 + common.py:184-199
# --------------------------------------------------------------------------------------------
def median_filter3(data: np.ndarray, size: int = 3) -> np.ndarray:
    data = np.asarray(data)
    out = np.copy(data)
    invalid = np.isnan(out)
    ny_, nx_ = data.shape
    r = size // 2
    if ny_ >= size and nx_ >= size:
        shp = (ny_ - size + 1, nx_ - size + 1, size, size)
        win = np.lib.stride_tricks.as_strided(data, shape=shp, strides=data.strides + data.strides)
        with warnings.catch_warnings():
            warnings.filterwarnings("ignore", r"All-NaN (slice|axis) encountered")
            # the reference chunks by 100x100 only to bound memory; chunking does not change values
            for y0 in range(0, shp[0], 100):
                for x0 in range(0, shp[1], 100):
                    blk = win[y0 : y0 + 100, x0 : x0 + 100]
                    out[r + y0 : r + y0 + blk.shape[0], r + x0 : r + x0 + blk.shape[1]] = np.nanmedian(blk, axis=(2, 3))
    out[invalid] = np.nan
    return out


# --------------------------------------------------------------------------------------------
# CBCA: aggregation/cbca.py:90-295 + cpp/src/aggregation.cpp:28-355
# --------------------------------------------------------------------------------------------
def cross_support(img, len_arms: int, intensity: float) -> np.ndarray:
    img = _f32(img)
    H, W = img.shape
    out = np.empty((H, W, 4), dtype=np.int16)
    lib().pbo_cross_support(_p(img, ctypes.c_float), H, W, int(len_arms), float(intensity), _p(out, ctypes.c_int16))
    return out


def computes_cross_supports(left, right, offset: int, distance: int, intensity: float):
    """cbca.py:184-295 for subpix == 1 without masks: median, NaN->inf, crop by offset, arms."""
    res = []
    for img in (left, right):
        m = median_filter3(np.copy(np.asarray(img)))
        m = np.nan_to_num(m, copy=False, nan=np.inf)
        if offset:
            m = m[offset:-offset, offset:-offset]
        res.append(cross_support(m, distance, intensity))
    return res[0], res[1]


def cbca_slice(cost2d: np.ndarray, cross_l: np.ndarray, cross_r: np.ndarray, d: int):
    """One call of the reference's ``aggregation_cpp.cbca`` (valid columns derived from d)."""
    H, W = cost2d.shape
    assert cost2d.dtype == np.float32
    step4 = np.empty((H, W), dtype=np.float32)
    sum4 = np.empty((H, W), dtype=np.float32)
    rs, cs = (s // 4 for s in cost2d.strides)
    rc = lib().pbo_cbca_slice(
        _p(cost2d, ctypes.c_float), rs, cs, H, W, int(d), _p(cross_l, ctypes.c_int16), _p(cross_r, ctypes.c_int16),
        _p(step4, ctypes.c_float), _p(sum4, ctypes.c_float),
    )
    if rc:
        raise MemoryError("pbo_cbca_slice")
    return step4, sum4


def cbca_cost_volume(left, right, cv: np.ndarray, offset: int, dmin: int, distance: int = 5, intensity: float = 30.0,
                     cmax=None):
    """cbca.py:127-182: returns (aggregated copy of cv, new cmax)."""
    cross_l, cross_r = computes_cross_supports(left, right, offset, distance, intensity)
    out = np.array(cv, dtype=np.float32, copy=True)
    view = out[offset:-offset, offset:-offset] if offset else out
    src = np.ascontiguousarray(view)
    H, W, D = src.shape
    agg = np.empty_like(src)
    rc = lib().pbo_cbca_volume(
        _p(src, ctypes.c_float), H, W, D, int(dmin), _p(cross_l, ctypes.c_int16), _p(cross_r, ctypes.c_int16), _p(agg, ctypes.c_float)
    )
    if rc:
        raise MemoryError("pbo_cbca_volume")
    view[...] = agg
    return out, (None if cmax is None else cmax * ((distance * 2) - 1) ** 2)


# --------------------------------------------------------------------------------------------
# SGM (unpinned, see module docstring) and WTA
# --------------------------------------------------------------------------------------------
def sgm_invalid_value(cmax: float, p2: float) -> float:
    return float(cmax + p2 + 1)


def sgm_cost_volume(cv: np.ndarray, p1: float = 8, p2: float = 32, cmax: float = 25, type_measure: str = "min",
                    overcounting: bool = False, n_dirs: int = 8) -> np.ndarray:
    cv = _f32(cv)
    H, W, D = cv.shape
    src = -cv if type_measure == "max" else cv
    out = np.empty_like(src)
    rc = lib().pbo_sgm(_p(src, ctypes.c_float), H, W, D, float(p1), float(p2), sgm_invalid_value(cmax, p2),
                       int(bool(overcounting)), int(n_dirs), _p(out, ctypes.c_float))
    if rc:
        raise ValueError(f"pbo_sgm failed: {rc}")
    return -out if type_measure == "max" else out


def sgm_min_cost_paths(cv: np.ndarray, p1: float = 8, p2: float = 32, cmax: float = 25, overcounting: bool = False):
    """SGM with the plugin's ``min_cost_paths`` option (plugin_libsgm.rst:411-413): (optimised volume, nb_of_directions (H, W))."""
    cv = _f32(cv)
    H, W, D = cv.shape
    out, nb = np.empty_like(cv), np.empty((H, W), dtype=np.float32)
    fn = lib().pbo_sgm_min_cost_paths
    fn.restype = ctypes.c_int
    rc = fn(_p(cv, ctypes.c_float), ctypes.c_int(H), ctypes.c_int(W), ctypes.c_int(D), ctypes.c_float(p1), ctypes.c_float(p2),
            ctypes.c_float(sgm_invalid_value(cmax, p2)), ctypes.c_int(int(bool(overcounting))), _p(out, ctypes.c_float), _p(nb, ctypes.c_float))
    if rc:
        raise ValueError(f"pbo_sgm_min_cost_paths failed: {rc}")
    return out, nb


def sgm_direction(C: np.ndarray, S: np.ndarray, p1: float, p2: float, direction: int, init: bool, halo_in=None, halo_out=None):
    """One direction (index into E, W, S, SE, SW, N, NE, NW) on a row tile, in place on S; C holds no NaN."""
    H, W, D = C.shape
    assert C.dtype == np.float32 and S.dtype == np.float32 and C.flags.c_contiguous and S.flags.c_contiguous
    hi = None if halo_in is None else _p(halo_in, ctypes.c_float)
    ho = None if halo_out is None else _p(halo_out, ctypes.c_float)
    rc = lib().pbo_sgm_direction(_p(C, ctypes.c_float), _p(S, ctypes.c_float), H, W, D, float(p1), float(p2), int(direction),
                                 int(bool(init)), hi, ho)
    if rc:
        raise ValueError(f"pbo_sgm_direction failed: {rc}")


def wta(cv: np.ndarray, disps, type_measure: str = "min", invalid_disparity: float = -9999.0):
    """disparity.py:434-455, 483-553 restated with numpy itself (argmin/argmax over axis 2)."""
    cv = np.asarray(cv)
    disps = np.asarray(disps)
    nan = np.isnan(cv)
    work = cv.copy()
    if type_measure == "max":
        work[nan] = -np.inf
        idx = np.argmax(work, axis=2)
    else:
        work[nan] = np.inf
        idx = np.argmin(work, axis=2)
    disp = disps[idx].astype(np.float32)
    invalid_mc = np.min(nan, axis=2)
    disp[invalid_mc] = invalid_disparity
    return disp, invalid_mc


def wta_c(cv: np.ndarray, disps, type_measure: str = "min", invalid_disparity: float = -9999.0):
    """Same through the C restatement (used for the timed CPU baseline on large volumes)."""
    cv = _f32(cv)
    H, W, D = cv.shape
    dc = _f32(disps)
    disp = np.empty((H, W), dtype=np.float32)
    inv = np.empty((H, W), dtype=np.uint8)
    lib().pbo_wta(_p(cv, ctypes.c_float), H, W, D, _p(dc, ctypes.c_float), int(type_measure == "max"),
                  float(invalid_disparity), _p(disp, ctypes.c_float), _p(inv, ctypes.c_uint8))
    return disp, inv.astype(bool)


def wta_validity_mask(vm: np.ndarray, invalid_mc: np.ndarray) -> np.ndarray:
    """disparity.py:468-474."""
    out = vm.copy()
    new_inv = invalid_mc & ((out & MSK_INVALID) == 0)
    out[new_inv] = MSK_INVALID
    return out


def reverse_cost_volume(left_cv: np.ndarray, min_disp: int) -> np.ndarray:
    """matching_cost/cpp/src/matching_cost.cpp:26-57."""
    left_cv = _f32(left_cv)
    H, W, D = left_cv.shape
    out = np.empty_like(left_cv)
    lib().pbo_reverse_cost_volume(_p(left_cv, ctypes.c_float), H, W, D, int(min_disp), _p(out, ctypes.c_float))
    return out


def reverse_disp_range(left_min: np.ndarray, left_max: np.ndarray):
    """matching_cost/cpp/src/matching_cost.cpp:59-131 (called by state_machine.py:673-675)."""
    left_min, left_max = _f32(left_min), _f32(left_max)
    H, W = left_min.shape
    rmin, rmax = np.empty_like(left_min), np.empty_like(left_max)
    lib().pbo_reverse_disp_range(_p(left_min, ctypes.c_float), _p(left_max, ctypes.c_float), H, W, _p(rmin, ctypes.c_float), _p(rmax, ctypes.c_float))
    return rmin, rmax


# --------------------------------------------------------------------------------------------
# input masks and variable disparity grids.
# criteria.py:36-63, 66-288 ; matching_cost/matching_cost.py:484-602, 770-872 ; cpp/src/criteria.cpp:27-110
# --------------------------------------------------------------------------------------------
def binary_dilation_nodata(msk: np.ndarray, no_data: int, window: int) -> np.ndarray:
    """criteria.py:36-63: scipy.ndimage.binary_dilation of ``msk == no_data`` with a full window x window structure
    (odd window: the centred square neighbourhood; outside the image counts as background)."""
    nd = np.asarray(msk) == no_data
    H, W = nd.shape
    half = (window - 1) // 2
    pad = np.zeros((H + 2 * half, W + 2 * half), dtype=bool)
    pad[half : half + H, half : half + W] = nd
    out = np.zeros((H, W), dtype=bool)
    for dy in range(window):
        for dx in range(window):
            out |= pad[dy : dy + H, dx : dx + W]
    return out


def partially_missing_variable_ranges(grid_min, grid_max, right_invalid: np.ndarray) -> np.ndarray:
    """cpp/src/criteria.cpp:27-110: True where [col + dmin, col + dmax] is not inside one run of unmasked right pixels."""
    right_invalid = np.asarray(right_invalid, dtype=bool)
    H, W = right_invalid.shape
    out = np.ones((H, W), dtype=bool)
    for r in range(H):
        runs, last = [], True
        for c in range(W):
            if right_invalid[r, c] != last:
                runs.append(c)
                last = bool(right_invalid[r, c])
        if not last:
            runs.append(W)
        for c in range(W):
            lo, hi = int(grid_min[r, c]) + c, int(grid_max[r, c]) + c
            out[r, c] = not any(runs[i] <= lo and hi < runs[i + 1] for i in range(0, len(runs) - 1, 2))
    return out


def validity_mask_with_masks(H: int, W: int, dmin: int, dmax: int, offset: int, window: int, left_msk=None, right_msk=None,
                             valid_pixels: int = 0, no_data: int = 1, grid_min=None, grid_max=None) -> np.ndarray:
    """criteria.validity_mask (criteria.py:66-158) with optional left / right ``msk`` (allocate_left_mask :178-213,
    allocate_right_mask :216-288) and, with a right mask and disparity grids, mask_partially_missing_variable_ranges
    (:161-175).  ``dmin`` / ``dmax`` are the cost volume's global range."""
    vm = validity_mask(H, W, dmin, dmax, offset)
    col = np.arange(W)
    if dmax < 0:
        bit_1 = np.where((col + dmax) < offset)[0]
    elif dmin > 0:
        bit_1 = np.where((col + dmin) > (W - 1 - offset))[0]
    else:
        bit_1 = np.array([], dtype=int)
    if left_msk is not None:
        left_msk = np.asarray(left_msk)
        vm += binary_dilation_nodata(left_msk, no_data, window).astype(np.uint16) * np.uint16(MSK_LEFT_NODATA_OR_BORDER)
        vm += np.where((left_msk != no_data) & (left_msk != valid_pixels), MSK_IN_VALIDITY_MASK_LEFT, 0).astype(np.uint16)
    if right_msk is not None:
        right_msk = np.asarray(right_msk)
        dil = binary_dilation_nodata(right_msk, no_data, window)
        r_mask = ((right_msk != no_data) & (right_msk != valid_pixels)).astype(np.int64)
        b_2_7 = np.zeros((H, W), dtype=np.int64)
        no_data_right = np.zeros((H, W), dtype=np.int64)
        nd = dmax - dmin + 1
        for dsp in range(dmin, dmax + 1):
            col_d = col + dsp
            ok = (col_d >= offset) & (col_d <= W - 1 - offset)
            b_2_7[:, col[ok]] += r_mask[:, col_d[ok]]
            b_2_7[:, col[~ok]] += 1
            no_data_right[:, col[ok]] += dil[:, col_d[ok]]
            no_data_right[:, col[~ok]] += 1
            b_2_7[:, bit_1] = 0
            no_data_right[:, bit_1] = 0
            vm[b_2_7 == nd] += MSK_IN_VALIDITY_MASK_RIGHT
            vm[no_data_right == nd] += MSK_RIGHT_NODATA_OR_DISPARITY_RANGE_MISSING
        if grid_min is not None:
            miss = partially_missing_variable_ranges(grid_min, grid_max, right_msk != valid_pixels)
            vm[miss] |= MSK_INCOMPLETE_VARIABLE_DISPARITY_RANGE
    return vm


def cv_masked_full(cv: np.ndarray, vm: np.ndarray, offset: int, window: int, dmin: int, left_msk=None, right_msk=None,
                   valid_pixels: int = 0, no_data: int = 1, grid_min=None, grid_max=None) -> None:
    """matching_cost.py:770-872 for subpix == 1, step == 1, in place on ``cv`` and ``vm``: NaN for cells whose left pixel
    or right pixel (col + d) is invalid or has a no_data in its window (masks_dilatation :484-602), NaN outside the
    per-pixel [grid_min, grid_max], then mask_invalid_variable_disparity_range and mask_border."""
    H, W, D = cv.shape

    def dilated(msk):
        out = np.zeros((H, W), dtype=np.float64)
        if msk is not None:
            msk = np.asarray(msk)
            out[(msk != valid_pixels) & (msk != no_data)] = np.nan
            out[binary_dilation_nodata(msk, no_data, window)] = np.nan
        return out

    m_left, m_right = dilated(left_msk), dilated(right_msk)
    for k in range(D):
        d = dmin + k
        p0, p1 = max(0 - d, 0), min(W - 1 - d, W - 1)                  # mask_column_interval_without_step :655-708
        if p0 > p1:
            continue
        cv[:, p0 : p1 + 1, k] = (cv[:, p0 : p1 + 1, k] + m_right[:, p0 + d : p1 + d + 1] + m_left[:, p0 : p1 + 1]).astype(np.float32)
    if grid_min is not None:
        gmin, gmax = np.asarray(grid_min)[:H, :W], np.asarray(grid_max)[:H, :W]
        for k in range(D):
            d = dmin + k
            cv[:, :, k][(d < gmin) | (d > gmax)] = np.nan
    cv_masked(cv, vm, offset)


# --------------------------------------------------------------------------------------------
# right disparity map of the fast cross-checking + the consistency check
# state_machine.py:436-448 ; validation/validation.py:226-371
# --------------------------------------------------------------------------------------------
def right_disparity_fast(left_cv: np.ndarray, dmin: int, dmax: int, type_measure: str = "min", invalid_disparity: float = -9999.0):
    """state_machine.py:438-448: reverse_cost_volume(left_cv, -dmax) followed by WTA on the right disparity range."""
    right_cv = reverse_cost_volume(left_cv, -dmax)
    return wta(right_cv, np.arange(-dmax, -dmin + 1), type_measure, invalid_disparity)


def cross_checking(disp_left, mask_left, disp_right, threshold: float, dmin: int, dmax: int, offset: int = 0):
    """CrossCheckingAccurate.disparity_checking, validation/validation.py:226-371, restated line by line (row loop and
    fancy indexing included).  Returns (updated validity mask, left-right distance map).  ``dmin`` / ``dmax`` = the
    left map's disparity interval (disparity.py:334-347).  Like the reference it assumes that a pixel whose disparity
    is NaN is flagged invalid; its "outside the right image" branch (validation.py:355-357) can never fire
    (``(col_right < 0) & (col_right >= nb_col)``) and is restated as such."""
    disp_left = np.asarray(disp_left, dtype=np.float32)
    disp_right = np.asarray(disp_right, dtype=np.float32)
    vm = np.array(mask_left, dtype=np.uint16, copy=True)
    nb_row, nb_col = disp_left.shape
    disparity_range = np.arange(dmin, dmax + 1)
    conf = np.full((nb_row, nb_col), np.nan, dtype=np.float32)
    thr = np.float32(threshold)
    for row in range(nb_row):
        valid_pixel = np.where((vm[row, :] & MSK_INVALID) == 0)
        col_left = np.arange(nb_col, dtype=np.int64)[valid_pixel]
        col_right = col_left + disp_left[row, col_left]
        col_right = col_right[np.logical_not(np.isnan(col_right))]
        col_right = np.rint(col_right).astype(int)
        inside_right = np.where((col_right >= 0) & (col_right < nb_col))
        right_disp = disp_right[row, col_right[inside_right]]
        right_disp[np.isnan(right_disp)] = np.inf
        left_disp = disp_left[row, col_left[inside_right]]
        left_disp[np.isnan(left_disp)] = np.inf
        conf[row, col_left[inside_right]] = np.abs(right_disp + left_disp)
        invalid = np.abs(right_disp + left_disp) > thr
        cols_inv = col_left[inside_right][invalid]
        index = np.tile(disparity_range, (len(cols_inv), 1)).astype(np.float32) + np.tile(cols_inv, (len(disparity_range), 1)).transpose()
        inside_col_disp = np.where((index >= 0) & (index < nb_col))
        dr = np.full(index.shape, np.inf, dtype=np.float32)
        dr[inside_col_disp] = disp_right[row, index[inside_col_disp].astype(int)]
        comp = np.rint(dr) == np.tile(-1 * disparity_range, (len(cols_inv), 1)).astype(np.float32)
        comp = np.sum(comp, axis=1)
        comp[comp > 1] = 1
        vm[row, cols_inv] += np.uint16(MSK_OCCLUSION)
        vm[row, cols_inv] += (MSK_MISMATCH * comp).astype(np.uint16)
        vm[row, cols_inv] -= (MSK_OCCLUSION * comp).astype(np.uint16)
    if offset > 0:
        vm[:offset, :] = MSK_LEFT_NODATA_OR_BORDER
        vm[-offset:, :] = MSK_LEFT_NODATA_OR_BORDER
        vm[offset:-offset, :offset] = MSK_LEFT_NODATA_OR_BORDER
        vm[offset:-offset, -offset:] = MSK_LEFT_NODATA_OR_BORDER
    return vm, conf


# --------------------------------------------------------------------------------------------
# sub-pixel refinement (refinement/refinement.py:78-181 + refinement/cpp/src/*.cpp)
# --------------------------------------------------------------------------------------------
def refinement(cv, disp, mask, d_min: float, d_max: float, subpix: int = 1, type_measure: str = "min", method: str = "vfit",
               approximate: bool = False):
    """loop_refinement / loop_approximate_refinement with vfit or quadratic.  Returns (itp_coeff, disp, mask) like the
    reference's C++ (refinement.cpp:29-181); inputs are not modified."""
    cv = _f32(cv)
    H, W, D = cv.shape
    disp = np.array(disp, dtype=np.float32, copy=True)
    mask = np.array(mask, dtype=np.uint16, copy=True)
    itp = np.empty((H, W), dtype=np.float32)
    lib().pbo_refinement(_p(cv, ctypes.c_float), H, W, D, float(d_min), float(d_max), int(subpix), int(type_measure == "max"),
                         {"vfit": 0, "quadratic": 1}[method], int(bool(approximate)), _p(disp, ctypes.c_float),
                         _p(mask, ctypes.c_uint16), _p(itp, ctypes.c_float))
    return itp, disp, mask


# --------------------------------------------------------------------------------------------
# cost-volume confidence (cost_volume_confidence/cpp/src/{ambiguity,risk}.cpp)
# --------------------------------------------------------------------------------------------
def ambiguity(cv, etas, grids, disparity_range_, sampled: bool = False):
    """compute_ambiguity_and_sampled_ambiguity (ambiguity.cpp:28-142) on a min-type volume; etas are cast to float32
    like the pybind11 signature does.  grids = (2, H, W) integer [disp_min, disp_max] per pixel."""
    cv = _f32(cv)
    H, W, D = cv.shape
    et = _f32(etas)
    gr = np.ascontiguousarray(grids, dtype=np.int32)
    dr = _f32(disparity_range_)
    amb = np.empty((H, W), dtype=np.float32)
    samp = np.empty((H, W, len(et)), dtype=np.float32) if sampled else None
    lib().pbo_ambiguity(_p(cv, ctypes.c_float), H, W, D, _p(et, ctypes.c_float), len(et), _p(gr, ctypes.c_int32), _p(dr, ctypes.c_float),
                        _p(amb, ctypes.c_float), _p(samp, ctypes.c_float) if sampled else None)
    return (amb, samp) if sampled else amb


def risk(cv, sampled_ambiguity, etas, grids, disparity_range_, sampled: bool = False):
    """compute_risk_and_sampled_risk (risk.cpp:28-197): returns (risk_max, risk_min, disp_sup, disp_inf[, samp_max, samp_min])."""
    cv = _f32(cv)
    H, W, D = cv.shape
    sa = _f32(sampled_ambiguity)
    et = np.ascontiguousarray(etas, dtype=np.float64)
    gr = np.ascontiguousarray(grids, dtype=np.int32)
    dr = _f32(disparity_range_)
    outs = [np.empty((H, W), dtype=np.float32) for _ in range(4)]
    samp = [np.empty((H, W, len(et)), dtype=np.float32) for _ in range(2)] if sampled else [None, None]
    lib().pbo_risk(_p(cv, ctypes.c_float), _p(sa, ctypes.c_float), H, W, D, _p(et, ctypes.c_double), len(et), _p(gr, ctypes.c_int32),
                   _p(dr, ctypes.c_float), *[_p(o, ctypes.c_float) for o in outs],
                   *[(_p(o, ctypes.c_float) if o is not None else None) for o in samp])
    return tuple(outs) + (tuple(samp) if sampled else ())


# --------------------------------------------------------------------------------------------
# deterministic synthetic stereo pair used by tests and bench
# --------------------------------------------------------------------------------------------
def synthetic_pair(H: int, W: int, D: int, seed: int = 20240607, block: int = 64):
    """Integer-valued float32 pair in [0, 255]; right = left warped by a piecewise-constant
    disparity field g in [-(D-1), 0] (left(r,c) ~ right(r, c+g))."""
    rng = np.random.default_rng(seed)
    tex = rng.integers(0, 256, (H + 2, W + D + 2)).astype(np.int64)
    sm = sum(tex[dy : dy + H, dx : dx + W + D] for dy in range(3) for dx in range(3)) // 9
    g = -rng.integers(0, D, ((H + block - 1) // block, 1))      # one disparity per band of `block` rows
    gfull = np.kron(g, np.ones((block, W), dtype=np.int64))[:H, :W]
    cols = np.arange(W)[None, :]
    left = sm[:, D : D + W]
    # right(r, c) = T[r, D + c - g(r,c)]  so that left(r,c) == right(r, c + g) inside a block
    right = np.take_along_axis(sm, np.clip(D + cols - gfull, 0, W + D - 1), axis=1)
    return left.astype(np.float32), right.astype(np.float32), gfull.astype(np.float32)
