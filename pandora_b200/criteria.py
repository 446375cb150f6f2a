"""criteria: mirror of ``validity_mask`` (src/pandora/criteria.py:66-158) including the input-mask branches
(allocate_left_mask :178-213, allocate_right_mask :216-288, mask_partially_missing_variable_ranges :161-175)."""
from __future__ import annotations

import numpy as np

from ._common import get_engine, store_var


def image_mask_flags(eng, img, window: int):
    """Device flag byte per pixel of ``img["msk"]`` (None when the image carries no mask)."""
    if "msk" not in getattr(img, "data_vars", {}):
        return None
    return eng.mask_flags(np.asarray(img["msk"].data), int(img.attrs["valid_pixels"]), int(img.attrs["no_data_mask"]), window)


def validity_mask(img_left, img_right, cv):
    """Create ``cv["validity_mask"]`` (uint16): disparity-range bits per column, then the left / right image-mask bits."""
    eng = get_engine()
    disps = np.asarray(cv.coords["disp"].data)
    dmin, dmax = int(round(float(disps[0]))), int(round(float(disps[-1])))
    H, W = int(cv.sizes["row"]), int(cv.sizes["col"])
    offset, window = int(cv.attrs["offset_row_col"]), int(cv.attrs["window_size"])
    mask = eng.validity_mask_init(H, W, dmin, dmax, offset)
    fl, fr = image_mask_flags(eng, img_left, window), image_mask_flags(eng, img_right, window)
    if fl is not None or fr is not None:
        gmin = gmax = None
        if fr is not None and "disparity" in getattr(img_left, "data_vars", {}):
            grid = np.asarray(img_left["disparity"].data, dtype=np.float32)
            gmin, gmax = eng.to_device(grid[0]), eng.to_device(grid[1])
        eng.validity_mask_masks(mask, dmin, dmax, offset, fl, fr, gmin, gmax)
    store_var(cv, "validity_mask", mask, dtype="uint16")           # stays in HBM until somebody reads it
    return cv
