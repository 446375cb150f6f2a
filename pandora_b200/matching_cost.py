"""matching_cost step: mirror of the reference's AbstractMatchingCost plugin API
(src/pandora/matching_cost/matching_cost.py:45-950) with B200 kernels behind compute_cost_volume.

Same factory / registry / method names / config keys / error behaviour as the reference:
``AbstractMatchingCost(**cfg)`` dispatches on ``cfg["matching_cost_method"]`` (KeyError when unknown,
matching_cost.py:80-107); subclasses register with ``@AbstractMatchingCost.register_subclass``.
The compute happens in ``Engine`` (CUDA); nothing here falls back to a CPU implementation.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from . import constants as cst
from . import _native
from ._common import ConfigError, deferred_recipe, device_var, device_volume, store_var, get_engine, image_array, store_deferred_volume, store_volume
from .dataset import Dataset


class AbstractMatchingCost:
    """Abstract matching-cost step (reference: matching_cost.py:45-131)."""

    matching_cost_methods_avail: Dict[str, type] = {}
    _WINDOW_SIZE, _SUBPIX, _BAND, _STEP_COL, _SPLINE_ORDER = 5, 1, None, 1, 1
    _SUBPIX_KERNELS = False          # measures whose kernels take the list of shifted right images (subpix 2 / 4)
    _VALID_WINDOWS: Optional[Tuple[int, ...]] = None

    def __new__(cls, **cfg):
        if cls is AbstractMatchingCost:
            method = cfg.get("matching_cost_method")
            try:
                return super().__new__(cls.matching_cost_methods_avail[method])
            except (KeyError, TypeError):
                raise KeyError(f"No matching cost method named {method} supported") from None
        return super().__new__(cls)

    @classmethod
    def register_subclass(cls, short_name: str, *args):
        def decorator(subclass):
            cls.matching_cost_methods_avail[short_name] = subclass
            for arg in args:
                cls.matching_cost_methods_avail[arg] = subclass
            return subclass

        return decorator

    def __init__(self, **cfg):
        self.instantiate_class(**cfg)

    # -- configuration: matching_cost.py:140-184 ----------------------------------------------------
    def instantiate_class(self, **cfg) -> None:
        self.cfg = self.check_conf(**cfg)
        self._window_size = int(self.cfg["window_size"])
        self._subpix = int(self.cfg["subpix"])
        self._band = self.cfg["band"]
        self._step_col = int(self.cfg["step"])
        self._method = str(self.cfg["matching_cost_method"])
        self._spline_order = int(self.cfg["spline_order"])
        del self.cfg["spline_order"]

    def check_conf(self, **cfg) -> dict:
        cfg.setdefault("window_size", self._WINDOW_SIZE)
        cfg.setdefault("subpix", self._SUBPIX)
        cfg.setdefault("band", self._BAND)
        if "step" in cfg and cfg["step"] != 1:
            raise ValueError("Step parameter cannot be different from 1")
        cfg.setdefault("step", self._STEP_COL)
        cfg.setdefault("spline_order", self._SPLINE_ORDER)
        allowed = {"matching_cost_method", "window_size", "subpix", "band", "step", "spline_order"}
        for key in cfg:
            if key not in allowed:
                raise ConfigError(f"Unknown key {key!r} in the matching_cost configuration")
        w = cfg["window_size"]
        if not isinstance(w, int) or isinstance(w, bool):
            raise ConfigError(f"window_size must be an int, got {w!r}")
        if self._VALID_WINDOWS is not None and w not in self._VALID_WINDOWS:
            raise ConfigError(f"window_size {w} not in {self._VALID_WINDOWS}")
        if self._VALID_WINDOWS is None and (w < 1 or w % 2 == 0):
            raise ConfigError(f"window_size {w} must be odd and > 0")
        if cfg["subpix"] not in (1, 2, 4):
            raise ConfigError(f"subpix {cfg['subpix']} not in [1, 2, 4]")
        if cfg["subpix"] != 1 and not self._SUBPIX_KERNELS:
            raise ConfigError("subpix: only 1 is implemented by the B200 kernels of this measure (Census takes 2 and 4)")
        if not (cfg["band"] is None or isinstance(cfg["band"], str)):
            raise ConfigError("band must be a str or None")
        if not (isinstance(cfg["spline_order"], int) and 1 <= cfg["spline_order"] <= 5):
            raise ConfigError("spline_order must be an int in [1, 5]")
        return cfg

    def desc(self) -> None:
        print(f"{self._method} similarity measure")

    # -- cost-volume container: matching_cost.py:330-427 ---------------------------------------------
    @staticmethod
    def get_min_max_from_grid(disp_min, disp_max) -> Tuple[int, int]:
        return int(np.nanmin(disp_min)), int(np.nanmax(disp_max))

    @staticmethod
    def constant_range(image):
        """(dmin, dmax) when the image's disparity grids are constant BY CONSTRUCTION -- ``attrs["disparity_source"]`` is the
        [min, max] pair ``add_disparity`` built them from (img_tools.py:141-161) -- else None.  Saves four NaN-reductions over
        the (row, col) grids per step (70 ms of host time at 4096 x 4096)."""
        src = getattr(image, "attrs", {}).get("disparity_source")
        if isinstance(src, (list, tuple)) and len(src) == 2 and all(isinstance(v, (int, np.integer)) for v in src):
            return int(src[0]), int(src[1])
        return None

    @staticmethod
    def reverse_disp_range(left_min: np.ndarray, left_max: np.ndarray):
        """Right disparity grids from the left ones (matching_cost.py:937-950 -> matching_cost_cpp.reverse_disp_range,
        cpp/src/matching_cost.cpp:59-131), computed on the device: float32 (row, col) arrays in, (right_min, right_max) out."""
        eng = get_engine()
        rmin, rmax = eng.reverse_disp_range(eng.to_device(np.asarray(left_min, dtype=np.float32)),
                                            eng.to_device(np.asarray(left_max, dtype=np.float32)))
        return rmin.cpu().numpy(), rmax.cpu().numpy()

    @staticmethod
    def get_disparity_range(disparity_min: int, disparity_max: int, subpix: int = 1) -> np.ndarray:
        """matching_cost.py:410-427."""
        if subpix == 1:
            return np.arange(disparity_min, disparity_max + 1)
        return np.append(np.arange(disparity_min, disparity_max, 1 / float(subpix), dtype=np.float64), [disparity_max])

    def allocate_cost_volume(self, image, disparity_grids, cfg=None) -> Dataset:
        """Empty cost-volume dataset with the reference's coordinates and attributes (matching_cost.py:
        377-407).  The NaN-filled host array of the reference is NOT allocated: the volume is created on
        the device by compute_cost_volume (17 GB at 4096x4096x256 would be written twice otherwise)."""
        c_row = np.asarray(image.coords["row"].data)
        c_col = np.asarray(image.coords["col"].data)
        dmin, dmax = self.constant_range(image) or self.get_min_max_from_grid(*disparity_grids)
        disps = self.get_disparity_range(dmin, dmax, self._subpix)
        index_col = np.arange(c_col[0], c_col[-1] + 1, self._step_col)
        cv = Dataset(coords={"row": c_row, "col": index_col, "disp": disps}, attrs=dict(image.attrs))
        cv.attrs.update({
            "sampling_interval": self._step_col,
            "col_to_compute": index_col,
            "window_size": self._window_size,
            "subpixel": self._subpix,
            "band_correl": self._band,
            "offset_row_col": int((self._window_size - 1) / 2),
            "measure": self._method,
        })
        return cv

    # -- compute -------------------------------------------------------------------------------------
    def compute_cost_volume(self, img_left, img_right, cost_volume):
        raise NotImplementedError

    def _disp_bounds(self, cost_volume) -> Tuple[int, int]:
        disps = np.asarray(cost_volume.coords["disp"].data)
        return int(round(float(disps[0]))), int(round(float(disps[-1])))

    def cv_masked(self, img_left, img_right, cost_volume, disp_min, disp_max) -> None:
        """matching_cost.py:770-872 on the device: cells whose left / right pixel is invalid or sees a no_data in its
        window, and cells outside the per-pixel [disp_min, disp_max], become NaN in ONE pass over the volume (the
        reference loops over the disparities in Python twice); then ``mask_invalid_variable_disparity_range`` and
        ``mask_border`` update the validity mask.  Without masks and with a fixed range the volume is only read."""
        from .criteria import image_mask_flags  # noqa: PLC0415

        eng = get_engine()
        dmin, dmax = self._disp_bounds(cost_volume)
        H, W = (int(s) for s in cost_volume["cost_volume"].shape[:2])
        off = int(cost_volume.attrs["offset_row_col"])
        fl, fr = image_mask_flags(eng, img_left, self._window_size), image_mask_flags(eng, img_right, self._window_size)
        gmin_h, gmax_h = np.asarray(disp_min, dtype=np.float32)[:H, :W], np.asarray(disp_max, dtype=np.float32)[:H, :W]
        own_grids = "disparity" in getattr(img_left, "data_vars", {}) and np.may_share_memory(gmin_h, img_left["disparity"].data)
        if own_grids and self.constant_range(img_left) == (dmin, dmax):
            variable = False                                   # the grids of this very image, constant by construction
        else:
            variable = bool(np.nanmax(gmin_h) != np.nanmin(gmin_h) or np.nanmax(gmax_h) != np.nanmin(gmax_h)
                            or int(np.nanmin(gmin_h)) != dmin or int(np.nanmax(gmax_h)) != dmax)
        recipe = deferred_recipe(cost_volume)
        if recipe is not None and fl is None and fr is None and not variable:
            # nothing to mask: the step only needs the all-NaN pixels, which a deferred Census volume knows from its
            # geometry -- the volume stays deferred (a following SGM step then runs the fused Census -> SGM kernels)
            flags = recipe.all_nan_flags()
        else:
            if self._subpix > 1 and (fl is not None or fr is not None or variable):
                raise ConfigError("subpix > 1 together with image masks or variable disparity grids is not implemented by the B200 kernels")
            cv_t = device_volume(eng, cost_volume)
            gmin = eng.to_device(gmin_h) if variable else None
            gmax = eng.to_device(gmax_h) if variable else None
            flags = eng.cv_masked(cv_t, dmin, fl, fr, gmin, gmax)            # masks the volume, reports all-NaN pixels
            store_volume(cost_volume, cv_t)
        if "validity_mask" in cost_volume:
            mask = device_var(eng, cost_volume, "validity_mask", "uint16")
        else:
            mask = eng.validity_mask_init(H, W, dmin, dmax, off)
        mask = eng.validity_mask(H, W, dmin, dmax, off, flags, mask=mask)
        store_var(cost_volume, "validity_mask", mask, dtype="uint16")


def shift_right_img(right: np.ndarray, subpix: int, order: int = 1):
    """img_tools.shift_right_img (img_tools.py:713-752) on a plain array: [right, right resampled at +1/subpix, ...]; the
    resampled copies are one column shorter."""
    from scipy.ndimage import zoom  # noqa: PLC0415

    nx = right.shape[1]
    out = [right]
    for ind in range(1, subpix):
        out.append(zoom(right, (1, (nx * subpix - (subpix - 1)) / float(nx)), order=order)[:, ind::subpix])
    return out


class CensusRecipe:
    """A Census cost volume that has not been computed yet (``LazyVolume(recipe=...)``): ``compute()`` runs the fill,
    ``all_nan_flags()`` gives what ``cv_masked`` needs without it, and ``Sgm.optimize_cv`` hands the images to the fused
    Census -> SGM kernels (``Engine.census_sgm``) so that the float Census volume is never written."""

    kind = "census"

    def __init__(self, eng, left, right, window: int, dmin: int, dmax: int):
        self.eng, self.left, self.right, self.window, self.dmin, self.dmax = eng, left, right, int(window), int(dmin), int(dmax)

    def compute(self):
        return self.eng.census(self.left, self.right, self.window, self.dmin, self.dmax)

    def all_nan_flags(self):
        """uint8 (H, W), 1 where every cell of the pixel is NaN: the window leaves the left image, or no disparity of the
        range puts it inside the right one (census.cpp:97-180: half <= x + d < W - half)."""
        import torch  # noqa: PLC0415

        H, W = (int(s) for s in self.left.shape)
        half = self.window // 2
        dev = self.left.device
        ys, xs = torch.arange(H, device=dev), torch.arange(W, device=dev)
        row_ok = (ys >= half) & (ys < H - half)
        lo = torch.clamp(half - xs, min=self.dmin)
        hi = torch.clamp(W - half - 1 - xs, max=self.dmax)
        col_ok = (xs >= half) & (xs < W - half) & (lo <= hi)
        return (~(row_ok[:, None] & col_ok[None, :])).to(torch.uint8).contiguous()


@AbstractMatchingCost.register_subclass("census")
class Census(AbstractMatchingCost):
    """Census matching cost (reference: matching_cost/census.py:39-153, cpp/src/census.cpp)."""

    _VALID_WINDOWS = (3, 5, 7, 9, 11, 13)
    _SUBPIX_KERNELS = True

    def compute_cost_volume(self, img_left, img_right, cost_volume):
        eng = get_engine()
        dmin, dmax = self._disp_bounds(cost_volume)
        cost_volume.attrs.update({"type_measure": "min", "cmax": int(self._window_size**2)})     # census.py:116-122
        left = eng.to_device(image_array(img_left, self._band))
        if self._subpix > 1:
            # census.py:113 + img_tools.shift_right_img (img_tools.py:713-752): the right image and its subpix - 1 copies
            # resampled at column offsets i / subpix -- the same scipy.ndimage.zoom call as the reference, on the host (an
            # O(H W) pre-processing of the input image); the list goes to the device like compute_matching_costs takes it
            rights = [eng.to_device(np.ascontiguousarray(r, dtype=np.float32)) for r in shift_right_img(image_array(img_right, self._band),
                                                                                                     self._subpix, self._spline_order)]
            n_disp = len(np.asarray(cost_volume.coords["disp"].data))
            store_volume(cost_volume, eng.census_subpix(left, rights, self._window_size, dmin, n_disp))
            return cost_volume
        right = eng.to_device(image_array(img_right, self._band))
        if _native.get_option("fuse_census_sgm") != 0:
            # deferred: computed when something reads it; a directly following SGM step fuses it away (CensusRecipe)
            recipe = CensusRecipe(eng, left, right, self._window_size, dmin, dmax)
            if store_deferred_volume(cost_volume, recipe, (left.shape[0], left.shape[1], dmax - dmin + 1)):
                return cost_volume
        store_volume(cost_volume, eng.census(left, right, self._window_size, dmin, dmax))
        return cost_volume


@AbstractMatchingCost.register_subclass("sad", "ssd")
class SadSsd(AbstractMatchingCost):
    """SAD / SSD matching cost (reference: matching_cost/sad_ssd.py:39-368)."""

    def compute_cost_volume(self, img_left, img_right, cost_volume):
        eng = get_engine()
        dmin, dmax = self._disp_bounds(cost_volume)
        l_np, r_np = image_array(img_left, self._band), image_array(img_right, self._band)
        mx = max(abs(np.amax(l_np) - np.amin(r_np)), abs(np.amax(r_np) - np.amin(l_np)))          # sad_ssd.py:125-137
        cmax = int(mx * self._window_size**2) if self._method == "sad" else int(mx**2 * self._window_size**2)
        cost_volume.attrs.update({"type_measure": "min", "cmax": cmax})
        cv = eng.sad_ssd(eng.to_device(l_np), eng.to_device(r_np), self._window_size, dmin, dmax, squared=self._method == "ssd")
        store_volume(cost_volume, cv)
        return cost_volume


@AbstractMatchingCost.register_subclass("zncc")
class Zncc(AbstractMatchingCost):
    """ZNCC matching cost (reference: matching_cost/zncc.py:38-277)."""

    def compute_cost_volume(self, img_left, img_right, cost_volume):
        eng = get_engine()
        dmin, dmax = self._disp_bounds(cost_volume)
        cost_volume.attrs.update({"type_measure": "max", "cmax": 1})                             # zncc.py:171-176
        left = eng.to_device(image_array(img_left, self._band))
        right = eng.to_device(image_array(img_right, self._band))
        store_volume(cost_volume, eng.zncc(left, right, self._window_size, dmin, dmax))
        return cost_volume


def validity_mask_flags():
    """Re-export of the bit flags for callers that only import this module."""
    return cst
