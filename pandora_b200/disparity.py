"""disparity step: mirror of AbstractDisparity / WinnerTakesAll (src/pandora/disparity/disparity.py:42-553)."""
from __future__ import annotations

from typing import Dict

import numpy as np

from ._common import ConfigError, device_var, device_volume, fused_wta, get_engine, store_var
from .dataset import Dataset


class AbstractDisparity:
    disparity_methods_avail: Dict[str, type] = {}

    def __new__(cls, **cfg):
        if cls is AbstractDisparity:
            method = cfg.get("disparity_method")
            try:
                return super().__new__(cls.disparity_methods_avail[method])
            except (KeyError, TypeError):
                raise KeyError(f"No disparity method named {method} supported") from None
        return super().__new__(cls)

    @classmethod
    def register_subclass(cls, short_name: str):
        def decorator(subclass):
            cls.disparity_methods_avail[short_name] = subclass
            return subclass

        return decorator

    def to_disp(self, cv, img_left=None, img_right=None):
        raise NotImplementedError


@AbstractDisparity.register_subclass("wta")
class WinnerTakesAll(AbstractDisparity):
    _INVALID_DISPARITY = -9999

    def __init__(self, **cfg):
        self.cfg = self.check_conf(**cfg)
        self._invalid_disparity = self.cfg["invalid_disparity"]

    def check_conf(self, **cfg) -> dict:
        if "invalid_disparity" not in cfg:
            cfg["invalid_disparity"] = self._INVALID_DISPARITY
        elif cfg["invalid_disparity"] == "NaN":
            cfg["invalid_disparity"] = np.nan
        for key in cfg:
            if key not in ("disparity_method", "invalid_disparity"):
                raise ConfigError(f"Unknown key {key!r} in the disparity configuration")
        if not isinstance(cfg["invalid_disparity"], (int, float)):
            raise ConfigError("invalid_disparity must be an int, a float or 'NaN'")
        return cfg

    def desc(self) -> None:
        print("Winner takes all method")

    def to_disp(self, cv, img_left=None, img_right=None):
        """disparity.py:400-480: ``disparity_map`` float32, ``validity_mask`` and ``cv["disp_indices"]``."""
        eng = get_engine()
        cv_t = device_volume(eng, cv)
        H, W, D = (int(s) for s in cv_t.shape)
        disps = np.asarray(cv.coords["disp"].data)
        dmin, dmax = int(round(float(disps[0]))), int(round(float(disps[-1])))
        is_max = cv.attrs.get("type_measure") == "max"
        invalid = float(self._invalid_disparity)
        subpix = int(cv.attrs.get("subpixel", 1) or 1)
        if subpix > 1:
            return self._to_disp_subpix(eng, cv, cv_t, disps, dmin, dmax, is_max, invalid, subpix)
        cached = fused_wta(cv)
        if cached is not None and not is_max and cached[2] == dmin and (cached[3] == invalid or (cached[3] != cached[3] and invalid != invalid)):
            disp_t, flags = cached[0], cached[1]                   # the producing kernel already ran this argmin (fused WTA)
        else:
            disp_t, flags = eng.wta(cv_t, dmin, is_max, invalid)
        out = Dataset(coords={"row": cv.coords["row"].data, "col": cv.coords["col"].data}, attrs=cv.attrs)
        store_var(out, "disparity_map", disp_t)                    # device-resident: read `.data` to get the host copy
        out["disparity_interval"] = (("disparity",), np.asarray(disps)[[0, -1]])                # disparity.py:301-315, 456
        store_var(cv, "disp_indices", disp_t.clone())
        if "validity_mask" in cv:
            mask_t = device_var(eng, cv, "validity_mask", "uint16").clone()
            mask_t = eng.validity_mask(H, W, dmin, dmax, 0, flags, wta_invalidate=True, mask=mask_t)
            store_var(out, "validity_mask", mask_t, dtype="uint16")
        self._carry_confidence(cv, out)
        return out

    def _to_disp_subpix(self, eng, cv, cv_t, disps, dmin, dmax, is_max, invalid, subpix):
        """Sub-pixel volumes (disparity.py:434-455: ``disp["disp"].data[indices]``): the argmin index k maps to dmin + k / subpix."""
        import torch  # noqa: PLC0415

        H, W, D = (int(s) for s in cv_t.shape)
        idx_t, flags = eng.wta(cv_t, 0, is_max, -1.0)                     # index map, -1 where every cell is NaN
        disp_t = torch.where(flags != 0, torch.full_like(idx_t, invalid), float(disps[0]) + idx_t / float(subpix))
        out = Dataset(coords={"row": cv.coords["row"].data, "col": cv.coords["col"].data}, attrs=cv.attrs)
        store_var(out, "disparity_map", disp_t)
        out["disparity_interval"] = (("disparity",), np.asarray(disps)[[0, -1]])
        store_var(cv, "disp_indices", disp_t.clone())
        if "validity_mask" in cv:
            mask_t = device_var(eng, cv, "validity_mask", "uint16").clone()
            mask_t = eng.validity_mask(H, W, dmin, dmax, 0, flags, wta_invalidate=True, mask=mask_t)
            store_var(out, "validity_mask", mask_t, dtype="uint16")
        self._carry_confidence(cv, out)
        return out

    @staticmethod
    def _carry_confidence(cv, out) -> None:
        if "confidence_measure" in cv:
            # disparity.py:462-466: the confidence layers computed on the cost volume travel with their `indicator`
            # coordinate (cost_volume_confidence runs BEFORE disparity in every legal pipeline, state_machine.py:134-139)
            out["confidence_measure"] = cv["confidence_measure"]
            if "indicator" in cv.coords:
                out.coords["indicator"] = cv.coords["indicator"]
