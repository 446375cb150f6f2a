"""refinement step: mirror of AbstractRefinement / Vfit / Quadratic (src/pandora/refinement/refinement.py:38-226,
vfit.py:38-77, quadratic.py:38-79) with the per-pixel loops of refinement/cpp/src/refinement.cpp on the device."""
from __future__ import annotations

from typing import Dict

import numpy as np

from ._common import ConfigError, device_var, device_volume, get_engine, store_var


class NullMargins:
    """margins.NullMargins of the reference (refinement.py:45)."""

    left = up = right = down = 0

    def astuple(self):
        return (0, 0, 0, 0)


class AbstractRefinement:
    subpixel_methods_avail: Dict[str, type] = {}
    margins = NullMargins()
    _method_id = None

    def __new__(cls, **cfg):
        if cls is AbstractRefinement:
            method = cfg.get("refinement_method")
            try:
                return super().__new__(cls.subpixel_methods_avail[method])
            except (KeyError, TypeError):
                raise KeyError(f"No refinement method named {method} supported") from None
        return super().__new__(cls)

    @classmethod
    def register_subclass(cls, short_name: str):
        def decorator(subclass):
            cls.subpixel_methods_avail[short_name] = subclass
            return subclass

        return decorator

    def __init__(self, **cfg):
        self.cfg = self.check_conf(**cfg)
        self._refinement_method_name = str(self.cfg["refinement_method"])

    @staticmethod
    def check_conf(**cfg) -> dict:
        for key in cfg:
            if key != "refinement_method":
                raise ConfigError(f"Unknown key {key!r} in the refinement configuration")
        if not isinstance(cfg.get("refinement_method"), str):
            raise ConfigError("refinement_method must be a str")
        return cfg

    def _run(self, cv, disp, approximate):
        eng = get_engine()
        cv_t = device_volume(eng, cv)
        disps = np.asarray(cv.coords["disp"].data)
        d_min, d_max = float(disps[0]), float(disps[-1])
        if approximate == 2:
            d_min, d_max = -d_max, -d_min
        subpix = int(cv.attrs["subpixel"])
        is_max = cv.attrs["type_measure"] == "max"
        disp_t = device_var(eng, disp, "disparity_map").clone()
        mask_t = device_var(eng, disp, "validity_mask", "uint16").clone()
        itp = eng.refinement(cv_t, disp_t, mask_t, d_min, d_max, subpix, is_max, self._refinement_method_name, approximate)
        store_var(disp, "disparity_map", disp_t)
        store_var(disp, "validity_mask", mask_t, dtype="uint16")
        disp.attrs["refinement"] = self._refinement_method_name
        store_var(disp, "interpolated_coeff", itp)
        return disp

    def subpixel_refinement(self, cv, disp) -> None:
        """refinement.py:78-122: refines ``disparity_map`` / ``validity_mask`` in place, adds ``interpolated_coeff``."""
        self._run(cv, disp, approximate=False)

    def approximate_subpixel_refinement(self, cv_left, disp_right):
        """refinement.py:124-181: right disparities refined along the diagonal of the LEFT volume."""
        return self._run(cv_left, disp_right, approximate=True)

    def right_subpixel_refinement(self, cv_left, disp_right):
        """``subpixel_refinement(right_cv, right_disparity)`` of the cross_checking_fast mode (state_machine.py:488-490)
        where right_cv = reverse_cost_volume(left_cv): the reversed cells are read from the LEFT volume."""
        return self._run(cv_left, disp_right, approximate=2)


@AbstractRefinement.register_subclass("vfit")
class Vfit(AbstractRefinement):
    def desc(self) -> None:
        print("Vfit refinement method")


@AbstractRefinement.register_subclass("quadratic")
class Quadratic(AbstractRefinement):
    def desc(self) -> None:
        print("Quadratic refinement method")
