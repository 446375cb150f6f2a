"""filter step: mirror of AbstractFilter / MedianFilter (src/pandora/filter/filter.py:38-140, filter/median.py:38-179)
for the 3x3 median the sample pipelines use; the NaN-aware median itself is the kernel CBCA already runs on the images."""
from __future__ import annotations

from typing import Dict

import numpy as np

from ._common import ConfigError, device_var, get_engine, store_var


class AbstractFilter:
    filter_methods_avail: Dict[str, type] = {}

    def __new__(cls, cfg=None, **kwargs):
        if cls is AbstractFilter:
            method = (cfg or {}).get("filter_method")
            try:
                return super().__new__(cls.filter_methods_avail[method])
            except (KeyError, TypeError):
                raise KeyError(f"No filter method named {method} supported") from None
        return super().__new__(cls)

    @classmethod
    def register_subclass(cls, short_name: str):
        def decorator(subclass):
            cls.filter_methods_avail[short_name] = subclass
            return subclass

        return decorator


@AbstractFilter.register_subclass("median")
class MedianFilter(AbstractFilter):
    _FILTER_SIZE = 3

    def __init__(self, cfg=None, **kwargs):
        self.cfg = self.check_conf(dict(cfg or {}))
        self._filter_size = int(self.cfg["filter_size"])

    def check_conf(self, cfg: dict) -> dict:
        cfg.setdefault("filter_size", self._FILTER_SIZE)
        for key in cfg:
            if key not in ("filter_method", "filter_size"):
                raise ConfigError(f"Unknown key {key!r} in the filter configuration")
        size = cfg["filter_size"]
        if not (isinstance(size, int) and size >= 1 and size % 2 != 0):
            raise ConfigError("filter_size must be an odd int >= 1")
        if size != 3:
            raise ConfigError("filter_size: only 3 is implemented by the B200 kernels")
        return cfg

    def desc(self) -> None:
        print("Median filter description")

    def filter_disparity(self, disp, img_left=None, img_right=None, cv=None) -> None:
        """filter/median.py:96-132: median of the valid pixels, invalid pixels untouched, in place."""
        eng = get_engine()
        d = device_var(eng, disp, "disparity_map").clone()
        m = device_var(eng, disp, "validity_mask", "uint16")
        eng.filter_median3(d, m)
        store_var(disp, "disparity_map", d)
        disp.attrs["filter"] = "median"
