"""aggregation step: mirror of AbstractAggregation / CrossBasedCostAggregation
(src/pandora/aggregation/aggregation.py:34-133, aggregation/cbca.py:39-298)."""
from __future__ import annotations

from typing import Dict

from ._common import ConfigError, device_volume, get_engine, image_array, store_volume


class AbstractAggregation:
    aggregation_methods_avail: Dict[str, type] = {}

    def __new__(cls, **cfg):
        if cls is AbstractAggregation:
            method = cfg.get("aggregation_method")
            try:
                return super().__new__(cls.aggregation_methods_avail[method])
            except (KeyError, TypeError):
                raise KeyError(f"No aggregation method named {method} supported") from None
        return super().__new__(cls)

    @classmethod
    def register_subclass(cls, short_name: str):
        def decorator(subclass):
            cls.aggregation_methods_avail[short_name] = subclass
            return subclass

        return decorator

    def desc(self):
        print("Aggregation method description")

    def cost_volume_aggregation(self, img_left, img_right, cv, **cfg) -> None:
        raise NotImplementedError


@AbstractAggregation.register_subclass("cbca")
class CrossBasedCostAggregation(AbstractAggregation):
    """Cross-based cost aggregation (Zhang 2009), reference defaults intensity 30.0 / distance 5 (cbca.py:46-47)."""

    _CBCA_INTENSITY = 30.0
    _CBCA_DISTANCE = 5

    def __init__(self, **cfg):
        self.cfg = self.check_conf(**cfg)
        self._cbca_intensity = self.cfg["cbca_intensity"]
        self._cbca_distance = self.cfg["cbca_distance"]

    def check_conf(self, **cfg) -> dict:
        cfg.setdefault("cbca_intensity", self._CBCA_INTENSITY)
        cfg.setdefault("cbca_distance", self._CBCA_DISTANCE)
        for key in cfg:
            if key not in ("aggregation_method", "cbca_intensity", "cbca_distance"):
                raise ConfigError(f"Unknown key {key!r} in the aggregation configuration")
        if not isinstance(cfg["cbca_intensity"], float) or not cfg["cbca_intensity"] > 0:
            raise ConfigError("cbca_intensity must be a float > 0")
        if not isinstance(cfg["cbca_distance"], int) or isinstance(cfg["cbca_distance"], bool) or not cfg["cbca_distance"] > 0:
            raise ConfigError("cbca_distance must be an int > 0")
        return cfg

    def desc(self):
        print("CrossBasedCostAggregation method")

    def cost_volume_aggregation(self, img_left, img_right, cv, **cfg) -> None:
        """In place on ``cv`` like cbca.py:90-182: aggregated volume, ``aggregation`` attr, rescaled ``cmax``."""
        if int(cv.attrs.get("subpixel", 1)) != 1:
            raise NotImplementedError("CBCA with subpix > 1 is not implemented by the B200 kernels")
        if "msk" in getattr(img_left, "data_vars", {}) or "msk" in getattr(img_right, "data_vars", {}):
            raise NotImplementedError("input masks are not on the B200 hot path yet (SURVEY.md 8f rank 1)")
        eng = get_engine()
        cv_t = device_volume(eng, cv)
        offset = int(cv.attrs["offset_row_col"])
        dmin = int(round(float(cv.coords["disp"].data[0])))
        left = eng.to_device(image_array(img_left))
        right = eng.to_device(image_array(img_right))
        out = eng.cbca(left, right, cv_t, offset, dmin, self._cbca_distance, self._cbca_intensity)
        store_volume(cv, out)
        cv.attrs["aggregation"] = "cbca"
        cv.attrs["cmax"] = cv.attrs["cmax"] * ((self._cbca_distance * 2) - 1) ** 2
