"""cost_volume_confidence step: mirror of AbstractCostVolumeConfidence / Ambiguity / Risk
(src/pandora/cost_volume_confidence/cost_volume_confidence.py:38-250, ambiguity.py:36-248, risk.py:36-233) with the
C++ loops of cost_volume_confidence/cpp/src/{ambiguity,risk}.cpp fused into one device pass."""
from __future__ import annotations

from typing import Dict

import numpy as np

from ._common import ConfigError, device_volume, get_engine
from .dataset import DataArray


class AbstractCostVolumeConfidence:
    confidence_methods_avail: Dict[str, type] = {}
    _indicator = ""

    def __new__(cls, **cfg):
        if cls is AbstractCostVolumeConfidence:
            method = cfg.get("confidence_method")
            try:
                return super().__new__(cls.confidence_methods_avail[method])
            except (KeyError, TypeError):
                raise KeyError(f"No confidence method named {method} supported") from None
        return super().__new__(cls)

    @classmethod
    def register_subclass(cls, short_name: str):
        def decorator(subclass):
            cls.confidence_methods_avail[short_name] = subclass
            return subclass

        return decorator

    @staticmethod
    def normalize_with_extremum(confidence: np.ndarray, dataset, nbr_etas: int, subpix: int = 1) -> np.ndarray:
        """cost_volume_confidence.py:115-138: confidence / ((global_disp_max - global_disp_min) * nbr_etas * subpix)."""
        global_disp_min, global_disp_max = dataset.attrs["global_disparity"][0], dataset.attrs["global_disparity"][1]
        return np.copy(confidence) / ((global_disp_max - global_disp_min) * nbr_etas * subpix)

    @staticmethod
    def allocate_confidence_map(name_confidence_measure: str, confidence_map: np.ndarray, disp=None, cv=None):
        """cost_volume_confidence.py:129-250: append one (row, col) indicator to ``confidence_measure`` of both datasets."""
        if "disp_min" not in name_confidence_measure and "disp_max" not in name_confidence_measure:
            name_confidence_measure = "confidence_from_" + name_confidence_measure
        layer = np.asarray(confidence_map, dtype=np.float32)[:, :, np.newaxis]

        def append(ds):
            if "confidence_measure" in ds.data_vars:
                old = np.asarray(ds["confidence_measure"].data, dtype=np.float32)
                ds["confidence_measure"] = (("row", "col", "indicator"), np.concatenate([old, layer], axis=2))
                ind = list(np.asarray(ds.coords["indicator"].data)) + [name_confidence_measure]
            else:
                ds["confidence_measure"] = (("row", "col", "indicator"), layer.copy())
                ind = [name_confidence_measure]
            ds.coords["indicator"] = DataArray(np.array(ind), ("indicator",))

        if cv is not None:
            append(cv)
        if disp is not None:
            if "confidence_measure" not in disp.data_vars and cv is not None:
                disp["confidence_measure"] = cv["confidence_measure"]
                disp.coords["indicator"] = cv.coords["indicator"]
            else:
                append(disp)
        return disp, cv

    # -- shared configuration of ambiguity.py:70-98 / risk.py:76-103 --
    _ETA_MIN, _ETA_MAX, _ETA_STEP, _PERCENTILE = 0.0, 0.7, 0.01, 1.0

    def _check_etas(self, cfg: dict, extra=()) -> dict:
        cfg.setdefault("eta_max", self._ETA_MAX)
        cfg.setdefault("eta_step", self._ETA_STEP)
        cfg.setdefault("indicator", self._indicator)
        for key in cfg:
            if key not in ("confidence_method", "eta_max", "eta_step", "indicator", *extra):
                raise ConfigError(f"Unknown key {key!r} in the confidence configuration")
        for key in ("eta_max", "eta_step"):
            if not (isinstance(cfg[key], float) and 0 < cfg[key] < 1):
                raise ConfigError(f"{key} must be a float in ]0, 1[")
        if not isinstance(cfg["indicator"], str):
            raise ConfigError("indicator must be a str")
        return cfg

    @staticmethod
    def _inputs(img_left, cv):
        grids = np.array([img_left["disparity"].data[0], img_left["disparity"].data[1]], dtype=np.int64)
        return grids, np.asarray(cv.coords["disp"].data).astype(np.float32), cv.attrs["type_measure"] == "max"


@AbstractCostVolumeConfidence.register_subclass("ambiguity")
class Ambiguity(AbstractCostVolumeConfidence):
    _NORMALIZATION = True
    _method = "ambiguity"

    def __init__(self, **cfg):
        self.cfg = self.check_conf(**cfg)
        self._normalization = self.cfg["normalization"]
        self._eta_max, self._eta_step = float(self.cfg["eta_max"]), float(self.cfg["eta_step"])
        self._indicator = self._method + str(self.cfg["indicator"])
        self._etas = np.arange(self._ETA_MIN, self._eta_max, self._eta_step)
        self._nbr_etas = self._etas.shape[0]
        self._percentile = self._PERCENTILE

    def check_conf(self, **cfg) -> dict:
        cfg.setdefault("normalization", self._NORMALIZATION)
        cfg = self._check_etas(cfg, extra=("normalization",))
        if not isinstance(cfg["normalization"], bool):
            raise ConfigError("normalization must be a bool")
        return cfg

    def desc(self) -> None:
        print("Ambiguity confidence method")

    @staticmethod
    def compute_ambiguity(cv, etas, nbr_etas, grids, disparity_range) -> np.ndarray:
        """ambiguity.py:188-216 on a host or device min-type volume."""
        eng = get_engine()
        cv_t = cv if not isinstance(cv, np.ndarray) else eng.to_device(cv)
        return eng.confidence(cv_t, etas[:nbr_etas], grids, disparity_range)["ambiguity"].cpu().numpy()

    @staticmethod
    def compute_ambiguity_and_sampled_ambiguity(cv, etas, nbr_etas, grids, disparity_range):
        eng = get_engine()
        cv_t = cv if not isinstance(cv, np.ndarray) else eng.to_device(cv)
        out = eng.confidence(cv_t, etas[:nbr_etas], grids, disparity_range, sampled_ambiguity=True)
        return out["ambiguity"].cpu().numpy(), out["sampled_ambiguity"].cpu().numpy()

    def normalize_with_percentile(self, ambiguity: np.ndarray) -> np.ndarray:
        """ambiguity.py:172-186."""
        norm_amb = np.copy(ambiguity)
        perc_min = np.percentile(norm_amb, self._percentile)
        perc_max = np.percentile(norm_amb, 100 - self._percentile)
        np.clip(norm_amb, perc_min, perc_max, out=norm_amb)
        return (norm_amb - np.min(norm_amb)) / (np.max(norm_amb) - np.min(norm_amb))

    def confidence_prediction(self, disp, img_left=None, img_right=None, cv=None):
        """ambiguity.py:107-170 (the max-type negation happens on the fly inside the kernel, the volume is untouched)."""
        eng = get_engine()
        grids, disparity_range, is_max = self._inputs(img_left, cv)
        amb = eng.confidence(device_volume(eng, cv), self._etas, grids, disparity_range, is_max=is_max)["ambiguity"].cpu().numpy()
        if self._normalization:
            subpix = int(cv.attrs.get("subpixel", 1))
            if "global_disparity" in img_left.attrs:                                   # ambiguity.py:148-162
                amb = self.normalize_with_extremum(amb, img_left, self._nbr_etas, subpix)
            elif img_right is not None and "global_disparity" in img_right.attrs:
                amb = self.normalize_with_extremum(amb, img_right, self._nbr_etas, subpix)
            else:
                amb = self.normalize_with_percentile(amb)
        return self.allocate_confidence_map(self._indicator, 1 - amb, disp, cv)


@AbstractCostVolumeConfidence.register_subclass("risk")
class Risk(AbstractCostVolumeConfidence):
    _method_max, _method_min = "risk_max", "risk_min"
    _method_disp_inf, _method_disp_sup = "disp_inf_from_risk", "disp_sup_from_risk"

    def __init__(self, **cfg):
        self.cfg = self._check_etas(dict(cfg))
        self._eta_max, self._eta_step = float(self.cfg["eta_max"]), float(self.cfg["eta_step"])
        ind = str(self.cfg["indicator"])
        self._indicator_max, self._indicator_min = self._method_max + ind, self._method_min + ind
        self._indicator_disp_sup, self._indicator_disp_inf = self._method_disp_sup + ind, self._method_disp_inf + ind
        self._etas = np.arange(self._ETA_MIN, self._eta_max, self._eta_step)
        self._nbr_etas = self._etas.shape[0]

    def desc(self) -> None:
        print("Risk method")

    @staticmethod
    def compute_risk(cv, sampled_ambiguity, etas, nbr_etas, grids, disparity_range):
        """risk.py:165-197: (risk_max, risk_min, disp_sup, disp_inf)."""
        eng = get_engine()
        cv_t = cv if not isinstance(cv, np.ndarray) else eng.to_device(cv)
        out = eng.confidence(cv_t, etas[:nbr_etas], grids, disparity_range, ambiguity=False, risk=True, sampled_ambiguity_in=sampled_ambiguity)
        return tuple(out[k].cpu().numpy() for k in ("risk_max", "risk_min", "disp_sup", "disp_inf"))

    @staticmethod
    def compute_risk_and_sampled_risk(cv, sampled_ambiguity, etas, nbr_etas, grids, disparity_range):
        eng = get_engine()
        cv_t = cv if not isinstance(cv, np.ndarray) else eng.to_device(cv)
        out = eng.confidence(cv_t, etas[:nbr_etas], grids, disparity_range, ambiguity=False, risk=True, sampled_risk=True,
                             sampled_ambiguity_in=sampled_ambiguity)
        return tuple(out[k].cpu().numpy() for k in ("risk_max", "risk_min", "disp_sup", "disp_inf", "sampled_risk_max", "sampled_risk_min"))

    def confidence_prediction(self, disp, img_left=None, img_right=None, cv=None):
        """risk.py:110-163: sampled ambiguity and risk from one pass over the volume."""
        eng = get_engine()
        grids, disparity_range, is_max = self._inputs(img_left, cv)
        out = eng.confidence(device_volume(eng, cv), self._etas, grids, disparity_range, is_max=is_max, ambiguity=False, risk=True)
        for name, key in ((self._indicator_max, "risk_max"), (self._indicator_min, "risk_min"),
                          (self._indicator_disp_sup, "disp_sup"), (self._indicator_disp_inf, "disp_inf")):
            disp, cv = self.allocate_confidence_map(name, out[key].cpu().numpy(), disp, cv)
        return disp, cv
