"""optimization step: mirror of AbstractOptimization (src/pandora/optimization/optimization.py:34-123)
plus the SGM implementation the reference gets from the external libSGM plugin
(docs/source/userguide/plugins/plugin_libsgm.rst; un-vendored, so numeric parity is pinned against
oracle/pandora_oracle.c only -- "parity unpinned" against libSGM itself)."""
from __future__ import annotations

from typing import Dict

import numpy as np

from ._common import ConfigError, deferred_recipe, device_volume, get_engine, store_volume
from ._native import SGM_MAX_DISP


class UniformMargins:
    """margins.UniformMargins(40) of the reference (optimization.py:43, marge.py:85-101) as a plain tuple holder."""

    def __init__(self, value: int):
        self.left = self.up = self.right = self.down = value

    def astuple(self):
        return (self.left, self.up, self.right, self.down)


class AbstractOptimization:
    optimization_methods_avail: Dict[str, type] = {}
    margins = UniformMargins(40)

    def __new__(cls, _img=None, **cfg):
        if cls is AbstractOptimization:
            method = cfg.get("optimization_method")
            try:
                return super().__new__(cls.optimization_methods_avail[method])
            except (KeyError, TypeError):
                raise KeyError(f"No optimization method named {method} supported") from None
        return super().__new__(cls)

    @classmethod
    def register_subclass(cls, short_name: str):
        def decorator(subclass):
            cls.optimization_methods_avail[short_name] = subclass
            return subclass

        return decorator

    def desc(self):
        print("Optimization method description")

    def optimize_cv(self, cv, img_left, img_right):
        raise NotImplementedError


@AbstractOptimization.register_subclass("sgm")
class Sgm(AbstractOptimization):
    """8-path SGM with constant penalties (plugin_libsgm.rst:88-211: P1 = 8, P2 = 32 defaults for Census), optionally on
    confidence-weighted costs (``use_confidence``, plugin_libsgm.rst:38-47)."""

    _P1, _P2 = 8, 32

    def __init__(self, _img=None, **cfg):
        self.cfg = self.check_conf(**cfg)
        pen = self.cfg["penalty"]
        self._p1, self._p2 = float(pen["P1"]), float(pen["P2"])
        self._overcounting = bool(self.cfg["overcounting"])

    def check_conf(self, **cfg) -> dict:
        cfg.setdefault("overcounting", False)
        cfg.setdefault("min_cost_paths", False)
        cfg.setdefault("use_confidence", None)
        pen = dict(cfg.get("penalty") or {})
        pen.setdefault("penalty_method", "sgm_penalty")
        pen.setdefault("p2_method", "constant")
        pen.setdefault("P1", self._P1)
        pen.setdefault("P2", self._P2)
        cfg["penalty"] = pen
        for key in cfg:
            if key not in ("optimization_method", "overcounting", "min_cost_paths", "use_confidence", "penalty", "geometric_prior"):
                raise ConfigError(f"Unknown key {key!r} in the optimization configuration")
        if pen["penalty_method"] != "sgm_penalty" or pen["p2_method"] != "constant":
            raise ConfigError("penalty: only penalty_method='sgm_penalty' with p2_method='constant' is implemented")
        if not pen["P1"] > 0 or not pen["P2"] > pen["P1"]:
            raise ConfigError("penalty: P1 > 0 and P2 > P1 are required")
        if not isinstance(cfg["min_cost_paths"], bool):
            raise ConfigError("min_cost_paths must be a bool")
        if cfg["use_confidence"] is not None and not isinstance(cfg["use_confidence"], str):
            raise ConfigError("use_confidence must be the name of a cost_volume_confidence step (a str)")
        return cfg

    def desc(self):
        print("Semi-global matching optimization (B200)")

    def _confidence_map(self, cv):
        """``use_confidence`` = name of a previous cost_volume_confidence step ("cost_volume_confidence[.suffix]"): its ambiguity
        confidence band ``confidence_from_ambiguity[.suffix]`` (state_machine.py:566-576 builds the suffix), or None when that band
        does not exist -- "default confidence values equal to 1 will be used" (plugin_libsgm.rst:47)."""
        name = self.cfg["use_confidence"]
        if not name or "confidence_measure" not in getattr(cv, "data_vars", {}):
            return None
        parts = str(name).split(".")
        band = "confidence_from_ambiguity" + ("." + parts[1] if len(parts) == 2 else "")
        indicators = [str(v) for v in np.asarray(cv.coords["indicator"].data)]
        if band not in indicators:
            return None
        return np.ascontiguousarray(np.asarray(cv["confidence_measure"].data)[:, :, indicators.index(band)], dtype=np.float32)

    NB_OF_DIRECTIONS = "optimization_plugin_libsgm_nb_of_directions"      # band name: docs/source/userguide/output.rst:22

    @classmethod
    def _append_band(cls, cv, layer: np.ndarray) -> None:
        """The plugin's own confidence band, appended without the "confidence_from_" prefix of allocate_confidence_map."""
        from .dataset import DataArray  # noqa: PLC0415

        layer = np.asarray(layer, dtype=np.float32)[:, :, np.newaxis]
        if "confidence_measure" in getattr(cv, "data_vars", {}):
            old = np.asarray(cv["confidence_measure"].data, dtype=np.float32)
            ind = [str(v) for v in np.asarray(cv.coords["indicator"].data)] + [cls.NB_OF_DIRECTIONS]
            cv["confidence_measure"] = (("row", "col", "indicator"), np.concatenate([old, layer], axis=2))
        else:
            ind = [cls.NB_OF_DIRECTIONS]
            cv["confidence_measure"] = (("row", "col", "indicator"), layer.copy())
        cv.coords["indicator"] = DataArray(np.array(ind), ("indicator",))

    def optimize_cv(self, cv, img_left, img_right):
        n_disp = len(np.asarray(cv.coords["disp"].data))
        if n_disp > SGM_MAX_DISP:                      # before any kernel runs (PB200_SGM_MAX_DISP, include/pandora_b200.h)
            raise ConfigError(f"sgm: {n_disp} disparities exceed the kernels' maximum of {SGM_MAX_DISP}")
        eng = get_engine()
        cmax = float(cv.attrs["cmax"])
        is_max = cv.attrs.get("type_measure") == "max"
        recipe = deferred_recipe(cv)
        confidence = self._confidence_map(cv)
        if confidence is None and not self.cfg["min_cost_paths"] and recipe is not None and getattr(recipe, "kind", None) == "census" and not is_max and cmax == float(recipe.window**2):
            # the Census volume was never computed: fused Census -> SGM (same bits, no float Census volume)
            fused = eng.census_sgm(recipe.left, recipe.right, recipe.window, recipe.dmin, recipe.dmax, self._p1, self._p2,
                                   self._overcounting, fuse_wta=True, invalid_disparity=-9999.0)
            if fused is not None:
                store_volume(cv, fused[0])
                # the last pass ran the argmin on the fly: a `disparity` step that follows directly takes it from here
                cv["cost_volume"]._data.wta_cache = (fused[1], fused[2], recipe.dmin, -9999.0)
                cv.attrs["optimization"] = "sgm"
                return cv
        cv_t = device_volume(eng, cv)
        if confidence is not None:
            # E(D) = sum_p C(p, D_p) * Confidence(p) + ... (plugin_libsgm.rst:38-47); confidences lie in [0, 1], so cmax still bounds the costs
            cv_t = eng.scale_volume(cv_t, eng.to_device(confidence))
        src = -cv_t if is_max else cv_t
        if self.cfg["min_cost_paths"]:
            # one launch per direction, each recording where its own path cost is minimal (plugin_libsgm.rst:411-413)
            out, nb = eng.sgm_min_cost_paths(src, self._p1, self._p2, cmax + self._p2 + 1.0, self._overcounting)
            self._append_band(cv, nb.cpu().numpy())
        else:
            out = eng.sgm(src, self._p1, self._p2, cmax + self._p2 + 1.0, self._overcounting)
        if is_max:
            out = -out
        store_volume(cv, out)
        cv.attrs["optimization"] = "sgm"
        return cv
