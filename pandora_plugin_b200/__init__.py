"""pandora_plugin_b200 -- registers the B200 step implementations with CNES/Pandora's own factories.

Pandora loads every entry point of the group ``pandora.plugin`` once at start-up (``import_plugin``,
src/pandora/__init__.py:141-148; recipe docs/source/developer_guide/your_plugin.rst:70-89); loading this module runs
``register()`` below, after which a user configuration selects the steps by name::

    "matching_cost": {"matching_cost_method": "census_b200", "window_size": 5, "subpix": 1},
    "optimization":  {"optimization_method": "sgm_b200", "penalty": {"P1": 8, "P2": 32}},
    "disparity":     {"disparity_method": "wta_b200", "invalid_disparity": -9999}

Names: ``census_b200`` / ``sad_b200`` / ``ssd_b200`` / ``zncc_b200`` (matching cost), ``cbca_b200`` (aggregation),
``sgm_b200`` (optimisation), ``wta_b200`` (disparity).  The classes derive from Pandora's abstract step classes
(``AbstractMatchingCost.register_subclass`` matching_cost.py:109-131, ``AbstractAggregation`` aggregation.py:72-91,
``AbstractOptimization`` optimization.py:75-94, ``AbstractDisparity`` disparity.py:79-98) and delegate to the mirrors in
``pandora_b200``, which call the sm_100a kernels through ``libpandora_b200.so``.

Residency.  Pandora's datasets are xarray objects whose arrays live on the host.  A B200 step leaves its result in HBM
and records it in ``dataset.attrs["pb200_resident"]``; the next B200 step takes it from there, any other reader gets the
host copy: by default every step also copies its result back (``keep_host_copy = True``), so mixing B200 and stock steps is
always correct.  ``pandora_plugin_b200.keep_host_copy = False`` skips those copies for pipelines made of B200 steps only
(a 4096 x 4096 x 256 volume is 17 GB): the host array is then a zero-stride NaN placeholder of the right shape.

Importing this module never needs Pandora or a GPU: without Pandora ``register()`` returns False and nothing happens.
"""
from __future__ import annotations

import numpy as np

keep_host_copy = True
REGISTERED = {}


def _resident(ds, name="cost_volume"):
    """The device tensor a previous B200 step left for variable ``name`` of ``ds``, or None."""
    entry = getattr(ds, "attrs", {}).get("pb200_resident")
    return entry.get(name) if isinstance(entry, dict) else None


def _leave(ds, name, tensor, dims=("row", "col", "disp")):
    """Record ``tensor`` as the resident copy of ``ds[name]`` and give the host side what it needs."""
    ds.attrs.setdefault("pb200_resident", {})[name] = tensor
    ds.attrs["pb200_resident_trusted"] = not keep_host_copy     # with host copies around, the host array stays authoritative
    if keep_host_copy:
        host = tensor.detach().cpu().numpy()
    else:
        host = np.broadcast_to(np.float32(np.nan), tuple(int(s) for s in tensor.shape))     # zero-stride placeholder: no memory
    if name in ds:
        ds[name].data = host
    else:
        ds[name] = (tuple(dims), host)


def _volume(eng, cv):
    t = _resident(cv)
    return t if t is not None else eng.to_device(np.ascontiguousarray(cv["cost_volume"].data, dtype=np.float32))


def build_classes(mc_base, agg_base, opt_base, disp_base):
    """Create and register the step classes on the given abstract bases (Pandora's, or stand-ins with the same
    ``register_subclass`` decorators in the tests).  Returns {name: class}."""
    import pandora_b200 as pb
    from pandora_b200._common import get_engine

    out = {}

    def matching_cost(name, method, impl_cls):
        @mc_base.register_subclass(name)
        class _MatchingCostB200(mc_base):
            _impl_cls, _method = impl_cls, method

            def __init__(self, **cfg):
                self._impl = self._impl_cls(**{**cfg, "matching_cost_method": self._method})
                self.cfg = dict(self._impl.cfg, matching_cost_method=name)
                self._window_size, self._subpix, self._band = self._impl._window_size, self._impl._subpix, self._impl._band
                self._step_col, self._method_name = self._impl._step_col, name

            def compute_cost_volume(self, img_left, img_right, cost_volume):       # matching_cost.py:233-267
                self._impl.compute_cost_volume(img_left, img_right, cost_volume)
                lazy = getattr(cost_volume["cost_volume"], "_data", None)
                if getattr(lazy, "deferred", False):
                    return cost_volume                    # Census left a recipe: a following sgm_b200 fuses it away
                tensor = lazy.tensor if hasattr(lazy, "tensor") else get_engine().to_device(cost_volume["cost_volume"].data)
                _leave(cost_volume, "cost_volume", tensor)
                return cost_volume

            def cv_masked(self, img_left, img_right, cost_volume, disp_min, disp_max):   # matching_cost.py:770-872, on the device
                self._impl.cv_masked(img_left, img_right, cost_volume, disp_min, disp_max)

        _MatchingCostB200.__name__ = f"{method.capitalize()}B200"
        out[name] = _MatchingCostB200

    matching_cost("census_b200", "census", pb.Census)
    matching_cost("sad_b200", "sad", pb.SadSsd)
    matching_cost("ssd_b200", "ssd", pb.SadSsd)
    matching_cost("zncc_b200", "zncc", pb.Zncc)

    @agg_base.register_subclass("cbca_b200")
    class CbcaB200(agg_base):
        def __init__(self, **cfg):
            self._impl = pb.CrossBasedCostAggregation(**{**cfg, "aggregation_method": "cbca"})
            self.cfg = dict(self._impl.cfg, aggregation_method="cbca_b200")

        def cost_volume_aggregation(self, img_left, img_right, cv, **cfg):          # aggregation.py:101-133, in place
            self._impl.cost_volume_aggregation(img_left, img_right, cv, **cfg)
            lazy = getattr(cv["cost_volume"], "_data", None)
            if hasattr(lazy, "tensor"):
                _leave(cv, "cost_volume", lazy.tensor)

    out["cbca_b200"] = CbcaB200

    @opt_base.register_subclass("sgm_b200")
    class SgmB200(opt_base):
        def __init__(self, _img=None, **cfg):
            self._impl = pb.Sgm(None, **{**cfg, "optimization_method": "sgm"})
            self.cfg = dict(self._impl.cfg, optimization_method="sgm_b200")        # read by state_machine.py:872-873

        def optimize_cv(self, cv, img_left, img_right):                             # optimization.py:104-123
            out_cv = self._impl.optimize_cv(cv, img_left, img_right)
            lazy = getattr(out_cv["cost_volume"], "_data", None)
            if hasattr(lazy, "tensor"):
                _leave(out_cv, "cost_volume", lazy.tensor)
            return out_cv

    out["sgm_b200"] = SgmB200

    @disp_base.register_subclass("wta_b200")
    class WtaB200(disp_base):
        def __init__(self, **cfg):
            self._impl = pb.WinnerTakesAll(**{**cfg, "disparity_method": "wta"})
            self.cfg = dict(self._impl.cfg, disparity_method="wta_b200")
            self._invalid_disparity = self._impl._invalid_disparity

        def to_disp(self, cv, img_left=None, img_right=None):                       # disparity.py:400-480
            return self._impl.to_disp(cv, img_left, img_right)

    out["wta_b200"] = WtaB200
    return out


def register() -> bool:
    """Register with Pandora when it is importable; False (and no side effect) otherwise."""
    try:
        from pandora.aggregation import aggregation
        from pandora.disparity import disparity
        from pandora.matching_cost import matching_cost
        from pandora.optimization import optimization
    except ImportError:
        return False
    REGISTERED.update(build_classes(matching_cost.AbstractMatchingCost, aggregation.AbstractAggregation,
                                    optimization.AbstractOptimization, disparity.AbstractDisparity))
    return True


register()
