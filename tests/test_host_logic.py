"""CPU tests: host-side mirror of the reference plugin API, the dataset shim, and the C-ABI library's
symbol table (no compute call is made without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

import pandora_b200 as pb
from pandora_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "pandora_b200.h")).read()
    declared = set(re.findall(r"PB200_API[^;(]*?\b(pb200_\w+)\s*\(", header))
    assert len(declared) >= 20
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/pandora_b200.h but not exported"
    assert declared == set(_native.PROTOTYPES), declared ^ set(_native.PROTOTYPES)


def test_library_loads_and_reports_no_device_without_gpu():
    lib = _native.load()
    assert lib.pb200_version() >= 100
    assert lib.pb200_device_count() >= 0
    assert isinstance(lib.pb200_last_error(), bytes)


def test_no_cpu_fallback_when_cuda_is_missing():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pb.get_engine()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mc = pb.AbstractMatchingCost(matching_cost_method="census", window_size=5)
        ds = pb.create_image_dataset(np.zeros((8, 8), np.float32), disparity=[-1, 1])
        cv = mc.allocate_cost_volume(ds, (ds["disparity"].data[0], ds["disparity"].data[1]))
        mc.compute_cost_volume(ds, ds, cv)


def test_bad_arguments_are_rejected_before_any_cuda_call():
    lib = _native.load()
    assert lib.pb200_wta(None, 4, 4, 4, 0, 0, 0.0, None, None, None) == _native.ERR_BAD_ARG
    assert b"pb200_wta" in lib.pb200_last_error()
    assert lib.pb200_census_cost_volume(1, 1, 8, 8, 4, 0, 4, 1, 1, 1 << 20, None, 0.0, None, None) == _native.ERR_UNSUPPORTED
    assert lib.pb200_sgm(1, 2, 4, 4, 1000, 8.0, 32.0, 58.0, 0, 0xFF, 3, None, None, None, None, None, 0, 0.0, None, None, 0, None) == _native.ERR_UNSUPPORTED
    assert lib.pb200_census_workspace_bytes(10, 10, 5) == 2 * 10 * 16 * 4


# ---- plugin registries / factories (reference: matching_cost.py:80-131 etc.) -----------------------
def test_registries_and_factories():
    assert set(pb.AbstractMatchingCost.matching_cost_methods_avail) >= {"census", "sad", "ssd", "zncc"}
    assert isinstance(pb.AbstractMatchingCost(matching_cost_method="ssd", window_size=3), pb.SadSsd)
    assert isinstance(pb.AbstractAggregation(aggregation_method="cbca"), pb.CrossBasedCostAggregation)
    assert isinstance(pb.AbstractOptimization(None, optimization_method="sgm"), pb.Sgm)
    assert isinstance(pb.AbstractDisparity(disparity_method="wta"), pb.WinnerTakesAll)
    for factory, key in [(pb.AbstractMatchingCost, "matching_cost_method"), (pb.AbstractAggregation, "aggregation_method"),
                         (pb.AbstractDisparity, "disparity_method")]:
        with pytest.raises(KeyError, match="supported"):
            factory(**{key: "nope"})
    with pytest.raises(KeyError, match="No optimization method named sgmm supported"):     # tests/test_plugins.py:96-123
        pb.AbstractOptimization(None, optimization_method="sgmm")

    @pb.AbstractMatchingCost.register_subclass("my_cost", "alias_cost")
    class MyCost(pb.AbstractMatchingCost):
        pass

    assert isinstance(pb.AbstractMatchingCost(matching_cost_method="alias_cost"), MyCost)
    del pb.AbstractMatchingCost.matching_cost_methods_avail["my_cost"], pb.AbstractMatchingCost.matching_cost_methods_avail["alias_cost"]


@pytest.mark.parametrize("window_size", [3, 5, 7, 9, 11, 13])
def test_census_nominal_window_size(window_size):                 # test_matching_cost_census.py:42-46
    assert pb.AbstractMatchingCost(matching_cost_method="census", window_size=window_size).cfg["window_size"] == window_size


@pytest.mark.parametrize("window_size", [-5, -1, 0, 1, 2, 4, 6, 8, 14, 15])
def test_census_invalid_window_size(window_size):                 # test_matching_cost_census.py:48-52
    with pytest.raises(pb.ConfigError) as err:
        pb.AbstractMatchingCost(matching_cost_method="census", window_size=window_size)
    assert "window_size" in err.value.args[0]


def test_matching_cost_defaults_and_step():
    mc = pb.AbstractMatchingCost(matching_cost_method="zncc")
    assert mc.cfg == {"matching_cost_method": "zncc", "window_size": 5, "subpix": 1, "band": None, "step": 1}
    with pytest.raises(ValueError, match="Step parameter cannot be different from 1"):     # matching_cost.py:176-178
        pb.AbstractMatchingCost(matching_cost_method="sad", step=2)
    with pytest.raises(pb.ConfigError):
        pb.AbstractMatchingCost(matching_cost_method="sad", subpix=3)


def test_cbca_sgm_wta_config():
    agg = pb.AbstractAggregation(aggregation_method="cbca")
    assert agg.cfg == {"aggregation_method": "cbca", "cbca_intensity": 30.0, "cbca_distance": 5}        # cbca.py:46-47
    with pytest.raises(pb.ConfigError):
        pb.AbstractAggregation(aggregation_method="cbca", cbca_distance=0)
    with pytest.raises(pb.ConfigError):
        pb.AbstractAggregation(aggregation_method="cbca", cbca_intensity=-1.0)
    opt = pb.AbstractOptimization(None, optimization_method="sgm")
    assert opt.cfg["penalty"]["P1"] == 8 and opt.cfg["penalty"]["P2"] == 32                              # plugin_libsgm.rst:198-211
    assert pb.AbstractOptimization.margins.astuple() == (40, 40, 40, 40)                                  # tests/test_optimization.py:28-30
    with pytest.raises(pb.ConfigError):
        pb.AbstractOptimization(None, optimization_method="sgm", penalty={"P1": 8, "P2": 4})
    wta = pb.AbstractDisparity(disparity_method="wta")
    assert wta.cfg["invalid_disparity"] == -9999                                                         # disparity.py:357
    assert np.isnan(pb.AbstractDisparity(disparity_method="wta", invalid_disparity="NaN").cfg["invalid_disparity"])


def test_allocate_cost_volume_container():
    """matching_cost.py:377-407: coordinates and attributes of the (empty) cost-volume dataset."""
    img = pb.create_image_dataset(np.zeros((6, 9), np.float32), disparity=[-3, 2], row0=10, col0=20)
    mc = pb.AbstractMatchingCost(matching_cost_method="census", window_size=5)
    cv = mc.allocate_cost_volume(img, (img["disparity"].data[0], img["disparity"].data[1]))
    np.testing.assert_array_equal(cv.coords["disp"].data, np.arange(-3, 3))
    np.testing.assert_array_equal(cv.coords["row"].data, np.arange(10, 16))
    np.testing.assert_array_equal(cv.coords["col"].data, np.arange(20, 29))
    for k, v in {"window_size": 5, "subpixel": 1, "band_correl": None, "offset_row_col": 2, "measure": "census",
                 "sampling_interval": 1}.items():
        assert cv.attrs[k] == v
    assert cv.sizes == {"row": 6, "col": 9, "disp": 6}


def test_dataset_shim_lazy_volume():
    class FakeTensor:
        shape = (2, 3, 4)
        calls = 0

        def detach(self):
            return self

        def cpu(self):
            FakeTensor.calls += 1
            return self

        def numpy(self):
            return np.ones((2, 3, 4), np.float32)

    ds = pb.Dataset(coords={"row": np.arange(2), "col": np.arange(3), "disp": np.arange(4)})
    ds["cost_volume"] = (("row", "col", "disp"), pb.LazyVolume(FakeTensor()))
    assert ds["cost_volume"].device_tensor() is not None and FakeTensor.calls == 0
    assert ds["cost_volume"].data.sum() == 24 and FakeTensor.calls == 1
    assert ds["cost_volume"].device_tensor() is None            # host copy is authoritative once handed out
    assert "cost_volume" in ds and "msk" not in ds.data_vars


def test_run_rejects_steps_outside_the_hot_path():
    img = pb.create_image_dataset(np.zeros((6, 9), np.float32), disparity=[-1, 1])
    with pytest.raises(NotImplementedError, match="outside the B200 hot path"):
        pb.run(img, img, {"pipeline": {"multiscale": {"multiscale_method": "fixed_zoom_pyramid"}}})
    with pytest.raises(NotImplementedError, match="interpolated_disparity"):
        pb.run(img, img, {"pipeline": {"validation": {"validation_method": "cross_checking_accurate", "interpolated_disparity": "sgm"}}})
    with pytest.raises(KeyError):
        pb.AbstractFilter({"filter_method": "bilateral_b200"})
    with pytest.raises(pb.ConfigError):
        pb.AbstractFilter({"filter_method": "median", "filter_size": 5})
    assert pb.AbstractFilter({"filter_method": "median"}).cfg["filter_size"] == 3


def test_next_row_registries_and_config_errors():
    """refinement / validation / cost_volume_confidence mirrors: same factories and error behaviour as the reference
    (refinement.py:52-74, validation.py:59-84, cost_volume_confidence.py:52-75)."""
    import pandora_b200 as pb

    assert isinstance(pb.AbstractRefinement(refinement_method="vfit"), pb.Vfit)
    assert isinstance(pb.AbstractRefinement(refinement_method="quadratic"), pb.Quadratic)
    with pytest.raises(KeyError):
        pb.AbstractRefinement(refinement_method="cubic")
    for name in ("cross_checking_fast", "cross_checking_accurate"):                    # tests/test_validation.py:92-103
        assert isinstance(pb.AbstractValidation(validation_method=name), pb.CrossCheckingAccurate)
    with pytest.raises(KeyError):
        pb.AbstractValidation(**{"something's wrong": "blue"})                         # tests/test_validation.py:78-83
    with pytest.raises(KeyError):
        pb.AbstractValidation(validation_method="hello")
    assert pb.AbstractValidation(validation_method="cross_checking_fast").cfg["cross_checking_threshold"] == 1.0
    amb = pb.AbstractCostVolumeConfidence(confidence_method="ambiguity")
    assert isinstance(amb, pb.Ambiguity) and amb._nbr_etas == 70 and amb.cfg["normalization"] is True
    risk = pb.AbstractCostVolumeConfidence(confidence_method="risk", eta_max=0.5, eta_step=0.3)
    assert isinstance(risk, pb.Risk) and list(risk._etas) == [0.0, 0.3]
    with pytest.raises(pb.ConfigError):
        pb.AbstractCostVolumeConfidence(confidence_method="ambiguity", eta_max=1.5)
    with pytest.raises(KeyError):
        pb.AbstractCostVolumeConfidence(confidence_method="variance")


def test_allocate_confidence_map_appends_indicators():
    import pandora_b200 as pb

    cv = pb.Dataset(coords={"row": np.arange(2), "col": np.arange(3)})
    disp = pb.Dataset(coords={"row": np.arange(2), "col": np.arange(3)})
    first, second = np.ones((2, 3), np.float32), np.full((2, 3), 2, np.float32)
    disp, cv = pb.AbstractCostVolumeConfidence.allocate_confidence_map("ambiguity", first, disp, cv)
    disp, cv = pb.AbstractCostVolumeConfidence.allocate_confidence_map("risk_max", second, disp, cv)
    for ds in (disp, cv):
        assert ds["confidence_measure"].data.shape == (2, 3, 2)
        assert list(ds.coords["indicator"].data) == ["confidence_from_ambiguity", "confidence_from_risk_max"]
        np.testing.assert_array_equal(ds["confidence_measure"].data[:, :, 1], second)


def test_normalize_with_extremum():                                # tests/test_confidence/test_ambiguity.py:204-231
    from pandora_b200.cost_volume_confidence import AbstractCostVolumeConfidence

    class _Img:
        attrs = {"disp_min": 0, "disp_max": 1, "global_disparity": [-2, 2]}

    ambiguity_ = AbstractCostVolumeConfidence(confidence_method="ambiguity", eta_max=0.2, eta_step=0.1)
    ambiguity = np.ones((4, 4))
    got = ambiguity_.normalize_with_extremum(ambiguity, _Img(), ambiguity_._nbr_etas)
    nbr_etas = np.arange(0.0, 0.2, 0.1).shape[0]
    np.testing.assert_array_equal(got, np.copy(ambiguity) / ((2 - (-2)) * nbr_etas))


@pytest.mark.parametrize("H,W,w,dmin,dmax", [(9, 30, 5, -7, 3), (7, 12, 3, -20, -4), (6, 15, 5, 4, 11), (5, 9, 7, -2, 2), (8, 40, 3, -63, 0),
                                             (4, 4, 5, -1, 1), (10, 11, 13, 0, 0)])
def test_census_recipe_all_nan_geometry(H, W, w, dmin, dmax):
    """CensusRecipe.all_nan_flags (what cv_masked uses instead of reading a deferred volume) against the all-NaN pixels
    of the oracle's Census volume: pure index arithmetic, checked on CPU tensors."""
    import torch

    from oracle import oracle as orc
    from pandora_b200.matching_cost import CensusRecipe

    g = np.random.default_rng(H * W + w)
    left = g.integers(0, 50, (H, W)).astype(np.float32)
    right = g.integers(0, 50, (H, W)).astype(np.float32)
    cv, _ = orc.census_cost_volume(left, right, w, dmin, dmax)
    rec = CensusRecipe(None, torch.from_numpy(left), torch.from_numpy(right), w, dmin, dmax)
    np.testing.assert_array_equal(rec.all_nan_flags().numpy().astype(bool), np.isnan(cv).all(axis=2))


def test_deferred_volume_protocol():
    """LazyVolume(recipe=...): nothing is computed until somebody asks for the tensor or ``.data``; afterwards the
    recipe is gone; a dataset variable that is not the shim's DataArray refuses the deferral."""
    import torch

    from pandora_b200._common import deferred_recipe, device_volume, store_deferred_volume
    from pandora_b200.dataset import Dataset

    calls = []

    class _Recipe:
        kind = "census"

        def compute(self):
            calls.append(1)
            return torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)

    cv = Dataset(coords={"row": np.arange(2), "col": np.arange(3), "disp": np.arange(4)})
    rec = _Recipe()
    assert store_deferred_volume(cv, rec, (2, 3, 4))
    assert cv["cost_volume"].shape == (2, 3, 4) and not calls
    assert deferred_recipe(cv) is rec and not calls                 # asking for the recipe computes nothing
    t = device_volume(None, cv)                                     # the first reader computes it, once
    assert calls == [1] and tuple(t.shape) == (2, 3, 4)
    assert deferred_recipe(cv) is None
    np.testing.assert_array_equal(cv["cost_volume"].data, np.arange(24, dtype=np.float32).reshape(2, 3, 4))
    assert calls == [1]

    class _Foreign(dict):                                           # e.g. a real xarray.Dataset: holds arrays only
        pass

    foreign = _Foreign()
    foreign["cost_volume"] = np.zeros((2, 3, 4), dtype=np.float32)
    assert not store_deferred_volume(foreign, rec, (2, 3, 4))


def test_run_refuses_step_orders_the_state_machine_refuses():
    """PandoraMachine._transitions_run (state_machine.py:75-140): volume steps only before `disparity`, map steps only
    after it.  The check happens before any kernel is needed, so it runs without a GPU."""
    img = pb.create_image_dataset(np.zeros((6, 9), np.float32), disparity=[-1, 1])
    with pytest.raises(pb.MachineError, match="cost_volume_confidence from state begin"):
        pb.run(img, img, {"pipeline": {"cost_volume_confidence": {"confidence_method": "ambiguity"}}})
    with pytest.raises(pb.MachineError, match="refinement from state begin"):
        pb.run(img, img, {"pipeline": {"refinement": {"refinement_method": "vfit"}}})
    with pytest.raises(pb.MachineError, match="disparity from state begin"):
        pb.run(img, img, {"pipeline": {"disparity": {"disparity_method": "wta"}}})


def test_library_options_and_path_record():
    """pb200_set_option / pb200_get_option / pb200_last_path: unknown names are refused, values round-trip, the context
    manager restores the previous value; nothing is read from the environment."""
    from pandora_b200 import _native

    lib = _native.load()
    assert _native.get_option("cbca.pipe") == -1
    with _native.option("cbca.pipe", 1):
        assert _native.get_option("cbca.pipe") == 1
    assert _native.get_option("cbca.pipe") == -1
    assert lib.pb200_set_option(b"no.such.option", 1) == _native.ERR_BAD_ARG
    assert b"no.such.option" in lib.pb200_last_error()
    assert _native.last_path("sgm") == ("none", 0)
    assert lib.pb200_last_path(b"nothing", None) == _native.ERR_BAD_ARG
    src = open(os.path.join(ROOT, "pandora_b200", "csrc", "sgm_packed.cuh")).read()
    for name in os.listdir(os.path.join(ROOT, "pandora_b200", "csrc")):
        if name.endswith((".cu", ".cuh")):
            text = open(os.path.join(ROOT, "pandora_b200", "csrc", name)).read()
            outside = text.split("#ifdef PB200_DEBUG_SWITCHES")[0] + "".join(t.split("#endif", 1)[-1] for t in text.split("#ifdef PB200_DEBUG_SWITCHES")[1:])
            assert "getenv" not in outside, f"{name} reads the environment in a release build"
    assert "PB200_DEBUG_SWITCHES" in src


def test_wta_keeps_the_indicator_coordinate_of_the_confidence_layers():
    """disparity.py:462-466 -- every legal pipeline computes cost_volume_confidence BEFORE disparity, so to_disp must hand
    both `confidence_measure` and its `indicator` coordinate to the disparity dataset (a later validation step appends to it)."""
    import inspect

    from pandora_b200 import disparity

    body = inspect.getsource(disparity.WinnerTakesAll)
    assert 'out.coords["indicator"] = cv.coords["indicator"]' in body and body.count("self._carry_confidence(cv, out)") == 2
    # the append that used to fail with KeyError('indicator') on a dataset built like to_disp builds it
    cv = pb.Dataset(coords={"row": np.arange(2), "col": np.arange(3)})
    _, cv = pb.AbstractCostVolumeConfidence.allocate_confidence_map("ambiguity", np.ones((2, 3), np.float32), None, cv)
    out = pb.Dataset({"disparity_map": (("row", "col"), np.zeros((2, 3), np.float32))}, coords={"row": np.arange(2), "col": np.arange(3)})
    out["confidence_measure"] = cv["confidence_measure"]
    out.coords["indicator"] = cv.coords["indicator"]
    out, _ = pb.AbstractCostVolumeConfidence.allocate_confidence_map("left_right_consistency", np.zeros((2, 3), np.float32), out, None)
    assert list(out.coords["indicator"].data) == ["confidence_from_ambiguity", "confidence_from_left_right_consistency"]


def test_sgm_rejects_too_many_disparities_before_any_work():
    cv = pb.Dataset(coords={"row": np.arange(2), "col": np.arange(3), "disp": np.arange(600)}, attrs={"cmax": 25, "type_measure": "min"})
    with pytest.raises(pb.ConfigError, match="exceed"):
        pb.AbstractOptimization(None, optimization_method="sgm").optimize_cv(cv, None, None)


def test_pandora_plugin_module_imports_and_registers_on_stand_in_factories():
    """pandora_plugin_b200 (the module the `pandora.plugin` entry point of pyproject.toml names): importing it without
    Pandora has no effect; on abstract bases with Pandora's `register_subclass` API it registers the seven B200 steps under
    their names, and the factories dispatch to them (src/pandora/__init__.py:141-148, matching_cost.py:88-131)."""
    import pandora_plugin_b200 as plug

    assert plug.register() is False and plug.REGISTERED == {}        # Pandora is not installed here
    text = open(os.path.join(ROOT, "pyproject.toml")).read()
    assert '[project.entry-points."pandora.plugin"]' in text and 'pandora_b200 = "pandora_plugin_b200"' in text

    def factory(key):
        class Base:
            avail = {}

            def __new__(cls, *args, **cfg):
                if cls is Base:
                    return super().__new__(cls.avail[cfg[key]])
                return super().__new__(cls)

            @classmethod
            def register_subclass(cls, name, *aliases):
                def deco(sub):
                    for n in (name,) + aliases:
                        cls.avail[n] = sub
                    return sub
                return deco
        return Base

    mc, agg, opt, disp = factory("matching_cost_method"), factory("aggregation_method"), factory("optimization_method"), factory("disparity_method")
    classes = plug.build_classes(mc, agg, opt, disp)
    assert sorted(classes) == ["cbca_b200", "census_b200", "sad_b200", "sgm_b200", "ssd_b200", "wta_b200", "zncc_b200"]
    step = mc(matching_cost_method="census_b200", window_size=5, subpix=1)
    assert isinstance(step, classes["census_b200"]) and step.cfg["matching_cost_method"] == "census_b200" and step.cfg["window_size"] == 5
    with pytest.raises(pb.ConfigError):
        mc(matching_cost_method="census_b200", window_size=4)                      # same config errors as the mirror classes
    o = opt(None, optimization_method="sgm_b200", penalty={"P1": 4, "P2": 20})
    assert o.cfg["optimization_method"] == "sgm_b200" and o.cfg["penalty"]["P2"] == 20
    assert agg(aggregation_method="cbca_b200", cbca_distance=3).cfg["cbca_distance"] == 3
    assert np.isnan(disp(disparity_method="wta_b200", invalid_disparity="NaN").cfg["invalid_disparity"])


def test_sgm_options_configuration_and_confidence_band_lookup():
    """use_confidence names a cost_volume_confidence step; its ambiguity band is confidence_from_ambiguity[.suffix]
    (state_machine.py:566-576), a missing band means confidence 1 (plugin_libsgm.rst:47); min_cost_paths is a bool."""
    import numpy as np

    from pandora_b200 import ConfigError, Dataset
    from pandora_b200.dataset import DataArray
    from pandora_b200.optimization import Sgm

    with pytest.raises(ConfigError):
        Sgm(None, optimization_method="sgm", use_confidence=3)
    with pytest.raises(ConfigError):
        Sgm(None, optimization_method="sgm", min_cost_paths="yes")
    with pytest.raises(ConfigError):
        Sgm(None, optimization_method="sgm", penalty={"P1": 8, "P2": 8})
    sgm = Sgm(None, optimization_method="sgm", use_confidence="cost_volume_confidence.before", min_cost_paths=True)
    assert sgm.cfg["min_cost_paths"] is True and sgm.cfg["penalty"]["P1"] == 8 and sgm.cfg["penalty"]["P2"] == 32
    cv = Dataset(coords={"row": np.arange(2), "col": np.arange(3), "disp": np.arange(-1, 1)})
    assert sgm._confidence_map(cv) is None                                  # no confidence measure at all
    layers = np.stack([np.full((2, 3), 0.25, np.float32), np.full((2, 3), 0.75, np.float32)], axis=2)
    cv["confidence_measure"] = (("row", "col", "indicator"), layers)
    cv.coords["indicator"] = DataArray(np.array(["confidence_from_ambiguity", "confidence_from_ambiguity.before"]), ("indicator",))
    np.testing.assert_array_equal(sgm._confidence_map(cv), np.full((2, 3), 0.75, np.float32))
    plain = Sgm(None, optimization_method="sgm", use_confidence="cost_volume_confidence")
    np.testing.assert_array_equal(plain._confidence_map(cv), np.full((2, 3), 0.25, np.float32))
    other = Sgm(None, optimization_method="sgm", use_confidence="cost_volume_confidence.after")
    assert other._confidence_map(cv) is None                                # band does not exist: confidence 1
    Sgm._append_band(cv, np.full((2, 3), 8.0, np.float32))
    assert list(cv.coords["indicator"].data)[-1] == "optimization_plugin_libsgm_nb_of_directions"
    assert cv["confidence_measure"].data.shape == (2, 3, 3)
