"""Shared pieces of the step classes: config errors, engine cache, device-residency helpers."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .dataset import DataArray, Dataset, LazyVolume


class ConfigError(ValueError):
    """Bad step configuration (the reference raises json_checker's DictCheckerError here; the message
    names the offending key the same way, e.g. tests/test_matching_cost/test_matching_cost_census.py:48-52)."""


class MachineError(Exception):
    """A step triggered in a state that has no transition for it (``transitions.MachineError`` in the reference:
    state_machine.py:75-140 lists the legal ones, ``PandoraMachine.run`` re-raises it, state_machine.py:718-720)."""


_ENGINES: Dict[str, object] = {}


def get_engine(device: Optional[str] = None):
    """One Engine per device, created on first use.  Raises when CUDA / the native library is missing."""
    import torch  # noqa: PLC0415

    from .engine import Engine  # noqa: PLC0415

    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("pandora_b200 needs a CUDA device: there is no CPU fallback")
        device = f"cuda:{torch.cuda.current_device()}"
    eng = _ENGINES.get(device)
    if eng is None:
        eng = Engine(device)
        _ENGINES[device] = eng
    return eng


def image_array(ds, band: Optional[str] = None) -> np.ndarray:
    """``ds["im"].data`` (optionally one band) as float32, like census.py:124-147."""
    data = ds["im"].data
    if band is not None:
        idx = list(ds.coords["band_im"].data).index(band)
        data = data[idx, :, :]
    return np.ascontiguousarray(data, dtype=np.float32)


def device_volume(engine, cv):
    """The cost volume of dataset ``cv`` as a device tensor: the resident copy when a previous
    pandora_b200 step left one, else an upload of the host array."""
    var = cv["cost_volume"]
    if isinstance(var, DataArray):
        t = var.device_tensor()
        if t is not None:
            return t
    # a real xarray dataset driven through pandora_plugin_b200 with keep_host_copy = False: the host array is a placeholder,
    # the volume is the tensor the previous B200 step recorded
    if cv.attrs.get("pb200_resident_trusted"):
        t = (cv.attrs.get("pb200_resident") or {}).get("cost_volume")
        if t is not None:
            return t
    return engine.to_device(np.ascontiguousarray(var.data, dtype=np.float32))


def store_volume(cv, tensor, keep_on_device: bool = True) -> None:
    """Put a device tensor back into ``cv["cost_volume"]``: lazily for the shim dataset, with an
    immediate D2H copy for a real xarray dataset (xarray coerces to numpy anyway)."""
    var = cv["cost_volume"] if "cost_volume" in cv else None
    if keep_on_device and (var is None or isinstance(var, DataArray)):
        if var is None:
            cv["cost_volume"] = (("row", "col", "disp"), LazyVolume(tensor))
        else:
            var.data = LazyVolume(tensor)
    else:
        cv["cost_volume"].data = tensor.detach().cpu().numpy()


def device_var(engine, ds, name: str, dtype: str = "float32"):
    """Variable ``name`` of dataset ``ds`` as a device tensor: the resident copy when a previous pandora_b200 step left one
    (no host round trip between steps), else an upload.  ``dtype`` "uint16" gives the int16 view the mask kernels take."""
    var = ds[name]
    if isinstance(var, DataArray):
        t = var.device_tensor()
        if t is not None:
            return t
    if dtype == "uint16":
        return engine.to_device(np.ascontiguousarray(var.data).astype(np.uint16).view(np.int16), dtype=None)
    return engine.to_device(np.ascontiguousarray(var.data, dtype=np.float32))


def store_var(ds, name: str, tensor, dims=("row", "col"), dtype: str = "float32") -> None:
    """Leave ``tensor`` as variable ``name`` of ``ds`` WITHOUT copying it to the host (shim datasets; the copy happens
    when somebody reads ``.data``); a real xarray dataset gets the numpy array at once."""
    host_dtype = np.uint16 if dtype == "uint16" else None
    if isinstance(ds, Dataset):
        ds[name] = (tuple(dims), LazyVolume(tensor, host_dtype=host_dtype))
    else:
        arr = tensor.detach().cpu().numpy()
        ds[name] = (tuple(dims), arr.view(np.uint16) if host_dtype is not None else arr)


def fused_wta(cv):
    """(disparity, all-NaN flags, dmin, invalid_disparity) when the kernel that produced ``cv["cost_volume"]`` also ran the
    winner-takes-all on it (fused Census -> SGM -> WTA) and nothing has replaced the volume since, else None."""
    var = cv["cost_volume"] if "cost_volume" in cv else None
    lazy = getattr(var, "_data", None)
    return getattr(lazy, "wta_cache", None) if isinstance(lazy, LazyVolume) else None


def deferred_recipe(cv):
    """The recipe of ``cv["cost_volume"]`` when the matching-cost step deferred it and nothing has read it yet."""
    var = cv["cost_volume"] if "cost_volume" in cv else None
    return var.deferred_recipe() if isinstance(var, DataArray) else None


def store_deferred_volume(cv, recipe, shape) -> bool:
    """Leave ``recipe`` in ``cv["cost_volume"]`` instead of a tensor (shim datasets only: a real xarray dataset wants a
    numpy array at once).  Returns False when the dataset cannot hold it."""
    var = cv["cost_volume"] if "cost_volume" in cv else None
    if var is not None and not isinstance(var, DataArray):
        return False
    try:
        if var is None:
            cv["cost_volume"] = (("row", "col", "disp"), LazyVolume(recipe=recipe, shape=shape))
            if not isinstance(cv["cost_volume"], DataArray):      # a real xarray.Dataset coerced it
                return False
        else:
            var.data = LazyVolume(recipe=recipe, shape=shape)
    except Exception:                                             # noqa: BLE001 -- xarray refuses non-array data
        return False
    return True
