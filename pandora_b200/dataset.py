"""Duck-typed stand-in for the few xarray features Pandora's hot path touches.

xarray is not installable in the build image, so the step classes are written against the
protocol the reference uses on its datasets (``ds["im"].data``, ``ds.coords["disp"].data``,
``ds.attrs``, ``ds.sizes``, ``"msk" in ds.data_vars``, assignable ``ds["cost_volume"].data``) and
run unchanged on real ``xarray.Dataset`` objects.  This shim implements exactly that protocol and
adds one thing xarray cannot do: a cost volume may stay **device resident** (``LazyVolume``) and is
only copied to the host when somebody reads ``.data``.
"""
from __future__ import annotations

import copy
from typing import Any, Dict, Iterable, Optional

import numpy as np


PINNED_D2H_MAX_BYTES = 1 << 30


def to_host(t) -> np.ndarray:
    """Device tensor -> numpy.  Maps up to 1 GiB (disparity maps, masks, confidence) land in page-locked memory from
    torch's caching host allocator: a fresh pageable array costs a page fault per 4 KB and a staged copy (44 ms for the
    100 MB of C3's two maps, measured), a recycled pinned block one DMA (4 ms).  The numpy array owns its block until the
    caller drops it.  Larger volumes (a 17 GB cost volume read through ``.data``) take the pageable path."""
    if not getattr(t, "is_cuda", False) or t.numel() * t.element_size() > PINNED_D2H_MAX_BYTES:
        return t.cpu().numpy()
    import torch  # noqa: PLC0415

    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t)
    return host.numpy()


class LazyVolume:
    """A float32 (row, col, disp) volume living in HBM; ``materialize()`` performs the D2H copy once.

    The volume may also be *deferred*: ``recipe`` is an object with ``compute() -> tensor`` that is only run when
    somebody asks for ``tensor`` (or ``.data``).  The Census step leaves such a recipe so that a directly following SGM
    step can run the fused Census -> SGM kernels and never write the Census volume (``matching_cost.CensusRecipe``)."""

    def __init__(self, tensor=None, recipe=None, shape=None, host_dtype=None):
        assert tensor is not None or (recipe is not None and shape is not None)
        self._tensor = tensor
        self.recipe = recipe
        self.shape = tuple(int(s) for s in (tensor.shape if tensor is not None else shape))
        # any (row, col[, ...]) variable can stay in HBM this way: `host_dtype` is the dtype the host sees when the device
        # tensor stores it under another view (the uint16 validity mask lives in an int16 tensor)
        self.host_dtype = None if host_dtype is None else np.dtype(host_dtype)
        self.dtype = self.host_dtype or np.dtype(np.float32)
        self._host: Optional[np.ndarray] = None
        self.wta_cache = None                          # (disparity, all-NaN flags, dmin, invalid_disparity) of a fused WTA

    @property
    def tensor(self):
        if self._tensor is None:
            self._tensor = self.recipe.compute()
        return self._tensor

    @property
    def deferred(self) -> bool:
        return self._tensor is None

    def materialize(self) -> np.ndarray:
        if self._host is None:
            self._host = to_host(self.tensor.detach())
            if self.host_dtype is not None and self._host.dtype != self.host_dtype:
                self._host = self._host.view(self.host_dtype)
        return self._host

    def __array__(self, dtype=None, copy=None):  # noqa: A002
        arr = self.materialize()
        return arr if dtype is None else arr.astype(dtype)


class DataArray:
    def __init__(self, data, dims: Iterable[str] = (), coords: Optional[Dict[str, Any]] = None):
        self._data = data
        self.dims = tuple(dims)
        self.coords = dict(coords or {})

    @property
    def data(self) -> np.ndarray:
        if isinstance(self._data, LazyVolume):
            self._data = self._data.materialize()      # from here on the host copy is authoritative
        return self._data

    @data.setter
    def data(self, value) -> None:
        self._data = value

    @property
    def values(self):
        return self.data

    def device_tensor(self):
        """The HBM copy if the volume has not been handed to the host yet, else None (a deferred volume is computed now)."""
        return self._data.tensor if isinstance(self._data, LazyVolume) else None

    def deferred_recipe(self):
        """The recipe of a volume that has not been computed yet, else None."""
        return self._data.recipe if isinstance(self._data, LazyVolume) and self._data.deferred else None

    @property
    def shape(self):
        return tuple(self._data.shape)

    def copy(self, deep: bool = True) -> "DataArray":
        return DataArray(np.array(self.data, copy=True) if deep else self.data, self.dims, dict(self.coords))

    def __array__(self, dtype=None, copy=None):  # noqa: A002
        return np.asarray(self.data, dtype=dtype)


class _Coords(dict):
    def __getitem__(self, key) -> DataArray:
        return super().__getitem__(key)


class Dataset:
    def __init__(self, data_vars: Optional[Dict[str, Any]] = None, coords: Optional[Dict[str, Any]] = None,
                 attrs: Optional[Dict[str, Any]] = None):
        self.coords = _Coords()
        for name, val in (coords or {}).items():
            self.coords[name] = val if isinstance(val, DataArray) else DataArray(np.asarray(val), (name,))
        self.data_vars: Dict[str, DataArray] = {}
        self.attrs: Dict[str, Any] = dict(attrs or {})
        for name, val in (data_vars or {}).items():
            self[name] = val

    def __getitem__(self, name: str) -> DataArray:
        if name in self.data_vars:
            return self.data_vars[name]
        return self.coords[name]

    def __setitem__(self, name: str, value) -> None:
        if isinstance(value, DataArray):
            self.data_vars[name] = value
        elif isinstance(value, tuple):              # (dims, array) like xarray
            dims, arr = value[0], value[1]
            self.data_vars[name] = DataArray(arr, dims)
        else:
            self.data_vars[name] = DataArray(value)

    def __contains__(self, name: str) -> bool:
        return name in self.data_vars or name in self.coords

    @property
    def sizes(self) -> Dict[str, int]:
        return {k: int(np.asarray(v.data).shape[0]) for k, v in self.coords.items()}

    def copy(self, deep: bool = True) -> "Dataset":
        out = Dataset(coords={k: v.copy(deep) for k, v in self.coords.items()}, attrs=copy.deepcopy(self.attrs) if deep else self.attrs)
        for k, v in self.data_vars.items():
            out.data_vars[k] = v.copy(deep)
        return out


def create_image_dataset(img, disparity=None, msk=None, row0: int = 0, col0: int = 0) -> Dataset:
    """What img_tools.create_dataset_from_inputs (img_tools.py:345-437) hands to the pipeline, minus
    the rasterio metadata: ``im`` float32 (row, col) and optionally ``disparity`` (2, row, col)."""
    img = np.asarray(img, dtype=np.float32)
    H, W = img.shape
    ds = Dataset({"im": (("row", "col"), img)}, coords={"row": np.arange(row0, row0 + H), "col": np.arange(col0, col0 + W)},
                 attrs={"valid_pixels": 0, "no_data_mask": 1, "crs": None, "transform": None})
    if msk is not None:
        ds["msk"] = (("row", "col"), np.asarray(msk))
    if disparity is not None:
        add_disparity(ds, disparity)
    return ds


def add_disparity(ds: Dataset, disparity) -> Dataset:
    """img_tools.add_disparity for an integer [min, max] pair: constant (2, row, col) grids."""
    dmin, dmax = int(disparity[0]), int(disparity[1])
    H, W = ds["im"].shape[-2:]
    grid = np.empty((2, H, W), dtype=np.float32)
    grid[0], grid[1] = dmin, dmax
    ds["disparity"] = (("band_disp", "row", "col"), grid)
    ds.coords["band_disp"] = DataArray(np.array(["min", "max"]), ("band_disp",))
    ds.attrs["disparity_source"] = [dmin, dmax]        # img_tools.py:141-161: the [min, max] pair the grids were built from
    return ds
