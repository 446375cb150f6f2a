"""ctypes binding of the C-ABI shared library ``libpandora_b200.so`` (include/pandora_b200.h).

The library is built in-tree by ``pandora_b200/csrc/build.sh`` (nvcc, sm_100a).  There is no CPU
fallback: if the shared object is missing ``load()`` raises, and every compute call returns an
error code (turned into ``Pb200Error``) when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libpandora_b200.so")
_lib = None

OK, ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE = 0, -1, -2, -3, -4
SGM_MAX_DISP = 512


class Pb200Error(RuntimeError):
    """A pb200_* entry point returned a negative status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"pandora_b200 native error {code}: {message}")
        self.code = code


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a (cross-compiles without a GPU)."""
    script = os.path.join(_HERE, "csrc", "build.sh")
    out = subprocess.run(["bash", script], capture_output=True, text=True)
    if verbose or out.returncode:
        print(out.stdout, out.stderr)
    if out.returncode:
        raise RuntimeError("building libpandora_b200.so failed")
    return LIB_PATH


_vp, _ci, _cf, _sz, _cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_double

# name -> (restype, argtypes); mirrors include/pandora_b200.h one to one
PROTOTYPES = {
    "pb200_version": (_ci, []),
    "pb200_last_error": (ctypes.c_char_p, []),
    "pb200_device_count": (_ci, []),
    "pb200_kernel_launches": (ctypes.c_uint64, []),
    "pb200_set_option": (_ci, [ctypes.c_char_p, _ci]),
    "pb200_get_option": (_ci, [ctypes.c_char_p]),
    "pb200_last_path": (_ci, [ctypes.c_char_p, _vp]),
    "pb200_census_workspace_bytes": (_sz, [_ci, _ci, _ci]),
    "pb200_census_cost_volume": (_ci, [_vp, _vp, _ci, _ci, _ci, _ci, _ci, _vp, _vp, _sz, _vp, _cf, _vp, _vp]),
    "pb200_census_cost_volume_rows": (_ci, [_vp, _vp, _ci, _ci, _ci, _ci, _ci, _vp, _vp, _sz, _vp, _cf, _vp, _ci, _ci, _vp]),
    "pb200_census_subpix_workspace_bytes": (_sz, [_ci, _ci, _ci, _ci]),
    "pb200_census_cost_volume_subpix": (_ci, [_vp, _vp, _ci, _ci, _ci, _ci, _ci, _ci, _vp, _vp, _sz, _vp]),
    "pb200_census_cost_volume_multi_host": (_ci, [_vp, _vp, _ci, _ci, _ci, _ci, _vp, _ci, _vp]),
    "pb200_census_descriptors_rows": (_ci, [_vp, _vp, _ci, _ci, _ci, _vp, _sz, _ci, _ci, _vp]),
    "pb200_census_sgm_workspace_bytes": (_sz, [_ci, _ci, _ci, _ci, _ci]),
    "pb200_census_sgm_descriptors": (_ci, [_vp, _vp, _ci, _ci, _ci, _ci, _ci, _cf, _cf, _vp, _sz, _vp, _vp]),
    "pb200_census_sgm": (_ci, [_vp, _vp, _ci, _ci, _ci, _ci, _ci, _cf, _cf, _ci, _vp, _vp, _sz, _vp, _sz, _vp, _cf, _vp, _ci, _vp, _vp]),
    "pb200_sgm_paths_workspace_bytes": (_sz, [_ci, _ci]),
    "pb200_sgm_min_cost_paths": (_ci, [_vp, _vp, _ci, _ci, _ci, _cf, _cf, _cf, _ci, _vp, _vp, _sz, _vp]),
    "pb200_scale_volume": (_ci, [_vp, _vp, _ci, _ci, _ci, _vp, _vp]),
    "pb200_census_sgm_batch": (_ci, [_vp, _vp, _ci, _ci, _ci, _ci, _ci, _ci, _cf, _cf, _ci, _vp, _vp, _sz, _vp, _sz, _vp, _cf, _vp, _vp, _vp]),
    "pb200_tile_link_bytes": (_sz, [_ci]),
    "pb200_census_sgm_tile": (_ci, [_vp, _vp, _ci, _ci, _ci, _ci, _ci, _ci, _cf, _cf, _ci, _ci, _ci, _vp, _vp, _sz, _vp, _sz, _vp, _cf, _vp, _vp,
                                    _vp, _vp, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, _ci, _vp]),
    "pb200_ipc_alloc": (_ci, [_sz, _vp, _vp]),
    "pb200_ipc_open": (_ci, [_vp, _vp]),
    "pb200_ipc_close": (_ci, [_vp]),
    "pb200_ipc_free": (_ci, [_vp]),
    "pb200_sad_ssd_cost_volume": (_ci, [_vp, _vp, _ci, _ci, _ci, _ci, _ci, _ci, _vp, _vp]),
    "pb200_zncc_workspace_bytes": (_sz, [_ci, _ci]),
    "pb200_zncc_cost_volume": (_ci, [_vp, _vp, _ci, _ci, _ci, _ci, _ci, _vp, _vp, _sz, _vp]),
    "pb200_reverse_cost_volume": (_ci, [_vp, _ci, _ci, _ci, _ci, _vp, _vp]),
    "pb200_reverse_disp_range": (_ci, [_vp, _vp, _ci, _ci, _vp, _vp, _vp]),
    "pb200_reverse_disp_range_host": (_ci, [_vp, _vp, _ci, _ci, _vp, _vp]),
    "pb200_median3": (_ci, [_vp, _ci, _ci, _vp, _vp]),
    "pb200_cross_support": (_ci, [_vp, _ci, _ci, _ci, _ci, _cf, _ci, _vp, _vp]),
    "pb200_cbca_aggregate": (_ci, [_vp, _vp, _ci, _ci, _ci, _ci, _ci, _vp, _vp, _ci, _vp]),
    "pb200_sgm_workspace_bytes": (_sz, [_ci, _ci, _ci]),
    "pb200_sgm_flag_offset": (_sz, [_ci, _ci]),
    "pb200_sgm": (_ci, [_vp, _vp, _ci, _ci, _ci, _cf, _cf, _cf, _ci, _ci, _ci, _vp, _vp, _vp, _vp, _vp, _ci, _cf, _vp, _vp, _sz, _vp]),
    "pb200_wta": (_ci, [_vp, _ci, _ci, _ci, _ci, _ci, _cf, _vp, _vp, _vp]),
    "pb200_validity_mask_init": (_ci, [_vp, _ci, _ci, _ci, _ci, _ci, _vp]),
    "pb200_validity_mask": (_ci, [_vp, _vp, _ci, _ci, _ci, _ci, _vp]),
    "pb200_mask_flags": (_ci, [_vp, _ci, _ci, _ci, _ci, _ci, _vp, _vp]),
    "pb200_validity_mask_masks": (_ci, [_vp, _ci, _ci, _ci, _ci, _ci, _vp, _vp, _vp, _vp, _vp]),
    "pb200_cv_masked": (_ci, [_vp, _ci, _ci, _ci, _ci, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pb200_wta_right": (_ci, [_vp, _ci, _ci, _ci, _ci, _ci, _cf, _vp, _vp, _vp]),
    "pb200_cross_checking": (_ci, [_vp, _vp, _vp, _ci, _ci, _cf, _ci, _ci, _ci, _vp, _vp]),
    "pb200_refinement": (_ci, [_vp, _ci, _ci, _ci, _cd, _cd, _ci, _ci, _ci, _ci, _vp, _vp, _vp, _vp]),
    "pb200_filter_median3": (_ci, [_vp, _vp, _ci, _ci, _vp, _vp]),
    "pb200_confidence_workspace_bytes": (_sz, [_ci, _ci, _ci]),
    "pb200_confidence": (_ci, [_vp, _ci, _ci, _ci, _ci, _vp, _ci, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pb200_census_cost_volume_host": (_ci, [_vp, _vp, _ci, _ci, _ci, _vp, _ci, _vp]),
    "pb200_reverse_cost_volume_host": (_ci, [_vp, _ci, _ci, _ci, _ci, _vp]),
    "pb200_cross_support_host": (_ci, [_vp, _ci, _ci, _ci, _cf, _vp]),
    "pb200_cbca_host": (_ci, [_vp, _ci, _ci, _vp, _vp, _vp, _vp, _ci, _vp, _vp]),
    "pb200_disparity_host": (_ci, [_vp, _vp, _ci, _ci, _ci, _ci, _ci, _ci, _ci, _cf, _cf, _cf, _ci, _cf, _vp, _vp, _vp]),
}


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with pandora_b200/csrc/build.sh (python -c 'import __graft_entry__ as g; g.build()'). "
                "pandora_b200 has no CPU fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        raise Pb200Error(rc, load().pb200_last_error().decode(errors="replace"))


def kernel_launches() -> int:
    return int(load().pb200_kernel_launches())


def set_option(name: str, value: int) -> None:
    """Kernel-selection option of the library (include/pandora_b200.h: pb200_set_option); -1 restores the default."""
    check(load().pb200_set_option(name.encode(), int(value)))


def get_option(name: str) -> int:
    return int(load().pb200_get_option(name.encode()))


class option:
    """``with option("cbca.pipe", 1): ...`` -- run a block with one kernel-selection option pinned."""

    def __init__(self, name: str, value: int):
        self.name, self.value = name, value

    def __enter__(self):
        self.old = int(load().pb200_get_option(self.name.encode()))
        set_option(self.name, self.value)
        return self

    def __exit__(self, *exc):
        set_option(self.name, self.old)
        return False


PATHS = {0: "none", 1: "sgm_float", 2: "sgm_packed4", 3: "sgm_wave2", 4: "sgm_wave2_census", 5: "sgm_wave1", 6: "sgm_wave1_census",
         10: "cbca_reg", 11: "cbca_pipe", 12: "cbca_staged", 20: "census_tma", 21: "census_direct", 22: "census_subpix",
         30: "reverse_tiled", 31: "reverse_gather", 40: "sad_taps", 41: "sad_running"}


def last_path(stage: str):
    """(name, detail) of the kernel family that served the last call of ``stage`` on this thread."""
    detail = ctypes.c_int(0)
    code = int(load().pb200_last_path(stage.encode(), ctypes.byref(detail)))
    return PATHS.get(code, str(code)), int(detail.value)
