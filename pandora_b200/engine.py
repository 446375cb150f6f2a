"""Device-side engine: owns nothing but torch tensors used as HBM buffers and calls the C-ABI kernels.

PyTorch is plumbing here (allocation, streams, ``torch.distributed``); every operation below is one or
a few launches of the hand-written sm_100a kernels in ``pandora_b200/csrc`` through
``libpandora_b200.so``.  All volumes are float32 ``(row, col, disp)`` tensors, disparity fastest --
the layout of Pandora's cost volume (``matching_cost/matching_cost.py:394-397`` in the reference).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _native

METHODS = {"census": 0, "sad": 1, "ssd": 2, "zncc": 3}


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Engine:
    """One engine per CUDA device.  Methods are asynchronous on the current torch stream."""

    def __init__(self, device: Optional[torch.device | str | int] = None):
        self.lib = _native.load()
        if not torch.cuda.is_available() or self.lib.pb200_device_count() <= 0:
            raise RuntimeError("pandora_b200 needs a CUDA device: there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._ws = {}

    # ---- helpers ------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def to_device(self, arr, dtype=torch.float32) -> torch.Tensor:
        if isinstance(arr, torch.Tensor):
            return arr.to(device=self.device, dtype=dtype).contiguous()
        host = np.ascontiguousarray(arr, dtype=np.float32 if dtype == torch.float32 else None)
        if host.ndim >= 2 and host.nbytes >= self.STAGED_H2D_MIN_BYTES and host.shape[0] >= 2 * self.STAGED_H2D_CHUNKS:
            return self._staged_upload(host)
        return torch.from_numpy(host).to(self.device, non_blocking=False)

    # A pageable numpy image (what Pandora hands over) reaches the device at ~11 GB/s through the driver's own staging
    # (12 ms for C3's two 67 MB images: the largest part of a plugin-level call).  Large arrays are cut into row chunks that
    # worker threads copy into a page-locked staging buffer (numpy releases the GIL for the copy) while the chunks already
    # staged travel over PCIe: the upload then runs at the speed of the slower of the two, not at their sum.
    STAGED_H2D_MIN_BYTES = 16 << 20
    STAGED_H2D_CHUNKS = 8

    def _staged_upload(self, host: np.ndarray) -> torch.Tensor:
        from concurrent.futures import ThreadPoolExecutor  # noqa: PLC0415

        t_host = torch.from_numpy(host)
        stage = getattr(self, "_h2d_stage", None)
        if stage is None or stage[0].numel() < t_host.numel() or stage[0].dtype != t_host.dtype:
            stage = (torch.empty(t_host.numel(), dtype=t_host.dtype, pin_memory=True), None)
        if stage[1] is not None:
            stage[1].synchronize()                            # the previous upload has left the staging buffer
        pinned = stage[0][: t_host.numel()].view(t_host.shape)
        pinned_np = pinned.numpy()
        out = torch.empty(t_host.shape, dtype=t_host.dtype, device=self.device)
        n = self.STAGED_H2D_CHUNKS
        rows = host.shape[0]
        edges = [rows * i // n for i in range(n + 1)]
        if getattr(self, "_h2d_pool", None) is None:
            self._h2d_pool = ThreadPoolExecutor(max_workers=4)
        futs = [self._h2d_pool.submit(np.copyto, pinned_np[edges[i]: edges[i + 1]], host[edges[i]: edges[i + 1]]) for i in range(n)]
        with torch.cuda.device(self.device):
            for i, f in enumerate(futs):
                f.result()
                out[edges[i]: edges[i + 1]].copy_(pinned[edges[i]: edges[i + 1]], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
        self._h2d_stage = (stage[0], ev)
        return out

    def empty(self, shape, dtype=torch.float32) -> torch.Tensor:
        return torch.empty(shape, dtype=dtype, device=self.device)

    def _workspace(self, key: str, nbytes: int) -> torch.Tensor:
        buf = self._ws.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=self.device)
            self._ws[key] = buf
        return buf

    @staticmethod
    def _hw(img: torch.Tensor) -> Tuple[int, int]:
        assert img.dim() == 2 and img.dtype == torch.float32 and img.is_contiguous()
        return int(img.shape[0]), int(img.shape[1])

    # ---- matching cost ------------------------------------------------------------------------------
    def census(self, left: torch.Tensor, right: torch.Tensor, window: int, dmin: int, dmax: int,
               out: Optional[torch.Tensor] = None, fuse_wta: bool = False, invalid_disparity: float = -9999.0,
               rows: Optional[Tuple[int, int]] = None, disp: Optional[torch.Tensor] = None, flags: Optional[torch.Tensor] = None):
        """Census cost volume; with ``fuse_wta`` also returns (disparity map, all-NaN flags).  ``rows`` = (begin, end)
        restricts the call to a band of rows of the volume (only the image rows within half a window of the band need
        to be resident yet -- see ``StereoPipeline.run_host``)."""
        H, W = self._hw(left)
        D = dmax - dmin + 1
        cv = self.empty((H, W, D)) if out is None else out
        nbytes = self.lib.pb200_census_workspace_bytes(H, W, window)
        ws = self._workspace("census", nbytes)
        if fuse_wta and disp is None:
            disp = self.empty((H, W))
            flags = self.empty((H, W), torch.uint8)
        r0, r1 = (0, H) if rows is None else rows
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_census_cost_volume_rows(
                _ptr(left), _ptr(right), H, W, window, dmin, D, _ptr(cv), _ptr(ws), ws.numel(),
                _ptr(disp) if fuse_wta else None, float(invalid_disparity), _ptr(flags) if fuse_wta else None, int(r0), int(r1),
                self._stream()))
        return (cv, disp, flags) if fuse_wta else cv

    def census_subpix(self, left: torch.Tensor, rights, window: int, dmin: int, n_disp: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Sub-pixel Census volume (``pb200_census_cost_volume_subpix``): ``rights`` = [right (H, W), shifted copies (H, W - 1) ...]."""
        import ctypes  # noqa: PLC0415

        H, W = self._hw(left)
        for i, r in enumerate(rights):
            assert r.dtype == torch.float32 and r.is_contiguous() and tuple(r.shape) == (H, W if i == 0 else W - 1)
        cv = self.empty((H, W, n_disp)) if out is None else out
        ws = self._workspace("census", self.lib.pb200_census_subpix_workspace_bytes(H, W, window, len(rights)))
        ptrs = (ctypes.c_void_p * len(rights))(*[r.data_ptr() for r in rights])
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_census_cost_volume_subpix(_ptr(left), ptrs, len(rights), H, W, window, int(dmin), int(n_disp), _ptr(cv),
                                                                   _ptr(ws), ws.numel(), self._stream()))
        return cv

    def census_descriptors(self, left: torch.Tensor, right: torch.Tensor, window: int, rows: Optional[Tuple[int, int]] = None) -> None:
        """Census transform of both images (or of a band of rows) into the engine's descriptor workspace."""
        H, W = self._hw(left)
        ws = self._workspace("census", self.lib.pb200_census_workspace_bytes(H, W, window))
        r0, r1 = (0, H) if rows is None else rows
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_census_descriptors_rows(_ptr(left), _ptr(right), H, W, window, _ptr(ws), ws.numel(), int(r0), int(r1),
                                                                 self._stream()))

    def census_sgm_descriptors(self, left: torch.Tensor, right: torch.Tensor, window: int, dmin: int, dmax: int, p1: float, p2: float) -> bool:
        """The census descriptors of both images in the layout ``census_sgm(..., descriptors_ready=True)`` will read for the
        same arguments (``pb200_census_sgm_descriptors``).  False when the configuration is not eligible for the fused stage."""
        import ctypes  # noqa: PLC0415

        H, W = self._hw(left)
        D = dmax - dmin + 1
        ws = self._workspace("census", self.lib.pb200_census_sgm_workspace_bytes(H, W, window, dmin, D))
        ok = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_census_sgm_descriptors(_ptr(left), _ptr(right), H, W, window, dmin, D, float(p1), float(p2), _ptr(ws),
                                                                ws.numel(), ctypes.addressof(ok), self._stream()))
        return bool(ok.value)

    def census_sgm(self, left: torch.Tensor, right: torch.Tensor, window: int, dmin: int, dmax: int, p1: float, p2: float,
                   overcounting: bool = False, out: Optional[torch.Tensor] = None, fuse_wta: bool = True, invalid_disparity: float = -9999.0,
                   disp: Optional[torch.Tensor] = None, flags: Optional[torch.Tensor] = None, descriptors_ready: bool = False):
        """Fused Census -> SGM [-> WTA] (``pb200_census_sgm``): the Hamming costs go from the census descriptors straight
        into the first SGM pass, the float32 Census volume is never materialised.  Returns ``None`` when the configuration
        is not eligible (nothing was computed: run ``census`` + ``sgm``), else (SGM volume, disparity, all-NaN flags)."""
        import ctypes  # noqa: PLC0415

        H, W = self._hw(left)
        D = dmax - dmin + 1
        res = self.empty((H, W, D)) if out is None else out
        if fuse_wta and disp is None:
            disp = self.empty((H, W))
            flags = self.empty((H, W), torch.uint8)
        cws = self._workspace("census", self.lib.pb200_census_sgm_workspace_bytes(H, W, window, dmin, D))
        sws = self._workspace("sgm", self.lib.pb200_sgm_workspace_bytes(H, W, D))
        ran = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_census_sgm(
                _ptr(left), _ptr(right), H, W, window, dmin, D, float(p1), float(p2), int(bool(overcounting)), _ptr(res), _ptr(cws),
                cws.numel(), _ptr(sws), sws.numel(), _ptr(disp) if fuse_wta else None, float(invalid_disparity),
                _ptr(flags) if fuse_wta else None, int(bool(descriptors_ready)), ctypes.addressof(ran), self._stream()))
        if not ran.value:
            return None
        return (res, disp, flags) if fuse_wta else (res, None, None)

    def census_sgm_batch(self, left: torch.Tensor, right: torch.Tensor, window: int, dmin: int, dmax: int, p1: float, p2: float,
                         overcounting: bool = False, out: Optional[torch.Tensor] = None, invalid_disparity: float = -9999.0,
                         disp: Optional[torch.Tensor] = None, flags: Optional[torch.Tensor] = None):
        """``pb200_census_sgm_batch``: a batch (n, H, W) of pairs through ONE wave per pass of the fused Census -> SGM -> WTA
        stage (the fill and drain of the wave across the SMs are paid once per batch); the results equal ``n`` calls of
        ``census_sgm`` bit for bit.  Returns ``None`` when not eligible, else (volumes (n, H, W, D), disparities, all-NaN flags)."""
        import ctypes  # noqa: PLC0415

        assert left.dim() == 3 and left.shape == right.shape and left.is_contiguous() and right.is_contiguous() and left.dtype == torch.float32
        n, H, W = (int(v) for v in left.shape)
        D = dmax - dmin + 1
        res = self.empty((n, H, W, D)) if out is None else out
        if disp is None:
            disp = self.empty((n, H, W))
            flags = self.empty((n, H, W), torch.uint8)
        cws = self._workspace("census", n * self.lib.pb200_census_sgm_workspace_bytes(H, W, window, dmin, D))
        sws = self._workspace("sgm", self.lib.pb200_sgm_workspace_bytes(n * H, W, D))
        ran = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_census_sgm_batch(
                _ptr(left), _ptr(right), n, H, W, window, dmin, D, float(p1), float(p2), int(bool(overcounting)), _ptr(res), _ptr(cws),
                cws.numel(), _ptr(sws), sws.numel(), _ptr(disp), float(invalid_disparity), _ptr(flags), ctypes.addressof(ran), self._stream()))
        return (res, disp, flags) if ran.value else None

    def sad_ssd(self, left, right, window: int, dmin: int, dmax: int, squared: bool = False, out=None) -> torch.Tensor:
        H, W = self._hw(left)
        D = dmax - dmin + 1
        cv = self.empty((H, W, D)) if out is None else out
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_sad_ssd_cost_volume(_ptr(left), _ptr(right), H, W, window, dmin, D, int(squared),
                                                             _ptr(cv), self._stream()))
        return cv

    def zncc(self, left, right, window: int, dmin: int, dmax: int, out=None) -> torch.Tensor:
        H, W = self._hw(left)
        D = dmax - dmin + 1
        cv = self.empty((H, W, D)) if out is None else out
        ws = self._workspace("zncc", self.lib.pb200_zncc_workspace_bytes(H, W))
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_zncc_cost_volume(_ptr(left), _ptr(right), H, W, window, dmin, D, _ptr(cv), _ptr(ws),
                                                          ws.numel(), self._stream()))
        return cv

    def matching_cost(self, method: str, left, right, window: int, dmin: int, dmax: int, out=None) -> torch.Tensor:
        if method == "census":
            return self.census(left, right, window, dmin, dmax, out=out)
        if method in ("sad", "ssd"):
            return self.sad_ssd(left, right, window, dmin, dmax, squared=(method == "ssd"), out=out)
        if method == "zncc":
            return self.zncc(left, right, window, dmin, dmax, out=out)
        raise KeyError(f"No matching cost method named {method} supported")

    def reverse_cost_volume(self, left_cv: torch.Tensor, min_disp: int) -> torch.Tensor:
        H, W, D = (int(s) for s in left_cv.shape)
        out = torch.empty_like(left_cv)
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_reverse_cost_volume(_ptr(left_cv), H, W, D, int(min_disp), _ptr(out), self._stream()))
        return out

    def reverse_disp_range(self, left_min: torch.Tensor, left_max: torch.Tensor):
        """matching_cost.cpp:59-131 on the device: (right_min, right_max) float32 (H, W) grids."""
        H, W = self._hw(left_min)
        rmin, rmax = torch.empty_like(left_min), torch.empty_like(left_max)
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_reverse_disp_range(_ptr(left_min), _ptr(left_max), H, W, _ptr(rmin), _ptr(rmax), self._stream()))
        return rmin, rmax

    # ---- aggregation --------------------------------------------------------------------------------
    def median3(self, img: torch.Tensor) -> torch.Tensor:
        H, W = self._hw(img)
        out = torch.empty_like(img)
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_median3(_ptr(img), H, W, _ptr(out), self._stream()))
        return out

    def cross_support(self, img: torch.Tensor, len_arms: int, intensity: float, offset: int = 0, nan_as_inf: bool = False):
        """Arms of ``img[offset:-offset, offset:-offset]`` (a strided view: no copy), (H', W', 4) int16."""
        H, W = self._hw(img)
        Hi, Wi = H - 2 * offset, W - 2 * offset
        out = self.empty((Hi, Wi, 4), torch.int16)
        base = img.data_ptr() + (offset * W + offset) * 4
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_cross_support(base, Hi, Wi, W, int(len_arms), float(intensity), int(nan_as_inf),
                                                       _ptr(out), self._stream()))
        return out

    def cbca_supports(self, left, right, offset: int, distance: int, intensity: float):
        """computes_cross_supports (cbca.py:184-295) for subpix 1 without masks."""
        cl = self.cross_support(self.median3(left), distance, intensity, offset, nan_as_inf=True)
        cr = self.cross_support(self.median3(right), distance, intensity, offset, nan_as_inf=True)
        return cl, cr

    def cbca(self, left, right, cv: torch.Tensor, offset: int, dmin: int, distance: int = 5, intensity: float = 30.0,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
        H, W, D = (int(s) for s in cv.shape)
        cl, cr = self.cbca_supports(left, right, offset, distance, intensity)
        res = torch.empty_like(cv) if out is None else out
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_cbca_aggregate(_ptr(cv), _ptr(res), H, W, D, int(dmin), int(offset), _ptr(cl), _ptr(cr),
                                                        int(distance), self._stream()))
        return res

    # ---- optimisation -------------------------------------------------------------------------------
    def sgm(self, cv: torch.Tensor, p1: float, p2: float, invalid_value: float, overcounting: bool = False,
            out: Optional[torch.Tensor] = None, fuse_wta: bool = False, dmin: int = 0, invalid_disparity: float = -9999.0,
            dir_mask: int = 0xFF, init_final: int = 3, halo_in_top=None, halo_in_bottom=None, halo_out_bottom=None, halo_out_top=None,
            disp: Optional[torch.Tensor] = None, flags: Optional[torch.Tensor] = None, packed: bool = False,
            float_only: bool = False):
        """8-path SGM (or the directions in ``dir_mask``).  ``packed``: split calls may leave the exact 16-bit
        intermediate representation in ``out`` / the halo buffers (include/pandora_b200.h, init_final bit 2)."""
        H, W, D = (int(s) for s in cv.shape)
        if packed:
            init_final |= 4
        if float_only:
            init_final |= 8
        res = torch.empty_like(cv) if out is None else out
        if fuse_wta and disp is None:
            disp = self.empty((H, W))
            flags = self.empty((H, W), torch.uint8)
        ws = self._workspace("sgm", self.lib.pb200_sgm_workspace_bytes(H, W, D))
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_sgm(
                _ptr(cv), _ptr(res), H, W, D, float(p1), float(p2), float(invalid_value), int(bool(overcounting)), int(dir_mask), int(init_final),
                _ptr(halo_in_top), _ptr(halo_in_bottom), _ptr(halo_out_bottom), _ptr(halo_out_top),
                _ptr(disp) if fuse_wta else None, int(dmin), float(invalid_disparity), _ptr(flags) if fuse_wta else None,
                _ptr(ws), ws.numel(), self._stream()))
        return (res, disp, flags) if fuse_wta else res

    def sgm_path_flag(self, W: int, D: int) -> torch.Tensor:
        """int32 view (1 element) of the fast-path flag inside the SGM workspace: 0 = the packed integer path ran."""
        H = 1
        ws = self._workspace("sgm", self.lib.pb200_sgm_workspace_bytes(H, W, D))
        off = int(self.lib.pb200_sgm_flag_offset(W, D))
        return ws[off: off + 4].view(torch.int32)

    def sgm_min_cost_paths(self, cv: torch.Tensor, p1: float, p2: float, invalid_value: float, overcounting: bool = False):
        """SGM with ``min_cost_paths`` (plugin_libsgm.rst:411-413): (optimised volume, nb_of_directions (H, W) float32)."""
        H, W, D = (int(v) for v in cv.shape)
        out = torch.empty_like(cv)
        nb = self.empty((H, W))
        ws = self._workspace("sgm_paths", self.lib.pb200_sgm_paths_workspace_bytes(H, W))
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_sgm_min_cost_paths(_ptr(cv), _ptr(out), H, W, D, float(p1), float(p2), float(invalid_value),
                                                            int(bool(overcounting)), _ptr(nb), _ptr(ws), ws.numel(), self._stream()))
        return out, nb

    def scale_volume(self, cv: torch.Tensor, confidence: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """cv(p, d) * confidence(p) (``use_confidence`` of the SGM step, plugin_libsgm.rst:38-47)."""
        H, W, D = (int(v) for v in cv.shape)
        res = torch.empty_like(cv) if out is None else out
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_scale_volume(_ptr(cv), _ptr(confidence), H, W, D, _ptr(res), self._stream()))
        return res

    # ---- disparity ----------------------------------------------------------------------------------
    def wta(self, cv: torch.Tensor, dmin: int, is_max: bool = False, invalid_disparity: float = -9999.0):
        H, W, D = (int(s) for s in cv.shape)
        disp = self.empty((H, W))
        flags = self.empty((H, W), torch.uint8)
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_wta(_ptr(cv), H, W, D, int(dmin), int(is_max), float(invalid_disparity), _ptr(disp),
                                             _ptr(flags), self._stream()))
        return disp, flags

    def validity_mask_init(self, H: int, W: int, dmin: int, dmax: int, offset: int) -> torch.Tensor:
        """criteria.validity_mask without image masks (criteria.py:106-147): the per-column disparity-range bits."""
        mask = self.empty((H, W), torch.int16)           # uint16 bit flags stored in an int16 tensor
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_validity_mask_init(_ptr(mask), H, W, int(dmin), int(dmax), int(offset), self._stream()))
        return mask

    def validity_mask(self, H: int, W: int, dmin: int, dmax: int, offset: int, flags: Optional[torch.Tensor] = None,
                      wta_invalidate: bool = False, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        with torch.cuda.device(self.device):
            if mask is None:
                mask = self.empty((H, W), torch.int16)       # uint16 bit flags stored in an int16 tensor
                _native.check(self.lib.pb200_validity_mask_init(_ptr(mask), H, W, int(dmin), int(dmax), int(offset), self._stream()))
            _native.check(self.lib.pb200_validity_mask(_ptr(mask), _ptr(flags), H, W, int(offset), int(wta_invalidate),
                                                       self._stream()))
        return mask

    # ---- input masks / disparity grids -------------------------------------------
    def mask_flags(self, msk, valid_pixels: int, no_data: int, window: int) -> torch.Tensor:
        """Per-pixel flag byte of an image mask (dilated no_data | invalid | not valid), criteria.py:36-63."""
        if isinstance(msk, torch.Tensor):                        # already on the device (int16)
            d_msk = msk.to(device=self.device, dtype=torch.int16).contiguous()
            H, W = (int(v) for v in d_msk.shape)
        else:
            host = np.ascontiguousarray(np.asarray(msk), dtype=np.int16)
            H, W = host.shape
            d_msk = torch.from_numpy(host).to(self.device)
        flags = self.empty((H, W), torch.uint8)
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_mask_flags(_ptr(d_msk), H, W, int(valid_pixels), int(no_data), int(window), _ptr(flags), self._stream()))
        return flags

    def validity_mask_masks(self, mask: torch.Tensor, dmin: int, dmax: int, offset: int, flags_left=None, flags_right=None,
                            grid_min=None, grid_max=None) -> torch.Tensor:
        H, W = (int(s) for s in mask.shape)
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_validity_mask_masks(_ptr(mask), H, W, int(dmin), int(dmax), int(offset), _ptr(flags_left),
                                                             _ptr(flags_right), _ptr(grid_min), _ptr(grid_max), self._stream()))
        return mask

    def cv_masked(self, cv: torch.Tensor, dmin: int, flags_left=None, flags_right=None, grid_min=None, grid_max=None) -> torch.Tensor:
        """In-place masking of the volume (matching_cost.py:815-856); returns the all-NaN flags (H, W) uint8."""
        H, W, D = (int(s) for s in cv.shape)
        all_nan = self.empty((H, W), torch.uint8)
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_cv_masked(_ptr(cv), H, W, D, int(dmin), _ptr(flags_left), _ptr(flags_right), _ptr(grid_min),
                                                   _ptr(grid_max), _ptr(all_nan), self._stream()))
        return all_nan

    # ---- fast cross-checking --------------------------------------------------------
    def wta_right(self, left_cv: torch.Tensor, min_disp_right: int, is_max: bool = False, invalid_disparity: float = -9999.0):
        """Right disparity map from the LEFT volume (no right volume is materialised); falls back to
        reverse_cost_volume + wta only for shapes the fused kernel does not take (D % 4 != 0, D > 992)."""
        H, W, D = (int(s) for s in left_cv.shape)
        if D % 4 != 0 or D > 992 or left_cv.data_ptr() % 16 != 0:
            return self.wta(self.reverse_cost_volume(left_cv, min_disp_right), min_disp_right, is_max, invalid_disparity)
        disp = self.empty((H, W))
        flags = self.empty((H, W), torch.uint8)
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_wta_right(_ptr(left_cv), H, W, D, int(min_disp_right), int(is_max), float(invalid_disparity),
                                                   _ptr(disp), _ptr(flags), self._stream()))
        return disp, flags

    def cross_checking(self, disp_left: torch.Tensor, mask_left: torch.Tensor, disp_right: torch.Tensor, threshold: float, dmin: int,
                       dmax: int, offset: int = 0):
        """validation.py:226-371; ``mask_left`` (int16 storage of the uint16 flags) is updated in place; returns the
        left-right distance map."""
        H, W = (int(s) for s in disp_left.shape)
        conf = self.empty((H, W))
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_cross_checking(_ptr(disp_left), _ptr(mask_left), _ptr(disp_right), H, W,
                                                        float(np.float32(threshold)), int(dmin), int(dmax), int(offset), _ptr(conf),
                                                        self._stream()))
        return conf

    # ---- sub-pixel refinement --------------------------------------------------------
    def refinement(self, cv: torch.Tensor, disp: torch.Tensor, mask: torch.Tensor, d_min: float, d_max: float, subpix: int = 1,
                   is_max: bool = False, method: str = "vfit", approximate=False) -> torch.Tensor:
        """``disp`` (float32) and ``mask`` (int16 storage) are refined in place; returns the interpolated coefficients.
        ``approximate``: False / 0 = the volume's own map, True / 1 = loop_approximate_refinement, 2 = right map refined
        on the reversed volume read from the LEFT one (d_min / d_max = right coordinates)."""
        H, W, D = (int(s) for s in cv.shape)
        itp = self.empty((H, W))
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_refinement(_ptr(cv), H, W, D, float(d_min), float(d_max), int(subpix), int(is_max),
                                                    {"vfit": 0, "quadratic": 1}[method], int(approximate), _ptr(disp), _ptr(mask),
                                                    _ptr(itp), self._stream()))
        return itp

    # ---- cost-volume confidence -------------------------------------------------------
    def confidence(self, cv: torch.Tensor, etas, grids=None, disparity_range=None, is_max: bool = False, ambiguity: bool = True,
                   sampled_ambiguity: bool = False, risk: bool = False, sampled_risk: bool = False, sampled_ambiguity_in=None) -> dict:
        """Ambiguity and / or risk of a volume in one pass (ambiguity.cpp:28-142, risk.cpp:28-197).  ``grids``: (2, H, W)
        integer [disp_min, disp_max] per pixel (host or device) or None; ``disparity_range``: the disp coordinates."""
        import ctypes  # noqa: PLC0415

        H, W, D = (int(s) for s in cv.shape)
        et = np.ascontiguousarray(etas, dtype=np.float64)
        n = int(et.shape[0])
        dr = torch.as_tensor(np.ascontiguousarray(disparity_range, dtype=np.float32)).to(self.device)
        g = None
        if grids is not None:
            g = grids if isinstance(grids, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(grids, dtype=np.int32))
            g = g.to(device=self.device, dtype=torch.int32).contiguous()
        out = {}
        if ambiguity:
            out["ambiguity"] = self.empty((H, W))
        if sampled_ambiguity:
            out["sampled_ambiguity"] = self.empty((H, W, n))
        if risk:
            for k in ("risk_max", "risk_min", "disp_sup", "disp_inf"):
                out[k] = self.empty((H, W))
        if sampled_risk:
            out["sampled_risk_max"] = self.empty((H, W, n))
            out["sampled_risk_min"] = self.empty((H, W, n))
        sa_in = None if sampled_ambiguity_in is None else self.to_device(sampled_ambiguity_in)
        ws = self._workspace("confidence", self.lib.pb200_confidence_workspace_bytes(H, W, n))
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_confidence(
                _ptr(cv), H, W, D, int(is_max), et.ctypes.data_as(ctypes.c_void_p), n, _ptr(g), _ptr(dr), _ptr(out.get("ambiguity")),
                _ptr(out.get("sampled_ambiguity")), _ptr(sa_in), _ptr(out.get("risk_max")), _ptr(out.get("risk_min")),
                _ptr(out.get("disp_sup")), _ptr(out.get("disp_inf")), _ptr(out.get("sampled_risk_max")), _ptr(out.get("sampled_risk_min")),
                _ptr(ws), ws.numel(), self._stream()))
        return out

    # ---- disparity filter ------------------------------------------------------------------------------------
    def filter_median3(self, disp: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        """3x3 median of the valid pixels of a disparity map, in place (filter/median.py:96-179)."""
        H, W = (int(s) for s in disp.shape)
        scratch = self.empty((2, H, W))
        with torch.cuda.device(self.device):
            _native.check(self.lib.pb200_filter_median3(_ptr(disp), _ptr(mask), H, W, _ptr(scratch), self._stream()))
        return disp
