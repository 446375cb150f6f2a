"""Pipeline drivers.

``run`` mirrors ``pandora.run`` + the four hot-path callbacks of ``PandoraMachine``
(src/pandora/__init__.py:51-124, state_machine.py:292-448) for the steps this package implements:
the datasets go from step to step exactly like in the reference, the cost volume staying in HBM in
between.  ``StereoPipeline`` is the fused, allocation-free fast path for a fixed configuration
(what bench.py times): images in, disparity map out.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from ._common import get_engine
from .aggregation import AbstractAggregation
from .disparity import AbstractDisparity
from .matching_cost import AbstractMatchingCost
from .optimization import AbstractOptimization

HOT_PATH_STEPS = ("matching_cost", "aggregation", "optimization", "disparity")


def run(img_left, img_right, cfg: dict):
    """Run the hot-path steps named in ``cfg["pipeline"]`` in order; returns (left disparity dataset, cost volume).

    ``cfg`` is a Pandora user configuration (``{"pipeline": {"matching_cost": {...}, ...}}``); steps outside
    the hot path (refinement, filter, validation, ...) are rejected -- they belong to Pandora itself.
    """
    pipeline = cfg["pipeline"]
    disp_grids = (img_left["disparity"].data[0], img_left["disparity"].data[1])
    cv = None
    disp = None
    for step, step_cfg in pipeline.items():
        name = step.split(".")[0]
        if name not in HOT_PATH_STEPS:
            raise NotImplementedError(f"step {step!r} is outside the B200 hot path (use Pandora's own implementation)")
        if name == "matching_cost":                               # state_machine.py:292-364
            mc = AbstractMatchingCost(**step_cfg)
            cv = mc.allocate_cost_volume(img_left, disp_grids, cfg)
            cv = mc.compute_cost_volume(img_left, img_right, cv)
            mc.cv_masked(img_left, img_right, cv, *disp_grids)
        elif name == "aggregation":                               # state_machine.py:366-380
            AbstractAggregation(**step_cfg).cost_volume_aggregation(img_left, img_right, cv)
        elif name == "optimization":                              # state_machine.py:404-419
            cv = AbstractOptimization(img_left, **step_cfg).optimize_cv(cv, img_left, img_right)
        elif name == "disparity":                                 # state_machine.py:421-448
            disp = AbstractDisparity(**step_cfg).to_disp(cv, img_left, img_right)
    return disp, cv


class StereoPipeline:
    """Fixed-configuration device pipeline: matching cost -> [CBCA] -> [SGM] -> WTA, buffers allocated once.

    ``method`` in {"census", "sad", "ssd", "zncc"}; ``cbca`` = None or (distance, intensity);
    ``sgm`` = None or (P1, P2[, overcounting]).  WTA is fused into the producing kernel when the last
    volume step is Census or SGM.
    """

    def __init__(self, H: int, W: int, dmin: int, dmax: int, method: str = "census", window: int = 5,
                 cbca: Optional[Tuple[int, float]] = None, sgm: Optional[Tuple] = None, invalid_disparity: float = -9999.0,
                 device: Optional[str] = None):
        import torch  # noqa: PLC0415

        self.torch = torch
        self.eng = get_engine(device)
        self.H, self.W, self.dmin, self.dmax = H, W, dmin, dmax
        self.D = dmax - dmin + 1
        self.method, self.window, self.cbca, self.sgm = method, window, cbca, sgm
        self.invalid_disparity = float(invalid_disparity)
        self.offset = (window - 1) // 2
        self.is_max = method == "zncc"
        e = self.eng
        self.cv_a = e.empty((H, W, self.D))
        self.cv_b = e.empty((H, W, self.D)) if (cbca or sgm) else None
        self.disp = e.empty((H, W))
        self.flags = e.empty((H, W), torch.uint8)
        self.d_left = e.empty((H, W))
        self.d_right = e.empty((H, W))
        self.h_left = torch.empty((H, W), dtype=torch.float32, pin_memory=True)
        self.h_right = torch.empty((H, W), dtype=torch.float32, pin_memory=True)
        self.h_disp = torch.empty((H, W), dtype=torch.float32, pin_memory=True)
        self.final_cv = None
        if method == "census":
            self.cmax = float(window * window)
        elif method == "zncc":
            self.cmax = 1.0
        else:
            self.cmax = None                                   # data dependent: set per call

    def run_device(self, left, right):
        """``left`` / ``right``: float32 (H, W) device tensors.  Returns the disparity tensor (device)."""
        e = self.eng
        cur, other = self.cv_a, self.cv_b
        have_disp = False
        cmax = self.cmax
        if self.method == "census":
            fuse = not self.cbca and not self.sgm
            if fuse:
                _, self.disp, self.flags = self._census_fused(left, right, cur)
                have_disp = True
            else:
                e.census(left, right, self.window, self.dmin, self.dmax, out=cur)
        else:
            e.matching_cost(self.method, left, right, self.window, self.dmin, self.dmax, out=cur)
            if cmax is None:
                mx = self.torch.maximum((left.max() - right.min()).abs(), (right.max() - left.min()).abs()).item()
                cmax = float(int((mx if self.method == "sad" else mx * mx) * self.window**2))
        if self.cbca:
            dist, inten = self.cbca
            e.cbca(left, right, cur, self.offset, self.dmin, dist, inten, out=other)
            cur, other = other, cur
            cmax = cmax * ((2 * dist - 1) ** 2)
        if self.sgm:
            p1, p2 = float(self.sgm[0]), float(self.sgm[1])
            over = bool(self.sgm[2]) if len(self.sgm) > 2 else False
            src = -cur if self.is_max else cur
            e.sgm(src, p1, p2, cmax + p2 + 1.0, over, out=other, fuse_wta=not self.is_max, dmin=self.dmin,
                  invalid_disparity=self.invalid_disparity, disp=self.disp, flags=self.flags)
            cur, other = other, cur
            if self.is_max:
                cur.neg_()
            else:
                have_disp = True
        if not have_disp:
            self.disp, self.flags = e.wta(cur, self.dmin, self.is_max, self.invalid_disparity)
        self.final_cv = cur
        return self.disp

    def _census_fused(self, left, right, out):
        from . import _native  # noqa: PLC0415

        e = self.eng
        ws = e._workspace("census", e.lib.pb200_census_workspace_bytes(self.H, self.W, self.window))
        with self.torch.cuda.device(e.device):
            _native.check(e.lib.pb200_census_cost_volume(
                left.data_ptr(), right.data_ptr(), self.H, self.W, self.window, self.dmin, self.D, out.data_ptr(),
                ws.data_ptr(), ws.numel(), self.disp.data_ptr(), self.invalid_disparity, self.flags.data_ptr(), e._stream()))
        return out, self.disp, self.flags

    def run_host(self, left: np.ndarray, right: np.ndarray) -> np.ndarray:
        """Host images in, host disparity map out: pinned staging, H2D + kernels + D2H on one stream."""
        t = self.torch
        self.h_left.copy_(t.from_numpy(np.ascontiguousarray(left, dtype=np.float32)))
        self.h_right.copy_(t.from_numpy(np.ascontiguousarray(right, dtype=np.float32)))
        self.d_left.copy_(self.h_left, non_blocking=True)
        self.d_right.copy_(self.h_right, non_blocking=True)
        disp = self.run_device(self.d_left, self.d_right)
        self.h_disp.copy_(disp, non_blocking=True)
        t.cuda.current_stream(self.eng.device).synchronize()
        return self.h_disp.numpy()

    def validity_mask(self):
        """uint16 validity mask of the last run (criteria + WTA rules), as a device tensor (int16 storage)."""
        mask = self.eng.validity_mask(self.H, self.W, self.dmin, self.dmax, self.offset, self.flags)
        return self.eng.validity_mask(self.H, self.W, self.dmin, self.dmax, self.offset, self.flags, wta_invalidate=True, mask=mask)

    def bytes_h2d(self) -> int:
        return 2 * self.H * self.W * 4

    def bytes_d2h(self) -> int:
        return self.H * self.W * 4
