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

from . import _native
from ._common import MachineError, get_engine
from .aggregation import AbstractAggregation
from .disparity import AbstractDisparity
from .matching_cost import AbstractMatchingCost
from .optimization import AbstractOptimization

# PandoraMachine._transitions_run (state_machine.py:75-140): trigger -> (source state, destination state)
TRANSITIONS = {"matching_cost": ("begin", "cost_volume"), "aggregation": ("cost_volume", "cost_volume"),
               "optimization": ("cost_volume", "cost_volume"), "cost_volume_confidence": ("cost_volume", "cost_volume"),
               "disparity": ("cost_volume", "disp_map"), "filter": ("disp_map", "disp_map"), "refinement": ("disp_map", "disp_map"),
               "validation": ("disp_map", "disp_map")}

HOT_PATH_STEPS = ("matching_cost", "aggregation", "optimization", "disparity", "refinement", "filter", "validation",
                  "cost_volume_confidence")


def run(img_left, img_right, cfg: dict, return_right: bool = False):
    """Run the steps named in ``cfg["pipeline"]`` in order, like ``PandoraMachine`` does (state_machine.py:292-590);
    returns (left disparity dataset, left cost volume) -- plus the right disparity dataset with ``return_right``.

    ``cfg`` is a Pandora user configuration (``{"pipeline": {"matching_cost": {...}, ...}}``).  Implemented steps:
    matching_cost, aggregation, optimization, disparity, cost_volume_confidence, refinement, filter (3x3 median) and
    validation -- ``cross_checking_accurate`` (every step also runs on the right image with the roles swapped,
    state_machine.py:316-331) and ``cross_checking_fast`` (the right map comes from the left volume).  Anything else
    (multiscale, semantic_segmentation, disparity interpolation ...) is rejected: it stays Pandora's.
    """
    from .cost_volume_confidence import AbstractCostVolumeConfidence  # noqa: PLC0415
    from .criteria import validity_mask  # noqa: PLC0415
    from .dataset import DataArray, add_disparity  # noqa: PLC0415
    from .filter import AbstractFilter  # noqa: PLC0415
    from .refinement import AbstractRefinement  # noqa: PLC0415
    from .validation import AbstractValidation, right_disparity_fast  # noqa: PLC0415

    pipeline = cfg["pipeline"]
    disp_grids = (img_left["disparity"].data[0], img_left["disparity"].data[1])
    right_mode = pipeline.get("validation", {}).get("validation_method")          # state_machine.py:600-640
    if right_mode not in (None, "cross_checking_fast", "cross_checking_accurate"):
        raise NotImplementedError(f"validation method {right_mode!r} is outside the B200 hot path")
    if "interpolated_disparity" in pipeline.get("validation", {}):
        raise NotImplementedError("interpolated_disparity is outside the B200 hot path (use Pandora's own implementation)")
    accurate = right_mode == "cross_checking_accurate"
    right_grids = None
    if accurate:
        if "disparity" not in getattr(img_right, "data_vars", {}):                # state_machine.py:668-683
            img_right = img_right.copy(deep=False)
            const = AbstractMatchingCost.constant_range(img_left)
            if const is not None:
                # constant left range: reverse_disp_range only trims the right grids where the left column leaves the image, i.e.
                # where every cost is NaN already, and their extrema are (-max, -min) -- the constant right range is equivalent
                add_disparity(img_right, (-const[1], -const[0]))
            else:
                rmin, rmax = AbstractMatchingCost.reverse_disp_range(disp_grids[0], disp_grids[1])
                img_right["disparity"] = (("band_disp", "row", "col"), np.stack([rmin, rmax], axis=0))
                img_right.coords["band_disp"] = DataArray(np.array(["min", "max"]), ("band_disp",))
        right_grids = (img_right["disparity"].data[0], img_right["disparity"].data[1])
    cv = right_cv = None
    disp = right_disp = None
    state = "begin"
    for step, step_cfg in pipeline.items():
        name = step.split(".")[0]
        if name not in HOT_PATH_STEPS:
            raise NotImplementedError(f"step {step!r} is outside the B200 hot path (use Pandora's own implementation)")
        source, dest = TRANSITIONS[name]
        if state != source:                                       # the order PandoraMachine would refuse
            raise MachineError(f"Can't trigger event {name} from state {state}!")
        state = dest
        if name == "matching_cost":                               # state_machine.py:292-364
            mc = AbstractMatchingCost(**step_cfg)
            cv = mc.allocate_cost_volume(img_left, disp_grids, cfg)
            cv = validity_mask(img_left, img_right, cv)
            cv = mc.compute_cost_volume(img_left, img_right, cv)
            mc.cv_masked(img_left, img_right, cv, *disp_grids)
            if accurate:
                right_cv = mc.allocate_cost_volume(img_right, right_grids, cfg)
                right_cv = validity_mask(img_right, img_left, right_cv)
                right_cv = mc.compute_cost_volume(img_right, img_left, right_cv)
                mc.cv_masked(img_right, img_left, right_cv, *right_grids)
        elif name == "aggregation":                               # state_machine.py:366-380
            agg = AbstractAggregation(**step_cfg)
            agg.cost_volume_aggregation(img_left, img_right, cv)
            if accurate:
                agg.cost_volume_aggregation(img_right, img_left, right_cv)
        elif name == "optimization":                              # state_machine.py:404-419
            opt = AbstractOptimization(img_left, **step_cfg)
            cv = opt.optimize_cv(cv, img_left, img_right)
            if accurate:
                right_cv = opt.optimize_cv(right_cv, img_right, img_left)
        elif name == "disparity":                                 # state_machine.py:421-448
            disparity_ = AbstractDisparity(**step_cfg)
            disp = disparity_.to_disp(cv, img_left, img_right)
            if accurate:
                right_disp = disparity_.to_disp(right_cv, img_right, img_left)
            elif right_mode == "cross_checking_fast":
                right_disp = right_disparity_fast(cv, disparity_.cfg["invalid_disparity"])
        elif name == "cost_volume_confidence":                    # state_machine.py:566-587
            step_cfg = dict(step_cfg)
            if len(step.split(".")) == 2:
                step_cfg["indicator"] = "." + step.split(".")[1]
            confidence_ = AbstractCostVolumeConfidence(**step_cfg)
            disp, cv = confidence_.confidence_prediction(disp, img_left, img_right, cv)
            if accurate:
                right_disp, right_cv = confidence_.confidence_prediction(right_disp, img_right, img_left, right_cv)
        elif name == "refinement":                                # state_machine.py:474-490
            refinement_ = AbstractRefinement(**step_cfg)
            refinement_.subpixel_refinement(cv, disp)
            if accurate:
                refinement_.subpixel_refinement(right_cv, right_disp)
            elif right_disp is not None:
                refinement_.right_subpixel_refinement(cv, right_disp)
        elif name == "filter":                                    # state_machine.py:450-473
            filter_ = AbstractFilter(dict(step_cfg))
            filter_.filter_disparity(disp, img_left)
            if right_disp is not None:
                filter_.filter_disparity(right_disp, img_right)
        elif name == "validation":                                # state_machine.py:492-519
            validation_ = AbstractValidation(**step_cfg)
            disp = validation_.disparity_checking(disp, right_disp, img_left, img_right, cv)
            right_disp = validation_.disparity_checking(right_disp, disp, img_right, img_left, right_cv)
            if right_mode == "cross_checking_fast":
                right_disp = None                                 # state_machine.py:514-519
    return (disp, cv, right_disp) if return_right else (disp, cv)


class StereoPipeline:
    """Fixed-configuration device pipeline: matching cost -> [CBCA] -> [SGM] -> WTA, buffers allocated once.

    ``method`` in {"census", "sad", "ssd", "zncc"}; ``cbca`` = None or (distance, intensity);
    ``sgm`` = None or (P1, P2[, overcounting]).  WTA is fused into the producing kernel when the last
    volume step is Census or SGM.
    """

    def __init__(self, H: int, W: int, dmin: int, dmax: int, method: str = "census", window: int = 5,
                 cbca: Optional[Tuple[int, float]] = None, sgm: Optional[Tuple] = None, invalid_disparity: float = -9999.0,
                 device: Optional[str] = None, fuse_census_sgm: Optional[bool] = None):
        import torch  # noqa: PLC0415

        # Census directly followed by SGM: one fused stage (the Census volume is never written), when the shape is
        # eligible -- see Engine.census_sgm.  the library option "fuse_census_sgm" = 0 / fuse_census_sgm=False keeps the two steps apart.
        if fuse_census_sgm is None:
            fuse_census_sgm = _native.get_option("fuse_census_sgm") != 0
        self.fuse_census_sgm = bool(fuse_census_sgm) and method == "census" and sgm is not None and not cbca
        self.fused_ran = False

        self.torch = torch
        self.eng = get_engine(device)
        self.H, self.W, self.dmin, self.dmax = H, W, dmin, dmax
        self.D = dmax - dmin + 1
        if sgm is not None and self.D > _native.SGM_MAX_DISP:   # refuse before allocating the volumes
            raise ValueError(f"sgm: {self.D} disparities exceed the kernels' maximum of {_native.SGM_MAX_DISP}")
        self.method, self.window, self.cbca, self.sgm = method, window, cbca, sgm
        self.invalid_disparity = float(invalid_disparity)
        self.offset = (window - 1) // 2
        self.is_max = method == "zncc"
        e = self.eng
        self._cv_a = None if self.fuse_census_sgm else e.empty((H, W, self.D))     # fused runs only need the SGM volume
        self.cv_b = e.empty((H, W, self.D)) if (cbca or sgm) else None
        self.disp = e.empty((H, W))
        self.flags = e.empty((H, W), torch.uint8)
        self.d_left = e.empty((H, W))
        self.d_right = e.empty((H, W))
        self.h_left = torch.empty((H, W), dtype=torch.float32, pin_memory=True)
        self.h_right = torch.empty((H, W), dtype=torch.float32, pin_memory=True)
        self.h_disp = torch.empty((H, W), dtype=torch.float32, pin_memory=True)
        self.final_cv = None
        self._copy_stream = None
        self._slots, self._submitted, self._d2h_stream = None, 0, None
        if method == "census":
            self.cmax = float(window * window)
        elif method == "zncc":
            self.cmax = 1.0
        else:
            self.cmax = None                                   # data dependent: set per call

    @property
    def cv_a(self):
        if self._cv_a is None:
            self._cv_a = self.eng.empty((self.H, self.W, self.D))
        return self._cv_a

    def _fused(self, left, right, descriptors_ready=False) -> bool:
        """Census -> SGM -> WTA as one fused stage; False when the shape is not eligible (nothing computed)."""
        p1, p2 = float(self.sgm[0]), float(self.sgm[1])
        over = bool(self.sgm[2]) if len(self.sgm) > 2 else False
        out = self.eng.census_sgm(left, right, self.window, self.dmin, self.dmax, p1, p2, over, out=self.cv_b, fuse_wta=True,
                                  invalid_disparity=self.invalid_disparity, disp=self.disp, flags=self.flags,
                                  descriptors_ready=descriptors_ready)
        self.fused_ran = out is not None
        if out is not None:
            self.final_cv = self.cv_b
        return out is not None

    def run_device_batch(self, left, right):
        """A batch of pairs, ``left`` / ``right``: float32 (n, H, W) device tensors, as ONE wave per pass of the fused
        Census -> SGM -> WTA stage (stream throughput: the time the wave needs to cross the SMs is paid once per batch).
        Returns the (n, H, W) disparity tensor; ``final_cv`` is the (n, H, W, D) SGM volume.  Same results as ``run_device``
        per pair, bit for bit; falls back to exactly that when the configuration is not eligible for the batched stage."""
        n = int(left.shape[0])
        bufs = getattr(self, "_batch_bufs", None)
        if bufs is None or bufs[0] < n:                        # one set of buffers for the largest batch seen; smaller batches use a prefix
            e = self.eng
            self._batch_bufs = bufs = None                     # release the smaller set before allocating the larger one
            bufs = (n, e.empty((n, self.H, self.W, self.D)) if self.sgm else None, e.empty((n, self.H, self.W)),
                    e.empty((n, self.H, self.W), self.torch.uint8))
            self._batch_bufs = bufs
        cv = bufs[1][:n] if bufs[1] is not None else None
        disp, flags = bufs[2][:n], bufs[3][:n]
        if self.fuse_census_sgm and n > 1:
            p1, p2 = float(self.sgm[0]), float(self.sgm[1])
            over = bool(self.sgm[2]) if len(self.sgm) > 2 else False
            out = self.eng.census_sgm_batch(left, right, self.window, self.dmin, self.dmax, p1, p2, over, out=cv,
                                            invalid_disparity=self.invalid_disparity, disp=disp, flags=flags)
            if out is not None:
                self.fused_ran, self.batched_ran = True, True
                self.final_cv, self.disp_batch, self.flags_batch = cv, disp, flags
                return disp
        self.batched_ran = False
        for i in range(n):                                     # not eligible: one pair at a time
            disp[i].copy_(self.run_device(left[i], right[i]))
            flags[i].copy_(self.flags)
            if cv is not None:
                cv[i].copy_(self.final_cv)
        self.final_cv, self.disp_batch, self.flags_batch = cv, disp, flags
        return disp

    def run_device(self, left, right):
        """``left`` / ``right``: float32 (H, W) device tensors.  Returns the disparity tensor (device)."""
        if self.fuse_census_sgm and self._fused(left, right):
            return self.disp
        have_disp, cmax = self._matching_cost(left, right)
        return self._after_cost(left, right, have_disp, cmax)

    def _matching_cost(self, left, right, rows=None):
        """Matching-cost step into ``cv_a`` (optionally one band of rows, Census only); returns (have_disp, cmax)."""
        e = self.eng
        cur = self.cv_a
        cmax = self.cmax
        if self.method == "census":
            fuse = not self.cbca and not self.sgm
            e.census(left, right, self.window, self.dmin, self.dmax, out=cur, fuse_wta=fuse, invalid_disparity=self.invalid_disparity,
                     rows=rows, disp=self.disp, flags=self.flags)
            return fuse, cmax
        e.matching_cost(self.method, left, right, self.window, self.dmin, self.dmax, out=cur)
        if cmax is None:
            if not self.sgm:
                return False, 0.0                              # SAD / SSD: the data-dependent cmax (a host round trip) only feeds SGM's invalid value
            mx = self.torch.maximum((left.max() - right.min()).abs(), (right.max() - left.min()).abs()).item()
            cmax = float(int((mx if self.method == "sad" else mx * mx) * self.window**2))
        return False, cmax

    def _after_cost(self, left, right, have_disp, cmax):
        e = self.eng
        cur, other = self.cv_a, self.cv_b
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

    def run_host(self, left: np.ndarray, right: np.ndarray) -> np.ndarray:
        """Host images in, host disparity map out (the end-to-end call).  Images that are not already in pinned
        memory are staged through the pipeline's pinned buffers.  For Census the upload is cut into row bands on a
        copy stream and the cost-volume fill of a band starts as soon as its rows (plus half a window) have
        arrived, so most of the host-to-device copy hides behind the fill."""
        t = self.torch
        dev = self.eng.device
        hl, hr = self._pinned(left, self.h_left), self._pinned(right, self.h_right)
        cur = t.cuda.current_stream(dev)
        bands = max(1, min(8, self.H // 256)) if self.method == "census" else 1
        if self.fuse_census_sgm:
            # the fused stage needs every descriptor: upload, transform, run (the transform of a band could overlap the
            # upload of the next one, but it is 0.2 ms of a 2.4 ms copy)
            self.d_left.copy_(hl, non_blocking=True)
            self.d_right.copy_(hr, non_blocking=True)
            if self._fused(self.d_left, self.d_right):
                self.h_disp.copy_(self.disp, non_blocking=True)
                cur.synchronize()
                return self.h_disp.numpy()
            have_disp, cmax = self._matching_cost(self.d_left, self.d_right)
        elif bands == 1:
            self.d_left.copy_(hl, non_blocking=True)
            self.d_right.copy_(hr, non_blocking=True)
            have_disp, cmax = self._matching_cost(self.d_left, self.d_right)
        else:
            if self._copy_stream is None:
                self._copy_stream = t.cuda.Stream(device=dev)
            cs = self._copy_stream
            cs.wait_stream(cur)                                   # the previous step may still read d_left / d_right
            edges = [self.H * b // bands for b in range(bands + 1)]
            copied = 0
            have_disp, cmax = False, self.cmax
            for b in range(bands):
                r0, r1 = edges[b], edges[b + 1]
                upto = self.H if b == bands - 1 else min(self.H, r1 + self.offset)
                with t.cuda.stream(cs):
                    self.d_left[copied:upto].copy_(hl[copied:upto], non_blocking=True)
                    self.d_right[copied:upto].copy_(hr[copied:upto], non_blocking=True)
                    ev = t.cuda.Event()
                    ev.record(cs)
                copied = upto
                cur.wait_event(ev)
                have_disp, cmax = self._matching_cost(self.d_left, self.d_right, rows=(r0, r1))
        disp = self._after_cost(self.d_left, self.d_right, have_disp, cmax)
        self.h_disp.copy_(disp, non_blocking=True)
        cur.synchronize()
        return self.h_disp.numpy()

    # ---- streaming entry: a sequence of stereo pairs, copies overlapped with the kernels ---------------------------
    def submit_host(self, left, right) -> int:
        """Asynchronous end-to-end step for a stream of stereo pairs: enqueue upload -> pipeline -> download of one pair
        and return a ticket for ``result_host``.  Two buffer sets alternate, so the upload of pair k + 1 (copy stream)
        and the download of pair k - 1 (a second copy stream) overlap the kernels of pair k; at most two pairs may be in
        flight (collect ticket k - 1 before submitting pair k + 1).  ``final_cv`` belongs to the pair submitted last."""
        t = self.torch
        dev = self.eng.device
        if self._slots is None:
            with t.cuda.device(dev):
                self._slots = [dict(d_left=self.eng.empty((self.H, self.W)), d_right=self.eng.empty((self.H, self.W)),
                                    d_disp=self.eng.empty((self.H, self.W)),
                                    h_left=t.empty((self.H, self.W), dtype=t.float32, pin_memory=True),
                                    h_right=t.empty((self.H, self.W), dtype=t.float32, pin_memory=True),
                                    h_disp=t.empty((self.H, self.W), dtype=t.float32, pin_memory=True),
                                    ev_h2d=None, ev_compute=None, ev_d2h=None) for _ in range(2)]
                self._copy_stream = self._copy_stream or t.cuda.Stream(device=dev)
                self._d2h_stream = t.cuda.Stream(device=dev)
        ticket = self._submitted
        self._submitted += 1
        s = self._slots[ticket % 2]
        cur = t.cuda.current_stream(dev)
        if s["ev_h2d"] is not None:
            s["ev_h2d"].synchronize()                             # the slot's staging buffers are free again (pageable inputs)
        hl, hr = self._pinned(left, s["h_left"]), self._pinned(right, s["h_right"])
        cs = self._copy_stream
        with t.cuda.stream(cs):
            if s["ev_compute"] is not None:
                cs.wait_event(s["ev_compute"])                    # pair k - 2 has read the slot's device images
            else:
                cs.wait_stream(cur)
            s["d_left"].copy_(hl, non_blocking=True)
            s["d_right"].copy_(hr, non_blocking=True)
            s["ev_h2d"] = t.cuda.Event()
            s["ev_h2d"].record(cs)
        cur.wait_event(s["ev_h2d"])
        if s["ev_d2h"] is not None:
            cur.wait_event(s["ev_d2h"])                           # the download of pair k - 2 has read the slot's disparity map
        disp = self.run_device(s["d_left"], s["d_right"])
        s["d_disp"].copy_(disp, non_blocking=True)
        s["ev_compute"] = t.cuda.Event()
        s["ev_compute"].record(cur)
        ds = self._d2h_stream
        with t.cuda.stream(ds):
            ds.wait_event(s["ev_compute"])
            s["h_disp"].copy_(s["d_disp"], non_blocking=True)
            s["ev_d2h"] = t.cuda.Event()
            s["ev_d2h"].record(ds)
        return ticket

    def submit_host_batch(self, lefts, rights) -> int:
        """``submit_host`` for a batch of pairs that goes through ONE wave per pass (``run_device_batch``): ``lefts`` / ``rights``
        are sequences of n host images (pinned tensors are used in place).  Uploads and downloads of neighbouring batches
        overlap the kernels of the current one; at most two batches in flight.  Returns a ticket for ``result_host_batch``."""
        t = self.torch
        dev = self.eng.device
        n = len(lefts)
        if getattr(self, "_bslots", None) is None:
            self._bslots, self._bsubmitted, self._btickets = {}, 0, {}
        if n not in self._bslots:                                 # one pair of buffer sets per batch size, allocated once
            with t.cuda.device(dev):
                mk = lambda pin: (t.empty((n, self.H, self.W), dtype=t.float32, pin_memory=True) if pin  # noqa: E731
                                  else self.eng.empty((n, self.H, self.W)))
                self._bslots[n] = [dict(d_left=mk(False), d_right=mk(False), d_disp=mk(False), h_left=mk(True), h_right=mk(True),
                                        h_disp=mk(True), ev_h2d=None, ev_compute=None, ev_d2h=None, uses=0) for _ in range(2)]
                self._copy_stream = self._copy_stream or t.cuda.Stream(device=dev)
                self._d2h_stream = self._d2h_stream or t.cuda.Stream(device=dev)
        ticket = self._bsubmitted
        self._bsubmitted += 1
        slots = self._bslots[n]
        s = slots[0] if slots[0]["uses"] <= slots[1]["uses"] else slots[1]      # the set of this size used longest ago
        s["uses"] = max(slots[0]["uses"], slots[1]["uses"]) + 1
        self._btickets[ticket] = s
        self._btickets.pop(ticket - 2, None)
        cur = t.cuda.current_stream(dev)
        if s["ev_h2d"] is not None:
            s["ev_h2d"].synchronize()
        srcs = [(self._pinned(lefts[i], s["h_left"][i]), self._pinned(rights[i], s["h_right"][i])) for i in range(n)]
        cs = self._copy_stream
        with t.cuda.stream(cs):
            if s["ev_compute"] is not None:
                cs.wait_event(s["ev_compute"])
            else:
                cs.wait_stream(cur)
            for i, (hl, hr) in enumerate(srcs):
                s["d_left"][i].copy_(hl, non_blocking=True)
                s["d_right"][i].copy_(hr, non_blocking=True)
            s["ev_h2d"] = t.cuda.Event()
            s["ev_h2d"].record(cs)
        cur.wait_event(s["ev_h2d"])
        if s["ev_d2h"] is not None:
            cur.wait_event(s["ev_d2h"])
        disp = self.run_device_batch(s["d_left"], s["d_right"])
        s["d_disp"].copy_(disp, non_blocking=True)
        s["ev_compute"] = t.cuda.Event()
        s["ev_compute"].record(cur)
        ds = self._d2h_stream
        with t.cuda.stream(ds):
            ds.wait_event(s["ev_compute"])
            s["h_disp"].copy_(s["d_disp"], non_blocking=True)
            s["ev_d2h"] = t.cuda.Event()
            s["ev_d2h"].record(ds)
        return ticket

    def result_host_batch(self, ticket: int) -> np.ndarray:
        """(n, H, W) disparity maps of a submitted batch (host array, valid until two more batches have been submitted)."""
        if ticket not in getattr(self, "_btickets", {}):
            raise ValueError(f"batch ticket {ticket} is not in flight")
        s = self._btickets[ticket]
        s["ev_d2h"].synchronize()
        return s["h_disp"].numpy()

    def result_host(self, ticket: int) -> np.ndarray:
        """Disparity map of a submitted pair (host array, valid until two more pairs have been submitted)."""
        if not (self._submitted - 2 <= ticket < self._submitted):
            raise ValueError(f"ticket {ticket} is not in flight (submitted so far: {self._submitted})")
        s = self._slots[ticket % 2]
        s["ev_d2h"].synchronize()
        return s["h_disp"].numpy()

    def _pinned(self, arr, staging):
        """``arr`` as a pinned float32 host tensor: itself when it already is one, else a copy into ``staging``."""
        t = self.torch
        src = arr if isinstance(arr, t.Tensor) else t.from_numpy(np.ascontiguousarray(arr, dtype=np.float32))
        if src.dtype == t.float32 and src.is_contiguous() and src.is_pinned():
            return src
        staging.copy_(src)
        return staging

    def validity_mask(self):
        """uint16 validity mask of the last run (criteria + WTA rules), as a device tensor (int16 storage)."""
        mask = self.eng.validity_mask(self.H, self.W, self.dmin, self.dmax, self.offset, self.flags)
        return self.eng.validity_mask(self.H, self.W, self.dmin, self.dmax, self.offset, self.flags, wta_invalidate=True, mask=mask)

    def bytes_h2d(self) -> int:
        return 2 * self.H * self.W * 4

    def bytes_d2h(self) -> int:
        return self.H * self.W * 4
