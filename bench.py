#!/usr/bin/env python3
"""bench.py -- headline benchmark of the Pandora dense cost-volume hot path on B200.

Metric (BASELINE.json): disparity Mpix/s at D=256 and cost-volume HBM GB/s vs roofline.
Workload at N=1: C3 = BASELINE.json configs[3], a synthetic 4096x4096 pair, Census 5x5 -> SGM 8-path
(P1=8, P2=32) -> WTA over D=256 disparities ([-255, 0]).  At N>1 the image grows by 4096 rows per GPU
(weak scaling): every rank owns a 4096x4096 row tile, SGM path states cross tile borders over NCCL
(pandora_b200/tiling.py).

A "step" is one pass of the whole pipeline over one stereo pair.  `value` times it with the images
resident in HBM (CUDA events, K steps, max over ranks); `e2e` times the same step through the host
API (pinned host images -> H2D -> kernels -> D2H of the disparity map).  `--impl reference` times the
reference's own CPU code path (oracle/_ref census C++ + the oracle's SGM/WTA port, 1 thread: the
reference is single-threaded) on a bounded row band of the same workload.

One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H_TILE, W_IMG, D_DISP, WINDOW, P1, P2 = 4096, 4096, 256, 5, 8.0, 32.0
METRIC = "disparity Mpix/s at D=256 (Census 5x5 + SGM 8-path + WTA)"
UNIT = "Mpix/s"


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int = 0):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for n, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU code on a bounded row band of the same workload
# ----------------------------------------------------------------------------------------------------
def cpu_pipeline_once(left, right, dmin, dmax):
    """Census (unmodified reference C++ from oracle/_ref when present, else the oracle port) -> SGM (oracle
    port of the libSGM step) -> WTA (C port of np.argmin).  Returns (seconds, kind)."""
    from oracle import oracle as orc

    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    kind = "port"
    mc = None
    try:
        if ref_dir not in sys.path:
            sys.path.insert(0, ref_dir)
        import matching_cost_cpp as mc  # noqa: PLC0415

        kind = "reference"
    except Exception:
        mc = None
    D = dmax - dmin + 1
    t0 = time.perf_counter()
    if mc is not None:
        cv = np.full(left.shape + (D,), np.nan, dtype=np.float32)                 # census.py:138
        cv = mc.compute_matching_costs(left, [right], cv, np.arange(dmin, dmax + 1).astype(np.float32), WINDOW, WINDOW)
    else:
        cv, _ = orc.census_cost_volume(left, right, WINDOW, dmin, dmax)
    S = orc.sgm_cost_volume(cv, P1, P2, cmax=WINDOW * WINDOW)
    orc.wta_c(S, np.arange(dmin, dmax + 1))
    return time.perf_counter() - t0, kind


def cpu_sample(rows: int):
    from pandora_b200.synthetic import synthetic_pair

    left, right, _ = synthetic_pair(rows, W_IMG, D_DISP)
    return np.ascontiguousarray(left), np.ascontiguousarray(right)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    rows = 32
    left, right = cpu_sample(rows)
    dmin, dmax = -(D_DISP - 1), 0
    kind = "port"
    for _ in range(min(args.warmup, 1)):
        cpu_pipeline_once(left, right, dmin, dmax)
    times = []
    for _ in range(args.steps):
        t, kind = cpu_pipeline_once(left, right, dmin, dmax)
        times.append(t)
    mean_t = float(np.mean(times))
    value = rows * W_IMG / mean_t / 1e6
    sample = (f"{rows} rows x {W_IMG} cols x D={D_DISP} band of the C3 pair per step; census = "
              f"{'unmodified reference C++ (oracle/_ref)' if kind == 'reference' else 'oracle port'}, SGM + WTA = oracle C port "
              "(libSGM is not vendored); single thread like the reference")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": mean_t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C3: 4096x4096 synthetic pair, Census 5x5 + SGM 8-path P1=8 P2=32 + WTA, D=256 (CPU: row band sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                         "hypothetical_row_parallel": {"cores": os.cpu_count(), "value": value * (os.cpu_count() or 1),
                                                       "note": "1-core figure x host cores (Census / WTA are row-parallel, SGM is not): an upper "
                                                               "bound, the reference itself is single-threaded"}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cores_available": os.cpu_count(),
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
CONFIGS = {
    # name: (H, W of ONE tile / of the whole image, scaling).  C3 = BASELINE.json configs[3] (weak scaling: one 4096-column tile
    # per GPU of a 4096-row image); C4 = configs[4] read as W = 16384, H = 4096, C4T = the transposed reading: both a
    # FIXED image column-tiled over the ranks (strong scaling).
    "C3": dict(H=4096, W=4096, scaling="weak"),
    "C4": dict(H=4096, W=16384, scaling="strong"),
    "C4T": dict(H=16384, W=4096, scaling="strong"),
}
SYN_TILE = 4096          # the global synthetic image is a row of independent 4096-column synthetic pairs (seed + tile index)


def global_columns(H, Wg, lo, n, D):
    """Columns [lo, lo + n) (cyclic) of the global synthetic pair as {tile index: (left, right)} of the 4096-column tiles they touch."""
    from pandora_b200.synthetic import synthetic_pair

    tiles = {}
    ntiles = (Wg + SYN_TILE - 1) // SYN_TILE
    c = lo
    while c < lo + n:
        t = (c % Wg) // SYN_TILE
        if t not in tiles:
            w = min(SYN_TILE, Wg - t * SYN_TILE)
            left, right, _ = synthetic_pair(H, w, D, seed=20240607 + t)
            tiles[t] = (np.ascontiguousarray(left), np.ascontiguousarray(right))
        c = (c // SYN_TILE + 1) * SYN_TILE
    assert len(tiles) <= ntiles
    return tiles


def other_config_lines(pb, torch, steps):
    """C0 (cones size, SAD 5x5 -> WTA, D = 64), C1 (1024^2 Census 5x5 -> WTA, D = 128) and C2 (2048^2 Census 5x5 -> CBCA -> WTA, D = 192): device-resident pipeline
    times of BASELINE.json's other single-GPU configurations, measured in this run (CUDA events), with the algorithmic bytes of
    SURVEY 8(d) next to them."""
    from pandora_b200.synthetic import synthetic_pair

    out = {}
    peak, _ = measured_peak_gbs()
    for name, (H, W, D, method, cbca, alg) in {"C0": (375, 450, 64, "sad", None, lambda D: 8.0 * D + 14.0),      # SAD fill + stand-alone WTA
                                              "C1": (1024, 1024, 128, "census", None, lambda D: 4.0 * D + 12.0),
                                              "C2": (2048, 2048, 192, "census", (5, 30.0), lambda D: 12.0 * D + 36.0)}.items():
        left, right, _ = synthetic_pair(H, W, D)
        pipe = pb.StereoPipeline(H, W, -(D - 1), 0, method, WINDOW, cbca=cbca, device="cuda:0")
        dl, dr = pipe.eng.to_device(left), pipe.eng.to_device(right)
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda:0")      # > L2: the volumes of C1 nearly fit it
        for _ in range(3):
            pipe.run_device(dl, dr)
        ts = []
        for _ in range(max(steps, 5)):
            flush.fill_(1)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            pipe.run_device(dl, dr)
            ev[1].record()
            torch.cuda.synchronize()
            ts.append(ev[0].elapsed_time(ev[1]))
        ms = float(np.mean(ts))
        bytes_ = alg(D) * H * W
        out[name] = {"workload": f"{H}x{W} {'SAD' if method == 'sad' else 'Census'} 5x5" + (" + CBCA" if cbca else "") + f" + WTA, D={D}"
                                 + (" (cones-sized synthetic pair)" if method == "sad" else ""),
                     "ms_per_step": ms,
                     "mpix_per_s": H * W / ms / 1e3, "pipeline_algorithmic_bytes": bytes_, "achieved_gbs": bytes_ / ms / 1e6,
                     "frac_of_measured_hbm": bytes_ / ms / 1e6 / peak, "l2": "256 MB flush between iterations",
                     "paths": {"census": pb.last_path("census")[0] if method == "census" else None, "sad": pb.last_path("sad")[0] if method == "sad" else None,
                               "cbca": pb.last_path("cbca")[0] if cbca else None}}
        del pipe, dl, dr, flush
        torch.cuda.empty_cache()
    return out


def plugin_leg(pb, torch, left, right, dmin, dmax, steps):
    """C3 through the plugin-level call: host numpy datasets in, `pandora_b200.run(cfg)` (the step classes behind the reference's
    plugin API: Census.compute_cost_volume -> cv_masked -> Sgm.optimize_cv -> WinnerTakesAll.to_disp), host disparity map and
    validity mask out.  Wall clock around the call, everything inside (pageable H2D of the images, kernels, D2H of the maps)."""
    cfg = {"pipeline": {"matching_cost": {"matching_cost_method": "census", "window_size": WINDOW, "subpix": 1},
                        "optimization": {"optimization_method": "sgm", "penalty": {"P1": P1, "P2": P2}},
                        "disparity": {"disparity_method": "wta", "invalid_disparity": -9999}}}
    dl = pb.create_image_dataset(left, disparity=[dmin, dmax])
    dr = pb.create_image_dataset(right)

    def once():
        disp, _cv = pb.run(dl, dr, cfg)
        d = np.asarray(disp["disparity_map"].data)
        m = np.asarray(disp["validity_mask"].data)
        return d, m

    ref, _ = once()
    for _ in range(2):                      # steady state of the result arrays: the caching host allocator holds the three blocks per
        d, _m = once()                      # map a caller that keeps one result and the latest one needs (each is page-locked once)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        d, _m = once()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    H, W = left.shape
    return {"value": H * W / ms / 1e3, "unit": UNIT, "ms_per_step": ms, "h2d_bytes_per_step": 2 * H * W * 4, "d2h_bytes_per_step": H * W * 6,
            "call": "pandora_b200.run(img_left, img_right, cfg): host numpy datasets in (pageable H2D), host disparity_map + validity_mask out "
                    "(page-locked result arrays from the caching host allocator)",
            "sgm_path": pb.last_path("sgm")[0], "same_map_every_call": bool(np.array_equal(ref, d))}, ref


def run_ours(args):
    import torch

    import pandora_b200
    from pandora_b200.synthetic import synthetic_pair

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    dist = None
    if world > 1:
        import torch.distributed as dist  # noqa: PLC0415

        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")    # NCCL's version banner goes to stdout otherwise: keep the JSON line alone there
        dist.init_process_group("nccl", device_id=torch.device(device))
    cfg = CONFIGS[args.config]
    dmin, dmax = -(D_DISP - 1), 0
    D = D_DISP
    H = cfg["H"]
    Wg = cfg["W"] * world if cfg["scaling"] == "weak" else cfg["W"]       # global image width
    Wt = Wg // world                                                       # columns per GPU
    steps = args.steps

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, k, batched=False):
        sync_all()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        if batched:
            fn(k)
        else:
            for _ in range(k):
                fn()
        ev[1].record()
        sync_all()
        ms = ev[0].elapsed_time(ev[1])
        if dist is not None:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    parity = None
    if world == 1:
        left, right, _ = synthetic_pair(H, Wg, D) if Wg <= SYN_TILE else (None, None, None)
        if left is None:
            tiles = global_columns(H, Wg, 0, Wg, D)
            left = np.concatenate([tiles[t][0] for t in sorted(tiles)], axis=1)
            right = np.concatenate([tiles[t][1] for t in sorted(tiles)], axis=1)
        pipe = pandora_b200.StereoPipeline(H, Wg, dmin, dmax, "census", WINDOW, sgm=(P1, P2), device=device)
        d_left, d_right = pipe.eng.to_device(left), pipe.eng.to_device(right)
        h_left = torch.from_numpy(left).pin_memory()
        h_right = torch.from_numpy(right).pin_memory()

        def step_device():
            return pipe.run_device(d_left, d_right)

        # stream throughput, like at N > 1: the timed steps run as batches of up to BATCH pairs that follow each other in ONE wave
        # per pass (pb200_census_sgm_batch: the wave's fill and drain across the 148 SMs are paid once per batch); every pair
        # has its own images, its own 17 GB SGM volume and its own disparity map.  `one_pair_per_launch` is the same step with
        # one pair per call.
        BATCH = 4 if Wg <= 4144 else 1
        batch_in = {}

        def steps_device(k):
            done = 0
            while done < k:
                m = min(BATCH, k - done)
                if m == 1:
                    pipe.run_device(d_left, d_right)
                else:
                    if m not in batch_in:
                        batch_in[m] = (d_left.unsqueeze(0).expand(m, -1, -1).contiguous(), d_right.unsqueeze(0).expand(m, -1, -1).contiguous())
                    pipe.run_device_batch(*batch_in[m])
                done += m

        def step_host():
            return pipe.run_host(h_left, h_right)

        h2d_bytes, d2h_bytes = 2 * H * Wg * 4, H * Wg * 4
    else:
        from pandora_b200.tiling import ColumnTiledStereoPipeline  # noqa: PLC0415

        # ---- parity first: a small image column-tiled over the same ranks must equal its one-GPU run bit for bit ------------
        hs, ws, ds = 72, 96 * world, 64
        sl, sr, _ = synthetic_pair(hs, ws, ds)
        small = ColumnTiledStereoPipeline(hs, ws, -(ds - 1), 0, rank, world, dist, WINDOW, P1, P2, device=device)
        e = small.eng
        dsl, dsr = e.to_device(sl), e.to_device(sr)
        ok = True
        for _ in range(2):
            small.run(dsl, dsr)
            tile = small.unshear()
            whole = e.census_sgm(dsl, dsr, WINDOW, -(ds - 1), 0, P1, P2)
            ok = ok and whole is not None and bool(torch.equal(tile, whole[1][:, rank * (ws // world):(rank + 1) * (ws // world)]))
            vol = small.unshear(small.cv)
            ok = ok and bool(torch.equal(torch.nan_to_num(vol, nan=-7.0),
                                         torch.nan_to_num(whole[0][:, rank * (ws // world):(rank + 1) * (ws // world)], nan=-7.0)))
        flag = torch.tensor([1 if ok else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity = {"checked": True, "ok": bool(flag.item()), "case": f"{hs}x{ws}x{ds} column-tiled over {world} ranks vs the one-GPU run of the "
                  "same image (disparity map and SGM volume, two images back to back), checked on every rank before timing"}
        small.close()
        if not parity["ok"]:
            raise SystemExit(f"bench.py: column-tiled result differs from the one-GPU result on rank {rank}")

        pipe = ColumnTiledStereoPipeline(H, Wg, dmin, dmax, rank, world, dist, WINDOW, P1, P2, device=device)
        lo, n = pipe.visited_columns(min(6, max(1, 65534 // H), steps))
        lo_img, n_img = (lo - D - 8) % Wg, min(Wg, n + D + 16)          # + the right-image windows
        tiles = global_columns(H, Wg, lo_img, n_img, D)
        d_left, d_right = pipe.eng.empty((H, Wg)).zero_(), pipe.eng.empty((H, Wg)).zero_()
        h_left = torch.zeros((H, Wg), dtype=torch.float32).pin_memory()
        h_right = torch.zeros((H, Wg), dtype=torch.float32).pin_memory()
        for t, (tl, tr) in tiles.items():
            h_left[:, t * SYN_TILE:t * SYN_TILE + tl.shape[1]] = torch.from_numpy(tl)
            h_right[:, t * SYN_TILE:t * SYN_TILE + tr.shape[1]] = torch.from_numpy(tr)
        # column ranges (not cyclic) this rank uploads per step
        ranges = [(lo_img, min(Wg, lo_img + n_img))] + ([(0, lo_img + n_img - Wg)] if lo_img + n_img > Wg else [])
        for a, b in ranges:
            d_left[:, a:b].copy_(h_left[:, a:b])
            d_right[:, a:b].copy_(h_right[:, a:b])
        h_disp = torch.empty((H, Wt), dtype=torch.float32).pin_memory()

        def step_device():
            pipe.run(d_left, d_right)
            return pipe.unshear()

        BATCH = min(6, max(1, 65534 // H))                    # images that follow each other in one wave (one 17 GB volume each)
        batch_in = {}

        def steps_device(k):
            """k images as a stream: batches of up to BATCH images that follow each other in ONE wave (the first row of image
            i + 1 enters the pipeline behind the last row of image i), every image's disparity tile brought back to image layout."""
            done = 0
            while done < k:
                m = min(BATCH, k - done)
                if m not in batch_in:
                    batch_in[m] = (d_left.unsqueeze(0).expand(m, -1, -1).contiguous(), d_right.unsqueeze(0).expand(m, -1, -1).contiguous())
                pipe.run(*batch_in[m])
                pipe.unshear()
                done += m

        # the host side of a step (ONE image per call): the image columns this rank's sheared tile visits for one image -- its
        # own Wt columns plus the H - 1 columns the shear drifts over, plus the right-image windows -- as CONTIGUOUS pinned
        # arrays (a strided pinned -> device copy takes a slow path), copied to a device staging buffer and from there into place
        lo1, n1 = pipe.visited_columns(1)
        lo1_img, n1_img = (lo1 - D - 8) % Wg, min(Wg, n1 + D + 16)
        ranges1 = [(lo1_img, min(Wg, lo1_img + n1_img))] + ([(0, lo1_img + n1_img - Wg)] if lo1_img + n1_img > Wg else [])
        h_parts = [(torch.from_numpy(np.ascontiguousarray(h_left[:, a:b].numpy())).pin_memory(),
                    torch.from_numpy(np.ascontiguousarray(h_right[:, a:b].numpy())).pin_memory()) for a, b in ranges1]
        d_parts = [(torch.empty_like(hl, device=device), torch.empty_like(hr, device=device)) for hl, hr in h_parts]
        host_ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

        def step_host():
            host_ev[0].record()
            for (a, b), (hl, hr), (dl, dr) in zip(ranges1, h_parts, d_parts):
                dl.copy_(hl, non_blocking=True)
                dr.copy_(hr, non_blocking=True)
                d_left[:, a:b].copy_(dl)
                d_right[:, a:b].copy_(dr)
            host_ev[1].record()
            pipe.run(d_left, d_right)
            host_ev[2].record()
            h_disp.copy_(pipe.unshear(), non_blocking=True)
            host_ev[3].record()
            torch.cuda.current_stream().synchronize()
            return h_disp

        # the same host step as a stream of images, two in flight: the upload of image k + 1 (copy stream, second set of device
        # images) and the download of image k - 1 (second copy stream) overlap the wave of image k
        d_sets = [(d_left, d_right), (d_left.clone(), d_right.clone())]
        d_out = [torch.empty((H, Wt), dtype=torch.float32, device=device) for _ in range(2)]
        h_out = [torch.empty((H, Wt), dtype=torch.float32).pin_memory() for _ in range(2)]
        cs_up, cs_dn = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)

        def stream_host(k):
            cur = torch.cuda.current_stream()
            ev_up, ev_run, ev_dn = [None, None], [None, None], [None, None]

            def upload(i):
                s_ = i % 2
                with torch.cuda.stream(cs_up):
                    if ev_run[s_] is not None:
                        cs_up.wait_event(ev_run[s_])               # the wave of image i - 2 has read this set
                    else:
                        cs_up.wait_stream(cur)
                    for (a, b), (hl, hr), (dl, dr) in zip(ranges1, h_parts, d_parts):
                        dl.copy_(hl, non_blocking=True)
                        dr.copy_(hr, non_blocking=True)
                        d_sets[s_][0][:, a:b].copy_(dl)
                        d_sets[s_][1][:, a:b].copy_(dr)
                    ev_up[s_] = torch.cuda.Event()
                    ev_up[s_].record(cs_up)

            upload(0)
            for i in range(k):
                s_ = i % 2
                if i + 1 < k:
                    upload(i + 1)
                cur.wait_event(ev_up[s_])
                if ev_dn[s_] is not None:
                    cur.wait_event(ev_dn[s_])                      # the download of image i - 2 has read d_out[s_]
                pipe.run(*d_sets[s_])
                d_out[s_].copy_(pipe.unshear())
                ev_run[s_] = torch.cuda.Event()
                ev_run[s_].record(cur)
                with torch.cuda.stream(cs_dn):
                    cs_dn.wait_event(ev_run[s_])
                    h_out[s_].copy_(d_out[s_], non_blocking=True)
                    ev_dn[s_] = torch.cuda.Event()
                    ev_dn[s_].record(cs_dn)
            for e_ in ev_dn:
                if e_ is not None:
                    e_.synchronize()
            return h_out[(k - 1) % 2]

        h2d_bytes = sum(b - a for a, b in ranges1) * H * 4 * 2 * world
        d2h_bytes = H * Wt * 4 * world

    for _ in range(max(args.warmup, 3)):
        step_device()
    steps_device(steps)                                        # allocates the buffers of the batches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = pandora_b200.kernel_launches()
    total_ms = timed(steps_device, steps, batched=True)
    launches = pandora_b200.kernel_launches() - launches0
    single_ms = timed(step_device, steps)                      # one image per call (latency view)
    batched_ran = bool(getattr(pipe, "batched_ran", False)) if world == 1 else None
    clocks = sampler.stop() if rank == 0 else None
    sgm_path = pandora_b200.last_path("sgm")

    # ---- per-stage device times inside the same kind of step (events on the launching stream) ----------
    stage = {}
    if world == 1 and getattr(pipe, "fused_ran", False):
        e = pipe.eng
        fev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        torch.cuda.synchronize()
        for i in range(steps):
            fev[i][0].record()
            e.census_sgm_descriptors(d_left, d_right, WINDOW, dmin, dmax, P1, P2)
            fev[i][1].record()
            e.census_sgm(d_left, d_right, WINDOW, dmin, dmax, P1, P2, False, out=pipe.cv_b, fuse_wta=True, invalid_disparity=-9999.0,
                         disp=pipe.disp, flags=pipe.flags, descriptors_ready=True)
            fev[i][2].record()
        torch.cuda.synchronize()
        stage["transform_ms"] = float(np.mean([a.elapsed_time(b) for a, b, _ in fev]))
        stage["fused_ms"] = float(np.mean([b.elapsed_time(c) for _, b, c in fev]))
        if batched_ran and BATCH in batch_in:
            # the same two launches over a batch of BATCH pairs: (batch time - the transforms of its pairs) / pairs
            bev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(max(2, steps // BATCH))]
            for a, b in bev:
                a.record()
                pipe.run_device_batch(*batch_in[BATCH])
                b.record()
            torch.cuda.synchronize()
            stage["batch_ms"] = float(np.mean([a.elapsed_time(b) for a, b in bev]))
            stage["fused_batch_ms_per_pair"] = stage["batch_ms"] / BATCH - stage["transform_ms"]

    for _ in range(2):
        step_host()
    sync_ms = timed(step_host, steps)                          # one synchronous host call per pair (latency view)
    host_parts = None
    if world > 1:
        host_parts = {"upload_ms": host_ev[0].elapsed_time(host_ev[1]), "kernels_ms": host_ev[1].elapsed_time(host_ev[2]),
                      "unshear_d2h_ms": host_ev[2].elapsed_time(host_ev[3]), "note": "rank 0, last timed call"}
    e2e_ms, e2e_mode = sync_ms, "synchronous host call per pair (H2D -> kernels -> D2H back to back)"
    e2e_wall_ms = None
    if world > 1:
        stream_host(2)
        sync_all()
        t_wall = time.perf_counter()
        e2e_ms = timed(stream_host, steps, batched=True)
        e2e_wall_ms = (time.perf_counter() - t_wall) * 1e3
        e2e_mode = ("stream of images, 2 in flight: every rank uploads the image columns its sheared tile visits (pinned host memory, copy "
                    "stream, second set of device images) while the wave of the previous image runs, and downloads its disparity tile on "
                    "a second copy stream; every image's copies are inside the timed region; one image per wave")
    if world == 1:
        # throughput view: a stream of pairs through StereoPipeline.submit_host / result_host -- every pair is uploaded
        # from pinned host memory and its disparity map downloaded inside the timed region; two buffer sets let the
        # copies of the neighbouring pairs overlap the kernels of the current one
        def stream_steps(k):
            # pair by pair (submit_host_batch exists, but a stream of whole batches pays the upload of its first batch and the
            # download of its last one in the open: 1007 vs 1107 Mpix/s over 5 pairs, the same over 12 -- gpurun_out/r3t_*)
            prev = None
            for _ in range(k):
                tk = pipe.submit_host(h_left, h_right)
                if prev is not None:
                    pipe.result_host(prev)
                prev = tk
            return pipe.result_host(prev)

        stream_steps(2)
        sync_all()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        t_wall = time.perf_counter()
        ev[0].record()
        stream_steps(steps)                                     # returns after the last disparity map is in host memory
        ev[1].record()
        sync_all()
        e2e_wall_ms = (time.perf_counter() - t_wall) * 1e3
        e2e_ms = max(ev[0].elapsed_time(ev[1]), 0.0)
        e2e_mode = ("stream of pairs, 2 in flight (StereoPipeline.submit_host / result_host): H2D of pair k+1 and D2H of pair k-1 "
                    "overlap the kernels of pair k; every pair's copies are inside the timed region; one pair per wave")

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    pix = float(H) * Wg
    ms_per_step = total_ms / steps
    value = pix / (ms_per_step * 1e-3) / 1e6
    e2e_value = pix / (e2e_ms / steps * 1e-3) / 1e6
    peak, peak_src = measured_peak_gbs()
    workload = {"C3": f"C3: {H}x{cfg['W']} synthetic pair per GPU, Census 5x5 + SGM 8-path P1=8 P2=32 + WTA, D=256, disp [-255, 0]",
                "C4": f"C4: {H} rows x {Wg} columns (BASELINE configs[4], W = 16384 reading), Census 5x5 + SGM 8-path + WTA, D=256",
                "C4T": f"C4 transposed reading: {H} rows x {Wg} columns, Census 5x5 + SGM 8-path + WTA, D=256"}[args.config]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload, "rows": H, "cols_total": Wg, "cols_per_gpu": Wt,
                   "parallelism": "1 GPU" if world == 1 else
                   f"column tiles x{world}: one skewed wavefront across all GPUs, SGM path states cross the tile borders as NVLink peer "
                   "stores issued by the kernels (no collective on the data path); the timed step ends with the disparity tile back "
                   "in image layout (one NCCL neighbour exchange)",
                   "sgm_path": list(sgm_path),
                   "l2": "the cost volume (17.2 GB per 4096x4096 tile, plus 12.9 GB of packed intermediates) exceeds the 126 MB L2; no flush needed"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": e2e_ms / steps, "mode": e2e_mode,
                "sync_call": {"value": pix / (sync_ms / steps * 1e-3) / 1e6, "unit": UNIT, "ms_per_call": sync_ms / steps}},
        "gpu_launches": int(launches),
    }
    if world == 1:
        line["config"]["mode"] = (f"stream throughput: the timed steps run as batches of up to {BATCH} pairs that follow each other in ONE wave per "
                                  "pass (pb200_census_sgm_batch; each pair has its own images, SGM volume and disparity map; results equal "
                                  "one call per pair bit for bit), so that the wave's fill and drain across the SMs are paid once per batch; "
                                  "`one_pair_per_launch` is the same step with one pair per call") if batched_ran else "one pair per call"
        line["one_pair_per_launch"] = {"ms_per_step": single_ms / steps, "value": pix / (single_ms / steps * 1e-3) / 1e6, "unit": UNIT}
    if parity is not None:
        line["parity_in_run"] = parity
        line["config"]["mode"] = ("stream throughput: the timed steps run as batches of up to 6 images that follow each other in ONE wave per pass "
                                  "(pb200_census_sgm_tile, nimg > 1), so that the time a wave needs to cross all GPUs is paid once per pass "
                                  "and batch; `one_image_at_a_time` is the same step with one image per call")
        line["one_image_at_a_time"] = {"ms_per_step": single_ms / steps, "value": pix / (single_ms / steps * 1e-3) / 1e6, "unit": UNIT}
        line["e2e"]["parts"] = host_parts
    if world == 1:
        if e2e_wall_ms is not None:
            line["e2e"]["wall_ms_per_step"] = e2e_wall_ms / steps
        sgm_alg = 8.0 * D * H * Wg          # SGM must read C (4D) and write S (4D) bytes per pixel
        if "fused_ms" in stage:
            # the dominant kernel pair of the step that was timed.  The SGM row of SURVEY 8(d) (8*D bytes per pixel: the cost C
            # and the result S) is kept as its algorithmic figure although pass 1 does not read a float C -- it computes the
            # Census costs from the descriptors -- so that the fraction stays comparable with earlier lines; what the stage
            # must move once that read is gone is 4*D written + the descriptors (see DESIGN.md 3.2 and the ncu file below)
            stage_ms = stage.get("fused_batch_ms_per_pair", stage["fused_ms"])      # per pair, inside the kind of step that was timed
            f_gbs = sgm_alg / (stage_ms * 1e-3) / 1e9
            kname = "sgm_wave1_kernel" if sgm_path[0].startswith("sgm_wave1") else "sgm_wave_kernel"
            line["roofline"] = {"bound": "hbm",
                                "kernel": f"fused Census+SGM stage = {kname} x2 (pass 1: Hamming costs from the census descriptors + "
                                          "E/SE/S/SW, writes C8 + P16; pass 2: W/NW/N/NE, writes float S + WTA); 2 launches timed as one "
                                          "unit against the SGM row's 8*D algorithmic bytes per pixel",
                                "achieved": f_gbs, "peak": peak, "unit": "GB/s", "frac": f_gbs / peak, "traffic": None,
                                "traffic_note": "dram__bytes per launch are in the ncu capture cited here, not re-measured in this run",
                                "traffic_ncu_file": "profiles/r2_ncu_fused_stage.txt",
                                "peak_source": peak_src, "algorithmic_bytes_per_stage": sgm_alg, "stage_ms": stage_ms,
                                "stage_ms_one_pair_per_launch": stage["fused_ms"],
                                "frac_one_pair_per_launch": sgm_alg / (stage["fused_ms"] * 1e-3) / 1e9 / peak}
            line["stages"] = {"census_transforms": {"ms": stage["transform_ms"], "in_step": True},
                              "census_sgm_fused_wta": {"ms": stage_ms, "algorithmic_bytes": sgm_alg, "in_step": True,
                                                       "achieved_gbs": f_gbs, "frac": f_gbs / peak},
                              "pipeline_algorithmic_bytes": (12.0 * D + 12.0) * H * Wg,
                              "pipeline_frac": (12.0 * D + 12.0) * H * Wg / (ms_per_step * 1e-3) / 1e9 / peak}
        line["config"]["fused_census_sgm"] = bool(getattr(pipe, "fused_ran", False))
        if args.config == "C3":
            # the same configuration through the plugin-level call, and BASELINE.json's other one-GPU configurations
            del d_left, d_right
            pipe = None
            torch.cuda.empty_cache()
            try:
                line["e2e_plugin"], pmap = plugin_leg(pandora_b200, torch, left, right, dmin, dmax, max(2, min(steps, 3)))
            except Exception as exc:  # noqa: BLE001
                line["e2e_plugin"] = {"error": repr(exc)}
            torch.cuda.empty_cache()
            try:
                line["other_configs"] = other_config_lines(pandora_b200, torch, steps)
            except Exception as exc:  # noqa: BLE001
                line["other_configs"] = {"error": repr(exc)}
        # CPU baseline on a bounded row band of the same pair, same run
        rows = 32
        cl, cr = np.ascontiguousarray(left[:rows, :W_IMG]), np.ascontiguousarray(right[:rows, :W_IMG])
        t, kind = cpu_pipeline_once(cl, cr, dmin, dmax)
        line["cpu_baseline"] = {"value": rows * W_IMG / t / 1e6, "unit": UNIT, "cores": 1, "kind": kind,
                                "sample": f"first {rows} rows x {W_IMG} cols x D={D} of the same pair, census "
                                          f"{'= unmodified reference C++ (oracle/_ref)' if kind == 'reference' else '= oracle port'}, "
                                          "SGM/WTA = oracle C port; 1 thread (the reference is single-threaded)",
                                "host_cores_available": os.cpu_count(),
                                "hypothetical_row_parallel": {"cores": os.cpu_count(), "value": rows * W_IMG / t / 1e6 * (os.cpu_count() or 1),
                                                              "note": "1-core figure x host cores: an upper bound, the reference is single-threaded"}}
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS), help="C3 (default, BASELINE's metric) | C4 | C4T")
    args = ap.parse_args()
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
