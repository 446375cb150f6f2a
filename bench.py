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

        dist.init_process_group("nccl", device_id=torch.device(device))
    dmin, dmax = -(D_DISP - 1), 0
    H, W, D = H_TILE, W_IMG, D_DISP
    # every rank's tile of the tall image: same generator, different seed per rank (tile r = rows [r*H, (r+1)*H))
    left, right, _ = synthetic_pair(H, W, D, seed=20240607 + rank)
    if world == 1:
        pipe = pandora_b200.StereoPipeline(H, W, dmin, dmax, "census", WINDOW, sgm=(P1, P2), device=device)
        d_left, d_right = pipe.eng.to_device(left), pipe.eng.to_device(right)
        h_left = torch.from_numpy(left).pin_memory()
        h_right = torch.from_numpy(right).pin_memory()

        def step_device():
            return pipe.run_device(d_left, d_right)

        def step_host():
            return pipe.run_host(h_left, h_right)
    else:
        from pandora_b200.tiling import TiledStereoPipeline  # noqa: PLC0415

        pipe = TiledStereoPipeline(H, W, dmin, dmax, rank, world, dist, WINDOW, P1, P2, device=device)
        d_left, d_right = pipe.eng.to_device(left), pipe.eng.to_device(right)
        h_left = torch.from_numpy(left).pin_memory()
        h_right = torch.from_numpy(right).pin_memory()
        h_disp = torch.empty((H, W), dtype=torch.float32).pin_memory()
        s_left, s_right = torch.empty_like(d_left), torch.empty_like(d_right)

        def step_device():
            return pipe.run(d_left, d_right)

        def step_host():
            s_left.copy_(h_left, non_blocking=True)
            s_right.copy_(h_right, non_blocking=True)
            out = pipe.run(s_left, s_right)
            h_disp.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return h_disp

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, k):
        sync_all()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
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

    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = pandora_b200.kernel_launches()
    total_ms = timed(step_device, args.steps)
    launches = pandora_b200.kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-stage device times inside the same kind of step (events on the launching stream) ----------
    stage = {}
    if world == 1:
        e = pipe.eng
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
        torch.cuda.synchronize()
        for i in range(args.steps):
            evs[i][0].record()
            e.census(d_left, d_right, WINDOW, dmin, dmax, out=pipe.cv_a)
            evs[i][1].record()
            e.sgm(pipe.cv_a, P1, P2, WINDOW * WINDOW + P2 + 1.0, False, out=pipe.cv_b, fuse_wta=True, dmin=dmin,
                  invalid_disparity=-9999.0, disp=pipe.disp, flags=pipe.flags)
            evs[i][2].record()
        torch.cuda.synchronize()
        stage["census_ms"] = float(np.mean([a.elapsed_time(b) for a, b, _ in evs]))
        stage["sgm_ms"] = float(np.mean([b.elapsed_time(c) for _, b, c in evs]))
        if pipe.fused_ran:
            # the stage the pipeline actually runs: census transforms + the two wavefront passes, Census costs computed
            # inside pass 1 (pb200_census_sgm)
            fev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
            torch.cuda.synchronize()
            for i in range(args.steps):
                fev[i][0].record()
                e.census_sgm_descriptors(d_left, d_right, WINDOW, dmin, dmax, P1, P2)
                fev[i][1].record()
                e.census_sgm(d_left, d_right, WINDOW, dmin, dmax, P1, P2, False, out=pipe.cv_b, fuse_wta=True, invalid_disparity=-9999.0,
                             disp=pipe.disp, flags=pipe.flags, descriptors_ready=True)
                fev[i][2].record()
            torch.cuda.synchronize()
            stage["transform_ms"] = float(np.mean([a.elapsed_time(b) for a, b, _ in fev]))
            stage["fused_ms"] = float(np.mean([b.elapsed_time(c) for _, b, c in fev]))

    for _ in range(2):
        step_host()
    sync_ms = timed(step_host, args.steps)                     # one synchronous host call per pair (latency view)
    e2e_ms, e2e_mode = sync_ms, "synchronous host call per pair (H2D -> kernels -> D2H back to back)"
    if world == 1:
        # throughput view: a stream of pairs through StereoPipeline.submit_host / result_host -- every pair is uploaded
        # from pinned host memory and its disparity map downloaded inside the timed region; two buffer sets let the
        # copies of the neighbouring pairs overlap the kernels of the current one
        def stream_steps(k):
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
        stream_steps(args.steps)                                # returns after the last disparity map is in host memory
        ev[1].record()
        sync_all()
        wall_ms = (time.perf_counter() - t_wall) * 1e3
        e2e_ms = max(ev[0].elapsed_time(ev[1]), 0.0)
        e2e_mode = ("stream of pairs, 2 in flight (StereoPipeline.submit_host / result_host): H2D of pair k+1 and D2H of pair k-1 "
                    "overlap the kernels of pair k; every pair's copies are inside the timed region")
        e2e_wall_ms = wall_ms

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    pix = float(H) * W * world
    ms_per_step = total_ms / args.steps
    value = pix / (ms_per_step * 1e-3) / 1e6
    e2e_value = pix / (e2e_ms / args.steps * 1e-3) / 1e6
    peak, peak_src = measured_peak_gbs()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "C3: 4096x4096 synthetic pair per GPU, Census 5x5 + SGM 8-path P1=8 P2=32 + WTA, D=256, disp [-255, 0]",
                   "rows_per_gpu": H, "cols": W, "total_rows": H * world,
                   "parallelism": "1 GPU" if world == 1 else f"row tiles x{world}, SGM path-state halo over NCCL p2p",
                   "l2": "the cost volume (17.2 GB per GPU, plus 12.9 GB of packed intermediates) exceeds the 126 MB L2; no flush needed"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * H * W * 4 * world, "d2h_bytes_per_step": H * W * 4 * world,
                "ms_per_step": e2e_ms / args.steps, "mode": e2e_mode,
                "sync_call": {"value": pix / (sync_ms / args.steps * 1e-3) / 1e6, "unit": UNIT, "ms_per_call": sync_ms / args.steps}},
        "gpu_launches": int(launches),
    }
    if world == 1:
        line["e2e"]["wall_ms_per_step"] = e2e_wall_ms / args.steps
        sgm_alg = 8.0 * D * H * W          # SGM must read C (4D) and write S (4D) bytes per pixel
        census_alg = (4.0 * D + 8.0) * H * W
        sgm_gbs = sgm_alg / (stage["sgm_ms"] * 1e-3) / 1e9
        fused = "fused_ms" in stage
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "sgm_stage_traffic.json")) as fh:
                traffic = json.load(fh).get("dram_bytes_per_fused_stage" if fused else "dram_bytes_per_stage")
                traffic = None if traffic is None else float(traffic)
        except Exception:
            pass
        if fused:
            # the dominant kernel pair of the step that was timed: the SGM row of SURVEY 8(d) (8*D bytes per pixel: the
            # cost C and the result S) is kept as its algorithmic figure although pass 1 no longer reads a float C --
            # it computes the Census costs from the descriptors -- so the fraction stays comparable with earlier lines
            f_gbs = sgm_alg / (stage["fused_ms"] * 1e-3) / 1e9
            line["roofline"] = {"bound": "hbm",
                                "kernel": "fused Census+SGM stage = sgm_wave_kernel x2 (pass 1: Hamming costs from the census "
                                          "descriptors + E/SE/S/SW, writes C8 + P16; pass 2: W/NW/N/NE, writes float S + WTA); 2 launches "
                                          "timed as one unit against the SGM row's 8*D algorithmic bytes per pixel",
                                "achieved": f_gbs, "peak": peak, "unit": "GB/s", "frac": f_gbs / peak, "traffic": traffic,
                                "peak_source": peak_src, "algorithmic_bytes_per_stage": sgm_alg, "stage_ms": stage["fused_ms"]}
        else:
            line["roofline"] = {"bound": "hbm",
                                "kernel": "SGM stage = sgm_wave_kernel x2 (pass 1: E/SE/S/SW reading float C; pass 2: W/NW/N/NE writing "
                                          "float S + WTA); 2 launches timed as one unit (the stage's 8*D algorithmic bytes per pixel "
                                          "= 4*D read by pass 1 + 4*D written by pass 2)",
                                "achieved": sgm_gbs, "peak": peak, "unit": "GB/s", "frac": sgm_gbs / peak, "traffic": traffic,
                                "peak_source": peak_src, "algorithmic_bytes_per_stage": sgm_alg, "stage_ms": stage["sgm_ms"]}
        cen_gbs = census_alg / (stage["census_ms"] * 1e-3) / 1e9
        line["stages"] = {"census_fill": {"ms": stage["census_ms"], "algorithmic_bytes": census_alg, "achieved_gbs": cen_gbs, "frac": cen_gbs / peak,
                                          "in_step": not fused},
                          "sgm_8path_wta": {"ms": stage["sgm_ms"], "algorithmic_bytes": sgm_alg, "achieved_gbs": sgm_gbs, "frac": sgm_gbs / peak,
                                            "in_step": not fused},
                          "pipeline_algorithmic_bytes": (12.0 * D + 12.0) * H * W,
                          "pipeline_frac": (12.0 * D + 12.0) * H * W / (ms_per_step * 1e-3) / 1e9 / peak}
        if fused:
            line["stages"]["census_transform_x2"] = {"ms": stage["transform_ms"], "in_step": True}
            line["stages"]["census_sgm_fused_wta"] = {"ms": stage["fused_ms"], "algorithmic_bytes": sgm_alg, "in_step": True,
                                                      "achieved_gbs": sgm_alg / (stage["fused_ms"] * 1e-3) / 1e9,
                                                      "frac": sgm_alg / (stage["fused_ms"] * 1e-3) / 1e9 / peak}
        line["config"]["fused_census_sgm"] = bool(fused)
        # CPU baseline on a bounded row band of the same pair, same run
        rows = 32
        cl, cr = np.ascontiguousarray(left[:rows]), np.ascontiguousarray(right[:rows])
        t, kind = cpu_pipeline_once(cl, cr, dmin, dmax)
        line["cpu_baseline"] = {"value": rows * W / t / 1e6, "unit": UNIT, "cores": 1, "kind": kind,
                                "sample": f"first {rows} rows x {W} cols x D={D} of the same pair, census "
                                          f"{'= unmodified reference C++ (oracle/_ref)' if kind == 'reference' else '= oracle port'}, "
                                          "SGM/WTA = oracle C port; 1 thread (the reference is single-threaded)",
                                "host_cores_available": os.cpu_count(),
                                "hypothetical_row_parallel": {"cores": os.cpu_count(), "value": rows * W / t / 1e6 * (os.cpu_count() or 1),
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
    args = ap.parse_args()
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
