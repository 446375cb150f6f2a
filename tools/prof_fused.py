#!/usr/bin/env python3
"""Time (CUDA events) the fused Census -> SGM -> WTA stage alone: python tools/prof_fused.py H W D [reps] [kernels].
`kernels` = comma list of "sgm.wave_kernel" values (1 = one column per warp, 2 = two columns per warp); with both, the
outputs are also compared bit for bit."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pandora_b200  # noqa: E402
from pandora_b200.synthetic import synthetic_pair  # noqa: E402

H, W, D = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
kernels = [int(v) for v in sys.argv[5].split(",")] if len(sys.argv) > 5 else [1, 2]
eng = pandora_b200.get_engine("cuda:0")
left, right, _ = synthetic_pair(H, W, D)
dl, dr = eng.to_device(left), eng.to_device(right)
disp = eng.empty((H, W))
flags = eng.empty((H, W), torch.uint8)
outs = {}
for k in kernels:
    out = eng.empty((H, W, D))
    with pandora_b200.option("sgm.wave_kernel", k):
        eng.census_sgm_descriptors(dl, dr, 5, -(D - 1), 0, 8, 32)
        for _ in range(2 if reps > 1 else 0):
            eng.census_sgm(dl, dr, 5, -(D - 1), 0, 8, 32, out=out, disp=disp, flags=flags, descriptors_ready=True)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        ev[0].record()
        for i in range(reps):
            assert eng.census_sgm(dl, dr, 5, -(D - 1), 0, 8, 32, out=out, disp=disp, flags=flags, descriptors_ready=True) is not None
            ev[i + 1].record()
        torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
    print(f"fused census+sgm+wta {H}x{W}x{D} kernel={k} path={pandora_b200.last_path('sgm')}: {min(ts):.3f} ms (min of {reps}), "
          f"{sum(ts) / len(ts):.3f} ms mean; {H * W / min(ts) / 1e3:.1f} Mpix/s; us/row (both passes) = {min(ts) * 1e3 / H:.3f}", flush=True)
    outs[k] = (out, disp.clone())
if len(outs) == 2:
    (a, da), (b, db) = outs.values()
    same = torch.equal(da, db) and torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0))
    print("outputs identical:", same, flush=True)

# batches of pairs through one wave (two-column kernels): time per image
if os.environ.get("PROF_BATCH"):
    for n in [int(v) for v in os.environ["PROF_BATCH"].split(",")]:
        bl = dl.unsqueeze(0).expand(n, -1, -1).contiguous()
        br = dr.unsqueeze(0).expand(n, -1, -1).contiguous()
        out = eng.empty((n, H, W, D))
        bd, bf = eng.empty((n, H, W)), eng.empty((n, H, W), torch.uint8)
        for _ in range(2):
            assert eng.census_sgm_batch(bl, br, 5, -(D - 1), 0, 8, 32, out=out, disp=bd, flags=bf) is not None
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.census_sgm_batch(bl, br, 5, -(D - 1), 0, 8, 32, out=out, disp=bd, flags=bf)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print(f"batch of {n}: {min(ts):.3f} ms = {min(ts) / n:.3f} ms per image (incl. the transforms), {n * H * W / min(ts) / 1e3:.1f} Mpix/s; "
              f"image 0 == single call: {torch.equal(bd[0], outs[2][1]) if 2 in outs else None}", flush=True)
        del out, bl, br, bd, bf
        torch.cuda.empty_cache()
