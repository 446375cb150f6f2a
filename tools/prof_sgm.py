#!/usr/bin/env python3
"""Time (CUDA events) the SGM stage alone on a random integer cost volume: python tools/prof_sgm.py H W D [reps] [dir_mask]."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pandora_b200  # noqa: E402

H, W, D = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
mask = int(sys.argv[5], 0) if len(sys.argv) > 5 else 0xFF
eng = pandora_b200.get_engine("cuda:0")
g = torch.Generator(device="cuda").manual_seed(1)
cv = torch.randint(0, 26, (H, W, D), device="cuda", generator=g).float()
out = torch.empty_like(cv)
first = min(r for r in range(8) if mask >> r & 1)
for _ in range(2):
    eng.sgm(cv, 8, 32, 58.0, out=out, dir_mask=mask, init_final=3)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for i in range(reps):
    eng.sgm(cv, 8, 32, 58.0, out=out, dir_mask=mask, init_final=3)
    ev[i + 1].record()
torch.cuda.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
print(f"sgm {H}x{W}x{D} mask={mask:#x}: {min(ts):.3f} ms (min of {reps}), {sum(ts)/len(ts):.3f} ms mean; "
      f"{H*W/min(ts)/1e3:.1f} Mpix/s; rows/us = {H/min(ts)/1e3:.4f}")
