#!/usr/bin/env python3
"""reverse_cost_volume timing: python tools/prof_reverse.py H W D -- tiled kernel vs the plain gather (PB200_REVERSE_GATHER=1)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pandora_b200  # noqa: E402

H, W, D = (int(a) for a in sys.argv[1:4])
eng = pandora_b200.get_engine("cuda:0")
cv = torch.rand((H, W, D), device="cuda")
outs = []
for name, env in (("tiled", None), ("gather", "1")):
    if env:
        os.environ["PB200_REVERSE_GATHER"] = env
    else:
        os.environ.pop("PB200_REVERSE_GATHER", None)
    ts = []
    for i in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = eng.reverse_cost_volume(cv, -(D - 1))
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    outs.append(torch.nan_to_num(out, nan=-7.0))
    print(f"reverse_cost_volume {H}x{W}x{D} {name}: {min(ts[1:]):.3f} ms; {8.0 * D * H * W / min(ts[1:]) / 1e6:.0f} GB/s", flush=True)
    del out
os.environ.pop("PB200_REVERSE_GATHER", None)
print("identical:", bool(torch.equal(outs[0], outs[1])))
