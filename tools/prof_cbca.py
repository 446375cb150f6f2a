#!/usr/bin/env python3
"""CBCA stage on C2-like input: python tools/prof_cbca.py H W D [reps] -- times the register kernel (default for
cbca_distance <= 5) and the staged kernel (option cbca.pipe = 1) with CUDA events and checks that they agree."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pandora_b200  # noqa: E402
from pandora_b200.synthetic import synthetic_pair  # noqa: E402

H, W, D = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
eng = pandora_b200.get_engine("cuda:0")
l, r, _ = synthetic_pair(H, W, D)
l, r = eng.to_device(l), eng.to_device(r)
cv = eng.census(l, r, 5, -(D - 1), 0)
cl, cr = eng.cbca_supports(l, r, 2, 5, 30.0)
outs = {}
for name, env in (("register kernel", None), ("staged kernel", "1")):
    pandora_b200.set_option("cbca.pipe", 1 if env else -1)       # kernel selection is a library option, not an environment variable
    out = torch.empty_like(cv)

    def agg():
        pandora_b200._native.check(eng.lib.pb200_cbca_aggregate(cv.data_ptr(), out.data_ptr(), H, W, D, -(D - 1), 2, cl.data_ptr(), cr.data_ptr(), 5,
                                                                torch.cuda.current_stream().cuda_stream))

    for _ in range(2 if reps > 1 else 0):
        agg()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        agg()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    alg = (8 * D + 24) * H * W
    print(f"cbca aggregate {H}x{W}x{D} {name}: {min(ts):.3f} ms (min of {reps}); {alg / min(ts) / 1e6:.0f} GB/s algorithmic", flush=True)
    outs[name] = out
pandora_b200.set_option("cbca.pipe", -1)
a, b = (torch.nan_to_num(o, nan=-7.0) for o in outs.values())
print("identical:", bool(torch.equal(a, b)), "mismatching cells:", int((a != b).sum()))
