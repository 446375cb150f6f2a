#!/usr/bin/env python3
"""SAD / SSD / ZNCC fill at C0 / C1 sizes: running-sum kernel vs the tap-ordered kernel (CUDA events, min of 5)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pandora_b200  # noqa: E402
from pandora_b200.synthetic import synthetic_pair  # noqa: E402

eng = pandora_b200.get_engine("cuda:0")


def t(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


for H, W, D in ((375, 450, 64), (1024, 1024, 128), (2048, 2048, 192)):
    l, r, _ = synthetic_pair(H, W, D)
    l, r = eng.to_device(l), eng.to_device(r)
    cv = eng.empty((H, W, D))
    gb = (4 * D + 8) * H * W / 1e9
    for name, fn in (("sad", lambda: eng.sad_ssd(l, r, 5, -(D - 1), 0, out=cv)), ("ssd", lambda: eng.sad_ssd(l, r, 5, -(D - 1), 0, squared=True, out=cv)),
                     ("zncc", lambda: eng.zncc(l, r, 5, -(D - 1), 0, out=cv))):
        ms = t(fn)
        with pandora_b200.option("sad.taps", 1):
            ms0 = t(fn)
        print(f"{H}x{W}x{D} {name}: running {ms:.4f} ms ({gb / ms:.0f} GB/s algorithmic)   taps {ms0:.4f} ms", flush=True)
