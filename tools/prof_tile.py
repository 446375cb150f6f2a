#!/usr/bin/env python3
"""Time the column-tile entry on one GPU (ntiles = 1, self-linked): python tools/prof_tile.py H W D -- whole step, each pass
alone, a batch of 4 images, the un-shear."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pandora_b200  # noqa: E402
from pandora_b200.synthetic import synthetic_pair  # noqa: E402
from pandora_b200.tiling import ColumnTiledStereoPipeline  # noqa: E402

H, W, D = (int(a) for a in sys.argv[1:4])
left, right, _ = synthetic_pair(H, W, D)
pipe = ColumnTiledStereoPipeline(H, W, -(D - 1), 0, 0, 1, None, 5, 8.0, 32.0, device="cuda:0")
dl, dr = pipe.eng.to_device(left), pipe.eng.to_device(right)


def timeit(name, fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"{name}: {min(ts):.3f} ms (min of {reps})", flush=True)


timeit("tile step (transforms + both passes)", lambda: pipe.run(dl, dr))
timeit("un-shear of the disparity tile", lambda: pipe.unshear())


d4l, d4r = torch.stack([dl] * 4), torch.stack([dr] * 4)
timeit("batch of 4 images in one wave", lambda: pipe.run(d4l, d4r), reps=2)
eng = pipe.eng
timeit("one-GPU entry (pb200_census_sgm, default kernels)", lambda: eng.census_sgm(dl, dr, 5, -(D - 1), 0, 8, 32, out=pipe.cv, disp=pipe.disp, flags=pipe.flags))
with pandora_b200.option("sgm.wave_kernel", 1):
    timeit("one-GPU entry, skewed kernels", lambda: eng.census_sgm(dl, dr, 5, -(D - 1), 0, 8, 32, out=pipe.cv, disp=pipe.disp, flags=pipe.flags))
