#!/usr/bin/env python3
"""One call of a secondary stage for an ncu capture: prof_one.py sad|ssd|zncc|cbca H W D"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pandora_b200  # noqa: E402
from pandora_b200.synthetic import synthetic_pair  # noqa: E402

what, H, W, D = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
eng = pandora_b200.get_engine("cuda:0")
l, r, _ = synthetic_pair(H, W, D)
l, r = eng.to_device(l), eng.to_device(r)
cv = eng.empty((H, W, D))
for _ in range(3):
    if what in ("sad", "ssd"):
        eng.sad_ssd(l, r, 5, -(D - 1), 0, squared=(what == "ssd"), out=cv)
    elif what == "zncc":
        eng.zncc(l, r, 5, -(D - 1), 0, out=cv)
    elif what == "census":
        eng.census(l, r, 5, -(D - 1), 0, out=cv, fuse_wta=True)
    elif what == "wta":
        eng.census(l, r, 5, -(D - 1), 0, out=cv)
        eng.wta(cv, -(D - 1))
    elif what == "cbca":
        eng.census(l, r, 5, -(D - 1), 0, out=cv)
        eng.cbca(l, r, cv, 2, -(D - 1))
torch.cuda.synchronize()
