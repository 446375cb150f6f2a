#!/usr/bin/env python3
"""One batched fused stage for an ncu capture: prof_batch.py H W D nimg"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pandora_b200  # noqa: E402
from pandora_b200.synthetic import synthetic_pair  # noqa: E402

H, W, D, n = (int(a) for a in sys.argv[1:5])
eng = pandora_b200.get_engine("cuda:0")
l, r, _ = synthetic_pair(H, W, D)
l = eng.to_device(l).unsqueeze(0).expand(n, -1, -1).contiguous()
r = eng.to_device(r).unsqueeze(0).expand(n, -1, -1).contiguous()
for _ in range(2):
    assert eng.census_sgm_batch(l, r, 5, -(D - 1), 0, 8, 32) is not None
torch.cuda.synchronize()
