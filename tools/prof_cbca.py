#!/usr/bin/env python3
"""CBCA stage on C2-like input for ncu: python tools/prof_cbca.py H W D"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pandora_b200
from pandora_b200.synthetic import synthetic_pair
H, W, D = (int(a) for a in sys.argv[1:4])
eng = pandora_b200.get_engine("cuda:0")
l, r, _ = synthetic_pair(H, W, D)
l, r = eng.to_device(l), eng.to_device(r)
cv = eng.census(l, r, 5, -(D - 1), 0)
out = torch.empty_like(cv)
for _ in range(3):
    eng.cbca(l, r, cv, 2, -(D - 1), 5, 30.0, out=out)
torch.cuda.synchronize()
