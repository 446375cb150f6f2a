#!/usr/bin/env python3
"""Where does the fused Census -> SGM stage differ from the two separate calls?  (debug aid, GPU box only)"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import pandora_b200  # noqa: E402

eng = pandora_b200.get_engine("cuda:0")
for (H, W, D, dmin) in [(12, 64, 64, -63), (16, 300, 128, -127), (20, 300, 256, -255), (20, 300, 256, -100), (40, 4096, 256, -255)]:
    g = np.random.default_rng(W)
    left = g.integers(0, 255, (H, W)).astype(np.float32)
    right = g.integers(0, 255, (H, W)).astype(np.float32)
    dl, dr = eng.to_device(left), eng.to_device(right)
    dmax = dmin + D - 1
    S_ref, disp_ref, _ = eng.sgm(eng.census(dl, dr, 5, dmin, dmax), 8, 32, 58.0, fuse_wta=True, dmin=dmin)
    out = eng.census_sgm(dl, dr, 5, dmin, dmax, 8, 32)
    torch.cuda.synchronize()
    if out is None:
        print(H, W, D, dmin, "NOT FUSED")
        continue
    a, b = torch.nan_to_num(out[0], nan=-7.0), torch.nan_to_num(S_ref, nan=-7.0)
    bad = (a != b)
    n = int(bad.sum())
    print(f"{H}x{W}x{D} dmin={dmin}: mismatching cells {n} of {bad.numel()}, disp mismatches {int((out[1] != disp_ref).sum())}")
    if n:
        idx = bad.nonzero()[:12].cpu().numpy()
        for (y, x, d) in idx:
            print("   y,x,d =", y, x, d, "fused", float(a[y, x, d]), "ref", float(b[y, x, d]))
        print("   rows with mismatches:", bad.any(2).any(1).nonzero().flatten()[:20].cpu().numpy())
        print("   cols with mismatches:", bad.any(2).any(0).nonzero().flatten()[:40].cpu().numpy())
        print("   disps with mismatches:", bad.any(0).any(0).nonzero().flatten()[:40].cpu().numpy())
        print("   nan pattern differs:", int((torch.isnan(out[0]) != torch.isnan(S_ref)).sum()))
