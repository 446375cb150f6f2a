#!/usr/bin/env python3
"""cProfile of C3 through the plugin-level call pandora_b200.run(img_left, img_right, cfg) (host datasets in and out)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pandora_b200 as pb  # noqa: E402
from pandora_b200.synthetic import synthetic_pair  # noqa: E402

H = W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
D = 256
left, right, _ = synthetic_pair(H, W, D)
cfg = {"pipeline": {"matching_cost": {"matching_cost_method": "census", "window_size": 5, "subpix": 1},
                    "optimization": {"optimization_method": "sgm", "penalty": {"P1": 8, "P2": 32}},
                    "disparity": {"disparity_method": "wta", "invalid_disparity": -9999}}}
dl = pb.create_image_dataset(left, disparity=[-(D - 1), 0])
dr = pb.create_image_dataset(right)


def once():
    disp, _cv = pb.run(dl, dr, cfg)
    return np.asarray(disp["disparity_map"].data), np.asarray(disp["validity_mask"].data)


once(); once()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    once()
torch.cuda.synchronize()
print(f"plugin call: {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms per pair")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    once()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
