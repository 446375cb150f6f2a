#!/usr/bin/env python3
"""Row-tiled C1 / C2 (pipelines without SGM) under torchrun: every rank a tile of `rows` rows of a (rows * N) x W image
(weak scaling), static input halo (2 rows, 7 with CBCA), no collective on the data path.  One JSON line per configuration.
usage: torchrun --nproc-per-node N tools/bench_local_tiles.py [steps]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pandora_b200.synthetic import synthetic_pair  # noqa: E402
from pandora_b200.tiling import TiledLocalPipeline  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = f"cuda:{lr}"
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device(dev))
for name, rows, W, D, cbca in (("C1", 1024, 1024, 128, None), ("C2", 2048, 2048, 192, (5, 30.0))):
    left, right, _ = synthetic_pair(rows, W, D, seed=20240607 + rank)
    pipe = TiledLocalPipeline(rows, W, -(D - 1), 0, rank, world, dist if world > 1 else None, "census", 5, cbca=cbca, device=dev)
    eng = pipe.pipe.eng
    lt, rt = eng.to_device(left), eng.to_device(right)
    for _ in range(3):
        pipe.run(lt, rt)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(steps):
        pipe.run(lt, rt)
    ev[1].record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev[0].elapsed_time(ev[1]) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"config": name, "n_gpus": world, "tile": [rows, W, D], "cbca": bool(cbca), "halo_rows": pipe.halo, "ms_per_step": float(ms.item()),
                          "mpix_per_s": rows * world * W / float(ms.item()) / 1e3, "scaling": "weak (one tile per GPU)"}), flush=True)
    del pipe, lt, rt
    torch.cuda.empty_cache()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
