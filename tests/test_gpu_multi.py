"""Multi-GPU parity (needs >= 2 CUDA devices, NCCL): the row-tiled Census -> SGM -> WTA pipeline over 2 ranks must
equal the single-GPU run of the whole image bit-for-bit.  Runs standalone too: ``python tests/test_gpu_multi.py``."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _log(rank, msg):
    print(f"[rank {rank}] {msg}", flush=True)


def _worker(rank, world, port, H, W, D, tmpdir):
    import torch
    import torch.distributed as dist

    import pandora_b200
    from pandora_b200.synthetic import synthetic_pair
    from pandora_b200.tiling import TiledStereoPipeline, split_rows

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    _log(rank, "process group up")
    left, right, _ = synthetic_pair(H, W, D)
    rows = split_rows(H, world)[rank]
    pipe = TiledStereoPipeline(len(rows), W, -(D - 1), 0, rank, world, dist, 5, 8.0, 32.0, device=f"cuda:{rank}")
    _log(rank, "pipeline built, links warm")
    lt = pipe.eng.to_device(np.ascontiguousarray(left[rows.start: rows.stop]))
    rt = pipe.eng.to_device(np.ascontiguousarray(right[rows.start: rows.stop]))
    for it in range(2):                      # twice: back-to-back steps must not interfere
        disp = pipe.run(lt, rt)
        torch.cuda.synchronize()
        _log(rank, f"step {it} done")
    np.save(os.path.join(tmpdir, f"disp{rank}.npy"), disp.cpu().numpy())
    np.save(os.path.join(tmpdir, f"S{rank}.npy"), pipe.S.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def run_case(world, H, W, D, tmpdir):
    import torch
    import torch.multiprocessing as mp

    import pandora_b200
    from pandora_b200.synthetic import synthetic_pair

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, H, W, D, str(tmpdir)), nprocs=world, join=True)
    left, right, _ = synthetic_pair(H, W, D)
    pipe = pandora_b200.StereoPipeline(H, W, -(D - 1), 0, "census", 5, sgm=(8.0, 32.0), device="cuda:0")
    whole = pipe.run_host(left, right).copy()
    S = pipe.final_cv.cpu().numpy()
    tiled = np.concatenate([np.load(os.path.join(tmpdir, f"disp{r}.npy")) for r in range(world)])
    St = np.concatenate([np.load(os.path.join(tmpdir, f"S{r}.npy")) for r in range(world)])
    np.testing.assert_array_equal(St, S)
    np.testing.assert_array_equal(tiled, whole)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(96, 320, 64), (64, 700, 256)])
def test_two_rank_tiled_pipeline_equals_single_gpu(shape, tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    run_case(2, *shape, tmp_path)


if __name__ == "__main__":
    import tempfile

    import torch

    n = min(torch.cuda.device_count(), int(sys.argv[1]) if len(sys.argv) > 1 else 2)
    with tempfile.TemporaryDirectory() as d:
        run_case(n, 96 * n // 2, 320, 64, d)
        print(f"tiled == single GPU on {n} ranks: OK", flush=True)
