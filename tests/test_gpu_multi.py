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


# ----------------------------------------------------------------------------------------------------------------------
# column tiles: the skewed wavefront as one wave across the GPUs (pb200_census_sgm_tile)
# ----------------------------------------------------------------------------------------------------------------------
def _col_worker(rank, world, port, H, Wg, D, dmin, steps, tmpdir):
    import torch
    import torch.distributed as dist

    from pandora_b200.synthetic import synthetic_pair
    from pandora_b200.tiling import ColumnTiledStereoPipeline

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    left, right, _ = synthetic_pair(H, Wg, D)
    pipe = ColumnTiledStereoPipeline(H, Wg, dmin, dmin + D - 1, rank, world, dist, 5, 8.0, 32.0, device=f"cuda:{rank}")
    lt, rt = pipe.eng.to_device(np.ascontiguousarray(left)), pipe.eng.to_device(np.ascontiguousarray(right))
    for it in range(steps):                  # back-to-back images: the links are never cleared, only the epoch changes
        pipe.run(lt, rt)
        tile = pipe.unshear()
        torch.cuda.synchronize()
        _log(rank, f"column-tiled step {it} done")
    # a batch of three images in one wave (image 1 differs: its paths must not leak into its neighbours)
    l3, r3 = torch.stack([lt, rt, lt]), torch.stack([rt, lt, rt])
    pipe.run(l3, r3)
    out3 = pipe.unshear()
    assert torch.equal(out3[0], tile) and torch.equal(out3[2], tile), "batched image differs"
    vol3 = pipe.unshear(pipe.cv)
    pipe.run(rt, lt)                         # the swapped pair alone, after a batch (link epochs / credits of both passes)
    assert torch.equal(pipe.unshear(), out3[1]) and torch.equal(torch.nan_to_num(pipe.unshear(pipe.cv)), torch.nan_to_num(vol3[1]))
    pipe.run(lt, rt)
    np.save(os.path.join(tmpdir, f"cdisp{rank}.npy"), tile.cpu().numpy())
    np.save(os.path.join(tmpdir, f"cS{rank}.npy"), pipe.unshear(pipe.cv).cpu().numpy())
    dist.barrier()
    pipe.close()
    dist.destroy_process_group()


def run_column_case(world, H, Wg, D, dmin, tmpdir, steps=2):
    import torch.multiprocessing as mp

    import pandora_b200
    from pandora_b200.synthetic import synthetic_pair

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_col_worker, args=(world, port, H, Wg, D, dmin, steps, str(tmpdir)), nprocs=world, join=True)
    left, right, _ = synthetic_pair(H, Wg, D)
    eng = pandora_b200.get_engine("cuda:0")
    out = eng.census_sgm(eng.to_device(left), eng.to_device(right), 5, dmin, dmin + D - 1, 8.0, 32.0)
    assert out is not None
    S, whole = out[0].cpu().numpy(), out[1].cpu().numpy()
    tiled = np.concatenate([np.load(os.path.join(tmpdir, f"cdisp{r}.npy")) for r in range(world)], axis=1)
    St = np.concatenate([np.load(os.path.join(tmpdir, f"cS{r}.npy")) for r in range(world)], axis=1)
    np.testing.assert_array_equal(tiled, whole)
    np.testing.assert_array_equal(St, S)


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,D,dmin", [(40, 300, 64, -63), (33, 4096, 256, -255), (70, 36, 64, -20), (25, 610, 128, -100)])
def test_column_tile_self_linked_on_one_gpu_vs_oracle(H, W, D, dmin, oracle):
    """ntiles = 1: the wave's cyclic boundary goes through a LINK buffer (the code path of a GPU boundary: tagged words with
    an epoch, credits, sheared storage, the pass-2 tile origin) instead of the in-kernel L2 ring -- bit-exact against the
    oracle chain, three images back to back."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pandora_b200.tiling import ColumnTiledStereoPipeline

    g = np.random.default_rng(H * W)
    base = g.integers(0, 9, (H, W + 8)).astype(np.float32)
    left, right = np.ascontiguousarray(base[:, 4:4 + W]), np.ascontiguousarray(np.roll(base, 3, axis=1)[:, 4:4 + W])
    right[g.random((H, W)) < 0.2] += 1.0
    cv, attrs = oracle.census_cost_volume(left, right, 5, dmin, dmin + D - 1)
    S = oracle.sgm_cost_volume(cv, 8, 32, cmax=attrs["cmax"])
    disp, inv = oracle.wta(S, np.arange(dmin, dmin + D))
    pipe = ColumnTiledStereoPipeline(H, W, dmin, dmin + D - 1, 0, 1, None, 5, 8.0, 32.0, device="cuda:0")
    lt, rt = pipe.eng.to_device(left), pipe.eng.to_device(right)
    for _ in range(3):
        pipe.run(lt, rt)
        np.testing.assert_array_equal(pipe.unshear().cpu().numpy(), disp)
    pipe.run(torch.stack([lt, rt, lt]), torch.stack([rt, lt, rt]))            # a batch in one wave; image 1 is another pair
    out3 = pipe.unshear().cpu().numpy()
    np.testing.assert_array_equal(out3[0], disp)
    np.testing.assert_array_equal(out3[2], disp)
    np.testing.assert_array_equal(pipe.unshear(pipe.cv)[2].cpu().numpy(), S)
    pipe.run(lt, rt)
    np.testing.assert_array_equal(pipe.unshear(pipe.cv).cpu().numpy(), S)
    np.testing.assert_array_equal(pipe.unshear(pipe.flags).cpu().numpy().astype(bool), inv)
    pipe.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world,H,W,D,dmin", [(2, 96, 320, 64, -63), (2, 48, 1400, 256, -255), (2, 300, 128, 128, -100), (4, 64, 512, 64, -63),
                                              (8, 40, 2048, 256, -255)])
def test_column_tiled_pipeline_equals_single_gpu(world, H, W, D, dmin, tmp_path):
    """One wave across `world` GPUs (NVLink peer stores inside the kernels) == the one-GPU run, volume and disparity map, bit for
    bit; H > Wt exercises tiles that drift across more than one neighbour, two images back to back the epoch tags."""
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} CUDA devices")
    run_column_case(world, H, W, D, dmin, tmp_path)


# ----------------------------------------------------------------------------------------------------------------------
# row tiles of the pipelines without SGM (C1 / C2): static input halo, no collective on the data path
# ----------------------------------------------------------------------------------------------------------------------
def _local_worker(rank, world, port, H, W, D, method, cbca, tmpdir):
    import torch
    import torch.distributed as dist

    from pandora_b200.synthetic import synthetic_pair
    from pandora_b200.tiling import TiledLocalPipeline, split_rows

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    left, right, _ = synthetic_pair(H, W, D)
    rows = split_rows(H, world)[rank]
    pipe = TiledLocalPipeline(len(rows), W, -(D - 1), 0, rank, world, dist, method, 5, cbca=cbca, device=f"cuda:{rank}")
    lt = pipe.pipe.eng.to_device(np.ascontiguousarray(left[rows.start: rows.stop]))
    rt = pipe.pipe.eng.to_device(np.ascontiguousarray(right[rows.start: rows.stop]))
    for _ in range(2):
        disp = pipe.run(lt, rt)
        torch.cuda.synchronize()
    np.save(os.path.join(tmpdir, f"ldisp{rank}.npy"), disp.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("method,cbca,shape", [("census", (5, 30.0), (120, 260, 192)), ("census", None, (90, 200, 128)), ("sad", None, (90, 200, 64))])
def test_two_rank_local_pipeline_equals_single_gpu(method, cbca, shape, tmp_path):
    """C2 / C1 / C0-style pipelines row-tiled over 2 GPUs (7-row halo with CBCA) == the one-GPU run of the whole image."""
    import torch
    import torch.multiprocessing as mp

    import pandora_b200
    from pandora_b200.synthetic import synthetic_pair

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    H, W, D = shape
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_local_worker, args=(2, port, H, W, D, method, cbca, str(tmp_path)), nprocs=2, join=True)
    left, right, _ = synthetic_pair(H, W, D)
    pipe = pandora_b200.StereoPipeline(H, W, -(D - 1), 0, method, 5, cbca=cbca, device="cuda:0")
    whole = pipe.run_host(left, right).copy()
    tiled = np.concatenate([np.load(os.path.join(tmp_path, f"ldisp{r}.npy")) for r in range(2)])
    np.testing.assert_array_equal(tiled, whole)


def test_local_pipeline_single_rank_on_one_gpu():
    """world = 1: the tiled wrapper is the plain pipeline (runs on the driver's one-GPU box)."""
    import torch

    import pandora_b200
    from pandora_b200.synthetic import synthetic_pair
    from pandora_b200.tiling import TiledLocalPipeline

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    H, W, D = 64, 128, 64
    left, right, _ = synthetic_pair(H, W, D)
    tl = TiledLocalPipeline(H, W, -(D - 1), 0, 0, 1, None, "census", 5, cbca=(5, 30.0), device="cuda:0")
    eng = tl.pipe.eng
    got = tl.run(eng.to_device(left), eng.to_device(right)).cpu().numpy()
    ref = pandora_b200.StereoPipeline(H, W, -(D - 1), 0, "census", 5, cbca=(5, 30.0), device="cuda:0").run_host(left, right)
    np.testing.assert_array_equal(got, ref)


test_local_pipeline_single_rank_on_one_gpu = pytest.mark.gpu(test_local_pipeline_single_rank_on_one_gpu)
