"""CPU tests of the multi-rank schedule: world_size 2 and 3 over gloo, with the CPU oracle standing in for
the CUDA kernels (same per-direction interface), must reproduce the untiled result bit-for-bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pandora_b200 import tiling


def test_direction_order_is_deadlock_free():
    for world in (1, 2, 3, 4, 5, 8):
        orders = [tiling.group_order(r, world) for r in range(world)]
        for r, o in enumerate(orders):
            assert sorted(o) == [0, 1]
            dirs = tiling.direction_order(r, world)
            assert sorted(dirs) == [2, 3, 4, 5, 6, 7]
            assert [d for d in dirs if d in tiling.DOWN] == list(tiling.DOWN) and [d for d in dirs if d in tiling.UP] == list(tiling.UP)
        # simulate: a rank may run group g once its upstream neighbour has run g
        done = [set() for _ in range(world)]
        pos = [0] * world
        progressed = True
        while progressed:
            progressed = False
            for r in range(world):
                while pos[r] < 2:
                    g = orders[r][pos[r]]
                    up = r - 1 if g == 0 else r + 1
                    if 0 <= up < world and g not in done[up]:
                        break
                    done[r].add(g)
                    pos[r] += 1
                    progressed = True
        assert all(p == 2 for p in pos), (world, pos)


def test_split_rows():
    parts = tiling.split_rows(10, 3)
    assert [len(p) for p in parts] == [4, 3, 3] and parts[0][0] == 0 and parts[-1][-1] == 9


class OracleBackend(tiling.SgmBackend):
    """Group-wise SGM through the CPU oracle (test stand-in for EngineSgmBackend)."""

    def __init__(self, orc, C, S, p1, p2):
        self.orc, self.C, self.S, self.p1, self.p2 = orc, C, S, p1, p2
        self.calls = []

    def new_halo(self):
        return torch.zeros((3,) + tuple(self.C.shape[1:]), dtype=torch.float32)

    def run_horizontal(self, init):
        self.calls.append(("h", init, False))
        self.orc.sgm_direction(self.C, self.S, self.p1, self.p2, 0, init, None, None)
        self.orc.sgm_direction(self.C, self.S, self.p1, self.p2, 1, False, None, None)

    def run_group(self, group, final, halo_in, halo_out):
        self.calls.append((group, False, final))
        for k, direction in enumerate((tiling.DOWN, tiling.UP)[group]):
            hi = None if halo_in is None else halo_in[k].numpy()
            ho = None if halo_out is None else halo_out[k].numpy()
            self.orc.sgm_direction(self.C, self.S, self.p1, self.p2, direction, False, hi, ho)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, H, W, D, tmpdir):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle as orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pg_down = dist.new_group(list(range(world)))
    pg_up = dist.new_group(list(range(world)))
    tiling.warm_up_links(rank, world, dist, pg_down, pg_up)
    g = np.random.default_rng(5)
    C = g.integers(0, 26, (H, W, D)).astype(np.float32)          # every rank draws the same full volume
    rows = tiling.split_rows(H, world)[rank]
    Ct = np.ascontiguousarray(C[rows.start: rows.stop])
    St = np.zeros_like(Ct)
    backend = OracleBackend(orc, Ct, St, 8.0, 32.0)
    order = tiling.run_tiled_sgm(backend, rank, world, dist, pg_down, pg_up)
    assert backend.calls[0] == ("h", True, False) and backend.calls[-1][2] is True
    assert [c[0] for c in backend.calls[1:]] == order
    # image-halo exchange helper
    img = torch.arange(H * W, dtype=torch.float32).reshape(H, W)[rows.start: rows.stop].contiguous()
    ext, top = tiling.exchange_image_halo(img, 2, rank, world, dist)
    lo = max(rows.start - 2, 0)
    hi = min(rows.stop + 2, H)
    assert torch.equal(ext, torch.arange(H * W, dtype=torch.float32).reshape(H, W)[lo:hi]) and top == rows.start - lo
    np.save(os.path.join(tmpdir, f"S{rank}.npy"), St)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_tiled_sgm_over_gloo_matches_untiled(world, tmp_path, oracle):
    H, W, D = 11, 9, 6
    port = _free_port()
    mp.spawn(_worker, args=(world, port, H, W, D, str(tmp_path)), nprocs=world, join=True)
    g = np.random.default_rng(5)
    C = g.integers(0, 26, (H, W, D)).astype(np.float32)
    whole = oracle.sgm_cost_volume(C, 8, 32, cmax=25)
    tiled = np.concatenate([np.load(tmp_path / f"S{r}.npy") for r in range(world)])
    np.testing.assert_array_equal(tiled, whole)


def _oracle_local(orc, left, right, dmin, dmax, cbca):
    cv, _ = orc.census_cost_volume(left, right, 5, dmin, dmax)
    if cbca:
        cv = orc.cbca_cost_volume(left, right, cv, 2, dmin, cbca[0], cbca[1])
        cv = cv[0] if isinstance(cv, tuple) else cv
    return orc.wta(cv, np.arange(dmin, dmax + 1))[0]


def _local_worker(rank, world, port, H, W, dmin, dmax, cbca, tmpdir):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle as orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.random.default_rng(9)
    left = g.integers(0, 60, (H, W)).astype(np.float32)
    right = g.integers(0, 60, (H, W)).astype(np.float32)
    rows = tiling.split_rows(H, world)[rank]

    def compute(l_ext, r_ext):
        return torch.from_numpy(_oracle_local(orc, l_ext.numpy(), r_ext.numpy(), dmin, dmax, cbca))

    pipe = tiling.TiledLocalPipeline(len(rows), W, dmin, dmax, rank, world, dist, "census", 5, cbca=cbca, compute=compute)
    assert pipe.halo == (7 if cbca else 2)
    disp = pipe.run(torch.from_numpy(left[rows.start: rows.stop].copy()), torch.from_numpy(right[rows.start: rows.stop].copy()))
    np.save(os.path.join(tmpdir, f"D{rank}.npy"), disp.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,cbca", [(2, None), (2, (5, 30.0)), (3, (5, 30.0))])
def test_tiled_local_pipeline_over_gloo_matches_untiled(world, cbca, tmp_path, oracle):
    """Row tiles of Census [-> CBCA] -> WTA with the static input halo of SURVEY 8(e) (2 rows; 7 with CBCA): bit-identical
    to the untiled oracle chain -- the halo really covers the median, the arms and the window."""
    H, W, dmin, dmax = 33, 24, -6, 2
    port = _free_port()
    mp.spawn(_local_worker, args=(world, port, H, W, dmin, dmax, cbca, str(tmp_path)), nprocs=world, join=True)
    g = np.random.default_rng(9)
    left = g.integers(0, 60, (H, W)).astype(np.float32)
    right = g.integers(0, 60, (H, W)).astype(np.float32)
    whole = _oracle_local(oracle, left, right, dmin, dmax, cbca)
    tiled = np.concatenate([np.load(tmp_path / f"D{r}.npy") for r in range(world)])
    np.testing.assert_array_equal(tiled, whole)


def test_local_halo_rows():
    assert tiling.local_halo_rows(5) == 2 and tiling.local_halo_rows(5, 5) == 7 and tiling.local_halo_rows(3, 9) == 10


# ---- column tiles: shear / un-shear host logic (the kernels need a GPU; the index maps and the neighbour exchange do not) ----
def _sheared_tiles(img, ntiles):
    """What the skewed wavefront leaves on every rank: sheared column c of row y of tile r = image column (r * Wt + c - y) mod Wg."""
    H, Wg = img.shape
    Wt = Wg // ntiles
    y = np.arange(H)[:, None]
    return np.stack([img[y, (r * Wt + np.arange(Wt)[None, :] - y) % Wg] for r in range(ntiles)])


@pytest.mark.parametrize("H,Wg,ntiles", [(5, 12, 3), (16, 16, 2), (40, 24, 4), (7, 8, 1)])
def test_unshear_gathered_inverts_the_shear(H, Wg, ntiles):
    img = np.arange(H * Wg, dtype=np.float32).reshape(H, Wg)
    tiles = _sheared_tiles(img, ntiles)
    Wt = Wg // ntiles
    for g in range(ntiles):
        np.testing.assert_array_equal(tiling.unshear_gathered(tiles, Wg, g), img[:, g * Wt:(g + 1) * Wt])
        np.testing.assert_array_equal(tiling.unshear_gathered(torch.from_numpy(tiles), Wg, g).numpy(), img[:, g * Wt:(g + 1) * Wt])


def _nb_worker(rank, world, port, n, H, Wt, tmpdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Wg = Wt * world
    # a batch is one tall sheared image: row y of image i sits at sheared row i * H + y
    imgs = [np.arange(H * Wg, dtype=np.float32).reshape(H, Wg) + 1000.0 * i for i in range(n)]
    tall = np.concatenate(imgs, axis=0)
    mine = _sheared_tiles(tall, world)[rank].reshape(n, H, Wt)
    pipe = object.__new__(tiling.ColumnTiledStereoPipeline)           # the exchange only needs these attributes
    pipe.torch, pipe.dist, pipe.rank, pipe.world, pipe.H, pipe.Wt, pipe.Wg, pipe._cache = torch, dist, rank, world, H, Wt, Wg, {}
    for _ in range(2):                                                 # twice: cached receive buffers
        out = pipe._unshear_neighbours(torch.from_numpy(np.ascontiguousarray(mine)))
    np.save(os.path.join(tmpdir, f"U{rank}.npy"), out.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,H,Wt", [(2, 1, 4, 8), (2, 3, 4, 8), (4, 2, 3, 6), (4, 5, 2, 4)])
def test_unshear_neighbours_over_gloo(world, n, H, Wt, tmp_path):
    """The neighbour exchange that brings sheared disparity tiles back to image layout (Wt a multiple of H; batches drift over
    several ranks): every rank ends with its own image tile of every image of the batch."""
    port = _free_port()
    mp.spawn(_nb_worker, args=(world, port, n, H, Wt, str(tmp_path)), nprocs=world, join=True)
    Wg = Wt * world
    for r in range(world):
        got = np.load(tmp_path / f"U{r}.npy")
        for i in range(n):
            img = np.arange(H * Wg, dtype=np.float32).reshape(H, Wg) + 1000.0 * i
            np.testing.assert_array_equal(got[i], img[:, r * Wt:(r + 1) * Wt])


@pytest.mark.parametrize("world,nimg,H,Wt", [(2, 1, 4, 8), (4, 2, 3, 6), (8, 1, 5, 4), (4, 6, 4, 8)])
def test_visited_columns_cover_the_sheared_tile(world, nimg, H, Wt):
    """visited_columns = the cyclic range of image columns a rank's sheared tile touches over a batch (what the host uploads)."""
    Wg = Wt * world
    for rank in range(world):
        pipe = object.__new__(tiling.ColumnTiledStereoPipeline)
        pipe.rank, pipe.H, pipe.Wt, pipe.Wg = rank, H, Wt, Wg
        lo, n = pipe.visited_columns(nimg)
        claimed = {(lo + k) % Wg for k in range(n)}
        y = np.arange(nimg * H)[:, None]
        touched = set(((rank * Wt + np.arange(Wt)[None, :] - y) % Wg).ravel().tolist())
        assert touched <= claimed and (len(claimed) == len(touched) or n == Wg)
