"""Row-tiled multi-GPU execution of the hot path (one process per GPU, ``torch.distributed``).

The image is cut into horizontal tiles, one per rank (rank 0 = top).  Census / SAD / ZNCC / CBCA / WTA
shard by rows with a small halo of INPUT image rows fetched once from the neighbours
(``exchange_image_halo``) -- no collective on the data path.  SGM has one real exchange step: the
downward paths (S, SE, SW) need the last row's path states ``L_r`` of the tile above, the upward paths
(N, NE, NW) the first row's states of the tile below.  They travel as point-to-point messages of
``3 x W x D`` floats per wave between neighbours only (NCCL send/recv over NVLink; gloo in the CPU
tests), which keeps the result bit-identical to the single-GPU run for integer-valued costs --
unlike Pandora's own tiling convention, a fixed 40-pixel margin (marge.py:85-101 in the reference),
which is an approximation.

Schedule: the horizontal directions are tile-local and run first.  The two vertical waves start at
opposite ends (rank 0 downward, rank N-1 upward); every rank runs its two vertical groups (one strip-sweep
kernel launch each: S+SE+SW and N+NE+NW) in the order their halo can arrive (``group_order``), which makes
the dependency graph acyclic -- a rank never waits for a neighbour that is waiting for it.  Each wave uses
its own process group so that the two flows between a pair of neighbours are not serialised on one
communicator; one message per border and wave carries the three (W, D) state planes.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

DIR_NAMES = ("E", "W", "S", "SE", "SW", "N", "NE", "NW")
HORIZONTAL, DOWN, UP = (0, 1), (2, 3, 4), (5, 6, 7)


def split_rows(total_rows: int, world: int) -> List[range]:
    """Contiguous row ranges, as even as possible, top tile first."""
    base, rem = divmod(total_rows, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append(range(start, start + n))
        start += n
    return out


def group_order(rank: int, world: int) -> List[int]:
    """Order in which ``rank`` runs its two vertical groups (0 = downward S/SE/SW, 1 = upward N/NE/NW): by
    earliest possible halo arrival (the downward wave reaches rank r at step r, the upward wave at step
    world-1-r), downward first on ties.  A rank that runs "up" first sits in the lower half and one that runs
    "down" first in the upper half, so no two ranks can wait on each other (acyclic dependency graph)."""
    return [0, 1] if rank <= world - 1 - rank else [1, 0]


def direction_order(rank: int, world: int) -> List[int]:
    """The six vertical directions in execution order (groups stay together, see ``group_order``)."""
    return [d for g in group_order(rank, world) for d in (DOWN, UP)[g]]


class SgmBackend:
    """What the tile executor needs from a device: run the horizontal pair or one vertical group on the local tile."""

    def new_halo(self):  # (3, W, D) float32 buffer: the three path states of one boundary row
        raise NotImplementedError

    def run_horizontal(self, init: bool) -> None:
        raise NotImplementedError

    def agree_on_path(self, dist) -> None:
        """Make every rank take the same numeric path for the vertical groups (no-op for backends with one path)."""

    def run_group(self, group: int, final: bool, halo_in, halo_out) -> None:
        raise NotImplementedError


class EngineSgmBackend(SgmBackend):
    """CUDA backend: ``Engine.sgm`` restricted to one direction group per call (one strip-sweep launch each)."""

    def __init__(self, engine, cv, out, p1, p2, invalid_value, overcounting=False, disp=None, flags=None, dmin=0,
                 invalid_disparity=-9999.0):
        self.eng, self.cv, self.out = engine, cv, out
        self.args = (p1, p2, invalid_value, overcounting)
        self.disp, self.flags, self.dmin, self.invalid_disparity = disp, flags, dmin, invalid_disparity

    def new_halo(self):
        _, W, D = self.cv.shape
        return self.eng.empty((3, W, D))

    def run_horizontal(self, init: bool) -> None:
        self.eng.sgm(self.cv, *self.args, out=self.out, dir_mask=0x03, init_final=1 if init else 0, packed=init)

    def agree_on_path(self, dist) -> None:
        # the packed integer path qualifies per tile (device flag); halos are only compatible when all ranks agree
        if dist is not None:
            _, W, D = self.cv.shape
            dist.all_reduce(self.eng.sgm_path_flag(W, D), op=dist.ReduceOp.MAX)
        # float E + W, executed on the device only where the (agreed) flag says the packed path is off
        self.eng.sgm(self.cv, *self.args, out=self.out, dir_mask=0x03, init_final=1, packed=True, float_only=True)

    def run_group(self, group, final, halo_in, halo_out) -> None:
        if group == 0:
            kw = {"halo_in_top": halo_in, "halo_out_bottom": halo_out}
        else:
            kw = {"halo_in_bottom": halo_in, "halo_out_top": halo_out}
        fuse = final and self.disp is not None
        self.eng.sgm(self.cv, *self.args, out=self.out, fuse_wta=fuse, dmin=self.dmin, invalid_disparity=self.invalid_disparity,
                     dir_mask=0x1C if group == 0 else 0xE0, init_final=2 if final else 0, disp=self.disp, flags=self.flags,
                     packed=True, **kw)


def run_tiled_sgm(backend: SgmBackend, rank: int, world: int, dist=None, pg_down=None, pg_up=None, order: Optional[List[int]] = None):
    """Execute the 8 directions on this rank's tile with halo hand-over to the neighbours.

    ``dist`` is ``torch.distributed`` (or None when world == 1); ``pg_down`` / ``pg_up`` are two process
    groups spanning all ranks (one per wave, so the two flows between a pair of neighbours are not serialised
    on one communicator).  One message of (3, W, D) floats per tile border and wave.  Returns the group order.
    """
    order = group_order(rank, world) if order is None else order
    has_up_nb, has_down_nb = rank > 0, rank < world - 1
    send_work, keep = [], []
    backend.run_horizontal(True)
    backend.agree_on_path(dist if world > 1 else None)
    for i, g in enumerate(order):
        # The receive is posted only now, never ahead of local work: an NCCL receive kernel that sits on the GPU
        # waiting for its peer would keep the cooperative strip-sweep launch of the OTHER group from becoming
        # resident, and two ranks doing that to each other deadlock.  Posted here, the only kernel ordered behind
        # the receive is the sweep that needs its data anyway.
        halo_in = None
        if world > 1 and ((g == 0 and has_up_nb) or (g == 1 and has_down_nb)):
            halo_in = backend.new_halo()
            dist.irecv(halo_in, src=rank - 1 if g == 0 else rank + 1, group=pg_down if g == 0 else pg_up).wait()
        send_to = None
        if world > 1 and ((g == 0 and has_down_nb) or (g == 1 and has_up_nb)):
            send_to = rank + 1 if g == 0 else rank - 1
        halo_out = backend.new_halo() if send_to is not None else None
        backend.run_group(g, i == len(order) - 1, halo_in, halo_out)
        if send_to is not None:
            keep.append(halo_out)
            send_work.append(dist.isend(halo_out, dst=send_to, group=pg_down if g == 0 else pg_up))
    for w in send_work:
        w.wait()
    return order


def warm_up_links(rank: int, world: int, dist, pg_down, pg_up, device=None) -> None:
    """Establish the neighbour connections of both wave groups with one batched (grouped) dummy exchange each.

    The first send / receive between two ranks on an NCCL communicator sets the transport up and BLOCKS THE HOST
    until the peer makes the matching call.  In ``run_tiled_sgm`` the first calls of two neighbours are on
    different groups (rank r sends "down" while rank r+1 sends "up"), so without this warm-up both hosts would
    sit in their first send forever.  A batched exchange posts each rank's send and receive together."""
    import torch  # noqa: PLC0415

    if world == 1:
        return
    for pg, to, frm in ((pg_down, rank + 1, rank - 1), (pg_up, rank - 1, rank + 1)):
        ops, bufs = [], []
        if 0 <= to < world:
            bufs.append(torch.zeros(4, dtype=torch.float32, device=device))
            ops.append(dist.P2POp(dist.isend, bufs[-1], to, pg))
        if 0 <= frm < world:
            bufs.append(torch.zeros(4, dtype=torch.float32, device=device))
            ops.append(dist.P2POp(dist.irecv, bufs[-1], frm, pg))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if device is not None and torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)


def exchange_image_halo(tile, half: int, rank: int, world: int, dist, group=None):
    """Fetch ``half`` rows of INPUT image from each neighbour; returns (extended tile, rows added on top).

    ``tile`` is a (rows, W) tensor (CPU for gloo, CUDA for nccl).  Matching-cost windows that straddle a
    tile border then see exactly the pixels the untiled run sees."""
    import torch  # noqa: PLC0415

    if world == 1 or half == 0:
        return tile, 0
    W = tile.shape[1]
    top = torch.empty((half, W), dtype=tile.dtype, device=tile.device) if rank > 0 else None
    bot = torch.empty((half, W), dtype=tile.dtype, device=tile.device) if rank < world - 1 else None
    ops = []
    if rank > 0:
        ops += [dist.P2POp(dist.isend, tile[:half].contiguous(), rank - 1, group), dist.P2POp(dist.irecv, top, rank - 1, group)]
    if rank < world - 1:
        ops += [dist.P2POp(dist.isend, tile[-half:].contiguous(), rank + 1, group), dist.P2POp(dist.irecv, bot, rank + 1, group)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    parts = [p for p in (top, tile, bot) if p is not None]
    return torch.cat(parts, dim=0), (half if rank > 0 else 0)


class TiledStereoPipeline:
    """Census -> SGM -> WTA on this rank's row tile of a tall image (the C3/C4 configuration, row-tiled)."""

    def __init__(self, tile_rows: int, W: int, dmin: int, dmax: int, rank: int, world: int, dist=None, window: int = 5,
                 p1: float = 8.0, p2: float = 32.0, invalid_disparity: float = -9999.0, device: Optional[str] = None):
        import torch  # noqa: PLC0415

        from ._common import get_engine  # noqa: PLC0415

        self.torch, self.dist = torch, dist
        self.eng = get_engine(device)
        self.rank, self.world = rank, world
        self.rows, self.W, self.dmin, self.dmax, self.window = tile_rows, W, dmin, dmax, window
        self.D = dmax - dmin + 1
        self.p1, self.p2, self.invalid_disparity = p1, p2, invalid_disparity
        self.half = window // 2
        self.top = self.half if rank > 0 else 0
        self.bot = self.half if rank < world - 1 else 0
        ext = tile_rows + self.top + self.bot
        self.cv_ext = self.eng.empty((ext, W, self.D))
        self.S = self.eng.empty((tile_rows, W, self.D))
        self.disp = self.eng.empty((tile_rows, W))
        self.flags = self.eng.empty((tile_rows, W), torch.uint8)
        self.pg_down = dist.new_group(list(range(world))) if world > 1 else None
        self.pg_up = dist.new_group(list(range(world))) if world > 1 else None
        warm_up_links(rank, world, dist, self.pg_down, self.pg_up, self.eng.device)

    def run(self, left_tile, right_tile):
        """``left_tile`` / ``right_tile``: this rank's (rows, W) float32 device tensors.  Returns the disparity tile."""
        e = self.eng
        l_ext, _ = exchange_image_halo(left_tile, self.half, self.rank, self.world, self.dist)
        r_ext, _ = exchange_image_halo(right_tile, self.half, self.rank, self.world, self.dist)
        e.census(l_ext.contiguous(), r_ext.contiguous(), self.window, self.dmin, self.dmax, out=self.cv_ext)
        cv = self.cv_ext[self.top: self.top + self.rows]
        # rows that are image border for the whole image only: tile-internal borders were computed from the halo
        backend = EngineSgmBackend(e, cv, self.S, self.p1, self.p2, float(self.window**2) + self.p2 + 1.0, False, self.disp,
                                   self.flags, self.dmin, self.invalid_disparity)
        run_tiled_sgm(backend, self.rank, self.world, self.dist, self.pg_down, self.pg_up)
        return self.disp


def local_halo_rows(window: int, cbca_distance: Optional[int] = None) -> int:
    """Rows of INPUT image a row tile needs from each neighbour so that a pipeline WITHOUT SGM is bit-identical to the
    untiled run.  Matching cost / WTA: the half window.  CBCA on top: a pixel's vertical arms reach
    ``distance - 1`` rows (aggregation.cpp:259-313), the horizontal sums of those rows need their costs (half window
    more) and their horizontal arms, taken on the 3x3-median image (one more row): (distance - 1) + half + 1 -- 7 rows
    for the C2 configuration.  The vertical arms of the halo rows themselves are wrong (truncated) and never used."""
    half = window // 2
    return half if not cbca_distance else (cbca_distance - 1) + half + 1


class TiledLocalPipeline:
    """Row tiles of a pipeline without SGM (Census / SAD / SSD / ZNCC [-> CBCA] -> WTA; C1 / C2): every rank extends its
    tile by ``local_halo_rows`` input rows from each neighbour (one point-to-point exchange of the two images per pair,
    no collective on the data path), runs the single-GPU pipeline on the extended tile and keeps its own rows.

    ``compute(left_ext, right_ext) -> disparity_ext`` replaces the device pipeline in the CPU (gloo) tests."""

    def __init__(self, tile_rows: int, W: int, dmin: int, dmax: int, rank: int, world: int, dist=None, method: str = "census",
                 window: int = 5, cbca=None, invalid_disparity: float = -9999.0, device: Optional[str] = None, compute=None):
        self.rank, self.world, self.dist = rank, world, dist
        self.rows, self.W = tile_rows, W
        self.halo = local_halo_rows(window, cbca[0] if cbca else None)
        if world > 1 and tile_rows < self.halo:
            raise ValueError(f"row tiles of {tile_rows} rows are shorter than the {self.halo}-row halo")
        self.top = self.halo if rank > 0 else 0
        self.bot = self.halo if rank < world - 1 else 0
        self.compute = compute
        self.pipe = None
        if compute is None:
            from .pipeline import StereoPipeline  # noqa: PLC0415

            self.pipe = StereoPipeline(tile_rows + self.top + self.bot, W, dmin, dmax, method, window, cbca=cbca,
                                       invalid_disparity=invalid_disparity, device=device)

    def run(self, left_tile, right_tile):
        """``left_tile`` / ``right_tile``: this rank's (rows, W) float32 tensors.  Returns the (rows, W) disparity tile."""
        l_ext, top = exchange_image_halo(left_tile, self.halo, self.rank, self.world, self.dist)
        r_ext, _ = exchange_image_halo(right_tile, self.halo, self.rank, self.world, self.dist)
        assert top == self.top and l_ext.shape[0] == self.rows + self.top + self.bot
        disp = self.compute(l_ext, r_ext) if self.compute is not None else self.pipe.run_device(l_ext.contiguous(), r_ext.contiguous())
        return disp[self.top: self.top + self.rows]


# ====================================================================================================================
# Column tiles: the skewed wavefront as ONE wave across all GPUs (pb200_census_sgm_tile)
# ====================================================================================================================
def sheared_owner(Wg: int, Wt: int, tile: int, H: int):
    """Index maps that un-shear image tile ``tile``: pixel (y, j) of the tile -- image column tile * Wt + j -- lives in the
    sheared tile of rank ``owner[y, j]`` at column ``col[y, j]`` (sheared column = (image column + row) mod Wg)."""
    y = np.arange(H, dtype=np.int64)[:, None]
    t = (tile * Wt + np.arange(Wt, dtype=np.int64)[None, :] + y) % Wg
    return t // Wt, t % Wt


def unshear_gathered(gathered, Wg: int, tile: int):
    """``gathered``: (ntiles, H, Wt[, ...]) sheared tiles of every rank (numpy or torch) -> image tile ``tile`` (H, Wt[, ...])."""
    ntiles, H, Wt = gathered.shape[:3]
    owner, col = sheared_owner(Wg, Wt, tile, H)
    rows = np.broadcast_to(np.arange(H)[:, None], owner.shape)
    if isinstance(gathered, np.ndarray):
        return gathered[owner, rows, col]
    import torch  # noqa: PLC0415

    dev = gathered.device
    return gathered[torch.from_numpy(owner).to(dev), torch.from_numpy(np.ascontiguousarray(rows)).to(dev), torch.from_numpy(col).to(dev)]


class TileLinks:
    """This rank's link buffer (peer-visible device memory, ``pb200_ipc_alloc``) and the two neighbours' buffers mapped into
    this process.  The 64-byte CUDA IPC handles travel through ``dist.all_gather_object``; nothing else is exchanged."""

    def __init__(self, lib, dist, rank: int, world: int, D: int):
        import ctypes  # noqa: PLC0415

        from . import _native  # noqa: PLC0415

        self.lib, self.rank, self.world = lib, rank, world
        self.nbytes = int(lib.pb200_tile_link_bytes(D))
        ptr, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        _native.check(lib.pb200_ipc_alloc(self.nbytes, ctypes.byref(ptr), handle))
        self.local = ptr.value
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, bytes(handle))
        self._opened = {}

        def mapped(r):
            if r == rank:
                return self.local
            if r not in self._opened:
                p = ctypes.c_void_p()
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(handles[r])
                _native.check(lib.pb200_ipc_open(buf, ctypes.byref(p)))
                self._opened[r] = p.value
            return self._opened[r]

        self.left, self.right = mapped((rank - 1) % world), mapped((rank + 1) % world)

    def close(self):
        for p in self._opened.values():
            self.lib.pb200_ipc_close(p)
        self._opened = {}
        if self.local:
            self.lib.pb200_ipc_free(self.local)
            self.local = None


class ColumnTiledStereoPipeline:
    """Census -> SGM -> WTA on images ``Wg`` columns wide, column-tiled over the ranks of one node: every rank runs the
    two skewed-wavefront passes on its ``Wt = Wg / world`` sheared columns and the border path states cross the GPU
    boundaries inside the kernels (NVLink peer stores).  Bit-identical to the one-GPU result.

    ``run(left, right)`` takes WHOLE images on the device -- one (H, Wg) pair, or a batch (n, H, Wg) whose images follow
    each other in one wave (stream throughput: the time a wave needs to cross all GPUs is paid once per pass and batch) -- and
    returns this rank's disparity tile(s) in sheared layout; ``unshear`` turns them into normal-layout image tiles."""

    def __init__(self, H: int, Wg: int, dmin: int, dmax: int, rank: int, world: int, dist, window: int = 5, p1: float = 8.0,
                 p2: float = 32.0, overcounting: bool = False, invalid_disparity: float = -9999.0, device=None):
        import torch  # noqa: PLC0415

        from ._common import get_engine  # noqa: PLC0415

        if Wg % world:
            raise ValueError(f"the image width ({Wg}) must be a multiple of the number of ranks ({world})")
        self.torch, self.dist = torch, dist
        self.eng = get_engine(device)
        self.lib = self.eng.lib
        self.H, self.Wg, self.Wt, self.dmin, self.dmax, self.D = H, Wg, Wg // world, dmin, dmax, dmax - dmin + 1
        self.rank, self.world, self.window = rank, world, window
        self.p1, self.p2, self.over, self.invalid = float(p1), float(p2), bool(overcounting), float(invalid_disparity)
        with torch.cuda.device(self.eng.device):
            self.links = TileLinks(self.lib, dist, rank, world, self.D)
        self._bufs = {}
        self.sws = self.eng._workspace("sgm", self.lib.pb200_sgm_workspace_bytes(H, self.Wt, self.D))
        self.epoch, self._prev = 0, (0, 0)                 # (epoch, rows) of the call that used the links last
        self.nimg = 1
        self.cv = self.disp = self.flags = None
        self._cache = {}
        if world > 1:
            dist.barrier()                                 # every link buffer exists and is mapped before the first wave

    def _buffers(self, n: int):
        """Volume / disparity / flag tiles and the descriptor workspace of a batch of ``n`` images."""
        if n not in self._bufs:
            e, t = self.eng, self.torch
            self._bufs[n] = (e.empty((n, self.H, self.Wt, self.D)), e.empty((n, self.H, self.Wt)), e.empty((n, self.H, self.Wt), t.uint8))
        self.cws = self.eng._workspace("census", n * self.lib.pb200_census_sgm_workspace_bytes(self.H, self.Wg, self.window, self.dmin, self.D))
        return self._bufs[n]

    def run(self, left, right):
        """``left`` / ``right``: float32 (H, Wg) or (n, H, Wg) device tensors holding at least the image columns this tile visits.
        Returns this rank's disparity tile(s) (H, Wt) / (n, H, Wt) in sheared layout (``unshear`` gives the image tiles)."""
        from . import _native  # noqa: PLC0415

        t = self.torch
        single = left.dim() == 2
        n = 1 if single else int(left.shape[0])
        assert tuple(left.shape[-2:]) == (self.H, self.Wg) and left.is_contiguous() and right.is_contiguous() and right.shape == left.shape
        cv, disp, flags = self._buffers(n)
        self.epoch = self.epoch % 65535 + 1                # 1 .. 65535: every ring word carries it, 0 means "never written"
        with t.cuda.device(self.eng.device):
            _native.check(self.lib.pb200_census_sgm_tile(
                left.data_ptr(), right.data_ptr(), n, self.H, self.Wg, self.window, self.dmin, self.D, self.p1, self.p2, int(self.over),
                self.rank, self.world, cv.data_ptr(), self.cws.data_ptr(), self.cws.numel(), self.sws.data_ptr(), self.sws.numel(),
                disp.data_ptr(), self.invalid, flags.data_ptr(), self.links.local, self.links.left, self.links.right,
                self.epoch, self._prev[0], self._prev[1], 3, t.cuda.current_stream(self.eng.device).cuda_stream))
        self._prev = (self.epoch, n * self.H)
        self.nimg = n
        self.cv, self.disp, self.flags = (cv[0], disp[0], flags[0]) if single else (cv, disp, flags)
        return self.disp

    def visited_columns(self, nimg: int = 1):
        """(lo, n): the cyclic range of image columns [lo, lo + n) mod Wg whose pixels this tile processes in a batch of ``nimg``."""
        rows = nimg * self.H
        n = min(self.Wg, self.Wt + rows - 1)
        return (self.rank * self.Wt - (rows - 1)) % self.Wg, n

    def unshear(self, tile=None):
        """Sheared tile(s) (default: the disparity tiles of the last run) -> normal-layout image tile(s) of this rank.  A batch is
        one tall sheared image: row y of image i sits at sheared row i * H + y."""
        t, dist = self.torch, self.dist
        src = self.disp if tile is None else tile
        batched = tuple(src.shape[:2]) != (self.H, self.Wt)       # (n, H, Wt[, ...]) rather than (H, Wt[, ...])
        assert tuple(src.shape[1:3] if batched else src.shape[:2]) == (self.H, self.Wt), "not a sheared tile of this pipeline"
        n = int(src.shape[0]) if batched else 1
        rows = n * self.H
        flat = src.reshape((rows, self.Wt) + tuple(src.shape[3 if batched else 2:]))
        if self.world > 1 and flat.dim() == 2 and self.Wt % self.H == 0:
            out = self._unshear_neighbours(flat.view(n, self.H, self.Wt))
            return out if batched else out[0]
        # general case (tiles that drift across several neighbours inside one image, or a volume): all-gather + one indexed read
        key = (tuple(flat.shape), flat.dtype)
        if key not in self._cache:
            owner, col = sheared_owner(self.Wg, self.Wt, self.rank, rows)
            idx = (owner * rows + np.arange(rows, dtype=np.int64)[:, None]) * self.Wt + col             # into (world, rows, Wt)
            self._cache[key] = (t.empty((self.world,) + tuple(flat.shape), dtype=flat.dtype, device=flat.device),
                                t.from_numpy(idx.reshape(-1)).to(flat.device))
        buf, idx = self._cache[key]
        if self.world > 1:
            dist.all_gather_into_tensor(buf, flat.contiguous())
        else:
            buf[0].copy_(flat)
        inner = int(np.prod(flat.shape[2:])) if flat.dim() > 2 else 1
        out = buf.view(self.world * rows * self.Wt, inner).index_select(0, idx).view(flat.shape)
        return out.view(src.shape)

    def _unshear_neighbours(self, tiles):
        """(n, H, Wt) sheared disparity tiles with Wt a multiple of H: the sheared rows of image i all lie in one block of Wt rows,
        so image tile g of image i = columns [o + y, o + y + Wt) of ranks g + q and g + q + 1, q = (i * H) // Wt, o = (i * H) % Wt:
        every rank sends the tile of image i to the two ranks that need it (NCCL point-to-point over NVLink) and reads its own
        image tile out of the two it receives."""
        t, dist = self.torch, self.dist
        n, N, r = int(tiles.shape[0]), self.world, self.rank
        key = ("nb", n)
        if key not in self._cache:
            y = t.arange(self.H, device=tiles.device)[:, None]
            j = t.arange(self.Wt, device=tiles.device)[None, :]
            self._cache[key] = (t.empty((n, 2, self.H, self.Wt), dtype=tiles.dtype, device=tiles.device),
                                [((i * self.H) % self.Wt + y + j) for i in range(n)])
        recv, idx = self._cache[key]
        ops, local = [], []
        for i in range(n):
            q = (i * self.H) // self.Wt
            for which, dst in enumerate(((r - q) % N, (r - q - 1) % N)):       # the image tiles my sheared tile contributes to
                src_rank = (r + q + which) % N                                # ... and who contributes part `which` of mine
                if dst == r:
                    local.append((i, which))
                else:
                    ops.append(dist.P2POp(dist.isend, tiles[i], dst))
                if src_rank != r:
                    ops.append(dist.P2POp(dist.irecv, recv[i, which], src_rank))
        for i, which in local:
            recv[i, which].copy_(tiles[i])
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return t.stack([t.gather(t.cat([recv[i, 0], recv[i, 1]], dim=1), 1, idx[i]) for i in range(n)])

    def close(self):
        self.links.close()
