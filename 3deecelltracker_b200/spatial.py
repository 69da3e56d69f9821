"""Spatial decomposition of ONE volume over the GPUs of a box (config 3: 1024 x 1024 x 96 cut 2 x 2 x 2).

What is sharded is the segmentation half of `Tracker._save_unet_regions` (tracker.py:662-669):
`_normalize_image` (preprocess.py:170-188) followed by `unet3_prediction` (unet3d.py:203-256).  Each rank owns a
contiguous block of the RAW stack and a contiguous sub-grid of the U-Net tile grid.  Per volume the ranks exchange

  * 2 (uint16) or 4 (float32) all-reduces of a 512-word histogram: `np.median` is a global order statistic
    (`ct_select_*` in ct3d.h), and
  * one round of send/recv of raw halo boxes: the voxels a rank's tiles read (reflect padding resolved against the
    whole volume, unet3d.py:235) plus the 26-voxel reach of the two chained 27 x 27 x 1 LCN windows
    (preprocess.py:163-166: the std window reads avg, which reads its own window) that other ranks own.

Nothing else moves: tiles are independent given the normalised input, so each rank writes the centre windows of its
own tiles.  The union over ranks is bit-identical to the single-GPU result (same tiles, same kernels, window sums
in a fixed order).  `SpatialPlan` is pure host arithmetic (tested on CPU); `exchange_halo` works on any
torch.distributed backend (NCCL on GPUs, gloo on CPU tensors in the tests).
"""
import ctypes as C
import math

import torch
import torch.distributed as dist

# Reach of the LCN with its (27, 27, 1) window (preprocess.py:163-166,185): std[p] sums (v - avg)^2 over the window
# of p, and avg[q] sums v over the window of q, so one output voxel reads raw voxels up to 2 * 13 away in x and y.
LCN_RADIUS = (26, 26, 0)


def _reflect(j, n):
    """numpy.pad(mode='reflect') source index of padded coordinate j for an axis of length n."""
    if n == 1:
        return 0
    period = 2 * (n - 1)
    j %= period
    return j if j < n else period - j


def _split(n, parts):
    """np.array_split boundaries: `parts` contiguous ranges covering [0, n)."""
    base, extra = divmod(n, parts)
    edges = [0]
    for p in range(parts):
        edges.append(edges[-1] + base + (1 if p < extra else 0))
    return edges


def _intersect(a, b):
    lo = tuple(max(x, y) for x, y in zip(a[0], b[0]))
    hi = tuple(min(x, y) for x, y in zip(a[1], b[1]))
    return (lo, hi) if all(l < h for l, h in zip(lo, hi)) else None


class SpatialPlan:
    """Who owns, reads and writes what.  Boxes are (lo, hi) tuples of global voxel coordinates, hi exclusive."""

    def __init__(self, shape, grid, model_input=(160, 160, 16), shrink=(24, 24, 2), lcn_radius=LCN_RADIUS):
        self.shape = tuple(int(s) for s in shape)
        self.grid = tuple(int(g) for g in grid)
        self.model_input = tuple(int(s) for s in model_input)
        self.shrink = tuple(int(s) for s in shrink)
        self.lcn_radius = tuple(int(s) for s in lcn_radius)
        if len(self.shape) != 3 or len(self.grid) != 3 or min(self.grid) < 1 or min(self.shape) < 1:
            raise ValueError("shape and grid must be positive 3-tuples")
        self.centre = tuple(i - 2 * s for i, s in zip(self.model_input, self.shrink))
        if min(self.centre) <= 0:
            raise ValueError(f"shrink {self.shrink} too large for input size {self.model_input}")
        self.num_tiles = tuple(int(math.ceil(s / c)) for s, c in zip(self.shape, self.centre))   # unet3d.py:259-279
        self.world = self.grid[0] * self.grid[1] * self.grid[2]
        self._own_edges = [_split(self.shape[a], self.grid[a]) for a in range(3)]
        self._tile_edges = [_split(self.num_tiles[a], self.grid[a]) for a in range(3)]

    def coords(self, rank):
        if not 0 <= rank < self.world:
            raise ValueError(f"rank {rank} outside world of {self.world}")
        return rank // (self.grid[1] * self.grid[2]), (rank // self.grid[2]) % self.grid[1], rank % self.grid[2]

    def owned_box(self, rank):
        """Raw voxels resident on `rank` before the exchange."""
        c = self.coords(rank)
        return (tuple(self._own_edges[a][c[a]] for a in range(3)), tuple(self._own_edges[a][c[a] + 1] for a in range(3)))

    def tile_box(self, rank):
        """Sub-grid of the tile grid run by `rank` (tile_lo, tile_hi)."""
        c = self.coords(rank)
        return (tuple(self._tile_edges[a][c[a]] for a in range(3)), tuple(self._tile_edges[a][c[a] + 1] for a in range(3)))

    def has_tiles(self, rank):
        lo, hi = self.tile_box(rank)
        return all(l < h for l, h in zip(lo, hi))

    def norm_box(self, rank):
        """Normalised voxels the rank's tiles read (bounding interval of the reflected coordinates per axis)."""
        if not self.has_tiles(rank):
            return None
        tlo, thi = self.tile_box(rank)
        lo, hi = [], []
        for a in range(3):
            first = tlo[a] * self.centre[a] - self.shrink[a]
            last = (thi[a] - 1) * self.centre[a] - self.shrink[a] + self.model_input[a]
            idx = [_reflect(j, self.shape[a]) for j in range(first, last)]
            lo.append(min(idx))
            hi.append(max(idx) + 1)
        return tuple(lo), tuple(hi)

    def raw_box(self, rank):
        """Raw voxels the rank needs: norm_box grown by the LCN reach, clipped to the volume."""
        nb = self.norm_box(rank)
        if nb is None:
            return None
        return (tuple(max(l - r, 0) for l, r in zip(nb[0], self.lcn_radius)),
                tuple(min(h + r, s) for h, r, s in zip(nb[1], self.lcn_radius, self.shape)))

    def out_box(self, rank):
        """Probability voxels the rank produces (union of its tiles' centre windows, clipped to the volume)."""
        if not self.has_tiles(rank):
            return None
        tlo, thi = self.tile_box(rank)
        return (tuple(l * c for l, c in zip(tlo, self.centre)),
                tuple(min(h * c, s) for h, c, s in zip(thi, self.centre, self.shape)))

    def transfers(self):
        """[(src, dst, box)]: raw voxels owned by src that dst needs; at most one box per ordered pair."""
        out = []
        for dst in range(self.world):
            need = self.raw_box(dst)
            if need is None:
                continue
            for src in range(self.world):
                if src == dst:
                    continue
                box = _intersect(need, self.owned_box(src))
                if box is not None:
                    out.append((src, dst, box))
        return out

    def halo_bytes(self, rank, itemsize=2):
        """Bytes `rank` receives per volume."""
        return sum(itemsize * math.prod(h - l for l, h in zip(*box)) for s, d, box in self.transfers() if d == rank)


def _slices(box, origin):
    return tuple(slice(l - o, h - o) for l, h, o in zip(box[0], box[1], origin))


def _wire(t):
    """The NCCL process group takes neither uint16 nor int16: ship the bytes (buffers are contiguous)."""
    return t.view(torch.uint8) if t.dtype in (torch.uint16, torch.int16) else t


def exchange_halo(owned, plan, rank, group=None):
    """owned: this rank's raw block (tensor shaped like plan.owned_box(rank), any device) -> the block
    plan.raw_box(rank) with the halo filled in from the other ranks (None when the rank runs no tiles; it still
    serves its neighbours).  One batch of point-to-point transfers (ncclSend/ncclRecv grouped on GPUs)."""
    return finish_halo(*start_halo(owned, plan, rank, group))


def start_halo(owned, plan, rank, group=None):
    """Post the sends / receives of `exchange_halo` and return without waiting, so that the caller can enqueue
    independent work (the global-median all-reduces run on a different NCCL communicator and stream) behind them."""
    own = plan.owned_box(rank)
    if tuple(owned.shape) != tuple(h - l for l, h in zip(*own)):
        raise ValueError(f"rank {rank} owns box {own}, got a block of shape {tuple(owned.shape)}")
    need = plan.raw_box(rank)
    ext = None
    if need is not None:
        ext = torch.empty(tuple(h - l for l, h in zip(*need)), dtype=owned.dtype, device=owned.device)
        mine = _intersect(need, own)
        if mine is not None:
            ext[_slices(mine, need[0])] = owned[_slices(mine, own[0])]
    ops, recvs, keep = [], [], []
    for src, dst, box in plan.transfers():
        if src == rank:
            buf = owned[_slices(box, own[0])].contiguous()
            keep.append(buf)
            ops.append(dist.P2POp(dist.isend, _wire(buf), dst, group=group))
        elif dst == rank:
            buf = torch.empty(tuple(h - l for l, h in zip(*box)), dtype=owned.dtype, device=owned.device)
            recvs.append((box, buf))
            ops.append(dist.P2POp(dist.irecv, _wire(buf), src, group=group))
    reqs = dist.batch_isend_irecv(ops) if ops else []
    return ext, need, recvs, reqs, keep


def finish_halo(ext, need, recvs, reqs, keep):
    for req in reqs:
        req.wait()
    for box, buf in recvs:
        ext[_slices(box, need[0])] = buf
    return ext


def distributed_median(owned_dev, total_count, group=None):
    """np.median over the voxels of all ranks (preprocess.py:181) -> 1-element float64 CUDA tensor, identical on
    every rank.  Lock-step radix select: one 512-word all-reduce per 8-bit digit."""
    from . import _lib
    from ._device import stream_ptr
    from .preprocess import _DTYPES
    lib = _lib.lib()
    dtype = _DTYPES[owned_dev.dtype]
    owned_dev = owned_dev.contiguous()
    words = (lib.ct_select_state_bytes() + 3) // 4
    state = torch.zeros(words + 64, dtype=torch.int32, device=owned_dev.device)
    off = lib.ct_select_hist_offset() // 4
    hist = state[off:off + 512]
    med = torch.empty(1, dtype=torch.float64, device=owned_dev.device)
    _lib.check(lib.ct_select_begin(state.data_ptr(), int(total_count), stream_ptr()))
    for p in range(lib.ct_select_passes(dtype)):
        _lib.check(lib.ct_select_hist(owned_dev.data_ptr(), dtype, owned_dev.numel(), state.data_ptr(), p, stream_ptr()))
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(hist, group=group)
        _lib.check(lib.ct_select_scan(state.data_ptr(), dtype, p, stream_ptr()))
    _lib.check(lib.ct_select_finish(state.data_ptr(), dtype, med.data_ptr(), stream_ptr()))
    return med


def segment_block(owned_dev, plan, rank, model, noise_level, group=None, median=None, marks=None):
    """The rank's share of `_normalize_image` + `unet3_prediction` for one volume.

    owned_dev: the rank's raw block on its GPU.  Returns (prob_block, out_box): float32 probabilities of
    plan.out_box(rank), or (None, None) for a rank without tiles.  The halo sends / receives are posted first and the
    lock-step radix select of the global median (histogram all-reduces) runs while they are in flight.
    `marks`: optional list that receives (name, CUDA event) pairs recorded on the current stream between the phases."""
    from .preprocess import normalize_block_device

    def mark(name):
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    mark("start")
    halo = start_halo(owned_dev, plan, rank, group)
    if median is None:
        median = distributed_median(owned_dev, math.prod(plan.shape), group)
    mark("median")
    ext = finish_halo(*halo)
    mark("halo")
    if ext is None:
        return None, None
    need, out = plan.raw_box(rank), plan.out_box(rank)
    norm = normalize_block_device(ext, noise_level, median)
    tlo, thi = plan.tile_box(rank)
    mark("lcn")
    prob = model.prediction_block_device(norm, need[0], plan.shape, plan.shrink, tlo, thi, out[0],
                                         tuple(h - l for l, h in zip(*out)))
    mark("unet")
    return prob, out
