"""Frame-parallel segment + track of a whole time-lapse over the GPUs of one box (BASELINE config 4: 256 frames of
512 x 512 x 35 over 8 GPUs; the loop of Tracker.track, tracker.py:1415-1431).

What depends on what in `track_one_vol` (tracker.py:1473-1536):
  * segmentation + watershed of volume t (tracker.py:605-650) depends on nothing but its raw stack;
  * the fit `_fit_ffn_prgls` (tracker.py:1224-1254) between volumes t-1 and t depends on their two SEGMENTED point sets;
  * only the replay `_predict_one_rep` (tracker.py:1269-1289) and the displacement bookkeeping
    (tracker.py:1179-1180, 1525-1534) carry state from volume to volume, and they are a few tiny kernels per volume.
So rank r takes a CONTIGUOUS block of volumes and streams it through its own FramePipeline (segmentation on the main
stream, fits on side streams); the only data that crosses ranks is (i) the point set of the FIRST volume of a block,
sent back to the rank that owns the previous block for the one fit that straddles the boundary (it runs beside that
block's last local fit), and (ii) the fitted transforms
(5 x (intermediate points, C) per volume, ~40 KB), gathered to rank 0, which replays them in volume order.  There is
no collective on the data path of the volumes themselves.  The result equals the single-GPU run bit for bit: the same
kernels see the same inputs (tests/test_timelapse_gloo.py for the plumbing, bench.py --verify-c4 on the GPUs).
"""
import torch
import torch.distributed as dist

from .pipeline import REP_NUM_PRGLS, FramePipeline
from .shard import block_for_rank


def _dist_on():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class TimelapseTracker:
    """`pipe`: a FramePipeline configured with `configure_watershed`.  `fit_fn`/`replay_fn`/`segment_step`/`finish`
    default to the pipeline's own stages; tests inject CPU stand-ins to exercise the sharding and gather logic."""

    def __init__(self, pipe, rank=None, world=None):
        self.pipe = pipe
        self.world = world if world is not None else (dist.get_world_size() if _dist_on() else 1)
        self.rank = rank if rank is not None else (dist.get_rank() if _dist_on() else 0)

    # ---- stages (overridable)
    def local_fits(self, frames, lo, hi, sink=None, boundary=None):
        """Stream volumes lo..hi-1 through the pipeline.  Returns (first point set, fits of the pairs (lo, lo+1) ..
        (hi-2, hi-1) and -- when `boundary` hands back the first point set of the next rank's block -- (hi-1, hi)).
        sink(t, prob, segmentation) sees every volume's device-resident results.  boundary(first point set) is called
        once, before the pipeline is flushed: the straddling fit then runs beside the block's last local fit."""
        p = self.pipe
        p.reset_raw()
        p.collect_fits = True
        try:
            for t in range(lo, hi):
                prob, seg, _ = p.step_raw(frames(t))
                if sink is not None:
                    sink(t, prob, seg)
                if t - lo == 3 and not p._reserved:              # all of the pipeline's streams exist by now
                    p.reserve_small_blocks()
            first = p._first_points
            if first is None:                                       # fewer volumes than the cell-count lag: resolve now
                p._submit_fits(p._resolve_segmented())
                first = p._first_points
            nxt = boundary(first) if boundary is not None else None
            p.flush_raw(extra_target=nxt)
            return first, list(p.collected)
        finally:
            p.collect_fits = False
            p.collected = []

    def fit(self, prev_pts, cur_pts):
        return self.pipe.fit(prev_pts, cur_pts)

    def replay(self, fit, tracked):
        return self.pipe.replay(fit, tracked)

    # ---- exchange helpers: padded (count, rows) transport over torch.distributed
    @staticmethod
    def _pack_points(pts, cap):
        buf = torch.zeros((cap + 1, 3), dtype=torch.float64, device=pts.device)
        buf[0, 0] = pts.shape[0]
        buf[1:1 + pts.shape[0]] = pts
        return buf

    @staticmethod
    def _unpack_points(buf):
        n = int(buf[0, 0].item())
        return buf[1:1 + n].clone()

    def _boundary_exchange(self, first_pts):
        """Send this block's FIRST point set to rank-1, receive rank+1's; returns the received point set or None.  The
        rank that owns volume hi-1 fits the pair (hi-1, hi): it knows the next block's first point set long before its own
        last one, so the straddling fit overlaps the tail of its block instead of following it."""
        if self.world == 1:
            return None
        device = first_pts.device
        cap = torch.tensor([first_pts.shape[0]], dtype=torch.int64, device=device)
        dist.all_reduce(cap, op=dist.ReduceOp.MAX)
        cap = int(cap.item())
        ops, recv = [], None
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, self._pack_points(first_pts, cap), self.rank - 1))
        if self.rank + 1 < self.world:
            recv = torch.zeros((cap + 1, 3), dtype=torch.float64, device=device)
            ops.append(dist.P2POp(dist.irecv, recv, self.rank + 1))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        return None if recv is None else self._unpack_points(recv)

    def _gather_fits(self, fits, cap, per_rank, device):
        """All fits of all ranks on every rank, in volume order: list of [(inter, beta, coef)] per volume pair."""
        n_rep = REP_NUM_PRGLS
        # one flat block per fit and repetition: [count, beta, 0 | inter (n,3) row-major in cap*3 | coef (3,n) row-major in
        # cap*3] -- both arrays travel in their own memory order, so unpacking is two views, no transposes
        blk = 3 + 6 * cap
        local = torch.zeros((per_rank, n_rep, blk), dtype=torch.float64, device=device)
        for k, fit in enumerate(fits):
            for i, (inter, beta, coef) in enumerate(fit):
                n = inter.shape[0]
                local[k, i, 0], local[k, i, 1] = n, beta
                local[k, i, 3:3 + 3 * n] = inter.reshape(-1)
                local[k, i, 3 + 3 * cap:3 + 3 * cap + 3 * n] = coef.reshape(-1)
        counts = torch.tensor([len(fits)], dtype=torch.int64, device=device)
        allc = [torch.zeros_like(counts) for _ in range(self.world)]
        dist.all_gather(allc, counts)
        alld = [torch.zeros_like(local) for _ in range(self.world)]
        dist.all_gather(alld, local)
        n_fits = torch.cat(allc).cpu().tolist()                           # one small download, then no more syncs
        heads = torch.stack([d[:, :, :2] for d in alld]).cpu()          # (world, per_rank, rep, [count, beta])
        out = []
        for r, d in enumerate(alld):
            for k in range(int(n_fits[r])):
                fit = []
                for i in range(n_rep):
                    n = int(heads[r, k, i, 0])
                    fit.append((d[k, i, 3:3 + 3 * n].view(n, 3), float(heads[r, k, i, 1]),
                                d[k, i, 3 + 3 * cap:3 + 3 * cap + 3 * n].view(3, n)))
                out.append(fit)
        return out

    def run(self, frames, n_frames, sink=None):
        """frames: callable t -> raw (x,y,z) CUDA tensor, asked only for the volumes of this rank's block.
        Returns on rank 0 the tracked coordinates of volumes 1..n_frames-1 as a list of (L,3) tensors (volume 0's
        centres are the tracked set, tracker.py:1124-1136); None on the other ranks."""
        if n_frames < self.world:
            raise ValueError(f"a time-lapse of {n_frames} volumes cannot be split over {self.world} ranks")
        lo, hi = block_for_rank(n_frames, self.rank, self.world)
        boundary = self._boundary_exchange if self.world > 1 else None
        first, fits = (self.local_fits(frames, lo, hi, boundary=boundary) if sink is None
                       else self.local_fits(frames, lo, hi, sink, boundary))
        device = first.device
        # padding capacity: the largest point set anywhere
        cap_local = max([first.shape[0]] + [f[0][0].shape[0] for f in fits] or [1])
        cap = torch.tensor([cap_local], dtype=torch.int64, device=device)
        if self.world > 1:
            dist.all_reduce(cap, op=dist.ReduceOp.MAX)
        cap = int(cap.item())
        # every rank but the last also carries the pair that straddles into the next block
        per_rank = 1 + max(block_for_rank(n_frames, r, self.world)[1] - block_for_rank(n_frames, r, self.world)[0]
                           for r in range(self.world))
        all_fits = fits if self.world == 1 else self._gather_fits(fits, cap, per_rank, device)
        if self.rank != 0:
            return None
        assert len(all_fits) == n_frames - 1, f"{len(all_fits)} fits for {n_frames} volumes"
        tracked, out = first, []
        for fit in all_fits:
            tracked = self.replay(fit, tracked)
            out.append(tracked)
        return out
