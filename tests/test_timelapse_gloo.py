"""CPU, world_size 2 and 3 over gloo: the frame-parallel time-lapse driver (timelapse.py, BASELINE config 4).
The GPU stages (segmentation + watershed, FFN + PR-GLS fit, replay) are replaced by deterministic CPU stand-ins
with the same data flow (a volume -> a point set whose size varies; a pair of point sets -> 5 x (points, beta, C);
replay = state carried from volume to volume), so what is tested is the sharding itself: contiguous blocks, the
boundary exchange of one point set per rank (a block's first point set travels back to the rank before it, which
fits the straddling pair), the padded all-gather of the fitted transforms and the sequential
replay on rank 0 -- the result must equal the single-process run bit for bit."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_pkg


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _points(t):
    g = torch.Generator().manual_seed(1000 + t)
    n = 21 + (7 * t) % 13                                     # point-set size varies from volume to volume
    return torch.rand((n, 3), dtype=torch.float64, generator=g) * 100.0


def _make_tracker(rank, world):
    tl = importlib.import_module("3deecelltracker_b200.timelapse")

    class Stub(tl.TimelapseTracker):
        def local_fits(self, frames, lo, hi, sink=None, boundary=None):
            pts = [frames(t) for t in range(lo, hi)]
            fits = [self.fit(a, b) for a, b in zip(pts, pts[1:])]
            nxt = boundary(pts[0]) if boundary is not None else None      # first point set of the next rank's block
            if nxt is not None:
                fits.append(self.fit(pts[-1], nxt))
            return pts[0], fits

        def fit(self, prev_pts, cur_pts):
            out, inter = [], prev_pts
            for i in range(5):
                beta = 300.0 * 0.8 ** i
                coef = (torch.sin(inter.t() * (i + 1)) + cur_pts.mean()) * 1e-3          # (3, N)
                out.append((inter, beta, coef))
                inter = inter + coef.t()
            return out

        def replay(self, fit, tracked):
            for inter, beta, coef in fit:
                w = torch.exp(-torch.cdist(tracked, inter) ** 2 / (2 * beta * beta))    # (L, N)
                tracked = tracked + w @ coef.t()
            return tracked

    return Stub(pipe=None, rank=rank, world=world)


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    load_pkg()
    asked = []

    def frames(t):
        asked.append(t)
        return _points(t)

    out = _make_tracker(rank, world).run(frames, n_frames)
    q.put((rank, asked, None if out is None else [o.clone() for o in out]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(2, 9), (3, 8), (2, 2), (3, 3)])
def test_timelapse_sharding_equals_single_process(world, n_frames):
    load_pkg()
    want = _make_tracker(0, 1).run(_points, n_frames)
    assert len(want) == n_frames - 1
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seen = []
    for rank, asked, out in res:
        seen += asked
        assert (out is not None) == (rank == 0)
    assert sorted(seen) == list(range(n_frames))                     # every volume segmented exactly once
    got = res[0][2]
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert torch.equal(a, b)
