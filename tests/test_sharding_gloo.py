"""CPU, world_size 2 over gloo: the N > 1 host logic (frame / tile sharding, max-over-ranks timing, result gather)."""
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


def _worker(rank, world, port, n_frames, n_tiles, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    load_pkg()
    shard = importlib.import_module("3deecelltracker_b200.shard")
    frames = shard.frames_for_rank(n_frames, rank, world, first=1)
    tiles = shard.tile_range_for_rank(n_tiles, rank, world)
    slowest = shard.max_over_ranks(0.5 + rank)                       # rank 1 is the slow one
    local = {f: torch.full((4, 3), float(f), dtype=torch.float64) for f in frames}
    merged = shard.gather_frame_results(local, dst=0)
    all_tiles = [None] * world
    dist.all_gather_object(all_tiles, tiles)
    q.put((rank, frames, tiles, slowest, None if merged is None else sorted(merged), all_tiles,
           None if merged is None else float(sum(v.sum() for v in merged.values()))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames,n_tiles", [(7, 75), (2, 3)])
def test_two_rank_sharding(n_frames, n_tiles):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, n_tiles, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, f0, t0, s0, m0, at0, sum0), (r1, f1, t1, s1, m1, at1, _) = res
    assert sorted(f0 + f1) == list(range(1, 1 + n_frames)) and not set(f0) & set(f1)
    assert t0[0] == 0 and t0[1] == t1[0] and t1[1] == n_tiles and abs((t0[1] - t0[0]) - (t1[1] - t1[0])) <= 1
    assert s0 == s1 == 1.5                                           # max over ranks on every rank
    assert m0 == list(range(1, 1 + n_frames)) and m1 is None         # gathered on rank 0 only
    assert sum0 == sum(12.0 * f for f in range(1, 1 + n_frames))
    assert at0 == at1 == [t0, t1]


def test_sharding_rules_single_process():
    load_pkg()
    shard = importlib.import_module("3deecelltracker_b200.shard")
    for n, w in [(75, 8), (800, 8), (5, 8), (0, 2)]:
        ranges = [shard.tile_range_for_rank(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        assert max(e - b for b, e in ranges) - min(e - b for b, e in ranges) <= 1
    assert shard.frames_for_rank(256, 3, 8) == list(range(3, 256, 8))
    with pytest.raises(ValueError):
        shard.tile_range_for_rank(10, 2, 2)
    assert shard.max_over_ranks(0.25) == 0.25
