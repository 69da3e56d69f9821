"""CPU: the spatial decomposition plan (config 3) and its halo exchange over gloo.

The plan is checked against a brute-force restatement of what unet3_prediction reads (unet3d.py:235-249: reflect
pre-pad, tiles at a stride of the centre size) and what the two chained 27 x 27 x 1 LCN windows reach (preprocess.py:163-166: 26 voxels);
the exchange is run for real with 2 and 8 gloo ranks and compared with slices of the whole volume."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_pkg


def _spatial():
    load_pkg()
    return importlib.import_module("3deecelltracker_b200.spatial")


def _brute_force_reads(shape, model_input, shrink, tile_lo, tile_hi):
    """Per axis: the set of source indices np.pad(..., 'reflect') + the tile slicing touch (unet3d.py:235,247)."""
    out = []
    for a in range(3):
        c = model_input[a] - 2 * shrink[a]
        num = -(-shape[a] // c)
        padded = np.pad(np.arange(shape[a]), (shrink[a], num * c - shape[a] + shrink[a]), mode="reflect")
        idx = set()
        for t in range(tile_lo[a], tile_hi[a]):
            idx.update(padded[t * c:t * c + model_input[a]].tolist())
        out.append(idx)
    return out


@pytest.mark.parametrize("shape,grid,model_input,shrink", [
    ((1024, 1024, 96), (2, 2, 2), (160, 160, 16), (24, 24, 2)),      # config 3
    ((512, 512, 35), (2, 2, 2), (160, 160, 16), (24, 24, 2)),
    ((300, 170, 21), (2, 1, 1), (160, 160, 16), (24, 24, 2)),
    ((100, 90, 40), (1, 2, 4), (64, 64, 64), (8, 8, 8)),
    ((50, 50, 5), (2, 2, 2), (160, 160, 16), (24, 24, 2)),           # fewer tiles than ranks: idle ranks
])
def test_plan_covers_what_the_tiles_read(shape, grid, model_input, shrink):
    sp = _spatial()
    plan = sp.SpatialPlan(shape, grid, model_input, shrink)
    owned = np.zeros(shape, np.int32)
    written = np.zeros(shape, np.int32)
    tiles = 0
    for r in range(plan.world):
        lo, hi = plan.owned_box(r)
        owned[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] += 1
        if not plan.has_tiles(r):
            assert plan.norm_box(r) is None and plan.raw_box(r) is None and plan.out_box(r) is None
            continue
        tlo, thi = plan.tile_box(r)
        tiles += int(np.prod([h - l for l, h in zip(tlo, thi)]))
        reads = _brute_force_reads(shape, model_input, shrink, tlo, thi)
        nlo, nhi = plan.norm_box(r)
        rlo, rhi = plan.raw_box(r)
        for a in range(3):
            assert min(reads[a]) == nlo[a] and max(reads[a]) + 1 == nhi[a]
            # LCN window of every normalised voxel the tiles read, clipped to the volume (zero padding beyond)
            assert rlo[a] == max(nlo[a] - plan.lcn_radius[a], 0) and rhi[a] == min(nhi[a] + plan.lcn_radius[a], shape[a])
        olo, ohi = plan.out_box(r)
        written[olo[0]:ohi[0], olo[1]:ohi[1], olo[2]:ohi[2]] += 1
    assert (owned == 1).all() and (written == 1).all()               # partitions of the volume
    assert tiles == int(np.prod(plan.num_tiles))
    for src, dst, (lo, hi) in plan.transfers():
        assert src != dst
        o, n = plan.owned_box(src), plan.raw_box(dst)
        for a in range(3):
            assert o[0][a] <= lo[a] < hi[a] <= o[1][a] and n[0][a] <= lo[a] < hi[a] <= n[1][a]
    pairs = [(s, d) for s, d, _ in plan.transfers()]
    assert len(pairs) == len(set(pairs))


def test_config3_numbers():
    """10 x 10 x 8 tiles, 100 per GPU; the halo is a few percent of the owned block."""
    plan = _spatial().SpatialPlan((1024, 1024, 96), (2, 2, 2))
    assert plan.num_tiles == (10, 10, 8)
    for r in range(8):
        tlo, thi = plan.tile_box(r)
        assert np.prod([h - l for l, h in zip(tlo, thi)]) == 100
    assert plan.owned_box(0) == ((0, 0, 0), (512, 512, 48))
    assert plan.norm_box(0) == ((0, 0, 0), (584, 584, 50)) and plan.raw_box(0) == ((0, 0, 0), (610, 610, 50))
    assert plan.norm_box(7) == ((536, 536, 46), (1024, 1024, 96)) and plan.raw_box(7) == ((510, 510, 46), (1024, 1024, 96))
    own_bytes = 512 * 512 * 48 * 2
    assert 0 < plan.halo_bytes(0) < 0.5 * own_bytes
    assert plan.halo_bytes(7) == (514 * 514 * 50 - 512 * 512 * 48) * 2      # 2 voxels in x, y and z


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shape, grid, model_input, shrink, dtype, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sp = _spatial()
    plan = sp.SpatialPlan(shape, grid, model_input, shrink)
    vol = np.random.default_rng(5).integers(0, 60000, shape).astype(dtype)     # same volume on every rank
    lo, hi = plan.owned_box(rank)
    owned = np.ascontiguousarray(vol[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]])
    t = torch.from_numpy(owned.view(np.int16)).view(torch.uint16) if dtype == np.uint16 else torch.from_numpy(owned)
    ext = sp.exchange_halo(t, plan, rank)
    need = plan.raw_box(rank)
    if need is None:
        ok = ext is None
    else:
        want = vol[need[0][0]:need[1][0], need[0][1]:need[1][1], need[0][2]:need[1][2]]
        got = ext.view(torch.int16).numpy().view(np.uint16) if dtype == np.uint16 else ext.numpy()
        ok = got.shape == want.shape and np.array_equal(got, want)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("shape,grid,model_input,shrink,dtype", [
    ((300, 170, 21), (2, 1, 1), (160, 160, 16), (24, 24, 2), np.uint16),
    ((120, 110, 40), (2, 2, 2), (64, 64, 16), (8, 8, 2), np.uint16),
    ((60, 50, 30), (1, 1, 2), (32, 32, 16), (4, 4, 2), np.float32),
])
def test_halo_exchange_over_gloo(shape, grid, model_input, shrink, dtype):
    world, port = int(np.prod(grid)), _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, grid, model_input, shrink, dtype, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(r, True) for r in range(world)]


def test_plan_rejects_bad_arguments():
    sp = _spatial()
    with pytest.raises(ValueError):
        sp.SpatialPlan((10, 10, 10), (2, 2, 2), (16, 16, 16), (8, 8, 8))
    plan = sp.SpatialPlan((64, 64, 16), (2, 1, 1), (32, 32, 16), (4, 4, 2))
    with pytest.raises(ValueError):
        plan.coords(2)
    with pytest.raises(ValueError):
        sp.exchange_halo(torch.zeros((3, 3, 3)), plan, 0)
