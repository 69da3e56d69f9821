"""Host-side sharding rules for N GPUs of one box (one process per GPU, torch.distributed for the plumbing).

Two independent ways the path shards, neither needs a data-path collective:
  * frames of a time-lapse (config 4: 256 frames over 8 GPUs) go to ranks in contiguous blocks (`block_for_rank`, used
    by timelapse.py) or round-robin (`frames_for_rank`); segmentation of a frame is independent of every other frame,
    and the per-frame fitted transforms (a few KB) are gathered to the rank that runs the sequential tracking tail
    (tracker.py:1179-1180 carries state frame to frame);
  * the tiles of ONE volume split into contiguous ranges of the (i, j, k) row-major tile grid of
    unet3_prediction (unet3d.py:246); `ct_unet3_prediction(tile_begin, tile_end)` writes only its tiles' centre
    windows, so the union over ranks is bit-identical to the single-GPU result.
"""
import torch
import torch.distributed as dist


def frames_for_rank(n_frames, rank, world, first=0):
    """Frame indices handled by `rank`: first + rank, first + rank + world, ..."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(first + rank, first + n_frames, world))


def block_for_rank(n_frames, rank, world):
    """Contiguous [lo, hi) block of a time-lapse handled by `rank` (np.array_split boundaries): the streaming pipeline
    of a rank needs consecutive volumes, so only one fit per rank straddles two blocks (timelapse.py)."""
    return tile_range_for_rank(n_frames, rank, world)


def tile_range_for_rank(n_tiles, rank, world):
    """Contiguous [begin, end) of the row-major tile list; ranges differ in length by at most one tile."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_tiles, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(seconds, device=None):
    """Timing rule of the bench contract: the job time is the slowest rank's."""
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_frame_results(local, dst=0):
    """Gather {frame index: (L,3) float64 tensor} dictionaries to rank `dst` (returns the merged dict there, None
    elsewhere).  Payloads are a few KB per frame; uses the default process group (NCCL on GPUs, gloo on CPU)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(local)
    world, rank = dist.get_world_size(), dist.get_rank()
    payload = {int(k): v.detach().cpu() for k, v in local.items()}
    out = [None] * world if rank == dst else None
    dist.gather_object(payload, out, dst=dst)
    if rank != dst:
        return None
    merged = {}
    for part in out:
        for k, v in part.items():
            if k in merged:
                raise RuntimeError(f"frame {k} produced by two ranks")
            merged[k] = v
    return merged
