#!/usr/bin/env python
"""bench.py -- headline benchmark of the segment-and-track hot path (contract in the task statement).

One "step" = one frame of BASELINE.json configs[1] (worm1 single mode, 512 x 512 x 35 uint16 stack, 164 cells):
    LCN normalise -> tiled 3D U-Net (unet3_a, 75 tiles) -> 5 x (FFN match + PR-GLS EM, 19 iterations) -> replay of the
    5 fitted transforms on the tracked cells -> trimmed mean.
The host watershed between segmentation and matching is outside SURVEY section 8 (row f-1); point sets are the
synthetic ground-truth centres, so the step is "segment + match + track without the host watershed stage".

metric  voxels/s = input-volume voxels (x*y*z) per second through the whole step (whole job, all GPUs).
value   inputs resident in HBM when the timed region starts.
e2e     same step through the public Python API with HOST buffers: pinned uint16 stack H2D, probability map
        and tracked coordinates D2H, all inside the timed region.
N > 1   frames shard one per GPU (weak scaling, no data-path collective); timing is max over ranks.

`--impl reference` times the CPU oracle (torch/oneDNN fp32 restatement of the Keras graphs + NumPy EM; TensorFlow is
not installable in this image) on a bounded sample of the same workload with all host threads.
"""
import os
import sys

if "--impl" in sys.argv and "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core (set before NumPy / torch load)
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SHAPE = (512, 512, 35)
N_CELLS = 164
Z_XY_RATIO = 9.2
NOISE_LEVEL = 20
BETA_TK, LAMBDA_TK, MAXITER_TK = 300, 0.1, 20       # single_mode_worm1-clear.ipynb:151
REP_NUM_PRGLS = 5
SHRINK = (24, 24, 2)
FLOP_PER_TILE = 35.573e9
WORKLOAD = "worm1 single-mode 512x512x35, unet3_a seg (75 tiles) + 5x(FFN match + PR-GLS 19 it) @164 cells"


def mod(name):
    return importlib.import_module("3deecelltracker_b200." + name)


def measured_traffic(tiles_per_batch):
    """dram bytes per conv launch from the newest committed `ncu --set full` capture (profiles/*_traffic.json); None
    when the capture was taken at a different batch size."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))
    if not files:
        return None
    with open(files[-1]) as f:
        t = json.load(f)
    if t.get("tiles_per_batch") != tiles_per_batch:
        return None
    return t["dram_bytes_per_launch_avg"]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], bf16_burst=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [v for v in sm if v > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------------
def make_inputs(frame):
    synth = mod("synth")
    centres0 = synth.blob_centres(SHAPE, N_CELLS, 1234)
    real0 = centres0 * np.array([1.0, 1.0, Z_XY_RATIO])
    real_t = synth.move_points(real0, 1234 + frame, affine_level=0.05, noise=0.002)
    centres_t = real_t / np.array([1.0, 1.0, Z_XY_RATIO])
    centres_t = np.clip(centres_t, 0, np.array(SHAPE) - 1)
    raw = synth.blob_stack(SHAPE, centres_t, 1234 + frame, z_xy_ratio=Z_XY_RATIO)
    return raw, real0, real_t


class Step:
    """One frame through the product's FramePipeline (pipeline.py; mirrors Tracker.track_one_vol's hot-path calls).
    Every step performs one segmentation (LCN + U-Net), submits one fit (5 x (FFN match + PR-GLS)) to a side stream
    and joins + replays the fit submitted `depth` steps earlier, so K timed steps contain K full frames of every stage;
    the fits still in flight after the last step are joined (flush) INSIDE the timed region.
    With overlap=False the stages run back to back on one stream."""

    def __init__(self, unet, ffn, overlap=True, reserve_sms=8, depth=2):
        self.pipe = mod("pipeline").FramePipeline(unet, ffn, NOISE_LEVEL, BETA_TK, LAMBDA_TK, MAXITER_TK, SHRINK,
                                                  overlap=overlap, reserve_sms=reserve_sms, depth=depth)

    def run(self, raw_dev, ref_dev, tgt_dev, tracked_dev):
        return self.pipe.step(raw_dev, ref_dev, tgt_dev, tracked_dev)

    def flush(self):
        return self.pipe.flush()


def gpu_main(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = mod("_lib")
    lib = L.lib()
    synth = mod("synth")
    unet = mod("unet3d").UNet3("a", weights=synth.unet_weights("a", 0), tiles_per_batch=args.tiles_per_batch,
                               engine=args.engine)
    ffn = mod("ffn").FFN(synth.ffn_weights(0))
    step = Step(unet, ffn, overlap=not args.no_overlap, reserve_sms=args.reserve_sms, depth=args.depth)
    serial = Step(unet, ffn, overlap=False)
    n_tiles, _ = unet.tile_count(SHAPE, SHRINK)

    raw, real0, real_t = make_inputs(frame=1 + rank)
    dev = torch.device("cuda", local_rank)
    raw_pinned = torch.from_numpy(raw.view(np.int16)).pin_memory()
    ref_pinned = torch.from_numpy(real0).pin_memory()
    tgt_pinned = torch.from_numpy(real_t).pin_memory()
    raw_dev = raw_pinned.to(dev).view(torch.uint16)
    ref_dev, tgt_dev = ref_pinned.to(dev), tgt_pinned.to(dev)
    tracked_dev = ref_dev.clone()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm.  The timed region is ONE interval around K steps plus the drain of the pipeline (the
    # fits still in flight are joined inside it), so it contains K full frames of every stage and nothing else; the
    # warm-up's own in-flight fits are drained before it starts.
    for _ in range(args.warmup):
        step.run(raw_dev, ref_dev, tgt_dev, tracked_dev)
    step.flush()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.ct_profile_enable(1)
    launches0 = lib.ct_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_tracked = 0
    for _ in range(args.steps):
        flush.zero_()                                            # flush L2 between timed iterations (inside the region)
        n_tracked += step.run(raw_dev, ref_dev, tgt_dev, tracked_dev)[1] is not None
    n_tracked += len(step.flush())
    e1.record()
    barrier()
    assert n_tracked == args.steps, f"{n_tracked} tracking results for {args.steps} steps"
    times = [(e0, e1)]
    launches = lib.ct_launch_count() - launches0
    lib.ct_profile_enable(0)
    dev_ms = sum(a.elapsed_time(b) for a, b in times)
    import ctypes as C
    prof = {}
    for tag, name in ((1, "conv"), (2, "em"), (3, "ffn"), (4, "lcn")):
        ms, cnt = C.c_double(), C.c_ulonglong()
        lib.ct_profile_read(tag, C.byref(ms), C.byref(cnt), 1)
        prof[name] = (ms.value, cnt.value)

    # ---- the same frame with the two stages back to back on one stream (frame latency; explains the overlap gain)
    serial_ms = None
    if not args.no_overlap:
        serial.run(raw_dev, ref_dev, tgt_dev, tracked_dev)
        barrier()
        ts = []
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            serial.run(raw_dev, ref_dev, tgt_dev, tracked_dev)
            e1.record()
            ts.append((e0, e1))
        barrier()
        serial_ms = sum(a.elapsed_time(b) for a, b in ts) / args.steps

    # ---- end-to-end arm: host buffers in, host results out, every step.  The inputs of step i+1 go up on a copy stream
    # while step i computes (the first step's upload is exposed); the results of step i (36.7 MB probability map + tracked coordinates) come down on a copy stream into one of two
    # pinned buffers while step i+1 computes; the host waits for step i's download before it submits step i+2, and for
    # everything at the end of the timed region.
    prob_host = [torch.empty(SHAPE, dtype=torch.float32).pin_memory() for _ in range(2)]
    out_host = torch.empty((N_CELLS, 3), dtype=torch.float64).pin_memory()
    copy_stream = torch.cuda.Stream()
    downloads = []

    uploads = []

    def upload():
        """One step's inputs, pinned host -> HBM on the copy stream."""
        with torch.cuda.stream(copy_stream):
            ts = (raw_pinned.to(dev, non_blocking=True).view(torch.uint16), ref_pinned.to(dev, non_blocking=True),
                  tgt_pinned.to(dev, non_blocking=True))
            ev = torch.cuda.Event()
            ev.record()
        return ts, ev

    def e2e_step(i, last=False):
        (r, a, b), up = uploads.pop(0) if uploads else upload()
        main = torch.cuda.current_stream()
        main.wait_event(up)
        for t_ in (r, a, b):
            t_.record_stream(main)
        if not last:
            uploads.append(upload())                  # the next step's inputs go up while this one computes
        prob, out = step.run(r, a, b, a)
        outs = ([] if out is None else [out]) + (step.flush() if last else [])
        ready = torch.cuda.Event()
        ready.record()
        copy_stream.wait_event(ready)
        with torch.cuda.stream(copy_stream):
            prob_host[i % 2].copy_(prob, non_blocking=True)
            for o in outs:
                out_host.copy_(o, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        for t_ in [prob] + outs:
            t_.record_stream(copy_stream)
        downloads.append(done)
        while len(downloads) > (0 if last else 1):
            downloads.pop(0).synchronize()
        return len(outs)

    for i in range(max(1, args.warmup // 2)):
        e2e_step(i, last=(i == max(1, args.warmup // 2) - 1))
    barrier()
    t0 = time.perf_counter()
    got = 0
    for i in range(args.steps):
        got += e2e_step(i, last=(i == args.steps - 1))
    barrier()
    e2e_s = time.perf_counter() - t0
    assert got == args.steps
    clocks = sampler.stop() if rank == 0 else None

    # max over ranks
    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        voxels = SHAPE[0] * SHAPE[1] * SHAPE[2]
        pk = peaks()
        conv_ms, conv_n = prof["conv"]
        conv_flops = n_tiles * FLOP_PER_TILE * args.steps
        achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        peak = pk["bf16_sustained"]
        line = {
            "metric": "voxels/s", "value": voxels * world * args.steps / (dev_ms * 1e-3), "unit": "voxels/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 hi/lo split (22-bit operands), fp32 accumulate (U-Net conv); f32 (LCN, FFN); f64 (PR-GLS EM)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": 1, "unet_tiles": n_tiles,
                       "unet_engine": args.engine, "tiles_per_batch": args.tiles_per_batch,
                       "l2": "flushed between timed iterations (256 MiB write)",
                       "sharding": "frames, one per GPU" if world > 1 else "single GPU",
                       "pipeline": ("serial: segmentation then tracking on one stream" if args.no_overlap else
                                    "segmentation on the main stream; %d fits (5x FFN+PR-GLS) in flight on side "
                                    "streams, each joined + replayed %d steps after submission; %d SMs kept out of the "
                                    "persistent conv grid; pipeline drained inside the timed region"
                                    % (args.depth, args.depth, args.reserve_sms)),
                       "host_watershed": "excluded (SURVEY 8f-1)"},
            "frames_per_s": world * args.steps / (dev_ms * 1e-3),
            "e2e": {"value": voxels * world * args.steps / (e2e_ms * 1e-3), "unit": "voxels/s",
                    "frames_per_s": world * args.steps / (e2e_ms * 1e-3),
                    "h2d_bytes_per_step": int(raw.nbytes + real0.nbytes + real_t.nbytes),
                    "d2h_bytes_per_step": int(prob_host[0].numel() * 4 + out_host.numel() * 8)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "unet 3x3x3 conv (%s engine)" % args.engine,
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": pk["source"] + ", bf16 dense sustained",
                         "launches": int(conv_n), "avg_launch_ms": conv_ms / max(conv_n, 1),
                         "share_of_step": conv_ms / dev_ms if dev_ms else None,
                         "algorithmic_flop_per_launch": conv_flops / max(conv_n, 1),
                         "traffic": measured_traffic(args.tiles_per_batch), "traffic_unit": "bytes/launch (ncu dram read+write)",
                         "note": "split-fp16 tcgen05 implicit GEMM, 3 MMA terms per fp32 product: bound by the tensor "
                                 "core's shared-memory operand reads (ncu: tc smem wavefronts 55-83% of peak), see DESIGN.md 3.2"},
            "stage_ms_per_step": {k: v[0] / args.steps for k, v in prof.items()},
            "serial_ms_per_step": serial_ms,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(sample_tiles=2, threads=None)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
# config 3: ONE 1024 x 1024 x 96 volume cut 2 x 2 x 2 over the GPUs (strong scaling of the segmentation half)
# --------------------------------------------------------------------------------------------------
C3_SHAPE, C3_CELLS = (1024, 1024, 96), 2048
_GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def gpu_spatial_main(args):
    """`--workload c3`: LCN + tiled U-Net of one zebrafish-heart-sized stack, spatially decomposed (spatial.py):
    histogram all-reduces for the global median, one round of halo send/recv, 800 tiles split over the ranks.
    With --verify every rank also runs the whole volume alone and checks its block bit for bit."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world not in _GRIDS:
        raise SystemExit("--workload c3 runs on 1, 2, 4 or 8 GPUs")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = mod("_lib").lib()
    synth, sp, pre = mod("synth"), mod("spatial"), mod("preprocess")
    shape = tuple(args.shape) if args.shape else C3_SHAPE
    unet = mod("unet3d").UNet3("a", weights=synth.unet_weights("a", 0), tiles_per_batch=args.tiles_per_batch,
                               engine=args.engine)
    plan = sp.SpatialPlan(shape, _GRIDS[world], unet.input_shape[1:4], SHRINK)
    cells = max(8, int(C3_CELLS * (shape[0] * shape[1] * shape[2]) / (1024 * 1024 * 96)))
    raw = synth.blob_stack(shape, synth.blob_centres(shape, cells, 4321), 4321)       # same volume on every rank
    lo, hi = plan.owned_box(rank)
    own_pinned = torch.from_numpy(np.ascontiguousarray(raw[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]).view(np.int16)).pin_memory()
    owned_dev = own_pinned.to(dev).view(torch.uint16)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(block):
        return sp.segment_block(block, plan, rank, unet, NOISE_LEVEL)

    if args.verify:
        prob, out = step(owned_dev)
        whole = pre._raw_to_device(raw)
        want = unet.prediction_device(pre.normalize_image_device(whole, NOISE_LEVEL), SHRINK)
        ok = prob is None or torch.equal(prob, want[out[0][0]:out[1][0], out[0][1]:out[1][1], out[0][2]:out[1][2]])
        flag = torch.tensor([1 if ok else 0], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag[0]) != 1:
            raise SystemExit(f"rank {rank}: decomposed result differs from the single-GPU result")
        del whole, want
    for _ in range(args.warmup):
        step(owned_dev)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.ct_profile_enable(1)
    launches0 = lib.ct_launch_count()
    times = []
    for _ in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(owned_dev)
        e1.record()
        times.append((e0, e1))
    barrier()
    launches = lib.ct_launch_count() - launches0
    lib.ct_profile_enable(0)
    dev_ms = sum(a.elapsed_time(b) for a, b in times)
    import ctypes as C
    ms, cnt = C.c_double(), C.c_ulonglong()
    lib.ct_profile_read(1, C.byref(ms), C.byref(cnt), 1)
    conv_ms, conv_n = ms.value, cnt.value

    out_box = plan.out_box(rank)
    prob_host = None if out_box is None else torch.empty(tuple(h - l for l, h in zip(*out_box)), dtype=torch.float32).pin_memory()

    def e2e_step():
        block = own_pinned.to(dev, non_blocking=True).view(torch.uint16)
        prob, _ = step(block)
        if prob is not None:
            prob_host.copy_(prob, non_blocking=True)
        torch.cuda.synchronize()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    my_tiles = int(np.prod([h - l for l, h in zip(*plan.tile_box(rank))]))
    t = torch.tensor([dev_ms, e2e_s * 1e3, conv_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, conv_ms = float(t[0]), float(t[1]), float(t[2])
    if rank == 0:
        voxels = shape[0] * shape[1] * shape[2]
        n_tiles = int(np.prod(plan.num_tiles))
        pk = peaks()
        achieved = n_tiles * FLOP_PER_TILE * args.steps / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        print(json.dumps({
            "metric": "voxels/s", "value": voxels * args.steps / (dev_ms * 1e-3), "unit": "voxels/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "fp16 hi/lo split, fp32 accumulate (U-Net conv); f32 (LCN)", "data": "synthetic",
            "config": {"workload": "zebrafish-heart %dx%dx%d stack, LCN + unet3_a seg (%d tiles), spatial %dx%dx%d "
                                   "decomposition + halo" % (shape + (n_tiles,) + plan.grid),
                       "tiles_rank0": my_tiles, "tiles_per_batch": args.tiles_per_batch,
                       "halo_bytes_rank0": plan.halo_bytes(0), "collectives": "2 x all-reduce(512 x u32) + 1 round send/recv",
                       "verified_bit_identical": bool(args.verify),
                       "l2": "flushed between timed iterations (256 MiB write)"},
            "e2e": {"value": voxels * args.steps / (e2e_ms * 1e-3), "unit": "voxels/s",
                    "h2d_bytes_per_step": int(own_pinned.numel() * 2),
                    "d2h_bytes_per_step": 0 if prob_host is None else int(prob_host.numel() * 4)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "unet 3x3x3 conv", "achieved": achieved,
                         "peak": pk["bf16_sustained"] * world, "unit": "TFLOP/s",
                         "frac": achieved / (pk["bf16_sustained"] * world), "launches": int(conv_n),
                         "note": "whole job: all tiles / slowest rank's conv time, against N x the per-GPU peak",
                         "share_of_step": conv_ms / dev_ms if dev_ms else None, "traffic": None},
            "clocks": clocks}))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
# CPU arm (oracle)
# --------------------------------------------------------------------------------------------------
def cpu_step(sample_tiles, model, ffn_model, norm_tiles, ref, tgt):
    """Bounded sample of one frame on the CPU oracle with the reference's control flow: `sample_tiles` U-Net tiles
    (one predict per tile, batch 1) + ONE of the five (FFN match + PR-GLS) repetitions.  Returns seconds for a
    whole frame, extrapolated: tiles * 75 / sample_tiles + 5 * rep."""
    from oracle import ffn as offn
    from oracle import prgls as oprgls
    t0 = time.perf_counter()
    for i in range(sample_tiles):
        model.predict(norm_tiles[i:i + 1])
    t_tiles = time.perf_counter() - t0
    t0 = time.perf_counter()
    corr = offn.initial_matching_quick(ffn_model, ref, tgt, 20)
    oprgls.pr_gls_quick(ref, tgt, corr, BETA=BETA_TK, max_iteration=MAXITER_TK, LAMBDA=LAMBDA_TK)
    t_rep = time.perf_counter() - t0
    return t_tiles, t_rep


def cpu_baseline(sample_tiles=2, threads=None, steps=1, warmup=0):
    import torch
    from oracle import ffn as offn
    from oracle import unet as ounet
    cores = os.cpu_count() or 1
    if threads:
        torch.set_num_threads(threads)
    used = torch.get_num_threads()
    model = ounet.UNetOracle("a", ounet.random_weights("a", 0))
    ffn_model = offn.FFNOracle(offn.random_weights(0))
    rng = np.random.default_rng(0)
    tiles = rng.normal(0, 1, (sample_tiles, 160, 160, 16, 1)).astype(np.float32)
    real0, real_t = make_points()
    for _ in range(warmup):
        cpu_step(1, model, ffn_model, tiles, real0, real_t)
    tt, tr = 0.0, 0.0
    for _ in range(steps):
        a, b = cpu_step(sample_tiles, model, ffn_model, tiles, real0, real_t)
        tt += a; tr += b
    tt /= steps; tr /= steps
    n_tiles = 75
    frame_s = tt * n_tiles / sample_tiles + REP_NUM_PRGLS * tr
    voxels = SHAPE[0] * SHAPE[1] * SHAPE[2]
    return {"value": voxels / frame_s, "unit": "voxels/s", "cores": used, "host_cores": cores, "kind": "port",
            "frame_seconds_extrapolated": frame_s, "unet_s_per_tile": tt / sample_tiles, "ffn_prgls_s_per_rep": tr,
            "sample": f"{sample_tiles} of 75 U-Net tiles (torch/oneDNN fp32, batch 1 per tile as unet3d.py:253) + 1 of 5 "
                      f"FFN+PR-GLS repetitions (NumPy fp64, dense (M*N,122) grid as ffn.py:320), extrapolated to one "
                      f"frame; LCN excluded (<1% of the CPU frame)",
            "note": "CPU restatement (torch/oneDNN) -- TensorFlow is not installable in this image"}


def make_points():
    synth = mod("synth")
    centres0 = synth.blob_centres(SHAPE, N_CELLS, 1234)
    real0 = centres0 * np.array([1.0, 1.0, Z_XY_RATIO])
    real_t = synth.move_points(real0, 1235, affine_level=0.05, noise=0.002)
    return real0, real_t


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_baseline(sample_tiles=2, threads=None, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    voxels = SHAPE[0] * SHAPE[1] * SHAPE[2]
    line = {"impl": "reference", "metric": "voxels/s", "value": base["value"], "unit": "voxels/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": base["frame_seconds_extrapolated"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (U-Net/FFN) + f64 (PR-GLS EM)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": base["sample"]},
            "frames_per_s": 1.0 / base["frame_seconds_extrapolated"],
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    assert voxels > 0
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--engine", default="auto", choices=["auto", "direct", "tcgen05", "tcgen05_classic", "tcgen05_stacked"])
    ap.add_argument("--tiles-per-batch", type=int, default=38)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--depth", type=int, default=2, help="fits (FFN + PR-GLS chains) in flight on side streams")
    ap.add_argument("--reserve-sms", type=int, default=8, help="SMs kept out of the persistent conv grid while overlapping")
    ap.add_argument("--no-overlap", action="store_true", help="run segmentation and tracking back to back on one stream")
    ap.add_argument("--workload", default="c1", choices=["c1", "c3"],
                    help="c1 (default, the contract's line): one worm1 frame per step; c3: one 1024x1024x96 volume "
                         "spatially decomposed over the GPUs")
    ap.add_argument("--shape", type=int, nargs=3, default=None, help="c3 only: override the volume shape")
    ap.add_argument("--verify", action="store_true", help="c3 only: check every rank's block against a single-GPU run")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_main(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        if args.workload == "c3":
            gpu_spatial_main(args)
        else:
            gpu_main(args)


if __name__ == "__main__":
    main()
