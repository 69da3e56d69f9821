#!/usr/bin/env python
"""bench.py -- headline benchmark of the segment-and-track hot path (contract in the task statement).

One "step" = one volume of BASELINE.json configs[1] (worm1 single mode, 512 x 512 x 35 uint16 stack, 164 cells) of a
synthetic time-lapse, through the whole Tracker.track_one_vol chain (tracker.py:1473-1536):
    LCN normalise -> tiled 3D U-Net (unet3_a, 75 tiles) -> watershed_2d + watershed_3d + centres of mass (on the GPU) ->
    5 x (FFN match + PR-GLS EM, 19 iterations) between the SEGMENTED point sets of consecutive volumes -> replay of the
    5 fitted transforms on the tracked cells -> trimmed mean.
The U-Net weights are the blob detector of synth.detector_unet_weights (no trained weights offline), so the
segmentation output really feeds the matcher.  `frames_per_s_without_watershed` (N = 1) is the round-1 workload
(ground-truth point sets, no watershed) for comparison.

metric  voxels/s = input-volume voxels (x*y*z) per second through the whole chain (whole job, all GPUs).
value   raw stacks resident in HBM when the timed region starts.
e2e     same volumes through the public Python API with HOST buffers: pinned uint16 stack H2D, probability map +
        label image (+ tracked coordinates) D2H, all inside the timed region.
N > 1   BASELINE configs[4]: a time-lapse of K x N volumes in contiguous blocks of K per GPU (timelapse.py; weak scaling,
        no collective on the volumes); timing is max over ranks.  The same line carries configs[3] under "c3":
        ONE 1024 x 1024 x 96 volume spatially decomposed over the N GPUs (strong scaling).

`--impl reference` times the CPU oracle (torch/oneDNN fp32 restatement of the Keras graphs, scipy + restated
scikit-image watershed, NumPy EM; TensorFlow is not installable in this image) on whole frames of the same workload
with all host threads.
"""
import os
import sys

if "--impl" in sys.argv and "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core (set before NumPy / torch load)
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import argparse
import gc
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
WORKLOAD = ("worm1 single-mode 512x512x35 time-lapse: LCN + unet3_a seg (75 tiles) + watershed/centroids + "
            "5x(FFN match + PR-GLS 19 it) + replay, 164 cells")


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
        if os.environ.get("CT3D_NO_SAMPLER"):                          # debugging only: the contract wants the samples
            return
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
MIN_SIZE = 40                                       # watershed min_size (voxels) for the synthetic blobs


def frame_centres(t):
    """Cell centres (voxel units) of volume t of the synthetic time-lapse: a slow affine drift of 164 cells."""
    synth = mod("synth")
    centres0 = synth.blob_centres(SHAPE, N_CELLS, 1234, margin=12)
    c = centres0.mean(axis=0)
    a = np.eye(3) + 0.0008 * t * np.array([[0.5, 1.0, 0.0], [-1.0, 0.4, 0.0], [0.0, 0.0, 0.0]])
    moved = (centres0 - c) @ a + c + 0.05 * t * np.array([1.0, -0.6, 0.0])
    return np.clip(moved, 6, np.array(SHAPE) - 7)


_NOISE = {}


def make_frame(t):
    """uint16 stack of volume t (N(100, 10^2) background + Gaussian blobs); the background noise field is
    drawn once per process and shifted by t voxels, the blobs are rendered per volume."""
    synth = mod("synth")
    if "bg" not in _NOISE:
        _NOISE["bg"] = np.random.default_rng(99).normal(100.0, 10.0, SHAPE).astype(np.float32)
    stack = synth.blob_stack(SHAPE, frame_centres(t), 1234, z_xy_ratio=Z_XY_RATIO, background=np.roll(_NOISE["bg"], t, axis=0))
    return stack


def make_inputs(frame):
    """Ground-truth point sets for the no-watershed comparison arm (round-1 workload)."""
    synth = mod("synth")
    centres0 = synth.blob_centres(SHAPE, N_CELLS, 1234)
    real0 = centres0 * np.array([1.0, 1.0, Z_XY_RATIO])
    real_t = synth.move_points(real0, 1234 + frame, affine_level=0.05, noise=0.002)
    return real0, real_t


def build_pipeline(args, overlap=True):
    synth = mod("synth")
    unet = mod("unet3d").UNet3("a", weights=synth.detector_unet_weights(0), tiles_per_batch=args.tiles_per_batch,
                               engine=args.engine)
    ffn = mod("ffn").FFN(synth.ffn_weights(0))
    pipe = mod("pipeline").FramePipeline(unet, ffn, NOISE_LEVEL, BETA_TK, LAMBDA_TK, MAXITER_TK, SHRINK,
                                         overlap=overlap, reserve_sms=args.reserve_sms, depth=args.depth, ws_lag=args.ws_lag)
    pipe.configure_watershed(Z_XY_RATIO, "min_size", MIN_SIZE, 0)
    return unet, ffn, pipe


def read_profile(lib):
    import ctypes as C
    prof = {}
    for tag, name in ((1, "conv"), (2, "em"), (3, "ffn"), (4, "lcn"), (6, "watershed")):
        ms, cnt = C.c_double(), C.c_ulonglong()
        lib.ct_profile_read(tag, C.byref(ms), C.byref(cnt), 1)
        prof[name] = (ms.value, cnt.value)
    return prof


def gpu_main(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = mod("_lib").lib()
    shard, tl = mod("shard"), mod("timelapse")
    unet, ffn, pipe = build_pipeline(args)
    n_tiles, _ = unet.tile_count(SHAPE, SHRINK)
    K = args.steps
    T = K * world                                        # volumes of the time-lapse: K per GPU (weak scaling)
    lo, hi = shard.block_for_rank(T, rank, world)
    warm = [torch.from_numpy(make_frame(T + i).view(np.int16)).pin_memory() for i in range(max(args.warmup, 3))]
    frames_pinned = {t: torch.from_numpy(make_frame(t).view(np.int16)).pin_memory() for t in range(lo, hi)}
    frames_dev = {t: p.to(dev).view(torch.uint16) for t, p in frames_pinned.items()}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dev_trace = [] if os.environ.get("CT3D_E2E_TRACE") else None  # host timestamps per volume (debugging)

    def resident(t):
        if dev_trace is not None:
            dev_trace.append(time.perf_counter())
        flush.zero_()                                    # flush L2 between timed iterations (inside the region)
        return frames_dev[t]

    tracker = tl.TimelapseTracker(pipe, rank, world)

    # ---- config 4 check on a short time-lapse: the sharded run equals the single-GPU run bit for bit
    c4_verified = None
    if world > 1 and not args.no_verify:
        tv = 3 * world
        vlo, vhi = shard.block_for_rank(tv, rank, world)
        vframes = {t: torch.from_numpy(make_frame(t).view(np.int16)).to(dev).view(torch.uint16)
                   for t in (range(tv) if rank == 0 else range(vlo, vhi))}
        got = tracker.run(lambda t: vframes[t], tv)
        ok = 1
        if rank == 0:
            want = tl.TimelapseTracker(pipe, 0, 1).run(lambda t: vframes[t], tv)
            ok = int(len(got) == len(want) and all(torch.equal(a, b) for a, b in zip(got, want)))
        flag = torch.tensor([ok], device=dev)
        dist.broadcast(flag, 0)
        if int(flag[0]) != 1:
            raise SystemExit("config 4: sharded time-lapse differs from the single-GPU run")
        c4_verified = tv
        del vframes

    # ---- warm-up (>= 3 volumes through every stage), then ONE timed interval around the whole time-lapse: K volumes
    # per GPU through segmentation + watershed + fit, the boundary exchange, the gather and the replay on rank 0
    wdev = [w.to(dev).view(torch.uint16) for w in warm]
    tl.TimelapseTracker(pipe, 0, 1).run(lambda t: wdev[t], len(wdev))
    pipe.reserve_small_blocks(64)                      # no cudaMalloc (a device-wide implicit sync) inside the timed regions
    # The host enqueues ~150 launches per volume from Python; a generation-2 pass of the cyclic garbage collector over the
    # interpreter's ~10^6 tracked objects (torch, numpy ...) stalls it for 40-140 ms [measured: host gaps of that size
    # between volumes, not spent in any CUDA wait], which empties the GPU's queues.  Everything allocated so far is moved
    # to the permanent generation and the collector is off while the timed regions run (reference counting still frees
    # every tensor at once).
    gc.collect()
    gc.freeze()
    gc.disable()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.ct_profile_enable(1)
    read_profile(lib)
    launches0 = lib.ct_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dev_trace is not None:
        pipe.trace = []
        mem0 = dict(torch.cuda.memory_stats())
    e0.record()
    tracked = tracker.run(resident, T)
    e1.record()
    barrier()
    launches = lib.ct_launch_count() - launches0
    lib.ct_profile_enable(0)
    dev_ms = e0.elapsed_time(e1)
    if dev_trace is not None and rank == 0:
        print("device-arm host trace (ms between volumes): " + " ".join(f"{(b - a) * 1e3:.1f}" for a, b in zip(dev_trace, dev_trace[1:]))
              + f" | total {dev_ms:.1f}", file=sys.stderr)
        mem1 = torch.cuda.memory_stats()
        print("  waits for cell counts (ms): " + " ".join(f"{w * 1e3:.1f}" for w in pipe.trace), file=sys.stderr)
        print("  allocator: " + ", ".join(f"{k} +{mem1[k] - mem0.get(k, 0)}" for k in ("num_device_alloc", "num_device_free", "num_alloc_retries",
                                                                                "reserved_bytes.all.current", "allocation.all.allocated")), file=sys.stderr)
        pipe.trace = None
    prof = read_profile(lib)
    n_cells = None
    if rank == 0:
        assert len(tracked) == T - 1, f"{len(tracked)} tracking results for {T} volumes"
        n_cells = int(tracked[-1].shape[0]) if tracked else 0        # --steps 1: one volume, nothing to track yet
        assert all(torch.isfinite(t).all() for t in tracked[-1:])

    # ---- comparison arms on one GPU: stages back to back on one stream, and the round-1 workload without watershed
    serial_ms = nows_ms = None
    if world == 1 and not args.no_compare:
        _, _, spipe = build_pipeline(args, overlap=False)
        st = tl.TimelapseTracker(spipe, 0, 1)
        st.run(lambda t: wdev[t], len(wdev))
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        st.run(resident, T)
        b.record()
        barrier()
        serial_ms = a.elapsed_time(b) / K
        real0, real_t = make_inputs(1)
        ref_dev, tgt_dev = torch.from_numpy(real0).to(dev), torch.from_numpy(real_t).to(dev)
        for _ in range(3):
            pipe.step(frames_dev[lo], ref_dev, tgt_dev, ref_dev)
        pipe.flush()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for t in range(lo, hi):
            pipe.step(resident(t), ref_dev, tgt_dev, ref_dev)
        pipe.flush()
        b.record()
        barrier()
        nows_ms = a.elapsed_time(b) / K

    # ---- end-to-end arm: host buffers in, host results out, every volume.  The raw stack of volume t+1 goes up on a
    # copy stream while volume t computes; probability map + label image of volume t come down on the copy stream into
    # one of three pinned buffers while volumes t+1, t+2 compute; tracked coordinates come down at the end (rank 0).
    prob_host = [torch.empty(SHAPE, dtype=torch.float32).pin_memory() for _ in range(3)]
    lab_host = [torch.empty(SHAPE, dtype=torch.int32).pin_memory() for _ in range(3)]
    # separate streams per direction: an upload must never queue behind a download that waits for a watershed
    up_stream, copy_stream = torch.cuda.Stream(), torch.cuda.Stream()
    state = {"next": None, "done": []}

    def upload(t):
        with torch.cuda.stream(up_stream):
            d = frames_pinned[t].to(dev, non_blocking=True).view(torch.uint16)
            ev = torch.cuda.Event()
            ev.record()
        return d, ev

    trace = [] if os.environ.get("CT3D_E2E_TRACE") else None      # host timestamps per volume (debugging the e2e arm)

    def e2e_frame(t):
        if trace is not None:
            trace.append(("frame", t, time.perf_counter()))
        nxt = state["next"]
        d, ev = (nxt[0], nxt[1]) if nxt is not None and nxt[2] == t else upload(t)
        main = torch.cuda.current_stream()
        main.wait_event(ev)
        d.record_stream(main)
        state["next"] = upload(t + 1) + (t + 1,) if t + 1 < hi else None
        return d

    def sink(t, prob, seg):
        if trace is not None:
            trace.append(("sink", t, time.perf_counter()))
        ready = torch.cuda.Event()
        ready.record()
        copy_stream.wait_event(ready)
        if seg.ready is not None:
            copy_stream.wait_event(seg.ready)            # the label image is produced on the pipeline's watershed stream
        with torch.cuda.stream(copy_stream):
            prob_host[t % 3].copy_(prob, non_blocking=True)
            lab_host[t % 3].copy_(seg.labels, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        prob.record_stream(copy_stream)
        seg.labels.record_stream(copy_stream)
        state["done"].append(done)
        while len(state["done"]) > 2:                    # three pinned buffers: volume t-2's download must be complete
            state["done"].pop(0).synchronize()

    # untimed warm-up of THIS arm's own machinery (copy streams, the allocator pools of the upload stream, pinned buffers)
    wpin = {i: w for i, w in enumerate(warm)}

    def warm_frame(t):
        with torch.cuda.stream(up_stream):
            d = wpin[t].to(dev, non_blocking=True).view(torch.uint16)
            ev = torch.cuda.Event()
            ev.record()
        torch.cuda.current_stream().wait_event(ev)
        d.record_stream(torch.cuda.current_stream())
        return d

    tl.TimelapseTracker(pipe, 0, 1).run(warm_frame, len(warm), sink=sink)
    for d in state["done"]:
        d.synchronize()
    state["done"].clear()
    if trace is not None:
        trace.clear()
    barrier()
    t0 = time.perf_counter()
    out = tracker.run(e2e_frame, T, sink=sink)
    coords_host = torch.stack(out).cpu() if out else None             # --steps 1: one volume, nothing tracked yet
    for d in state["done"]:
        d.synchronize()
    barrier()
    e2e_s = time.perf_counter() - t0
    if trace is not None and rank == 0:
        print("e2e trace (ms since start): " + " ".join(f"{k[0]}{n}@{(ts - t0) * 1e3:.1f}" for k, n, ts in trace)
              + f" end@{e2e_s * 1e3:.1f}", file=sys.stderr)
    clocks = sampler.stop() if rank == 0 else None

    # ---- config 3 (strong scaling of ONE 1024 x 1024 x 96 volume) on the same GPUs, same build
    c3 = None if args.no_c3 else c3_measure(args, unet, rank, world, dev, barrier)

    gc.enable()
    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        voxels = SHAPE[0] * SHAPE[1] * SHAPE[2]
        pk = peaks()
        conv_ms, conv_n = prof["conv"]
        conv_flops = n_tiles * FLOP_PER_TILE * K                       # rank 0's volumes
        achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        peak = pk["bf16_sustained"]
        lcn_ms, ws_ms = prof["lcn"][0] / K, prof["watershed"][0] / K
        line = {
            "metric": "voxels/s", "value": voxels * T / (dev_ms * 1e-3), "unit": "voxels/s",
            "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 hi/lo split (22-bit operands), fp32 accumulate (U-Net conv); f32 (LCN, FFN); f64 + int32 (watershed); f64 (PR-GLS EM)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": 1, "timelapse_volumes": T, "unet_tiles": n_tiles,
                       "unet_engine": args.engine, "tiles_per_batch": args.tiles_per_batch, "cells_tracked": n_cells,
                       "l2": "flushed between timed iterations (256 MiB write)",
                       "sharding": ("contiguous blocks of %d volumes per GPU (timelapse.py); per rank one point-set "
                                    "send/recv for the fit that straddles two blocks, one all-gather of the fitted "
                                    "transforms, sequential replay on rank 0; no collective on the volumes" % K)
                                   if world > 1 else "single GPU",
                       "pipeline": ("segmentation + watershed on the main stream; %d fits (5x FFN+PR-GLS) in flight on side "
                                    "streams; %d SMs kept out of the persistent conv grid; whole time-lapse incl. pipeline "
                                    "drain, gather and replay inside the timed region" % (args.depth, args.reserve_sms)),
                       "host_watershed": "included (runs on the GPU: ct_watershed_segment; segmentation output feeds the matcher)"},
            "frames_per_s": T / (dev_ms * 1e-3),
            "frames_per_s_without_watershed": None if nows_ms is None else 1e3 / nows_ms,
            "e2e": {"value": voxels * T / (e2e_ms * 1e-3), "unit": "voxels/s", "frames_per_s": T / (e2e_ms * 1e-3),
                    "h2d_bytes_per_step": int(frames_pinned[lo].numel() * 2),
                    "d2h_bytes_per_step": int(prob_host[0].numel() * 4 + lab_host[0].numel() * 4 +
                                              (0 if coords_host is None else coords_host.numel() * 8 // max(T - 1, 1)))},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "unet 3x3x3 conv (%s engine)" % args.engine,
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": pk["source"] + ", bf16 dense sustained",
                         "launches": int(conv_n), "avg_launch_ms": conv_ms / max(conv_n, 1),
                         "share_of_step": conv_ms / dev_ms if dev_ms else None,
                         "algorithmic_flop_per_launch": conv_flops / max(conv_n, 1),
                         "traffic": measured_traffic(args.tiles_per_batch), "traffic_unit": "bytes/launch (ncu dram read+write)",
                         "note": "split-fp16 tcgen05 implicit GEMM, 3 MMA terms per fp32 product: bound by the tensor "
                                 "core's shared-memory operand reads, see DESIGN.md 3.2"},
            "roofline_secondary": {
                "lcn": {"bound": "hbm", "algorithmic_bytes": 8 * voxels, "ms": lcn_ms,
                        "achieved_gbs": 8 * voxels / (lcn_ms * 1e-3) / 1e9 if lcn_ms else None, "peak_gbs": pk["hbm"]},
                "watershed": {"bound": "hbm (passes) + latency (flood)", "algorithmic_bytes": 8 * voxels, "ms": ws_ms,
                              "achieved_gbs": 8 * voxels / (ws_ms * 1e-3) / 1e9 if ws_ms else None, "peak_gbs": pk["hbm"],
                              "note": "algorithmic = read float32 probabilities + write int32 labels"},
                "ffn": {"bound": "hbm (factored pair stage)", "ms": prof["ffn"][0] / K,
                        "algorithmic_bytes_per_match": 4 * 512 * 2 * N_CELLS + 4 * N_CELLS * N_CELLS, "matches_per_step": 5},
                "em": {"bound": "latency (one SM per problem)", "ms": prof["em"][0] / K, "launches_per_step": prof["em"][1] / K,
                       "us_per_iteration": prof["em"][0] * 1e3 / max(prof["em"][1], 1) / (MAXITER_TK - 1)}},
            "stage_ms_per_step": {k: v[0] / K for k, v in prof.items()},
            "serial_ms_per_step": serial_ms,
            "c4": {"volumes": T, "verified_bit_identical_to_single_gpu_on_volumes": c4_verified},
            "c3": c3,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
# config 3: ONE 1024 x 1024 x 96 volume cut 2 x 2 x 2 over the GPUs (strong scaling of the segmentation half)
# --------------------------------------------------------------------------------------------------
C3_SHAPE, C3_CELLS = (1024, 1024, 96), 2048
_GRIDS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def c3_measure(args, unet, rank, world, dev, barrier, steps=3, warmup=3, verify=False):
    """LCN + tiled U-Net of one zebrafish-heart-sized stack, spatially decomposed (spatial.py): halo send/recv posted
    first, histogram all-reduces for the global median while it is in flight, 800 tiles split over the ranks.
    Returns the result dict on rank 0 (None elsewhere).  With verify every rank also runs the whole volume alone and
    checks its block bit for bit."""
    import torch
    import torch.distributed as dist

    if world not in _GRIDS:
        return None
    lib = mod("_lib").lib()
    synth, sp, pre = mod("synth"), mod("spatial"), mod("preprocess")
    shape = tuple(args.shape) if args.shape else C3_SHAPE
    plan = sp.SpatialPlan(shape, _GRIDS[world], unet.input_shape[1:4], SHRINK)
    cells = max(8, int(C3_CELLS * (shape[0] * shape[1] * shape[2]) / (1024 * 1024 * 96)))
    raw = synth.blob_stack(shape, synth.blob_centres(shape, cells, 4321), 4321)       # same volume on every rank
    lo, hi = plan.owned_box(rank)
    own_pinned = torch.from_numpy(np.ascontiguousarray(raw[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]).view(np.int16)).pin_memory()
    owned_dev = own_pinned.to(dev).view(torch.uint16)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step(block, marks=None):
        return sp.segment_block(block, plan, rank, unet, NOISE_LEVEL, marks=marks)

    if verify:
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
    del raw
    for _ in range(warmup):
        step(owned_dev)
    barrier()
    lib.ct_profile_enable(1)
    read_profile(lib)
    launches0 = lib.ct_launch_count()
    times, all_marks = [], []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = []
        e0.record()
        step(owned_dev, marks)
        e1.record()
        times.append((e0, e1))
        all_marks.append(marks)
    barrier()
    launches = lib.ct_launch_count() - launches0
    lib.ct_profile_enable(0)
    dev_ms = sum(a.elapsed_time(b) for a, b in times)
    prof = read_profile(lib)
    conv_ms, conv_n = prof["conv"]
    split = {}
    for marks in all_marks:
        for (n0, a), (n1, b) in zip(marks, marks[1:]):
            split[n1] = split.get(n1, 0.0) + a.elapsed_time(b) / steps

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
    for _ in range(steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    my_tiles = int(np.prod([h - l for l, h in zip(*plan.tile_box(rank))]))
    t = torch.tensor([dev_ms, e2e_s * 1e3, conv_ms, split.get("median", 0.0), split.get("halo", 0.0), split.get("lcn", 0.0),
                      split.get("unet", 0.0)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, conv_ms = float(t[0]), float(t[1]), float(t[2])
    del owned_dev, flush
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    voxels = shape[0] * shape[1] * shape[2]
    n_tiles = int(np.prod(plan.num_tiles))
    pk = peaks()
    achieved = n_tiles * FLOP_PER_TILE * steps / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    return {"metric": "voxels/s", "value": voxels * steps / (dev_ms * 1e-3), "unit": "voxels/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": dev_ms / steps, "scaling": "strong",
            "workload": "zebrafish-heart %dx%dx%d stack, LCN + unet3_a seg (%d tiles), spatial %dx%dx%d decomposition + halo"
                        % (shape + (n_tiles,) + plan.grid),
            "tiles_rank0": my_tiles, "tile_batches_rank0": "%d per launch" % unet._balanced_batch(my_tiles),
            "halo_bytes_per_rank": [plan.halo_bytes(r) for r in range(world)],
            "collectives": "1 round of send/recv (posted first) overlapped with 2 x all-reduce(512 x u32)",
            "ms_split_max_over_ranks": {"median_allreduce": float(t[3]), "halo_wait_and_unpack": float(t[4]),
                                        "lcn": float(t[5]), "unet": float(t[6]), "conv_kernels": conv_ms / steps},
            "verified_bit_identical": bool(verify),
            "e2e": {"value": voxels * steps / (e2e_ms * 1e-3), "unit": "voxels/s",
                    "h2d_bytes_per_step": int(own_pinned.numel() * 2),
                    "d2h_bytes_per_step": 0 if prob_host is None else int(prob_host.numel() * 4)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"] * world, "unit": "TFLOP/s",
                         "frac": achieved / (pk["bf16_sustained"] * world),
                         "note": "whole job: all tiles / slowest rank's conv time, against N x the per-GPU peak"}}


def gpu_spatial_main(args):
    """`--workload c3`: config 3 alone (see c3_measure), printed as its own JSON line."""
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
    unet = mod("unet3d").UNet3("a", weights=mod("synth").unet_weights("a", 0), tiles_per_batch=args.tiles_per_batch,
                               engine=args.engine)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    gc.collect()
    gc.freeze()
    gc.disable()                                      # see gpu_main: no collector pauses inside the timed steps
    res = c3_measure(args, unet, rank, world, dev, barrier, steps=args.steps, warmup=max(args.warmup, 3), verify=args.verify)
    gc.enable()
    if rank == 0:
        res.update({"higher_is_better": True, "vs_baseline": None, "data": "synthetic",
                    "dtype": "fp16 hi/lo split, fp32 accumulate (U-Net conv); f32 (LCN)",
                    "config": {"workload": res["workload"]}, "clocks": sampler.stop()})
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------
# CPU arm (oracle)
# --------------------------------------------------------------------------------------------------
class CpuFrame:
    """One WHOLE frame of the workload on the CPU oracle, with the reference's control flow (Tracker.track_one_vol,
    tracker.py:1473-1536): _normalize_image -> unet3_prediction (75 tiles, one predict per tile, batch 1,
    unet3d.py:253) -> watershed_2d + watershed_3d + centres (scipy for real, scikit-image restated) ->
    5 x (initial_matching_quick on the dense (M*N,122) grid + pr_gls_quick) -> 5 x _predict_one_rep -> trim_mean."""

    def __init__(self):
        import torch
        from oracle import ffn as offn
        from oracle import unet as ounet
        synth = mod("synth")
        self.torch = torch
        self.model = ounet.UNetOracle("a", synth.detector_unet_weights(0))
        self.ffn = offn.FFNOracle(synth.ffn_weights(0))
        self.prev = None
        self.tracked = None
        self.t = 0
        self.stage_s = {"lcn": 0.0, "unet": 0.0, "watershed": 0.0, "ffn_prgls": 0.0}

    def step(self):
        from oracle import ffn as offn
        from oracle import prgls as oprgls
        from oracle import unet as ounet
        from oracle import watershed as ows
        raw = make_frame(self.t)
        self.t += 1
        t0 = time.perf_counter()
        norm = ounet.normalize_image(raw.copy(), NOISE_LEVEL).astype(np.float32)
        t1 = time.perf_counter()
        prob = ounet.unet3_prediction(norm[None, ..., None], self.model, SHRINK)[0, ..., 0]
        t2 = time.perf_counter()
        _, centres, _, _ = ows.segment(prob, Z_XY_RATIO, "min_size", MIN_SIZE, 0)
        pts = centres * np.array([1.0, 1.0, Z_XY_RATIO])
        t3 = time.perf_counter()
        if self.prev is not None:
            inter, pred = self.prev, self.tracked
            for i in range(REP_NUM_PRGLS):
                beta = BETA_TK * 0.8 ** i
                corr = offn.initial_matching_quick(self.ffn, inter, pts, 20)
                _, tx, c = oprgls.pr_gls_quick(inter, pts, corr, BETA=beta, max_iteration=MAXITER_TK, LAMBDA=LAMBDA_TK)
                pred = oprgls.predict_one_rep(pred, inter, beta, c)
                inter = tx
            self.tracked = oprgls.trim_mean(pred[None], 0.1)
        else:
            self.tracked = pts.copy()
        self.prev = pts
        t4 = time.perf_counter()
        for k, v in zip(("lcn", "unet", "watershed", "ffn_prgls"), (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
            self.stage_s[k] += v
        return t4 - t0


def cpu_run(steps, warmup, budget_s=150.0):
    """Whole frames on the CPU oracle; the number of timed frames is capped so the run stays within a few minutes."""
    import torch
    cores = os.cpu_count() or 1
    used = torch.get_num_threads()
    cpu = CpuFrame()
    cpu.step()                                             # volume 1: no fit yet (initiate_tracking); also warms caches
    warm_done = 1
    first = cpu.step()
    warm_done += 1
    n = max(1, min(steps, int(budget_s / max(first, 1e-3))))
    for k in cpu.stage_s:
        cpu.stage_s[k] = 0.0
    total = 0.0
    for _ in range(n):
        total += cpu.step()
    frame_s = total / n
    voxels = SHAPE[0] * SHAPE[1] * SHAPE[2]
    return {"value": voxels / frame_s, "unit": "voxels/s", "cores": used, "host_cores": cores, "kind": "port",
            "frame_seconds": frame_s, "frames_timed": n, "warmup_frames": warm_done,
            "stage_seconds_per_frame": {k: v / n for k, v in cpu.stage_s.items()},
            "sample": "%d whole frames (LCN + 75 U-Net tiles + watershed + 5 x (FFN match + PR-GLS) + replay), nothing "
                      "extrapolated" % n,
            "note": "CPU restatement (torch/oneDNN fp32 for the Keras graphs, scipy + restated scikit-image for the "
                    "watershed, NumPy fp64 EM) -- TensorFlow is not installable in this image"}


def cpu_baseline():
    return cpu_run(steps=1, warmup=1, budget_s=20.0)


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_run(steps=max(1, args.steps), warmup=args.warmup)
    line = {"impl": "reference", "metric": "voxels/s", "value": base["value"], "unit": "voxels/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
            "steps_timed": base["frames_timed"], "ms_per_step": base["frame_seconds"] * 1e3,
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (U-Net/FFN) + f64 (watershed, PR-GLS EM)", "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "frames_per_s": 1.0 / base["frame_seconds"],
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
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
    ap.add_argument("--ws-lag", type=int, default=2, help="volumes the host runs ahead of the watershed it needs the cell count of")
    ap.add_argument("--reserve-sms", type=int, default=8, help="SMs kept out of the persistent conv grid while overlapping")
    ap.add_argument("--no-overlap", action="store_true", help="run segmentation and tracking back to back on one stream")
    ap.add_argument("--workload", default="c1", choices=["c1", "c3"],
                    help="c1 (default, the contract's line): one worm1 frame per step; c3: one 1024x1024x96 volume "
                         "spatially decomposed over the GPUs")
    ap.add_argument("--shape", type=int, nargs=3, default=None, help="c3 only: override the volume shape")
    ap.add_argument("--verify", action="store_true", help="c3 only: check every rank's block against a single-GPU run")
    ap.add_argument("--no-verify", action="store_true", help="skip the config-4 equality check (N > 1) before timing")
    ap.add_argument("--no-compare", action="store_true", help="skip the serial / no-watershed comparison arms (N = 1)")
    ap.add_argument("--no-c3", action="store_true", help="skip the config-3 strong-scaling measurement")
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
