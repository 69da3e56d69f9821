"""Frame pipeline of a time-lapse: the two hot paths of `Tracker.track_one_vol` (tracker.py:1473-1536) on separate streams.

Three facts of the reference make the loop pipelinable without changing a single result:
  * segmentation of volume t+1 (`_normalize_image` + `unet3_prediction`, tracker.py:662-669) does not depend on the
    tracking of volume t;
  * the expensive part of tracking, `_fit_ffn_prgls` (tracker.py:1224-1254: REP_NUM_PRGLS x (FFN match + PR-GLS)), fits
    transforms between the SEGMENTED point sets of two volumes -- it does not read any tracking result;
  * only the replay `_predict_one_rep` (tracker.py:1269-1289, five tiny Gaussian-kernel products) carries state from one
    volume to the next.
And the two stages want opposite things from the machine: the U-Net is a persistent tensor-core kernel that fills every
SM, the PR-GLS EM is ONE latency-bound fp64 CTA per problem (5 sequential repetitions x 19 iterations).  Back to back,
the EM leaves 147 SMs idle for a third of the frame.

`FramePipeline.step` therefore (1) joins the fit submitted `depth` steps ago and replays it on the tracked cells,
(2) enqueues this step's fit on a high-priority side stream (`depth` of them take turns, so `depth` fits are in
flight), (3) enqueues the segmentation on the caller's stream, with a few SMs kept out of the convolution's persistent
grid (`ct_set_reserved_sms`) so the EM CTAs and the small FFN kernels never queue behind a convolution CTA.  Every step
still performs exactly one segmentation, one fit and one replay; results are those of the serial order.
"""
import collections

import torch

from . import _lib
from . import watershed as _ws
from .preprocess import normalize_image_device
from .track import MODE_TRACK, EmProblem, replay_fit_device, run_em

REP_NUM_PRGLS = 5          # tracker.py:45
K_POINTS = 20              # tracker.py:1259


class FramePipeline:
    def __init__(self, unet_model, ffn_model, noise_level, beta_tk, lambda_tk, maxiter_tk, shrink=(24, 24, 2),
                 overlap=True, reserve_sms=8, depth=2, ws_lag=2):
        self.unet, self.ffn = unet_model, ffn_model
        self.noise_level, self.shrink = noise_level, tuple(shrink)
        self.beta_tk, self.lambda_tk, self.max_iteration = beta_tk, lambda_tk, maxiter_tk
        self.overlap = bool(overlap)
        self.reserve_sms = int(reserve_sms)
        self.depth = max(1, int(depth))
        # raw-stack mode: how many volumes the host runs ahead of the watershed whose cell count it needs (the count sizes
        # the fit's launches).  With a lag of 1 the host waits for the watershed of volume t-1 before it can enqueue the
        # segmentation of volume t+1 -- and that watershed, squeezed in beside the convolutions of volume t, finishes
        # about when they do, so the main stream runs dry; a lag of 2 keeps a whole volume of work enqueued.
        self.ws_lag = min(max(1, int(ws_lag)), 6)
        self._streams = None
        self._pending = collections.deque()          # (stream, fit, tracked_prev or None)
        self._tracked = None                         # running tracked coordinates when the caller does not pass them
        self._count = 0
        # raw-stack mode (step_raw): watershed results whose cell count is still on its way to the host, and the
        # point set of the last volume whose count is known
        self._ws_stream = None
        self._segmented = collections.deque()        # (Segmentation, pinned scalars, event)
        self._prev_points = None
        self._first_points = None
        self.collect_fits = False                    # True: joined fits are kept in `collected` instead of replayed
        self.collected = []                          # (timelapse.py replays them on the rank that owns the state)
        self.z_xy_ratio, self.ws_method, self.min_size, self.cell_num = 1.0, "min_size", 0, 0
        self.trace = None                            # debugging: a list collects the host's waits for the cell counts (s)
        self._pinned_ring, self._pinned_next = None, 0  # pinned cell-count scalars, reused every 8 volumes (lag <= 3)
        self._reserved = False                       # reserve_small_blocks has run

    # ---- the stages, each on the current stream
    def segment(self, raw_dev):
        """tracker.py:662-669: raw (x,y,z) CUDA tensor -> probability map (x,y,z) float32 CUDA tensor."""
        norm = normalize_image_device(raw_dev, self.noise_level, (27, 27, 1))
        return self.unet.prediction_device(norm, self.shrink)

    def fit(self, seg_prev_dev, seg_cur_dev):
        """tracker.py:1224-1254 for one source volume: REP_NUM_PRGLS x (FFN match + PR-GLS with beta * 0.8**i).
        (N,3), (M,3) float64 CUDA -> [(intermediate points, beta, C)] per repetition."""
        inter, out = seg_prev_dev, []
        for i in range(REP_NUM_PRGLS):
            beta = self.beta_tk * (0.8 ** i)
            corr = self.ffn.match_device(inter, seg_cur_dev, K_POINTS)
            p = run_em([EmProblem(inter, seg_cur_dev, corr)], MODE_TRACK, beta, self.lambda_tk, self.max_iteration,
                       1e8, 0.5)[0]
            out.append((inter, beta, p.coef))
            inter = p.ref_out
        return out

    def replay(self, fit, tracked_prev_dev):
        """tracker.py:1269-1289 + the single-mode trimmed mean of :1503-1507: (L,3) -> (L,3)."""
        return replay_fit_device([(i.contiguous(), b, c.contiguous()) for i, b, c in fit], tracked_prev_dev.contiguous(), 0.1)

    def track(self, seg_prev_dev, seg_cur_dev, tracked_prev_dev):
        return self.replay(self.fit(seg_prev_dev, seg_cur_dev), tracked_prev_dev)

    def reset(self, tracked0_dev=None):
        """Forget fits in flight; `tracked0_dev` seeds the running tracked coordinates."""
        self.flush()
        self._tracked = tracked0_dev

    def reset_raw(self):
        """Raw-stack mode: forget every volume seen so far (the next `step_raw` volume is volume 1 again)."""
        self.flush_raw()
        self._prev_points = self._first_points = self._tracked = None
        self.collected = []

    # ---- one pipelined step
    def _join_oldest(self):
        stream, fit, tracked_prev = self._pending.popleft()
        main = torch.cuda.current_stream()
        main.wait_stream(stream)
        for inter, _, coef in fit:                   # allocated on the side stream, read by the replay on this one
            inter.record_stream(main)
            coef.record_stream(main)
        if self.collect_fits:
            self.collected.append(fit)
            return None
        prev = tracked_prev if tracked_prev is not None else self._tracked
        if prev is None:
            raise ValueError("no tracked coordinates to replay onto: pass tracked_prev_dev or call reset(tracked0_dev)")
        tracked = self.replay(fit, prev)
        if tracked_prev is None:
            self._tracked = tracked
        return tracked

    def step(self, raw_next_dev, seg_prev_dev, seg_cur_dev, tracked_prev_dev=None):
        """One frame of each stage.  Segments `raw_next_dev`; submits the fit seg_prev -> seg_cur; returns
        (probability map, tracked coordinates of the fit submitted `depth` steps ago, or None while the pipeline
        fills).  tracked_prev_dev = None replays onto the pipeline's running coordinates (see `reset`).  Everything
        returned is safe to use on the current stream."""
        if not self.overlap:
            prob = self.segment(raw_next_dev)
            prev = tracked_prev_dev if tracked_prev_dev is not None else self._tracked
            if prev is None:
                raise ValueError("no tracked coordinates to replay onto: pass tracked_prev_dev or call reset(tracked0_dev)")
            tracked = self.track(seg_prev_dev, seg_cur_dev, prev)
            if tracked_prev_dev is None:
                self._tracked = tracked
            return prob, tracked
        if self._streams is None:
            self._streams = [torch.cuda.Stream(priority=-1) for _ in range(self.depth)]
        lib = _lib.lib()
        main = torch.cuda.current_stream()
        tracked = self._join_oldest() if len(self._pending) >= self.depth else None
        side = self._streams[self._count % self.depth]
        self._count += 1
        side.wait_stream(main)                                # inputs are ordered before the fit
        old = lib.ct_set_reserved_sms(self.reserve_sms)
        try:
            with torch.cuda.stream(side):
                fit = self.fit(seg_prev_dev, seg_cur_dev)
            self._pending.append((side, fit, tracked_prev_dev))
            prob = self.segment(raw_next_dev)
        finally:
            lib.ct_set_reserved_sms(old)
        return prob, tracked

    def flush(self):
        """Join every fit still in flight; returns their tracked coordinates, oldest first."""
        out = []
        while self._pending:
            out.append(self._join_oldest())
        return out

    def reserve_small_blocks(self, megabytes=64):
        """Grow the caching allocator's small-block pools of the pipeline's streams now.  A time-lapse keeps every fitted
        transform (5 x (points, coefficients) per volume) until the replay, so those pools grow while it runs, 2 MB
        segment by segment, and each cudaMalloc is an implicit synchronisation of the whole device: the host then stalls
        for the 2-3 volumes of work it has queued [measured: 50-140 ms gaps every few volumes in one run out of three].
        Call once after the first volumes (the streams exist by then)."""
        self._reserved = True
        streams = [torch.cuda.current_stream()] + list(self._streams or []) + ([self._ws_stream] if self._ws_stream else [])
        n = max(1, int(megabytes))
        for st in streams:
            with torch.cuda.stream(st):
                held = [torch.empty(1 << 20, dtype=torch.uint8, device="cuda") for _ in range(n)]
            del held

    # ---- raw-stack mode: segmentation -> watershed -> centres -> fit, the whole Tracker.track_one_vol chain
    def configure_watershed(self, z_xy_ratio, method="min_size", min_size=0, cell_num=0):
        """Parameters of Tracker._watershed (tracker.py:671-684) for `step_raw`."""
        self.z_xy_ratio, self.ws_method, self.min_size, self.cell_num = float(z_xy_ratio), method, min_size, cell_num

    def segment_cells(self, raw_dev):
        """tracker.py:605-650 on the device: LCN + U-Net + watershed + centres of mass; nothing leaves HBM except the
        four scalars (cell count ...), which go to pinned memory asynchronously.
        With `overlap` the watershed runs on its own stream: its priority floods are a few hundred latency-bound warps
        (one per connected component) that fit beside the next volume's tensor-core convolutions, exactly like the EM.
        `seg.ready` is the event that marks the label image / centres / pinned scalars complete."""
        prob = self.segment(raw_dev)
        main = torch.cuda.current_stream()
        if self.overlap:
            if self._ws_stream is None:
                self._ws_stream = torch.cuda.Stream(priority=-1)
            produced = torch.cuda.Event()
            produced.record(main)
            self._ws_stream.wait_event(produced)
            prob.record_stream(self._ws_stream)
            stream = self._ws_stream
        else:
            stream = main
        with torch.cuda.stream(stream):
            seg = _ws.segment_device(prob, self.z_xy_ratio, self.ws_method, self.min_size, self.cell_num)
            # A page-locked allocation is an implicit synchronisation point of the whole device (and the caching host
            # allocator does allocate whenever its few cached blocks are still tied to unfinished watersheds): one in the
            # loop stalled the host for 40-140 ms every few volumes [measured].  The scalars land in a ring allocated once.
            if self._pinned_ring is None:
                self._pinned_ring = [torch.empty(4, dtype=torch.int32).pin_memory() for _ in range(8)]
            pinned = self._pinned_ring[self._pinned_next % len(self._pinned_ring)]
            self._pinned_next += 1
            pinned.copy_(seg.scalars, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
        seg.ready = ev
        return prob, seg, pinned, ev

    def _resolve_segmented(self):
        """Point sets (n,3) of the volumes whose cell count has arrived, oldest first.  Waiting here does not idle
        the GPU: the caller has already enqueued the next volume's segmentation behind the one waited for."""
        out = []
        while self._segmented:
            seg, pinned, ev = self._segmented.popleft()
            if self.trace is not None:
                import time
                t0 = time.perf_counter()
                ev.synchronize()
                self.trace.append(time.perf_counter() - t0)
            else:
                ev.synchronize()
            n = int(pinned[0])
            if n > _ws.MAX_CELLS:
                raise ValueError(f"watershed found {n} cells; at most {_ws.MAX_CELLS} are supported")
            if n == 0:
                raise ValueError("No cell was detected by watershed! Try to reduce the min_size.")   # tracker.py:643
            out.append((seg.centres_real[:n], ev))
        return out

    def _submit_fits(self, point_sets):
        for pts, ev in point_sets:
            if self._prev_points is not None:
                if self._streams is None:
                    self._streams = [torch.cuda.Stream(priority=-1) for _ in range(self.depth)]
                side = self._streams[self._count % self.depth] if self.overlap else torch.cuda.current_stream()
                self._count += 1
                # the fit depends on the two watershed results only (the older one precedes `ev` on the same stream),
                # NOT on the segmentation of the next volume that is already enqueued behind them
                side.wait_event(ev)
                pts.record_stream(side)
                self._prev_points.record_stream(side)
                with torch.cuda.stream(side):
                    fit = self.fit(self._prev_points, pts)
                self._pending.append((side, fit, None))
            else:
                self._first_points = pts
                if self._tracked is None:
                    self._tracked = pts.clone()      # initiate_tracking: volume 1's centres are the tracked set
            self._prev_points = pts

    def step_raw(self, raw_dev):
        """One volume of the whole chain on a RAW stack: joins + replays the oldest fit when `depth` are in flight,
        enqueues LCN + U-Net + watershed of `raw_dev`, then -- while that runs -- takes the cell count of the volume
        before and submits its fit (previous point set -> that volume's point set) to a side stream.
        Returns (probability map, Segmentation, tracked coordinates or None while the pipeline fills)."""
        lib = _lib.lib()
        tracked = self._join_oldest() if len(self._pending) >= self.depth else None
        old = lib.ct_set_reserved_sms(self.reserve_sms if self.overlap else 0)
        try:
            prob, seg, pinned, ev = self.segment_cells(raw_dev)
            self._segmented.append((seg, pinned, ev))
            if not self.overlap:
                self._submit_fits(self._resolve_segmented())      # serial order: count now, fit now
                while self._pending:
                    tracked = self._join_oldest()
            else:
                ready = [self._segmented.popleft()] if len(self._segmented) > self.ws_lag else []
                keep = list(self._segmented)
                self._segmented = collections.deque(ready)
                pts = self._resolve_segmented()
                self._segmented = collections.deque(keep)
                self._submit_fits(pts)
        finally:
            lib.ct_set_reserved_sms(old)
        return prob, seg, tracked

    def flush_raw(self, extra_target=None):
        """Resolve the last segmented volume, submit its fit and join everything; tracked coordinates, oldest first.
        extra_target: a further point set (the first volume of the NEXT rank's block, timelapse.py) -- the fit from this
        block's last volume onto it is submitted on a side stream together with the last local fit, not after it."""
        self._submit_fits(self._resolve_segmented())
        if extra_target is not None:
            ev = torch.cuda.Event()
            ev.record()
            self._submit_fits([(extra_target, ev)])
        return self.flush()
