"""Frame pipeline of a time-lapse: the two hot paths of `Tracker.track_one_vol` (tracker.py:1473-1536) on two streams.

Segmentation of volume t+1 (`_normalize_image` + `unet3_prediction`, tracker.py:662-669) does not depend on the
tracking of volume t (`_fit_ffn_prgls` x REP_NUM_PRGLS + `_predict_one_rep` + `trim_mean`, tracker.py:1224-1289,1507),
and the two stages want opposite things from the machine: the U-Net is a persistent tensor-core kernel that fills
every SM, the PR-GLS EM is ONE latency-bound fp64 CTA per problem (5 sequential repetitions x 19 iterations).  Run back
to back the EM leaves 147 SMs idle for a third of the frame.  `FramePipeline.step` enqueues the tracking stage of the
previous volume on a high-priority side stream and the segmentation of the next volume on the caller's stream, with
one SM kept out of the convolution's persistent grid (`ct_set_reserved_sms`) so the EM CTA never waits behind a
convolution CTA; the streams join before `step` returns, so a step is still one frame of each stage.
"""
import torch

from . import _lib
from .preprocess import normalize_image_device
from .track import MODE_TRACK, EmProblem, predict_one_rep_device, run_em, trim_mean_device

REP_NUM_PRGLS = 5          # tracker.py:45
K_POINTS = 20              # tracker.py:1259


class FramePipeline:
    def __init__(self, unet_model, ffn_model, noise_level, beta_tk, lambda_tk, maxiter_tk, shrink=(24, 24, 2),
                 overlap=True, reserve_sms=1):
        self.unet, self.ffn = unet_model, ffn_model
        self.noise_level, self.shrink = noise_level, tuple(shrink)
        self.beta_tk, self.lambda_tk, self.max_iteration = beta_tk, lambda_tk, maxiter_tk
        self.overlap = bool(overlap)
        self.reserve_sms = int(reserve_sms)
        self._side = None

    # ---- the two stages, each on the current stream
    def segment(self, raw_dev):
        """tracker.py:662-669: raw (x,y,z) CUDA tensor -> probability map (x,y,z) float32 CUDA tensor."""
        norm = normalize_image_device(raw_dev, self.noise_level, (27, 27, 1))
        return self.unet.prediction_device(norm, self.shrink)

    def fit_predict(self, seg_prev_dev, seg_cur_dev, tracked_prev_dev):
        """tracker.py:1224-1289 for one source volume: REP_NUM_PRGLS x (FFN match + PR-GLS with beta * 0.8**i), the
        fitted transforms replayed on the tracked cells.  (N,3), (M,3), (L,3) float64 CUDA -> (L,3)."""
        inter, pred = seg_prev_dev, tracked_prev_dev
        for i in range(REP_NUM_PRGLS):
            beta = self.beta_tk * (0.8 ** i)
            corr = self.ffn.match_device(inter, seg_cur_dev, K_POINTS)
            p = run_em([EmProblem(inter, seg_cur_dev, corr)], MODE_TRACK, beta, self.lambda_tk, self.max_iteration,
                       1e8, 0.5)[0]
            pred = predict_one_rep_device(pred, inter, beta, p.coef)
            inter = p.ref_out
        return pred

    def track(self, seg_prev_dev, seg_cur_dev, tracked_prev_dev):
        """Single-mode prediction of tracker.py:1503-1507 (one source volume, trimmed mean over a stack of one)."""
        return trim_mean_device(self.fit_predict(seg_prev_dev, seg_cur_dev, tracked_prev_dev)[None], 0.1)

    # ---- one pipelined step
    def step(self, raw_next_dev, track_args):
        """Segment `raw_next_dev` (volume t+1) while volume t is tracked: track_args = (segmented points of t-1,
        segmented points of t, tracked points of t-1), all float64 CUDA tensors produced before this call on the
        current stream.  Returns (prob of t+1, predicted coordinates of t); both are safe to use on the current
        stream when the call returns."""
        if not self.overlap:
            return self.segment(raw_next_dev), self.track(*track_args)
        if self._side is None:
            self._side = torch.cuda.Stream(priority=-1)
        lib = _lib.lib()
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)                          # inputs (and the previous step) are ordered before
        old = lib.ct_set_reserved_sms(self.reserve_sms)
        try:
            with torch.cuda.stream(self._side):
                tracked = self.track(*track_args)
            prob = self.segment(raw_next_dev)
        finally:
            lib.ct_set_reserved_sms(old)
        main.wait_stream(self._side)                          # join
        tracked.record_stream(main)
        return prob, tracked
