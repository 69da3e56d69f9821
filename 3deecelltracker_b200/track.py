"""PR-GLS for the U-Net workflow -- drop-in for CellTracker/track.py (hot-path functions only).

pr_gls_quick (track.py:11-114), initial_matching_quick (track.py:117-178), get_reference_vols /
get_remote_vols (track.py:575-610).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._device import WORKSPACE, aligned_ptr, stream_ptr, to_device
from .ffn import FFN

MODE_TRACK, MODE_LITE = 0, 1


class EmProblem:
    """One EM problem with device-resident inputs; outputs are allocated here."""

    def __init__(self, ref, tgt, corr, tracked=None, prior_given=False):
        self.ref = to_device(ref, torch.float64) if not isinstance(ref, torch.Tensor) else ref
        self.tgt = to_device(tgt, torch.float64) if not isinstance(tgt, torch.Tensor) else tgt
        if isinstance(corr, torch.Tensor):
            self.corr = corr.contiguous()
        else:
            c = np.asarray(corr)
            self.corr = to_device(c, torch.float64 if c.dtype != np.float32 else torch.float32)
        if self.corr.dtype not in (torch.float32, torch.float64):
            self.corr = self.corr.to(torch.float64)
        self.tracked = None
        if tracked is not None:
            self.tracked = to_device(tracked, torch.float64) if not isinstance(tracked, torch.Tensor) else tracked
        self.prior_given = prior_given
        n, m = int(self.ref.shape[0]), int(self.tgt.shape[0])
        if tuple(self.corr.shape) != (m, n):
            raise ValueError(f"corr has shape {tuple(self.corr.shape)}, expected ({m}, {n})")
        dev = self.ref.device
        self.post = torch.empty((m, n), dtype=torch.float64, device=dev)
        self.ref_out = torch.empty((n, 3), dtype=torch.float64, device=dev)
        self.coef = torch.empty((3, n), dtype=torch.float64, device=dev)
        self.tracked_out = None if self.tracked is None else torch.empty_like(self.tracked)
        self.iterations = torch.zeros(1, dtype=torch.int32, device=dev)


def run_em(problems, mode, beta, lambda_, max_iteration, vol, threshold):
    """Launch a batch of EM problems (one persistent CTA each) on the current stream."""
    lib = _lib.lib()
    prm = _lib.CtPrglsParams(mode=mode, max_iteration=int(max_iteration), beta=float(beta), lambda_=float(lambda_),
                             vol=float(vol), threshold=float(threshold))
    arr = (_lib.CtPrglsProblem * len(problems))()
    total = 512
    for i, p in enumerate(problems):
        n, m = int(p.ref.shape[0]), int(p.tgt.shape[0])
        l = 0 if p.tracked is None else int(p.tracked.shape[0])
        a = arr[i]
        a.ref, a.tgt, a.corr = p.ref.data_ptr(), p.tgt.data_ptr(), p.corr.data_ptr()
        a.tracked = None if p.tracked is None else p.tracked.data_ptr()
        a.post, a.ref_out, a.coef = p.post.data_ptr(), p.ref_out.data_ptr(), p.coef.data_ptr()
        a.tracked_out = None if p.tracked_out is None else p.tracked_out.data_ptr()
        a.iterations = p.iterations.data_ptr()
        a.n_ref, a.n_tgt, a.n_tracked = n, m, l
        a.corr_is_f64 = 1 if p.corr.dtype == torch.float64 else 0
        a.prior_given = 1 if p.prior_given else 0
        total += lib.ct_prgls_workspace_bytes(n, m, l)
    ws = WORKSPACE.get("prgls", total)
    wp = aligned_ptr(ws)
    _lib.check(lib.ct_prgls(C.byref(prm), arr, len(problems), wp, ws.numel() - (wp - ws.data_ptr()), stream_ptr()))
    return problems


def pr_gls_quick(X, Y, corr, BETA=300, max_iteration=20, LAMBDA=0.1, vol=1E8):
    """Coherent movements from the initial matching by PR-GLS (track.py:11-114).
    Returns (P (M,N), T_X (N,3), C (3,N)) as float64 ndarrays."""
    p = EmProblem(np.asarray(X, dtype=np.float64), np.asarray(Y, dtype=np.float64), corr)
    run_em([p], MODE_TRACK, BETA, LAMBDA, max_iteration, vol, 0.5)
    return p.post.cpu().numpy(), p.ref_out.cpu().numpy(), p.coef.cpu().numpy()


def initial_matching_quick(ffn_model, ref, tgt, k_ptrs):
    """corr (M,N) between all pairs (track.py:117-178).  The reference feeds a legacy two-input Keras model;
    here the same FFN object serves both conventions."""
    if not isinstance(ffn_model, FFN):
        raise TypeError("initial_matching_quick needs a 3deecelltracker_b200.ffn.FFN model (no CPU fallback)")
    ref_dev = to_device(np.asarray(ref, dtype=np.float64), torch.float64)
    tgt_dev = to_device(np.asarray(tgt, dtype=np.float64), torch.float64)
    return ffn_model.match_device(ref_dev, tgt_dev, k_ptrs).cpu().numpy()


def predict_one_rep_device(pre_dev, inter_dev, beta, coef_dev):
    """tracker.py:1269-1289 on device tensors: post = pre + (C G)^T."""
    post = torch.empty_like(pre_dev)
    _lib.check(_lib.lib().ct_predict_one_rep(pre_dev.data_ptr(), int(pre_dev.shape[0]), inter_dev.data_ptr(),
                                             int(inter_dev.shape[0]), float(beta), coef_dev.data_ptr(),
                                             post.data_ptr(), stream_ptr()))
    return post


def replay_fit_device(fit, tracked_prev_dev, proportiontocut=0.1):
    """tracker.py:1269-1289 for every repetition of one fitted volume pair, then the single-mode trimmed mean of
    tracker.py:1503-1507, in ONE library call (the time-lapse driver replays every volume of the recording in order on one
    rank; six calls per volume were host-bound).  fit: [(intermediate points (n,3), beta, coef (3,n))] per repetition."""
    n_rep, count = len(fit), int(tracked_prev_dev.shape[0])
    inter = (C.c_void_p * n_rep)(*[f[0].data_ptr() for f in fit])
    coef = (C.c_void_p * n_rep)(*[f[2].data_ptr() for f in fit])
    n_ref = (C.c_int * n_rep)(*[int(f[0].shape[0]) for f in fit])
    beta = (C.c_double * n_rep)(*[float(f[1]) for f in fit])
    scratch = torch.empty((2,) + tuple(tracked_prev_dev.shape), dtype=torch.float64, device=tracked_prev_dev.device)
    out = torch.empty_like(tracked_prev_dev)
    _lib.check(_lib.lib().ct_replay_fit(tracked_prev_dev.data_ptr(), count, n_rep, inter, n_ref, beta, coef,
                                        float(proportiontocut), scratch.data_ptr(), out.data_ptr(), stream_ptr()))
    return out


def trim_mean_device(stack_dev, proportiontocut=0.1):
    """scipy.stats.trim_mean(stack, p, axis=0) for a (E, L, 3) float64 CUDA tensor."""
    e = int(stack_dev.shape[0])
    out = torch.empty(stack_dev.shape[1:], dtype=torch.float64, device=stack_dev.device)
    _lib.check(_lib.lib().ct_trim_mean(stack_dev.contiguous().data_ptr(), e, out.numel(), float(proportiontocut),
                                       out.data_ptr(), stream_ptr()))
    return out


def get_remote_vols(ensemble, vol):
    """track.py:605-610."""
    interval = (vol - 1) // ensemble
    start = (vol - 1) % ensemble + 1
    return list(range(start, vol - interval + 1, interval))


def get_reference_vols(ensemble, vol, adjacent=False):
    """track.py:575-602."""
    if not ensemble:
        return [vol - 1]
    if vol - 1 < ensemble:
        return list(range(1, vol))
    if adjacent:
        return list(range(vol - ensemble, vol))
    return get_remote_vols(ensemble, vol)
