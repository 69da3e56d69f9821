"""FFN match on the GPU -- drop-in for the inference half of CellTracker/ffn.py.

`FFN` (ffn.py:225-265) keeps `.predict(x122, batch_size)` and Keras-order `get_weights/set_weights`;
`initial_matching_ffn(ffn_model, ref, tgt, k_ptrs)` (ffn.py:268-327) and `normalize_points`
(ffn.py:330-374) keep their signatures.  The pair grid is never materialised (see csrc/ffn.cu).
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._device import WORKSPACE, aligned_ptr, require_cuda, stream_ptr, to_device

k_ptrs = 20            # ffn.py:20
NUMBER_FEATURES = 61   # ffn.py:26

_SHAPES = [(61, 512), (512,), (512,), (512,), (512,), (1024, 512), (512,), (512,), (512,), (512,), (512, 1), (1,)]


def _keras_default_weights(seed=0):
    rng = np.random.default_rng(seed)

    def glorot(fi, fo):
        lim = math.sqrt(6.0 / (fi + fo))
        return rng.uniform(-lim, lim, (fi, fo)).astype(np.float32)

    bn = [np.ones(512, np.float32), np.zeros(512, np.float32), np.zeros(512, np.float32), np.ones(512, np.float32)]
    return [glorot(61, 512)] + [b.copy() for b in bn] + [glorot(1024, 512)] + [b.copy() for b in bn] + \
           [glorot(512, 1), np.zeros(1, np.float32)]


class FFN:
    """GPU-resident FFN: 61 -> 512 (shared) x2 -> concat 1024 -> 512 -> 1 sigmoid (ffn.py:225-265)."""

    def __init__(self, weights=None):
        self._handle = None
        self._weights = None
        self.set_weights(weights if weights is not None else _keras_default_weights())

    def set_weights(self, weights):
        if len(weights) != len(_SHAPES):
            raise ValueError(f"expected {len(_SHAPES)} weight arrays, got {len(weights)}")
        ws = []
        for w, shp in zip(weights, _SHAPES):
            w = np.asarray(w, dtype=np.float32)
            if w.shape != shp:
                raise ValueError(f"weight shape {w.shape} does not match {shp}")
            ws.append(w)
        require_cuda()
        flat = np.ascontiguousarray(np.concatenate([w.reshape(-1) for w in ws]))
        handle = C.c_void_p()
        _lib.check(_lib.lib().ct_ffn_create(flat.ctypes.data, flat.size, C.byref(handle)))
        self._release()
        self._handle, self._weights = handle, ws

    def get_weights(self):
        return [w.copy() for w in self._weights]

    def save_weights(self, path):
        np.savez(path, *self._weights)

    def load_weights(self, path):
        """`.npz` (save_weights / io_formats.convert_keras_h5) or a Keras `.h5` file (tracker.py:1121; needs h5py)."""
        from .io_formats import load_weight_file
        try:
            self.set_weights(load_weight_file(path))
        except (OSError, KeyError) as e:
            raise ValueError(f"Failed to load the FFN model from {path}: {e}") from e

    def _release(self):
        if self._handle is not None:
            _lib.lib().ct_ffn_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def __call__(self, x):
        return self.predict(x)

    def predict(self, x, batch_size=1024, verbose=0):
        """Keras predict: (rows, 122) [or the legacy pair [a61, b61], track.py:175] -> (rows, 1) float32."""
        if isinstance(x, (list, tuple)):
            x = np.concatenate([np.asarray(x[0]), np.asarray(x[1])], axis=1)
        arr = np.asarray(x)
        if arr.ndim != 2 or arr.shape[1] != 2 * NUMBER_FEATURES:
            raise ValueError(f"expected input of shape (rows, 122), got {arr.shape}")
        if arr.shape[0] == 0:
            return np.zeros((0, 1), np.float32)
        dev = to_device(arr.astype(np.float32, copy=False), torch.float32)
        rows = int(dev.shape[0])
        out = torch.empty(rows, dtype=torch.float32, device=dev.device)
        lib = _lib.lib()
        ws = WORKSPACE.get("ffn_predict", lib.ct_ffn_predict_workspace_bytes(rows))
        wp = aligned_ptr(ws)
        _lib.check(lib.ct_ffn_predict(self._handle, dev.data_ptr(), rows, out.data_ptr(), wp,
                                      ws.numel() - (wp - ws.data_ptr()), stream_ptr()))
        return out.cpu().numpy()[:, None]

    # ---- device-side operator
    def match_device(self, ref_dev, tgt_dev, k=k_ptrs, out=None):
        """ref (N,3), tgt (M,3) float64 CUDA tensors -> corr (M,N) float32 CUDA tensor."""
        n, m = int(ref_dev.shape[0]), int(tgt_dev.shape[0])
        if out is None:
            out = torch.empty((m, n), dtype=torch.float32, device=ref_dev.device)
        lib = _lib.lib()
        ws = WORKSPACE.get("ffn_match", lib.ct_ffn_match_workspace_bytes(n, m))
        wp = aligned_ptr(ws)
        _lib.check(lib.ct_ffn_match(self._handle, ref_dev.data_ptr(), n, tgt_dev.data_ptr(), m, int(k), out.data_ptr(),
                                    wp, ws.numel() - (wp - ws.data_ptr()), stream_ptr()), ValueError)
        return out


def knn_features(points, k=k_ptrs):
    """(n,3) -> (n, 3k+1) float32 features (ffn.py:288-304), host in / host out."""
    dev = to_device(np.asarray(points, dtype=np.float64), torch.float64)
    n = int(dev.shape[0])
    out = torch.empty((n, 3 * k + 1), dtype=torch.float32, device=dev.device)
    _lib.check(_lib.lib().ct_knn_features(dev.data_ptr(), n, int(k), out.data_ptr(), stream_ptr()), ValueError)
    return out.cpu().numpy()


def initial_matching_ffn(ffn_model, ref, tgt, k_ptrs=k_ptrs):
    """corr (M,N) float32 between all pairs of reference and target points (ffn.py:268-327)."""
    if not isinstance(ffn_model, FFN):
        raise TypeError("initial_matching_ffn needs a 3deecelltracker_b200.ffn.FFN model (no CPU fallback)")
    ref_dev = to_device(np.asarray(ref, dtype=np.float64), torch.float64)
    tgt_dev = to_device(np.asarray(tgt, dtype=np.float64), torch.float64)
    return ffn_model.match_device(ref_dev, tgt_dev, k_ptrs).cpu().numpy()


def normalize_points(points, return_para=False):
    """ffn.py:330-374: centre, scale by 3 * std of the projection on the first principal axis.
    Host arithmetic (a 3x3 eigen-problem; SURVEY a-11 'negligible')."""
    points = np.asarray(points)
    if points.ndim != 2:
        raise ValueError(f"Points should be a 2D table, but get {points.ndim}D")
    if points.shape[1] != 3:
        raise ValueError(f"Points should have 3D coordinates, but get {points.shape[1]}D")
    mean = np.mean(points, axis=0)
    centred = points - mean
    cov = centred.T @ centred
    evals, evecs = np.linalg.eigh(cov)
    axis = evecs[:, -1]
    std = np.std(centred @ axis)
    norm_points = (points - mean) / (3 * std)
    if return_para:
        return norm_points, (mean, 3 * std)
    return norm_points
