"""CPU oracle for the FFN match path: k-NN features, FFN forward on the M x N pair grid,
normalize_points (NumPy float32/float64).

TEST INFRASTRUCTURE ONLY -- see oracle/prgls.py header for who may import this.

Parity status: feature building + pair-grid ordering are PINNED (tests/golden holds the exact
(M*N,122) batches the unmodified reference `initial_matching_ffn` / `initial_matching_quick` hand to
`ffn_model.predict`, captured through a recording duck-typed model by oracle/make_golden.py).
The FFN forward itself (ffn.py:225-265, Keras Dense/BatchNormalization/LeakyReLU) is UNPINNED:
TensorFlow is not installable here and the reference ships no weights or test vectors; the
restatement follows the source with Keras defaults written out (BN eps 1e-3 in inference form,
LeakyReLU alpha 0.3, Dense(1) + sigmoid with bias).

Weight container ("Keras order" of FFN.get_weights()):
  W1 (61,512); bn1 gamma, beta, mean, var (512 each); W2 (1024,512); bn2 gamma, beta, mean, var;
  W3 (512,1); b3 (1).
"""
import math

import numpy as np

K_PTRS = 20
NUMBER_FEATURES = 61
BN_EPS = 1e-3
LEAKY_ALPHA = 0.3


def knn_features(points_nx3, k_ptrs=K_PTRS):
    """(n, 3k+1) float32 features.  Reference: ffn.py:288-304 / track.py:137-155.
    For each point: the k+1 nearest points of its own set (itself first, distance 0); mean_dist is the
    mean of all k+1 distances INCLUDING the zero self distance; features = k neighbour offsets divided
    by mean_dist (row-major (k,3)) followed by mean_dist; stored as float32."""
    pts = np.asarray(points_nx3, dtype=np.float64)
    n = pts.shape[0]
    if n < k_ptrs + 1:
        raise ValueError(f"Expected n_neighbors <= n_samples, but n_samples = {n}, n_neighbors = {k_ptrs + 1}")
    out = np.zeros((n, 3 * k_ptrs + 1), dtype=np.float32)
    for i in range(n):
        diff = pts - pts[i]
        d2 = np.sum(diff * diff, axis=1)
        order = np.argsort(d2, kind="stable")[:k_ptrs + 1]
        dist = np.sqrt(d2[order])
        mean_dist = np.mean(dist)
        rel = (pts[order[1:]] - pts[order[0]]) / mean_dist
        row = np.zeros(3 * k_ptrs + 1)
        row[:3 * k_ptrs] = rel.reshape(-1)
        row[3 * k_ptrs] = mean_dist
        out[i] = row
    return out


def pair_grid(ref_feat_nxf, tgt_feat_mxf):
    """(M*N, 2F): row m*N+n = [ref n | tgt m].  Reference: ffn.py:306-307,319-324."""
    n, m = ref_feat_nxf.shape[0], tgt_feat_mxf.shape[0]
    a = np.broadcast_to(ref_feat_nxf[None], (m, n, ref_feat_nxf.shape[1]))
    b = np.broadcast_to(tgt_feat_mxf[:, None], (m, n, tgt_feat_mxf.shape[1]))
    return np.concatenate([a, b], axis=2).reshape(m * n, -1)


def random_weights(seed=0):
    """Seeded parity weights: Glorot-uniform Dense kernels (keras default), BN gamma U(0.5,1.5),
    beta/mean N(0,0.1^2), var U(0.5,1.5), b3 N(0,0.05^2)."""
    rng = np.random.default_rng(seed)

    def glorot(fi, fo):
        lim = math.sqrt(6.0 / (fi + fo))
        return rng.uniform(-lim, lim, (fi, fo)).astype(np.float32)

    def bn(c):
        return [rng.uniform(0.5, 1.5, c).astype(np.float32), (rng.standard_normal(c) * 0.1).astype(np.float32),
                (rng.standard_normal(c) * 0.1).astype(np.float32), rng.uniform(0.5, 1.5, c).astype(np.float32)]

    return [glorot(61, 512)] + bn(512) + [glorot(1024, 512)] + bn(512) + \
           [glorot(512, 1), (rng.standard_normal(1) * 0.05).astype(np.float32)]


def _bn(x, g, b, mu, var):
    return (x - mu) * (g / np.sqrt(var + x.dtype.type(BN_EPS))) + b


def _leaky(x):
    return np.where(x > 0, x, x.dtype.type(LEAKY_ALPHA) * x)


class FFNOracle:
    """Duck-type of the Keras FFN: .predict(x122, batch_size) (ffn.py:260-265) and the legacy
    two-input form .predict([a61, b61], batch_size) used by Tracker (track.py:175)."""

    def __init__(self, weights, dtype=np.float32):
        self.dtype = dtype
        self.w = [np.asarray(a, dtype=dtype) for a in weights]

    def _forward(self, x):
        W1, g1, b1, m1, v1, W2, g2, b2, m2, v2, W3, b3 = self.w
        x = np.asarray(x, dtype=self.dtype)
        f1 = _leaky(_bn(x[:, :61] @ W1, g1, b1, m1, v1))        # Dense -> BN -> LeakyReLU (ffn.py:240-245)
        f2 = _leaky(_bn(x[:, 61:] @ W1, g1, b1, m1, v1))
        h = _leaky(_bn(np.concatenate([f1, f2], axis=1) @ W2, g2, b2, m2, v2))
        z = h @ W3 + b3
        return 1.0 / (1.0 + np.exp(-z))

    def predict(self, x, batch_size=1024, verbose=0):
        if isinstance(x, (list, tuple)):
            x = np.concatenate([np.asarray(x[0]), np.asarray(x[1])], axis=1)
        outs = [self._forward(x[i:i + batch_size]) for i in range(0, x.shape[0], batch_size)]
        return np.concatenate(outs, axis=0).astype(np.float32 if self.dtype == np.float32 else np.float64)


def initial_matching_ffn(ffn_model, ref, tgt, k_ptrs=K_PTRS):
    """corr (M,N).  Reference: ffn.py:268-327."""
    fr = knn_features(ref, k_ptrs)
    ft = knn_features(tgt, k_ptrs)
    pred = ffn_model.predict(pair_grid(fr, ft), batch_size=1024)
    return np.reshape(pred, (tgt.shape[0], ref.shape[0]))


def initial_matching_quick(ffn_model, ref, tgt, k_ptrs=K_PTRS):
    """corr (M,N) through the legacy two-input model.  Reference: track.py:117-178."""
    fr = knn_features(ref, k_ptrs)
    ft = knn_features(tgt, k_ptrs)
    g = pair_grid(fr, ft)
    pred = ffn_model.predict([g[:, :fr.shape[1]], g[:, fr.shape[1]:]], batch_size=1024)
    return np.reshape(pred, (tgt.shape[0], ref.shape[0]))


def normalize_points(points, return_para=False):
    """Centre, scale by 3*std of the projection on the first principal axis.  Reference: ffn.py:330-374
    (sklearn PCA(1): projection of centred data on the leading right singular vector; np.std, ddof=0;
    the sign convention of the axis does not affect the std)."""
    points = np.asarray(points)
    if points.ndim != 2:
        raise ValueError(f"Points should be a 2D table, but get {points.ndim}D")
    if points.shape[1] != 3:
        raise ValueError(f"Points should have 3D coordinates, but get {points.shape[1]}D")
    mean = np.mean(points, axis=0)
    centred = points - mean
    _, _, vt = np.linalg.svd(centred, full_matrices=False)
    std = np.std(centred @ vt[0])
    norm = (points - mean) / (3 * std)
    return (norm, (mean, 3 * std)) if return_para else norm
