"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

TEST INFRASTRUCTURE ONLY.  Run as ``python -m oracle.make_golden`` from the repo root, in a container
where /root/reference is mounted.  The outputs are committed; nothing reads /root/reference at test,
smoke or bench time.

What is executed verbatim (via oracle/_ref_shim.py or by exec-ing function source read from the
reference file at generation time -- no reference source is copied into this repo):
  track.py       pr_gls_quick, initial_matching_quick, get_reference_vols
  trackerlite.py prgls_with_two_ref, prgls_quick, simple_match, estimate_posterior,
                 solve_movements_ref, get_volumes_list
  ffn.py         initial_matching_ffn, normalize_points (function source exec'd; module imports TF)
  unet3d.py      unet3_prediction + _get_sizes_padded_im (function source exec'd around a duck model)
  tracker.py     Tracker._predict_one_rep (function source exec'd)
"""
import ast
import itertools
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import _ref_shim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CSV = os.path.join(_ref_shim.REFERENCE_ROOT, "Examples", "use_stardist", "worm3_points_t1.csv")


def _exec_functions(rel_path, names, namespace):
    """exec the source of the named top-level functions / methods of a reference file."""
    path = os.path.join(_ref_shim.REFERENCE_ROOT, rel_path)
    src = open(path).read()
    tree = ast.parse(src)
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names:
            found[node.name] = ast.get_source_segment(src, node)
    for n in names:
        code = found[n]
        # methods are indented; dedent
        lines = code.split("\n")
        ind = len(lines[0]) - len(lines[0].lstrip())
        code = "\n".join(l[ind:] if l[:ind].strip() == "" else l for l in lines)
        exec(compile(code, path + ":" + n, "exec"), namespace)
    return namespace


class RecordingModel:
    """Duck-typed keras model: records what the reference passes to predict, returns a fixed pattern."""

    def __init__(self):
        self.calls = []

    def predict(self, x, batch_size=None):
        if isinstance(x, (list, tuple)):
            x = np.concatenate(x, axis=1)
        self.calls.append(np.array(x, copy=True))
        # deterministic pseudo-probabilities in (0,1) that depend on the row content
        s = np.sin(np.asarray(x, dtype=np.float64).sum(axis=1) * 12.9898) * 43758.5453
        return (s - np.floor(s)).astype(np.float32)[:, None]


def synth_target(ref, rng, affine_level=0.05, noise=0.002, drop=0.05, add=0.05):
    """Affine-perturbed copy of ref with dropped + added points (semantics of ffn.py:29-54)."""
    mean = ref.mean(axis=0)
    c = ref - mean
    scale = np.abs(c).max()
    a = np.eye(3) + (rng.random((3, 3)) - 0.5) * affine_level
    t = (c / scale) @ a + (rng.random(c.shape) - 0.5) * 4 * noise
    t = t * scale + mean
    keep = rng.random(len(t)) > drop
    t = t[keep]
    n_add = int(add * len(ref))
    extra = rng.uniform(ref.min(axis=0), ref.max(axis=0), (n_add, 3))
    t = np.concatenate([t, extra], axis=0)
    return t[rng.permutation(len(t))]


def synth_corr(ref, tgt, rng, width):
    """A plausible FFN output: high for truly near pairs, noisy elsewhere, in (0,1)."""
    d2 = ((ref[None] - tgt[:, None]) ** 2).sum(axis=2)
    return np.clip(np.exp(-d2 / (2 * width ** 2)) * 0.95 + rng.random(d2.shape) * 0.3, 0, 1)


def adversarial():
    """tests/golden/pr_gls_adversarial.npz: pr_gls_quick (track.py:11-114, unmodified) on M-step systems that stress
    an elimination WITHOUT row exchanges (the GPU kernel's solver): reference points far from every target so that
    their posterior column sums are exactly 0 (pivot = lambda sigma^2 alone), near-duplicate reference points (almost
    equal Gram rows) and lambda = 1e-5.  Own RNG stream: the other golden files do not change."""
    track, _ = _ref_shim.load()
    pts = np.loadtxt(CSV)
    rng = np.random.default_rng(777)

    def corr_of(X, Y):
        d2 = ((X[None] - Y[:, None]) ** 2).sum(axis=2)
        return np.clip(np.exp(-d2 / 72.0) * 0.95 + rng.random(d2.shape) * 0.3, 0, 1)

    Xa = pts.copy()
    Xa[:12] += np.array([4000.0, -3000.0, 2500.0])
    Ya = pts[12:] + rng.normal(0, 1.5, pts[12:].shape)
    ca = corr_of(Xa, Ya)
    Xb = np.concatenate([pts[:100], pts[:20] + rng.normal(0, 1e-6, (20, 3))])
    Yb = pts[:110] + rng.normal(0, 2.0, (110, 3))
    cb = corr_of(Xb, Yb)
    out = {}
    for name, (X, Y, cr, kw) in {
        "far_cols": (Xa, Ya, ca, dict(BETA=300, max_iteration=20, LAMBDA=1e-5)),
        "far_cols_b1000": (Xa, Ya, ca, dict(BETA=1000, max_iteration=10, LAMBDA=1e-5)),
        "near_dup": (Xb, Yb, cb, dict(BETA=1000, max_iteration=10, LAMBDA=1e-5)),
    }.items():
        P, TX, C = track.pr_gls_quick(X.copy(), Y.copy(), cr.copy(), **kw)
        d = dict(X=X, Y=Y, corr=cr, P=P, T_X=TX, C=C, **{k: np.float64(v) for k, v in kw.items()})
        out.update({f"{name}__{k}": v for k, v in d.items()})
    np.savez_compressed(os.path.join(GOLD, "pr_gls_adversarial.npz"), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    if "--only-adversarial" in sys.argv:
        return adversarial()
    track, lite = _ref_shim.load()
    pts = np.loadtxt(CSV)            # worm3, 180 x 3
    rng = np.random.default_rng(20260101)

    # ---------------------------------------------------------------- pr_gls_quick (track.py:11)
    cases = {}
    worm_xyz = pts * np.array([1.0, 1.0, 1.0])
    tgt = synth_target(worm_xyz, rng)
    corr = synth_corr(worm_xyz, tgt, rng, 6.0)
    for name, (X, Y, cr, kw) in {
        "worm3_single": (worm_xyz, tgt, corr, dict(BETA=300, max_iteration=20, LAMBDA=0.1)),
        "worm3_ensemble": (worm_xyz, tgt, corr, dict(BETA=1000, max_iteration=10, LAMBDA=1e-5)),
        "small": (worm_xyz[:30], tgt[:25], synth_corr(worm_xyz[:30], tgt[:25], rng, 6.0),
                  dict(BETA=300 * 0.8 ** 2, max_iteration=5, LAMBDA=0.1)),
        "nomatch": (worm_xyz[:40], tgt[:33], rng.random((33, 40)) * 0.45,
                    dict(BETA=200, max_iteration=6, LAMBDA=0.1)),
    }.items():
        P, TX, C = track.pr_gls_quick(X.copy(), Y.copy(), cr.copy(), **kw)
        cases[name] = dict(X=X, Y=Y, corr=cr, P=P, T_X=TX, C=C, **{k: np.float64(v) for k, v in kw.items()})
    np.savez_compressed(os.path.join(GOLD, "pr_gls_quick.npz"),
                        **{f"{c}__{k}": v for c, d in cases.items() for k, v in d.items()})

    # ---------------------------------------------------------------- Tracker._predict_one_rep
    ns = _exec_functions("CellTracker/tracker.py", ["_predict_one_rep"], {"np": np})

    class _Self:
        cell_num_t0 = 0

    tracked = worm_xyz[:150] + rng.normal(0, 1.0, (150, 3))
    _Self.cell_num_t0 = tracked.shape[0]
    c0 = cases["worm3_single"]
    post, pre = ns["_predict_one_rep"](_Self, tracked.copy(), c0["X"], 300.0, c0["C"])
    np.savez_compressed(os.path.join(GOLD, "predict_one_rep.npz"), pre=tracked, inter=c0["X"], beta=300.0,
                        C=c0["C"], post=post)

    # ---------------------------------------------------------------- trackerlite EM (trackerlite.py:242-417)
    ffn_ns = _exec_functions("CellTracker/ffn.py", ["normalize_points", "initial_matching_ffn"],
                             {"np": np, "ndarray": np.ndarray, "Union": __import__("typing").Union,
                              "Tuple": __import__("typing").Tuple,
                              "PCA": __import__("sklearn.decomposition", fromlist=["PCA"]).PCA,
                              "NearestNeighbors": __import__("sklearn.neighbors",
                                                             fromlist=["NearestNeighbors"]).NearestNeighbors})
    normalize_points = ffn_ns["normalize_points"]
    ref_norm, (mean, scale) = normalize_points(pts, return_para=True)
    tgt_norm = (synth_target(pts, rng) - mean) / scale
    conf_norm = ref_norm[:170] + rng.normal(0, 0.002, (170, 3))
    corr_l = synth_corr(ref_norm, tgt_norm, rng, 0.02).astype(np.float32)
    prior, pairs = lite.simple_match(corr_l)
    out = dict(points=pts, ref_norm=ref_norm, mean=mean, scale=scale, tgt_norm=tgt_norm, conf_norm=conf_norm,
               corr=corr_l, prior=prior, pairs=pairs)
    for tag, (beta, lam) in {"b3l3": (3.0, 3.0), "b1l01": (1.0, 0.1)}.items():
        pred, post_ = lite.prgls_with_two_ref(prior, tgt_norm, ref_norm, conf_norm, beta=beta, lambda_=lam)
        out[f"{tag}__pred"] = pred
        out[f"{tag}__post"] = post_
        out[f"{tag}__beta"] = beta
        out[f"{tag}__lambda"] = lam
    pq, postq = lite.prgls_quick(prior, tgt_norm, ref_norm, beta=3.0, lambda_=3.0)
    out["quick__pred"], out["quick__post"] = pq, postq
    # single E / M step
    ep = lite.estimate_posterior(prior, 0.01, ref_norm, tgt_norm, 0.05)
    out["estep__post"] = ep
    out["mstep__C"] = lite.solve_movements_ref(0.01, 3.0, ep, ref_norm, tgt_norm,
                                               lite.gaussian_kernel(ref_norm, ref_norm, 9.0))
    # float64 corr variant + threshold variants for simple_match
    corr64 = synth_corr(ref_norm[:60], tgt_norm[:50], rng, 0.02)
    p64, pr64 = lite.simple_match(corr64, threshold=0.3)
    out["sm64__corr"], out["sm64__prior"], out["sm64__pairs"] = corr64, p64, pr64
    np.savez_compressed(os.path.join(GOLD, "trackerlite_em.npz"), **out)

    # ---------------------------------------------------------------- FFN feature/grid builder
    rec = RecordingModel()
    ref_s, tgt_s = ref_norm[:40], tgt_norm[:35]
    corr_rec = ffn_ns["initial_matching_ffn"](rec, ref_s, tgt_s, 20)
    rec2 = RecordingModel()
    corr_rec2 = track.initial_matching_quick(rec2, pts[:33], tgt[:45], 20)
    rec3 = RecordingModel()
    ffn_ns["initial_matching_ffn"](rec3, ref_norm, tgt_norm, 20)
    n_full = ref_norm.shape[0]
    np.savez_compressed(os.path.join(GOLD, "ffn_features.npz"),
                        ref_s=ref_s, tgt_s=tgt_s, grid_s=rec.calls[0], corr_s=corr_rec,
                        ref_q=pts[:33], tgt_q=tgt[:45], grid_q=rec2.calls[0], corr_q=corr_rec2,
                        ref_full=ref_norm, tgt_full=tgt_norm,
                        feat_ref_full=rec3.calls[0][:n_full, :61],
                        feat_tgt_full=rec3.calls[0][::n_full, 61:])

    # ---------------------------------------------------------------- ensemble scheduling
    sched = {}
    for ens in (0, 5, 10, 20):
        for vol in list(range(2, 70)) + [101, 200, 256]:
            for adj in (False, True):
                sched[f"ref_{ens}_{vol}_{int(adj)}"] = np.array(track.get_reference_vols(ens, vol, adjacent=adj))
    for cur in list(range(2, 70)) + [101, 256]:
        for samp in (5, 20):
            for adj in (False, True):
                for start in (1, 3):
                    if cur > start:
                        sched[f"lite_{cur}_{samp}_{int(adj)}_{start}"] = np.array(
                            lite.get_volumes_list(cur, [4, 7], samp, adj, start))
    np.savez_compressed(os.path.join(GOLD, "schedules.npz"), **sched)

    # ---------------------------------------------------------------- unet3_prediction tiling
    uns = _exec_functions("CellTracker/unet3d.py", ["unet3_prediction", "_get_sizes_padded_im"],
                          {"np": np, "itertools": itertools, "math": math})

    class DuckUNet:
        def __init__(self, shape):
            self.input_shape = (None,) + shape + (1,)
            self.output_shape = (None,) + shape + (1,)
            g = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
            self.ramp = (0.001 * g[0] + 0.01 * g[1] + 0.1 * g[2]).astype(np.float32)

        def predict(self, x):
            # position-dependent, input-dependent: pins both the read window and the crop/scatter
            return (x * 0.5 + self.ramp[None, ..., None]).astype(np.float32)

    tiles = {}
    for tag, (shape, tin, shrink) in {
        "a": ((30, 25, 11), (16, 16, 8), (2, 2, 1)),
        "b": ((20, 9, 5), (16, 12, 6), (3, 2, 1)),      # pad > size-1 along some axes (multi reflection)
        "c": ((24, 24, 8), (16, 16, 8), (4, 4, 2)),     # exact multiple of the centre
    }.items():
        img = rng.random((1,) + shape + (1,)).astype(np.float32)
        outp = uns["unet3_prediction"](img, DuckUNet(tin), shrink)
        tiles[f"{tag}__img"] = img
        tiles[f"{tag}__out"] = outp
        tiles[f"{tag}__tin"] = np.array(tin)
        tiles[f"{tag}__shrink"] = np.array(shrink)
    np.savez_compressed(os.path.join(GOLD, "unet_tiling.npz"), **tiles)
    adversarial()
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
