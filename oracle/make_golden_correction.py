"""tests/golden/accurate_correction.npz: the accurate-correction loop of the UNMODIFIED reference on a small case.

TEST INFRASTRUCTURE ONLY.  Run as ``python -m oracle.make_golden_correction`` where /root/reference is mounted.

Executed verbatim (function source read from the reference at generation time, nothing copied into this repo):
  track.py    gaussian_filter, get_subregions, _get_coordinates                      (imported through oracle/_ref_shim)
  tracker.py  Tracker.interpolate_seg, _interpolate, _relabel_separated_cells, cal_subregions, _transform_cells_quick,
              _correction_once_interp, _evaluate_correction, _accurate_correction, _transform_motion_to_image,
              _transform_disps, _transform_layer_to_real, _transform_real_to_interpolated
  watershed.py recalculate_cell_boundaries
scikit-image is not installable here; the three functions the code above calls from it are supplied by the oracle:
`skimage.filters.gaussian` = scipy.ndimage.gaussian_filter (what scikit-image itself calls for a float image),
`skimage.measure.label` on an integer image and `skimage.segmentation.watershed` = oracle/watershed.py restatements
(PARITY UNPINNED for those two, see oracle/watershed.py).
"""
import os
import sys
import types

import numpy as np
from scipy import ndimage as ndi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import _ref_shim  # noqa: E402
from oracle import watershed as ows  # noqa: E402
from oracle.make_golden import GOLD, _exec_functions  # noqa: E402


def label_int_image(img, connectivity=3):
    """skimage.measure.label(int image, connectivity = ndim): components of EQUAL non-zero value, raster numbering."""
    out = np.zeros(img.shape, dtype=np.int64)
    nxt = 0
    first = []
    for v in np.unique(img):
        if v == 0:
            continue
        lab, n = ndi.label(img == v, structure=np.ones((3,) * img.ndim, dtype=bool))
        for k in range(1, n + 1):
            idx = np.flatnonzero(lab.ravel() == k)
            first.append((idx[0], idx))
    first.sort(key=lambda t: t[0])
    flat = out.ravel()
    for _, idx in first:
        nxt += 1
        flat[idx] = nxt
    return out


def build_namespace():
    track, _ = _ref_shim.load()
    stub = types.ModuleType("skimage")
    stub.filters = types.ModuleType("skimage.filters")
    stub.filters.gaussian = lambda img, sigma, mode="constant": ndi.gaussian_filter(np.asarray(img, dtype=np.float64), sigma, mode=mode)
    sys.modules["skimage"] = stub
    sys.modules["skimage.filters"] = stub.filters
    ns = {"np": np, "ndm": ndi, "label": label_int_image, "gaussian_filter": track.gaussian_filter,
          "get_subregions": track.get_subregions, "distance_transform_edt": ndi.distance_transform_edt,
          "watershed": lambda image, markers, mask: ows.watershed(image, np.array(markers), mask),
          "save_img3ts": lambda *a, **k: None, "REP_NUM_CORRECTION": 20, "ndarray": np.ndarray}
    _exec_functions("CellTracker/watershed.py", ["recalculate_cell_boundaries"], ns)
    _exec_functions("CellTracker/tracker.py",
                    ["interpolate_seg", "_interpolate", "_relabel_separated_cells", "cal_subregions", "_transform_cells_quick",
                     "_correction_once_interp", "_evaluate_correction", "_accurate_correction", "_transform_motion_to_image",
                     "_transform_disps", "_transform_layer_to_real", "_transform_real_to_interpolated"], ns)
    return ns


def make_self(ns, seg_vol1, z_xy_ratio, z_scaling):
    class Self:
        pass
    s = Self()
    for name in ("interpolate_seg", "_interpolate", "cal_subregions", "_transform_cells_quick", "_correction_once_interp",
                 "_evaluate_correction", "_accurate_correction", "_transform_motion_to_image", "_transform_layer_to_real",
                 "_transform_real_to_interpolated"):
        setattr(Self, name, ns[name])
    Self._transform_disps = staticmethod(ns["_transform_disps"])
    Self._relabel_separated_cells = staticmethod(ns["_relabel_separated_cells"])
    s.x_siz, s.y_siz, s.z_siz = seg_vol1.shape
    s.z_xy_ratio, s.z_scaling = z_xy_ratio, z_scaling
    s.segmentation_manual_relabels = seg_vol1
    s.use_8_bit = True

    class P:
        track_results = ""
    s.paths = P()
    return s


def synthetic_case(rng, shape=(48, 44, 6), n_cells=7):
    """Label image of vol 1 (ellipsoids, two of them touching), probability map + raw image of the target volume."""
    X, Y, Z = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    seg = np.zeros(shape, dtype=np.int64)
    centres = []
    k = 0
    while k < n_cells:
        c = rng.uniform([8, 8, 1], [shape[0] - 8, shape[1] - 8, shape[2] - 1])
        if any(np.linalg.norm((c - q) * [1, 1, 3]) < 9 for q in centres):
            continue
        r = rng.uniform(3.5, 5.5)
        m = ((X - c[0]) ** 2 + (Y - c[1]) ** 2 + ((Z - c[2]) * 2.5) ** 2) <= r * r
        if m.sum() < 10:
            continue
        k += 1
        seg[m & (seg == 0)] = k
        centres.append(c)
    return seg, np.array(centres)


def main():
    ns = build_namespace()
    rng = np.random.default_rng(4242)
    out = {}
    for tag, (ratio, zs) in {"zs1": (3.0, 1), "zs3": (3.0, 3)}.items():
        seg, centres = synthetic_case(rng)
        s = make_self(ns, seg, ratio, zs)
        interp_raw, interp_cover = ns["gaussian_filter"](seg, z_scaling=zs, smooth_sigma=2.5)
        s.interpolate_seg()
        s.cal_subregions()
        L = s.cell_num_t0
        # target volume: cells moved by a smooth field; probability / raw images rendered at the moved positions
        move = rng.normal(0, 1.2, (L, 3)) * np.array([1.0, 1.0, 0.3])
        shape = seg.shape
        X, Y, Z = np.meshgrid(*[np.arange(v) for v in shape], indexing="ij")
        prob = np.zeros(shape, dtype=np.float32)
        t0_layer = s.r_coordinates_tracked_t0 / np.array([1, 1, ratio])
        for c in t0_layer + move:
            prob = np.maximum(prob, np.exp(-(((X - c[0]) ** 2 + (Y - c[1]) ** 2) / 18.0 + ((Z - c[2]) ** 2) / 1.5)).astype(np.float32))
        raw = (rng.normal(100, 10, shape) + 2000 * prob).clip(0, 65535).astype(np.uint16)

        class Seg:
            pass
        s.segresult = Seg()
        s.segresult.image_cell_bg = prob[None, ..., None]
        s.segresult.image_gcn = raw.copy() / 65536.0

        class Hist:
            pass
        s.history = Hist()
        s.history.r_displacements = [np.zeros((L, 3))]
        s.history.r_tracked_coordinates = [s.r_coordinates_tracked_t0.copy()]
        # prediction of FFN + PR-GLS = true motion + error (what accurate correction is there to remove)
        r_pred = s.r_coordinates_tracked_t0 + (move + rng.normal(0, 0.8, (L, 3)) * [1, 1, 0.2]) * np.array([1, 1, ratio])
        on_boundary = np.zeros(L, dtype=int)
        on_boundary[rng.integers(0, L)] = 1
        r_disp, i_disp = s._accurate_correction(on_boundary, r_pred.copy())
        one = s._correction_once_interp(s._transform_real_to_interpolated(r_pred - s.r_coordinates_tracked_t0), on_boundary)
        labels = s._transform_motion_to_image(on_boundary, i_disp)
        d = dict(seg_vol1=seg, interp_raw=interp_raw, interp_cover=interp_cover, z_xy_ratio=float(ratio), z_scaling=int(zs), seg_interp=s.seg_cells_interpolated_corrected,
                 relabels=s.segmentation_manual_relabels, r_tracked_t0=s.r_coordinates_tracked_t0, prob=prob, raw=raw,
                 r_pred=r_pred, on_boundary=on_boundary, r_disp=r_disp, i_disp=i_disp, once_r_disp=one[0],
                 once_i_disp=one[1], once_corr=one[2], tracked_labels=labels,
                 region_min=np.array(s.region_xyz_min), region_width=np.array(s.region_width))
        out.update({f"{tag}__{k}": v for k, v in d.items()})
        print(tag, "cells", L, "interp", s.seg_cells_interpolated_corrected.shape, "max |correction|", np.abs(one[2]).max())
    np.savez_compressed(os.path.join(GOLD, "accurate_correction.npz"), **out)
    print(os.path.getsize(os.path.join(GOLD, "accurate_correction.npz")), "bytes")


if __name__ == "__main__":
    main()
