"""CPU: the accurate-correction passes (csrc/correction_core.cuh), instantiated sequentially by tests/emul/ws_host.cpp,
and the host-side volume-1 preparation (correction.interpolate_labels) against goldens produced by running the
UNMODIFIED reference methods (oracle/make_golden_correction.py: Tracker._accurate_correction,
_correction_once_interp, _transform_cells_quick, _transform_motion_to_image, interpolate_seg, track.gaussian_filter)."""
import ctypes as C
import importlib
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden, load_pkg


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("corr_emul") / "libws_emul.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out,
                    os.path.join(ROOT, "tests", "emul", "ws_host.cpp")], check=True)
    return C.CDLL(out)


def case(tag):
    g = golden("accurate_correction.npz")
    return {k.split("__", 1)[1]: g[k] for k in g.files if k.startswith(tag + "__")}


def cells_of(seg_interp):
    """Voxel lists per label (what correction.CellRegions builds, without the device upload)."""
    L = int(seg_interp.max())
    idx = np.argwhere(seg_interp > 0)
    labs = seg_interp[seg_interp > 0]
    order = np.argsort(labs, kind="stable")
    idx, labs = idx[order], labs[order]
    start = np.searchsorted(labs, np.arange(1, L + 2)).astype(np.int32)
    vox4 = np.zeros((len(idx), 4), np.int16)
    vox4[:, :3] = idx
    lo = np.array([idx[start[k]:start[k + 1]].min(0) for k in range(L)], np.int32)
    hi = np.array([idx[start[k]:start[k + 1]].max(0) for k in range(L)], np.int32)
    return vox4, start, lo, (hi + 1 - lo).astype(np.int32)


vp = lambda a: a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("tag", ["zs1", "zs3"])
def test_correction_passes_match_reference(emul, tag):
    q = case(tag)
    seg_i, zs, ratio = q["seg_interp"], int(q["z_scaling"]), float(q["z_xy_ratio"])
    vox4, start, rmin, rw = cells_of(seg_i)
    assert np.array_equal(rmin, q["region_min"]) and np.array_equal(rw, q["region_width"])     # get_subregions
    L, pad = len(rmin), np.ascontiguousarray(rw.max(0), np.int32)
    prob, raw = np.ascontiguousarray(q["prob"], np.float32), np.ascontiguousarray(q["raw"])
    x, y, z = prob.shape
    t0, zero = np.ascontiguousarray(q["r_tracked_t0"]), np.zeros((L, 3))
    onb = np.ascontiguousarray(q["on_boundary"], np.int32)
    r_pred = np.ascontiguousarray(q["r_pred"])
    for max_rep, want_r, want_i in ((20, q["r_disp"], q["i_disp"]), (1, q["once_r_disp"], q["once_i_disp"])):
        r_disp, i_disp, reps = np.zeros((L, 3)), np.zeros((L, 3), np.int32), np.zeros(2, np.int32)
        emul.corr_emul_accurate_correction(vp(vox4), vp(start), vp(rmin), vp(rw), L, len(vox4), vp(pad), *seg_i.shape, zs,
                                           vp(prob), vp(raw), 0, x, y, z, C.c_double(ratio), vp(t0), vp(zero), vp(t0),
                                           vp(r_pred), vp(onb), max_rep, vp(r_disp), vp(i_disp), vp(reps))
        np.testing.assert_allclose(r_disp, want_r, rtol=0, atol=1e-10)
        assert np.array_equal(i_disp, want_i)
        assert reps[0] >= 1 and (max_rep == 1 or reps[1] == 1)
    lab = np.zeros(prob.shape, np.int32)
    emul.corr_emul_tracked_labels(vp(vox4), vp(start), vp(rmin), vp(rw), L, len(vox4), vp(pad), *seg_i.shape, zs,
                                  vp(np.ascontiguousarray(q["i_disp"], np.int32)), vp(onb), x, y, z, vp(lab))
    assert np.array_equal(lab, q["tracked_labels"])


@pytest.mark.parametrize("tag", ["zs1", "zs3"])
def test_interpolate_labels_matches_reference(tag):
    load_pkg()
    corr = importlib.import_module("3deecelltracker_b200.correction")
    q = case(tag)
    out, cover = corr.interpolate_labels(q["seg_vol1"], z_scaling=int(q["z_scaling"]), smooth_sigma=2.5)
    assert np.array_equal(out, q["interp_raw"]) and np.array_equal(cover, q["interp_cover"])
