"""GPU parity: accurate correction, tracked label image and the volume-1 preparation (through the C ABI) against goldens
produced by running the UNMODIFIED reference methods (oracle/make_golden_correction.py).

Bars: integer work (interpolated labels, integer displacements, tracked label image) bit-exact; real displacements
1e-9 (the centre-of-mass sums are fp64 but added in a different order than scipy's)."""
import importlib

import numpy as np
import pytest
import torch

from conftest import golden, load_pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m():
    load_pkg()
    return {n: importlib.import_module("3deecelltracker_b200." + n) for n in ("correction", "tracker", "watershed")}


def case(tag):
    g = golden("accurate_correction.npz")
    return {k.split("__", 1)[1]: g[k] for k in g.files if k.startswith(tag + "__")}


@pytest.mark.parametrize("tag", ["zs1", "zs3"])
def test_accurate_correction_and_tracked_labels(m, tag):
    q = case(tag)
    corr = m["correction"]
    regions = corr.CellRegions(q["seg_interp"], int(q["z_scaling"]))
    assert np.array_equal(regions.region_xyz_min, q["region_min"]) and np.array_equal(regions.region_width, q["region_width"])
    prob = torch.from_numpy(np.ascontiguousarray(q["prob"], np.float32)).cuda()
    raw = torch.from_numpy(q["raw"].view(np.int16)).cuda().view(torch.uint16)
    L = regions.n_cells
    t0, zero = q["r_tracked_t0"], np.zeros((L, 3))
    for max_rep, want_r, want_i in ((20, q["r_disp"], q["i_disp"]), (1, q["once_r_disp"], q["once_i_disp"])):
        r_disp, i_disp, reps = corr.accurate_correction_device(regions, prob, raw, float(q["z_xy_ratio"]), t0, zero, t0,
                                                               q["r_pred"], q["on_boundary"], max_rep)
        np.testing.assert_allclose(r_disp.cpu().numpy(), want_r, rtol=0, atol=1e-9)
        assert np.array_equal(i_disp.cpu().numpy(), want_i)
        assert int(reps[0]) >= 1
    lab = corr.tracked_labels_device(regions, q["i_disp"], q["on_boundary"], prob.shape)
    assert np.array_equal(lab.cpu().numpy(), q["tracked_labels"])
    # repeated calls give identical results (atomics only count integers)
    again = corr.tracked_labels_device(regions, q["i_disp"], q["on_boundary"], prob.shape)
    assert torch.equal(lab, again)


@pytest.mark.parametrize("tag", ["zs1", "zs3"])
def test_tracker_interpolate_seg_matches_reference(m, tag):
    """Tracker.load_manual_seg -> interpolate_seg -> cal_subregions (tracker.py:934-950, 1044-1110) on the GPU helpers."""
    q = case(tag)
    x, y, z = q["seg_vol1"].shape
    t = m["tracker"].Tracker(volume_num=2, siz_xyz=(x, y, z), z_xy_ratio=float(q["z_xy_ratio"]), z_scaling=int(q["z_scaling"]),
                             noise_level=20, min_size=10, beta_tk=300, lambda_tk=0.1, maxiter_tk=20)
    t.load_manual_seg(q["seg_vol1"])
    t.interpolate_seg()
    assert np.array_equal(t.seg_cells_interpolated_corrected, q["seg_interp"])
    assert np.array_equal(t.segmentation_manual_relabels, q["relabels"])
    np.testing.assert_allclose(t.r_coordinates_tracked_t0, q["r_tracked_t0"], rtol=0, atol=1e-12)
    t.cal_subregions()
    assert np.array_equal(t.region_xyz_min, q["region_min"]) and np.array_equal(t.region_width, q["region_width"])


def test_label_components_and_recalculate_boundaries(m):
    corr = m["correction"]
    rng = np.random.default_rng(0)
    img = np.zeros((30, 28, 6), np.int64)
    img[2:9, 3:9, 1:4] = 4
    img[12:20, 3:9, 0:3] = 4          # same value, separate region -> two components
    img[9:12, 9:15, 2:5] = 2          # touches the first block diagonally but has another value
    img[22:28, 18:26, 1:6] = 7
    lab, n = corr.label_components(img)
    assert n == 4
    # raster order of the first voxel: (2,3,1) < (9,9,2) < (12,3,0) < (22,18,1)
    assert lab[2, 3, 1] == 1 and lab[9, 9, 2] == 2 and lab[12, 3, 0] == 3 and lab[22, 18, 1] == 4
    assert np.array_equal(lab > 0, img > 0)
    # recalculate_cell_boundaries: two cells whose overlap band is split by the distance watershed
    seg = np.zeros((24, 20, 3), np.int64)
    ov = np.zeros_like(seg)
    seg[3:11, 4:16, :] = 1
    seg[13:21, 4:16, :] = 2
    ov[seg > 0] = 1
    ov[11:13, 4:16, :] = 2            # overlap band between them
    out = corr.recalculate_cell_boundaries(seg.copy(), ov)
    assert np.array_equal(out[3:11], np.where(seg[3:11] > 0, 1, 0)) and np.array_equal(out[13:21], np.where(seg[13:21] > 0, 2, 0))
    assert set(np.unique(out[11:13, 4:16])) <= {1, 2} and (out[11, 4:16] == 1).all() and (out[12, 4:16] == 2).all()
    assert rng is not None
