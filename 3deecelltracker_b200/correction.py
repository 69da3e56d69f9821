"""Accurate correction of tracked positions and the tracked label image on the GPU -- drop-ins for
Tracker._accurate_correction / _correction_once_interp / _transform_cells_quick / _transform_motion_to_image
(tracker.py:1177-1191, 1310-1399), recalculate_cell_boundaries (watershed.py:111-151), the label interpolation of
volume 1 (track.py:322-363 `gaussian_filter`, tracker.py:1044-1110) and skimage.measure.label on a label image.

The per-volume loop (up to 20 repetitions of stamp -> overlap -> weighted centre of mass -> update) runs entirely on the
device (ct_accurate_correction); volume-1 preparation is host NumPy / SciPy plus two GPU helpers and runs once.
"""
import ctypes as C

import numpy as np
import torch
from scipy import ndimage as ndi

from . import _lib
from ._device import WORKSPACE, aligned_ptr, require_cuda, stream_ptr

REP_NUM_CORRECTION = 20    # tracker.py:46
_RAW_DTYPES = {torch.uint16: 0, torch.float32: 1, torch.uint8: 2}


def _ws(name, nbytes):
    buf = WORKSPACE.get(name, nbytes)
    wp = aligned_ptr(buf)
    return wp, buf.numel() - (wp - buf.data_ptr())


def label_components(image):
    """skimage.measure.label(label image, connectivity=3): components of equal non-zero value, raster numbering.
    image: (x,y,z) integer ndarray -> (labels int32 ndarray, n)."""
    dev = require_cuda()
    img = torch.from_numpy(np.ascontiguousarray(image, dtype=np.int32)).to(dev)
    x, y, z = (int(s) for s in img.shape)
    out = torch.empty_like(img)
    n = torch.zeros(1, dtype=torch.int32, device=dev)
    lib = _lib.lib()
    wp, wn = _ws("label_components", lib.ct_label_components_workspace_bytes(x, y, z))
    _lib.check(lib.ct_label_components(img.data_ptr(), x, y, z, out.data_ptr(), n.data_ptr(), wp, wn, stream_ptr()))
    return out.cpu().numpy(), int(n.item())


def recalculate_cell_boundaries(segmentation_xyz, cell_overlaps_mask, sampling_xy=(1, 1), print_message=False):
    """watershed.py:111-151 (all z slices in one call).  Like the reference, the voxels of `segmentation_xyz` that lie
    in an overlap are zeroed in place."""
    if tuple(sampling_xy) != (1, 1):
        raise ValueError("only sampling_xy = (1, 1) is supported (the reference never passes anything else)")
    dev = require_cuda()
    seg = torch.from_numpy(np.ascontiguousarray(segmentation_xyz, dtype=np.int32)).to(dev)
    ov = torch.from_numpy(np.ascontiguousarray(cell_overlaps_mask, dtype=np.int32)).to(dev)
    x, y, z = (int(s) for s in seg.shape)
    out = torch.empty_like(seg)
    lib = _lib.lib()
    wp, wn = _ws("recalc_boundaries", lib.ct_recalculate_cell_boundaries_workspace_bytes(x, y, z))
    _lib.check(lib.ct_recalculate_cell_boundaries(seg.data_ptr(), ov.data_ptr(), x, y, z, out.data_ptr(), wp, wn, stream_ptr()))
    segmentation_xyz[np.asarray(cell_overlaps_mask) > 1] = 0
    return out.cpu().numpy().astype(np.int64)


def interpolate_labels(label_image, z_scaling=10, smooth_sigma=5):
    """Smoothed / z-interpolated label image of volume 1 (semantics of track.py:322-363): every cell is repeated
    z_scaling times along z, blurred with a Gaussian inside its own padded box and re-thresholded so that it keeps its
    voxel count.  Returns (labels summed where cells overlap, number of cells covering each voxel), both padded by 5 on
    every side like the reference's output."""
    rep = np.repeat(np.asarray(label_image), z_scaling, axis=2)
    out = np.zeros(tuple(s + 10 for s in rep.shape), dtype=np.int64)
    cover = np.zeros_like(out)
    for lab, box in enumerate(ndi.find_objects(rep), start=1):
        if box is None:
            raise ValueError(f"label {lab} is missing from the label image (labels must be 1..n)")
        inside = rep[box] == lab
        padded = np.zeros(tuple(s + 10 for s in inside.shape))
        padded[5:-5, 5:-5, 5:-5][inside] = 0.5
        keep_fraction = 1 - np.divide(int(inside.sum()), padded.size, dtype="float")
        smooth = ndi.gaussian_filter(padded, smooth_sigma, mode="constant")
        region = smooth > np.percentile(smooth, keep_fraction * 100)
        dst = tuple(slice(b.start, b.stop + 10) for b in box)
        out[dst] += region * lab
        cover[dst] += region
    return out, cover


class CellRegions:
    """Cells of volume 1 on the interpolated grid as voxel lists (cal_subregions, tracker.py:1093-1110 /
    get_subregions, track.py:501-533), resident on the device."""

    def __init__(self, seg_interpolated, z_scaling):
        seg = np.asarray(seg_interpolated)
        self.shape = tuple(int(s) for s in seg.shape)
        self.z_scaling = int(z_scaling)
        self.n_cells = int(seg.max())
        if self.n_cells < 1:
            raise ValueError("no cells in the interpolated segmentation")
        idx = np.argwhere(seg > 0)
        labs = seg[seg > 0]
        order = np.argsort(labs, kind="stable")
        idx, labs = idx[order], labs[order]
        start = np.searchsorted(labs, np.arange(1, self.n_cells + 2)).astype(np.int32)
        if np.any(np.diff(start) == 0):
            raise ValueError("labels of the interpolated segmentation must be 1..n without gaps")
        vox4 = np.zeros((len(idx), 4), dtype=np.int16)
        vox4[:, :3] = idx
        lo = np.array([idx[start[k]:start[k + 1]].min(axis=0) for k in range(self.n_cells)], dtype=np.int32)
        hi = np.array([idx[start[k]:start[k + 1]].max(axis=0) for k in range(self.n_cells)], dtype=np.int32)
        self.region_xyz_min, self.region_width = lo, (hi + 1 - lo).astype(np.int32)
        self.pad = np.ascontiguousarray(self.region_width.max(axis=0), dtype=np.int32)        # pad_x, pad_y, pad_z
        dev = require_cuda()
        self.n_vox = int(len(idx))
        self.vox4 = torch.from_numpy(vox4).to(dev)
        self.start = torch.from_numpy(start).to(dev)
        self.rmin = torch.from_numpy(lo).to(dev)
        self.rwidth = torch.from_numpy(self.region_width).to(dev)

    def _cell_args(self):
        return (self.vox4.data_ptr(), self.start.data_ptr(), self.rmin.data_ptr(), self.rwidth.data_ptr(), self.n_cells,
                self.n_vox, self.pad.ctypes.data, self.shape[0], self.shape[1], self.shape[2], self.z_scaling)


def _f64(a, dev):
    return a.to(device=dev, dtype=torch.float64).contiguous() if isinstance(a, torch.Tensor) else \
        torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)


def _i32(a, dev):
    return a.to(device=dev, dtype=torch.int32).contiguous() if isinstance(a, torch.Tensor) else \
        torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)


def accurate_correction_device(regions, prob_dev, raw_dev, z_xy_ratio, r_tracked_t0, r_disp_prev, r_tracked_prev, r_pred,
                               cells_on_boundary, max_rep=REP_NUM_CORRECTION):
    """tracker.py:1177-1191 on the device.  prob_dev (x,y,z) float32, raw_dev (x,y,z) uint16 / float32 / uint8 CUDA
    tensors; the (L,3) coordinate arrays and the (L,) boundary flags may be ndarrays or CUDA tensors.
    Returns (r_disp (L,3) float64, i_disp (L,3) int32, [repetitions, converged] int32) as CUDA tensors, no sync."""
    dev = require_cuda()
    lib = _lib.lib()
    x, y, z = (int(s) for s in prob_dev.shape)
    if (x, y, z * regions.z_scaling) != regions.shape:
        raise ValueError(f"probability map {(x, y, z)} does not match the interpolated cells {regions.shape}")
    if raw_dev.dtype not in _RAW_DTYPES:
        raw_dev = raw_dev.to(torch.float32)
    prob_dev, raw_dev = prob_dev.contiguous(), raw_dev.contiguous()
    t0, dp, tp, rp = (_f64(a, dev) for a in (r_tracked_t0, r_disp_prev, r_tracked_prev, r_pred))
    onb = _i32(cells_on_boundary, dev)
    L = regions.n_cells
    r_disp = torch.empty((L, 3), dtype=torch.float64, device=dev)
    i_disp = torch.empty((L, 3), dtype=torch.int32, device=dev)
    reps = torch.zeros(2, dtype=torch.int32, device=dev)
    wp, wn = _ws("correction", lib.ct_correction_workspace_bytes(x, y, z, L, regions.n_vox))
    _lib.check(lib.ct_accurate_correction(*regions._cell_args(), prob_dev.data_ptr(), raw_dev.data_ptr(),
                                          _RAW_DTYPES[raw_dev.dtype], x, y, z, float(z_xy_ratio), t0.data_ptr(),
                                          dp.data_ptr(), tp.data_ptr(), rp.data_ptr(), onb.data_ptr(), int(max_rep),
                                          r_disp.data_ptr(), i_disp.data_ptr(), reps.data_ptr(), wp, wn, stream_ptr()))
    return r_disp, i_disp, reps


def tracked_labels_device(regions, i_disp, cells_on_boundary, shape_xyz):
    """tracker.py:1391-1399 on the device -> (x,y,z) int32 CUDA tensor."""
    dev = require_cuda()
    lib = _lib.lib()
    x, y, z = (int(s) for s in shape_xyz)
    idp, onb = _i32(i_disp, dev), _i32(cells_on_boundary, dev)
    out = torch.empty((x, y, z), dtype=torch.int32, device=dev)
    wp, wn = _ws("tracked_labels", lib.ct_tracked_labels_workspace_bytes(x, y, z, regions.n_cells, regions.n_vox))
    _lib.check(lib.ct_tracked_labels(*regions._cell_args(), idp.data_ptr(), onb.data_ptr(), x, y, z, out.data_ptr(), wp, wn,
                                     stream_ptr()))
    return out
