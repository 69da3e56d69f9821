"""Watershed + centroid stage on the GPU -- drop-in for CellTracker/watershed.py:16-108 as used by
Tracker._watershed (tracker.py:671-684) and the centre-of-mass lines of Tracker._segment (tracker.py:646-648).

The reference runs this stage on the host between the U-Net and the matcher (scipy EDT + Gaussian, scikit-image
peak_local_max / label / watershed / find_boundaries, one Python iteration per z slice).  Here the probability map
stays in HBM: `segment_device` takes the CUDA tensor the U-Net produced and returns the label image and the cell
centres as CUDA tensors; `segment` is the same call with host arrays in and out.  Label images and centres are
bit-identical to the CPU path (tests/test_gpu_watershed.py).
"""
import numpy as np
import torch

from . import _lib
from ._device import WORKSPACE, aligned_ptr, require_cuda, stream_ptr

MAX_CELLS = 16384
_METHODS = {"min_size": 0, "cell_num": 1}


def _gaussian_weights(sigma, truncate=4.0):
    """One-sided weights of scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, int(truncate * sigma + 0.5))."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[radius:], dtype=np.float64)


_W_XY = _gaussian_weights(2.0)          # watershed.py:40,90: sigma 2 in the plane
_W_Z = _gaussian_weights(0.3)           # watershed.py:90: sigma 0.3 along z


class Segmentation:
    """Device-resident result of one call: `labels` (x,y,z) int32; `centres` (MAX_CELLS,3) float64 in voxel units and
    `centres_real` (z scaled by z_xy_ratio), of which the first n_cells rows are valid; `scalars` int32[4] = n_cells,
    min_size, cell_num, background voxels."""

    def __init__(self, labels, centres2, scalars):
        self.labels, self.centres, self.centres_real, self.scalars = labels, centres2[0], centres2[1], scalars
        self._host = None
        self.ready = None                # CUDA event set by callers that produce the result on a side stream

    def host_scalars(self):
        """(n_cells, min_size, cell_num): synchronises the current stream once."""
        if self._host is None:
            if self.ready is not None:
                torch.cuda.current_stream().wait_event(self.ready)
            self._host = [int(v) for v in self.scalars.cpu().tolist()]
        return self._host[0], self._host[1], self._host[2]

    def centres_host(self):
        n = self.host_scalars()[0]
        if n > MAX_CELLS:
            raise ValueError(f"watershed found {n} cells; at most {MAX_CELLS} are supported")
        return self.centres[:n].cpu().numpy()


def segment_device(prob_dev, z_xy_ratio, method="min_size", min_size=0, cell_num=0):
    """prob_dev: (x,y,z) float32 CUDA tensor -> Segmentation (all outputs on the device, no synchronisation)."""
    if method not in _METHODS:
        raise ValueError("The method parameter should be either min_size or cell_num")      # watershed.py:99-100
    if prob_dev.dim() != 3:
        raise ValueError(f"expected a 3D probability map (x, y, z), got {prob_dev.dim()}D")
    dev = require_cuda()
    prob_dev = prob_dev.to(device=dev, dtype=torch.float32).contiguous()
    lib = _lib.lib()
    x, y, z = (int(s) for s in prob_dev.shape)
    labels = torch.empty((x, y, z), dtype=torch.int32, device=dev)
    centres = torch.empty((2, MAX_CELLS, 3), dtype=torch.float64, device=dev)
    scalars = torch.empty(4, dtype=torch.int32, device=dev)
    ws = WORKSPACE.get("watershed", lib.ct_watershed_workspace_bytes(x, y, z, MAX_CELLS))
    wp = aligned_ptr(ws)
    _lib.check(lib.ct_watershed_segment(prob_dev.data_ptr(), x, y, z, float(z_xy_ratio), _METHODS[method],
                                        int(min_size or 0), int(cell_num or 0), _W_XY.ctypes.data, _W_Z.ctypes.data,
                                        labels.data_ptr(), centres.data_ptr(), MAX_CELLS, scalars.data_ptr(), wp,
                                        ws.numel() - (wp - ws.data_ptr()), stream_ptr()))
    return Segmentation(labels, centres, scalars)


def segment(image_cell_bg_xyz, z_xy_ratio, method="min_size", min_size=0, cell_num=0):
    """Host form: probability map ndarray (x,y,z) -> (segmentation_auto int32 (x,y,z), centres (n,3) float64 in voxel
    units, min_size, cell_num) -- what Tracker._watershed + center_of_mass return (tracker.py:646-648,671-684)."""
    dev = require_cuda()
    prob = torch.from_numpy(np.ascontiguousarray(image_cell_bg_xyz, dtype=np.float32)).to(dev)
    seg = segment_device(prob, z_xy_ratio, method, min_size, cell_num)
    n, min_size, cell_num = seg.host_scalars()
    return seg.labels.cpu().numpy(), seg.centres_host(), min_size, cell_num


def segment_centres(labels_xyz):
    """ndimage.center_of_mass(labels > 0, labels, range(1, max + 1)) (tracker.py:1063-1065) for a host label image."""
    from scipy import ndimage as ndi
    lab = np.asarray(labels_xyz)
    n = int(lab.max())
    return np.asarray(ndi.center_of_mass(lab > 0, lab, range(1, n + 1)), dtype=np.float64).reshape(n, 3)
