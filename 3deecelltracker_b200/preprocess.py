"""LCN normalisation on the GPU -- drop-in for CellTracker/preprocess.py:136-188.

`_normalize_image(image, noise_level)` keeps the reference signature (preprocess.py:170): it takes the
raw 3D stack (x, y, z) as an ndarray (uint16 / uint8 / float) and returns the normalised stack as a host
ndarray.  `normalize_image_device` is the same operator for callers that keep data in HBM.
"""
import numpy as np
import torch

from . import _lib
from ._device import WORKSPACE, aligned_ptr, require_cuda, stream_ptr, to_device

_DTYPES = {torch.uint16: 0, torch.float32: 1, torch.uint8: 2}     # kernel dtype codes (ct3d.h); anything else -> float32


def _raw_to_device(image):
    if isinstance(image, torch.Tensor):
        t = image
        if t.dtype not in _DTYPES:
            t = t.to(torch.float32)
        return t.to(require_cuda()).contiguous()
    arr = np.asarray(image)
    if arr.dtype == np.uint16:
        return to_device(arr, torch.uint16)
    if arr.dtype == np.uint8:
        return to_device(arr, torch.uint8)
    return to_device(arr.astype(np.float32, copy=False), torch.float32)


def normalize_image_device(raw_dev, noise_level, filter_size=(27, 27, 1), out=None, subtract_median=True):
    """raw_dev: CUDA tensor (x,y,z) uint16 / uint8 / float32 -> float32 CUDA tensor (x,y,z)."""
    if raw_dev.dim() != 3:
        raise ValueError(f"expected a 3D image (x, y, z), got {raw_dev.dim()}D")
    if filter_size[2] != 1:
        raise ValueError("only filters of z-extent 1 are supported (the reference uses (27, 27, 1))")
    lib = _lib.lib()
    x, y, z = (int(s) for s in raw_dev.shape)
    if out is None:
        out = torch.empty((x, y, z), dtype=torch.float32, device=raw_dev.device)
    ws = WORKSPACE.get("lcn", lib.ct_normalize_workspace_bytes(x, y, z))
    wp = aligned_ptr(ws)
    _lib.check(lib.ct_normalize_image(raw_dev.data_ptr(), _DTYPES[raw_dev.dtype], out.data_ptr(), x, y, z,
                                      float(noise_level), int(filter_size[0]), int(filter_size[1]),
                                      1 if subtract_median else 0,
                                      wp, ws.numel() - (wp - ws.data_ptr()), stream_ptr()))
    return out


def normalize_block_device(raw_block_dev, noise_level, median_dev, filter_size=(27, 27, 1), out=None):
    """LCN of a block of a larger volume with the volume's median supplied (1-element float64 CUDA tensor): the
    per-rank step of the spatially decomposed `_normalize_image` (spatial.py).  Output voxels within 2 * (filter // 2) of a
    block face that is not a face of the volume are not meaningful (ct3d.h)."""
    if raw_block_dev.dim() != 3 or filter_size[2] != 1:
        raise ValueError("expected a 3D block and a filter of z-extent 1")
    lib = _lib.lib()
    raw_block_dev = raw_block_dev.contiguous()
    x, y, z = (int(s) for s in raw_block_dev.shape)
    if out is None:
        out = torch.empty((x, y, z), dtype=torch.float32, device=raw_block_dev.device)
    ws = WORKSPACE.get("lcn", lib.ct_normalize_workspace_bytes(x, y, z))
    wp = aligned_ptr(ws)
    _lib.check(lib.ct_normalize_image_with_median(raw_block_dev.data_ptr(), _DTYPES[raw_block_dev.dtype], out.data_ptr(),
                                                  x, y, z, float(noise_level), int(filter_size[0]), int(filter_size[1]),
                                                  median_dev.data_ptr(), wp, ws.numel() - (wp - ws.data_ptr()),
                                                  stream_ptr()))
    return out


def median_device(raw_dev):
    """np.median of a CUDA tensor (uint16 / uint8 / float32), returned as a 1-element float64 CUDA tensor."""
    lib = _lib.lib()
    out = torch.empty(1, dtype=torch.float64, device=raw_dev.device)
    ws = WORKSPACE.get("median", 8192)
    wp = aligned_ptr(ws)
    _lib.check(lib.ct_median(raw_dev.data_ptr(), _DTYPES[raw_dev.dtype], raw_dev.numel(), out.data_ptr(), wp,
                             ws.numel() - (wp - ws.data_ptr()), stream_ptr()))
    return out


def lcn_gpu(img3d, noise_level=5, filter_size=(27, 27, 1)):
    """preprocess.py:136-167: local contrast normalisation of an already median-subtracted image."""
    dev = _raw_to_device(img3d)
    return normalize_image_device(dev, noise_level, filter_size, subtract_median=False).cpu().numpy()


def _normalize_image(image, noise_level):
    """preprocess.py:170-188: median subtract, clamp, local contrast normalisation (27, 27, 1)."""
    dev = _raw_to_device(image)
    return normalize_image_device(dev, noise_level, (27, 27, 1)).cpu().numpy()
