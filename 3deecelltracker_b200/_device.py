"""Device-side plumbing: PyTorch owns device memory and streams, nothing else (no torch math on the hot path)."""
import numpy as np
import torch

from . import _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("3deecelltracker_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def to_device(a, dtype, pinned_stage=None):
    """ndarray / tensor -> contiguous CUDA tensor of `dtype` (torch dtype)."""
    dev = require_cuda()
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=dtype).contiguous()
    arr = np.ascontiguousarray(a)
    if arr.dtype == np.uint16:
        # torch has limited uint16 support on older builds: ship the bytes
        if dtype != torch.uint16:
            # convert on the host: a device-side cast of the int16 view would turn values >= 32768 negative
            return torch.from_numpy(arr.astype(np.int64)).to(device=dev, dtype=dtype)
        t = torch.from_numpy(arr.view(np.int16)).to(dev, non_blocking=False)
        return t.view(torch.uint16)
    return torch.from_numpy(arr).to(device=dev, dtype=dtype)


def ptr(t):
    return t.data_ptr() if t is not None else None


class Workspace:
    """Grow-only cache of device scratch buffers, keyed by name AND by the stream they are used on (caller-owned
    workspaces of the C ABI): work enqueued on different streams may run concurrently and must not share scratch."""

    def __init__(self):
        self._bufs = {}

    def get(self, name, nbytes):
        dev = require_cuda()
        key = (name, dev.index, torch.cuda.current_stream().cuda_stream)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=dev)
            self._bufs[key] = buf
        return buf

    def clear(self):
        self._bufs.clear()


WORKSPACE = Workspace()


def aligned_ptr(buf):
    p = buf.data_ptr()
    return (p + 255) // 256 * 256
