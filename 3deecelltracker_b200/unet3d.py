"""3D U-Net forward and tiled prediction -- drop-in for the inference half of CellTracker/unet3d.py.

`unet3_a() / unet3_b() / unet3_c()` (unet3d.py:26-81) return a model object exposing what
`unet3_prediction` and the Tracker use from a Keras model: `.input_shape`, `.output_shape`, `.predict`,
`.get_weights()` / `.set_weights()` in Keras order, `.save_weights()` / `.load_weights()` (npz container;
Keras .h5 import needs h5py, which this image lacks -- see INTEGRATION.md).
`unet3_prediction(img, model, shrink)` keeps the reference signature (unet3d.py:203).
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._device import WORKSPACE, aligned_ptr, require_cuda, stream_ptr, to_device

_SPECS = {
    "a": dict(input=(160, 160, 16), pool=(2, 2, 1), relu=0, down=[(8, 16), (16, 32), (32, 64)],
              up=[(64, 64), (32, 32), (16, 16)], out=(8, 8)),          # unet3d.py:26-37,84-98
    "b": dict(input=(96, 96, 8), pool=(2, 2, 1), relu=1, down=[(64, 64), (128, 128)],
              up=[(256, 256), (128, 128)], out=(64, 64)),              # unet3d.py:40-67
    "c": dict(input=(64, 64, 64), pool=(2, 2, 2), relu=0, down=[(8, 16), (16, 32), (32, 64)],
              up=[(64, 64), (32, 32), (16, 16)], out=(8, 8)),          # unet3d.py:70-81
}

ENGINES = {"auto": 0, "direct": 1, "tcgen05": 2, "tcgen05_classic": 3, "tcgen05_stacked": 4}
# conv_block_device only: the tensor-core kernels between split-fp16 activation buffers, as they run inside the network
BLOCK_ENGINES = dict(ENGINES, tcgen05_split=5, tcgen05_split_dst=6, tcgen05_split_src=7,
                     # the plane-walk kernel (unet_tcz.cu) wherever it can run: split buffers / split source -> fp32
                     planewalk_split=8, planewalk_split_src=9)


def _conv_layers(spec):
    layers, c, skips = [], 1, []
    for f1, f2 in spec["down"]:
        layers += [(c, f1), (f1, f2)]
        skips.append(f2)
        c = f2
    for (f1, f2), skip in zip(spec["up"], reversed(skips)):
        layers += [(c, f1), (f1, f2)]
        c = f2 + skip
    layers += [(c, spec["out"][0]), (spec["out"][0], spec["out"][1])]
    return layers


def _keras_default_weights(spec, seed=None):
    """Keras defaults for an untrained model: glorot_uniform kernels, zero bias, BN gamma 1 / beta 0 /
    mean 0 / var 1."""
    rng = np.random.default_rng(seed)
    ws = []
    for cin, cout in _conv_layers(spec):
        lim = math.sqrt(6.0 / (27 * cin + 27 * cout))
        ws += [rng.uniform(-lim, lim, (3, 3, 3, cin, cout)).astype(np.float32), np.zeros(cout, np.float32),
               np.ones(cout, np.float32), np.zeros(cout, np.float32), np.zeros(cout, np.float32),
               np.ones(cout, np.float32)]
    c = spec["out"][1]
    lim = math.sqrt(6.0 / (c + 1))
    ws += [rng.uniform(-lim, lim, (1, 1, 1, c, 1)).astype(np.float32), np.zeros(1, np.float32)]
    return ws


class UNet3:
    """GPU-resident 3D U-Net with the slice of the Keras Model API the reference touches."""

    def __init__(self, variant="a", weights=None, tiles_per_batch=8, engine="auto"):
        if variant not in _SPECS:
            raise ValueError(f"unknown U-Net variant {variant!r}")
        self.variant = variant
        self._spec = _SPECS[variant]
        x, y, z = self._spec["input"]
        self.input_shape = (None, x, y, z, 1)
        self.output_shape = (None, x, y, z, 1)
        self.tiles_per_batch = int(tiles_per_batch)
        self._engine = engine
        self._handle = None
        self._weights = None
        self.set_weights(weights if weights is not None else _keras_default_weights(self._spec, seed=0))

    # ---- Keras-like weight API
    def _cspec(self):
        s = _lib.CtUNetSpec()
        sp = self._spec
        s.in_x, s.in_y, s.in_z = sp["input"]
        s.pool_x, s.pool_y, s.pool_z = sp["pool"]
        s.act_relu = sp["relu"]
        s.levels = len(sp["down"])
        for i, (a, b) in enumerate(sp["down"]):
            s.down[i][0], s.down[i][1] = a, b
        for i, (a, b) in enumerate(sp["up"]):
            s.up[i][0], s.up[i][1] = a, b
        s.out[0], s.out[1] = sp["out"]
        return s

    def set_weights(self, weights):
        layers = _conv_layers(self._spec)
        if len(weights) != 6 * len(layers) + 2:
            raise ValueError(f"expected {6 * len(layers) + 2} weight arrays, got {len(weights)}")
        shapes = []
        for cin, cout in layers:
            shapes += [(3, 3, 3, cin, cout)] + [(cout,)] * 5
        shapes += [(1, 1, 1, self._spec["out"][1], 1), (1,)]
        ws = []
        for w, shp in zip(weights, shapes):
            w = np.asarray(w, dtype=np.float32)
            if w.shape != shp:
                raise ValueError(f"weight shape {w.shape} does not match {shp}")
            ws.append(w)
        require_cuda()
        lib = _lib.lib()
        flat = np.ascontiguousarray(np.concatenate([w.reshape(-1) for w in ws]))
        spec = self._cspec()
        handle = C.c_void_p()
        _lib.check(lib.ct_unet_create(C.byref(spec), flat.ctypes.data, flat.size, C.byref(handle)))
        self._release()
        self._handle = handle
        self._weights = ws
        self.set_engine(self._engine)

    def get_weights(self):
        return [w.copy() for w in self._weights]

    def save_weights(self, path):
        np.savez(path, *self._weights)

    def load_weights(self, path):
        """`.npz` (save_weights / io_formats.convert_keras_h5) or a Keras `.h5` model / weights file (tracker.py:579;
        needs h5py)."""
        from .io_formats import load_weight_file
        self.set_weights(load_weight_file(path))

    def set_engine(self, engine):
        self._engine = engine
        _lib.check(_lib.lib().ct_unet_set_engine(self._handle, ENGINES[engine]))

    @property
    def flops_per_tile(self):
        return float(_lib.lib().ct_unet_flops_per_tile(self._handle))

    def _release(self):
        if self._handle is not None:
            _lib.lib().ct_unet_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _workspace(self, tiles_per_batch):
        lib = _lib.lib()
        ws = WORKSPACE.get("unet", lib.ct_unet_workspace_bytes(self._handle, tiles_per_batch))
        wp = aligned_ptr(ws)
        return wp, ws.numel() - (wp - ws.data_ptr())

    # ---- Keras Model.predict
    def predict_device(self, tiles_dev):
        """(B, x, y, z) float32 CUDA tensor -> (B, x, y, z) float32 CUDA tensor."""
        b = int(tiles_dev.shape[0])
        if tuple(tiles_dev.shape[1:]) != tuple(self.input_shape[1:4]):
            raise ValueError(f"expected tiles of shape {self.input_shape[1:4]}, got {tuple(tiles_dev.shape[1:])}")
        out = torch.empty_like(tiles_dev)
        tpb = max(1, min(self.tiles_per_batch, b))
        wp, wn = self._workspace(tpb)
        _lib.check(_lib.lib().ct_unet_predict_tiles(self._handle, tiles_dev.data_ptr(), out.data_ptr(), b, wp, wn,
                                                    tpb, stream_ptr()))
        return out

    def predict(self, x, batch_size=None, verbose=0):
        """(B, x, y, z, 1) ndarray -> (B, x, y, z, 1) float32 ndarray (unet3d.py:253)."""
        arr = np.asarray(x)
        if arr.ndim != 5 or arr.shape[4] != 1:
            raise ValueError(f"expected input of shape (batch, x, y, z, 1), got {arr.shape}")
        dev = to_device(arr[..., 0].astype(np.float32, copy=False), torch.float32)
        return self.predict_device(dev).cpu().numpy()[..., None]

    def conv_block_device(self, layer, x_dev, engine="tcgen05"):
        """One Conv3D + LeakyReLU/ReLU + BatchNorm block (unet3d.py:101-141) on a channels-last CUDA tensor
        (B, x, y, z, Cin) float32 -> (B, x, y, z, Cout) float32.  `layer` indexes the conv blocks in graph order."""
        cin, cout = _conv_layers(self._spec)[layer]
        b, x, y, z, c = (int(v) for v in x_dev.shape)
        if c != cin:
            raise ValueError(f"layer {layer} expects {cin} input channels, got {c}")
        lib = _lib.lib()
        x_dev = x_dev.contiguous()
        out = torch.empty((b, x, y, z, cout), dtype=torch.float32, device=x_dev.device)
        ws = WORKSPACE.get("unet_block", lib.ct_unet_conv_block_workspace_bytes(self._handle, layer, b, x, y, z))
        wp = aligned_ptr(ws)
        _lib.check(lib.ct_unet_conv_block(self._handle, int(layer), BLOCK_ENGINES[engine], x_dev.data_ptr(), out.data_ptr(),
                                          b, x, y, z, wp, ws.numel() - (wp - ws.data_ptr()), stream_ptr()))
        return out

    def _balanced_batch(self, n_tiles):
        """Tiles per launch: the fewest batches that respect `tiles_per_batch`, of (almost) equal size -- 100 tiles
        at 38 per batch run as 34 + 34 + 32 rather than 38 + 38 + 24 (a short last batch leaves SMs idle in every
        persistent convolution launch)."""
        n = max(1, int(n_tiles))
        cap = max(1, int(self.tiles_per_batch))
        batches = -(-n // cap)
        return -(-n // batches)

    # ---- unet3_prediction on a device-resident normalised volume
    def tile_count(self, shape_xyz, shrink):
        sh = (C.c_int * 3)(*[int(s) for s in shrink])
        counts = (C.c_int * 3)()
        n = _lib.lib().ct_unet_tile_count(self._handle, int(shape_xyz[0]), int(shape_xyz[1]), int(shape_xyz[2]),
                                          C.byref(sh), C.byref(counts))
        if n < 0:
            raise ValueError(_lib.lib().ct_last_error().decode())
        return n, tuple(counts)

    def prediction_device(self, vol_dev, shrink=(24, 24, 2), tile_range=None, out=None):
        """vol_dev (x,y,z) float32 CUDA -> prob (x,y,z) float32 CUDA; `tile_range` = (begin, end) shards tiles."""
        x, y, z = (int(s) for s in vol_dev.shape)
        n, _ = self.tile_count((x, y, z), shrink)
        begin, end = (0, n) if tile_range is None else tile_range
        if out is None:
            out = torch.zeros((x, y, z), dtype=torch.float32, device=vol_dev.device)
        tpb = self._balanced_batch(end - begin)
        wp, wn = self._workspace(tpb)
        sh = (C.c_int * 3)(*[int(s) for s in shrink])
        _lib.check(_lib.lib().ct_unet3_prediction(self._handle, vol_dev.data_ptr(), out.data_ptr(), x, y, z,
                                                  C.byref(sh), int(begin), int(end), wp, wn, tpb, stream_ptr()))
        return out


    def prediction_block_device(self, block_dev, block_lo, volume_shape, shrink, tile_lo, tile_hi, out_lo, out_dim,
                                out=None):
        """One rank's share of unet3_prediction under spatial decomposition (spatial.py): block_dev holds the
        normalised voxels of the box starting at block_lo; tiles tile_lo <= (i,j,k) < tile_hi of the volume's tile
        grid are run and their centre windows written into a block starting at out_lo of extent out_dim."""
        block_dev = block_dev.contiguous()
        x, y, z = (int(s) for s in volume_shape)
        if out is None:
            out = torch.zeros(tuple(int(s) for s in out_dim), dtype=torch.float32, device=block_dev.device)
        n = max(1, math.prod(int(h) - int(l) for l, h in zip(tile_lo, tile_hi)))
        tpb = self._balanced_batch(n)
        wp, wn = self._workspace(tpb)
        i3 = lambda v: (C.c_int * 3)(*[int(q) for q in v])
        _lib.check(_lib.lib().ct_unet3_prediction_block(
            self._handle, block_dev.data_ptr(), C.byref(i3(block_lo)), C.byref(i3(block_dev.shape)), out.data_ptr(),
            C.byref(i3(out_lo)), C.byref(i3(out.shape)), x, y, z, C.byref(i3(shrink)), C.byref(i3(tile_lo)),
            C.byref(i3(tile_hi)), wp, wn, tpb, stream_ptr()))
        return out


def unet3_a(**kw):
    """unet3d.py:26-37."""
    return UNet3("a", **kw)


def unet3_b(**kw):
    """unet3d.py:40-67."""
    return UNet3("b", **kw)


def unet3_c(**kw):
    """unet3d.py:70-81."""
    return UNet3("c", **kw)


def _get_sizes_padded_im(img_siz_i, out_centr_siz_i):
    """unet3d.py:259-279."""
    num_axis_i = int(math.ceil(img_siz_i * 1.0 / out_centr_siz_i))
    return num_axis_i * out_centr_siz_i, num_axis_i


def unet3_prediction(img, model, shrink=(24, 24, 2)):
    """Predict cell / non-cell regions tile by tile (unet3d.py:203-256).

    img: (1, x, y, z, 1) normalised image (ndarray).  Returns (1, x, y, z, 1) float32 ndarray.
    `model` must be a UNet3 (the reference passes a Keras model; its duck type is kept: input_shape,
    output_shape, predict)."""
    arr = np.asarray(img)
    if arr.ndim != 5 or arr.shape[0] != 1 or arr.shape[4] != 1:
        raise ValueError(f"expected img of shape (1, x, y, z, 1), got {arr.shape}")
    if not isinstance(model, UNet3):
        raise TypeError("unet3_prediction needs a UNet3 model (there is no CPU / Keras fallback)")
    dev = to_device(arr[0, :, :, :, 0].astype(np.float32, copy=False), torch.float32)
    out = model.prediction_device(dev, shrink)
    return out.cpu().numpy()[None, ..., None]
