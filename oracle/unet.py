"""CPU oracle for the 3D U-Net path: LCN normalisation, the three U-Net graphs and the tiled
prediction (torch-CPU, fp32 or fp64).

TEST INFRASTRUCTURE ONLY -- see oracle/prgls.py header for who may import this.

Parity status: UNPINNED for the Keras graphs.  TensorFlow/Keras cannot be installed in this image and
the reference ships no tests, golden vectors or weights for unet3d.py / preprocess.py, so there is
nothing to check this restatement against except the source text.  It follows
  unet3d.py:84-98   (_unet3_depth3 graph),  unet3d.py:40-67 (unet3_b graph),
  unet3d.py:101-141 (Conv3D 3x3x3 'same' -> LeakyReLU(0.3) -> BatchNorm(eval, eps=1e-3); ReLU variant),
  unet3d.py:144-200 (MaxPooling3D(pool), UpSampling3D(size) nearest, concatenate([up, skip])),
  unet3d.py:203-279 (reflect pre-pad, tile grid, centre crop / scatter),
  preprocess.py:117-188 (median subtract, clamp, two zero-padded 27x27x1 box filters in fp32)
with the Keras defaults written out.  The tiling control flow (`unet3_prediction`) is pure NumPy in the
reference and is PINNED: tests/golden holds outputs of the reference's own loop (executed by
oracle/make_golden.py around a duck-typed model).

Weight container ("Keras order"): for every Conv3D+BN block the arrays
  kernel (3,3,3,Cin,Cout), bias (Cout), gamma, beta, moving_mean, moving_var (Cout each)
and for the head kernel (1,1,1,C,1), bias (1) -- the order Keras `Model.get_weights()` returns.
"""
import itertools
import math

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3          # keras BatchNormalization default epsilon
LEAKY_ALPHA = 0.3      # keras LeakyReLU default alpha


# --------------------------------------------------------------------------------------------------
# architecture description shared by the three builders
# --------------------------------------------------------------------------------------------------
def unet_spec(variant):
    """Return dict(input, pool, act, down, up, out) for 'a' | 'b' | 'c'.
    a: unet3d.py:26-37,84-98   b: unet3d.py:40-67   c: unet3d.py:70-81."""
    if variant == "a":
        return dict(input=(160, 160, 16), pool=(2, 2, 1), act="leaky",
                    down=[(8, 16), (16, 32), (32, 64)], up=[(64, 64), (32, 32), (16, 16)], out=(8, 8))
    if variant == "b":
        return dict(input=(96, 96, 8), pool=(2, 2, 1), act="relu",
                    down=[(64, 64), (128, 128)], up=[(256, 256), (128, 128)], out=(64, 64))
    if variant == "c":
        return dict(input=(64, 64, 64), pool=(2, 2, 2), act="leaky",
                    down=[(8, 16), (16, 32), (32, 64)], up=[(64, 64), (32, 32), (16, 16)], out=(8, 8))
    raise ValueError(variant)


def conv_layers(spec):
    """List of (cin, cout) for the 3x3x3 blocks in graph (= weight) order."""
    layers = []
    c = 1
    skips = []
    for (f1, f2) in spec["down"]:
        layers += [(c, f1), (f1, f2)]
        skips.append(f2)
        c = f2
    for (f1, f2), skip in zip(spec["up"], reversed(skips)):
        layers += [(c, f1), (f1, f2)]
        c = f2 + skip
    layers += [(c, spec["out"][0]), (spec["out"][0], spec["out"][1])]
    return layers


def random_weights(variant, seed=0):
    """Seeded parity weights: He-normal kernels, bias N(0,0.05^2), BN gamma U(0.5,1.5),
    beta/mean N(0,0.1^2), var U(0.5,1.5).  Returns list of float32 arrays in Keras order."""
    rng = np.random.default_rng(seed)
    spec = unet_spec(variant)
    ws = []
    for cin, cout in conv_layers(spec):
        fan_in = 27 * cin
        ws.append((rng.standard_normal((3, 3, 3, cin, cout)) * math.sqrt(2.0 / fan_in)).astype(np.float32))
        ws.append((rng.standard_normal(cout) * 0.05).astype(np.float32))
        ws.append(rng.uniform(0.5, 1.5, cout).astype(np.float32))
        ws.append((rng.standard_normal(cout) * 0.1).astype(np.float32))
        ws.append((rng.standard_normal(cout) * 0.1).astype(np.float32))
        ws.append(rng.uniform(0.5, 1.5, cout).astype(np.float32))
    c_last = spec["out"][1]
    ws.append((rng.standard_normal((1, 1, 1, c_last, 1)) * math.sqrt(2.0 / c_last)).astype(np.float32))
    ws.append((rng.standard_normal(1) * 0.05).astype(np.float32))
    return ws


class UNetOracle:
    """Duck-type of the Keras model used by unet3_prediction: .input_shape, .output_shape, .predict."""

    def __init__(self, variant, weights, dtype=torch.float32, threads=None):
        self.spec = unet_spec(variant)
        self.dtype = dtype
        self.layers = conv_layers(self.spec)
        assert len(weights) == 6 * len(self.layers) + 2
        self.w = [torch.from_numpy(np.asarray(a)).to(dtype) for a in weights]
        x, y, z = self.spec["input"]
        self.input_shape = (None, x, y, z, 1)
        self.output_shape = (None, x, y, z, 1)
        if threads:
            torch.set_num_threads(threads)

    # Conv3D(3,'same') -> activation -> BN(eval): unet3d.py:117-119 / :139-140
    def _block(self, x, i):
        k, b, g, be, mu, var = self.w[6 * i:6 * i + 6]
        x = F.conv3d(x, k.permute(4, 3, 0, 1, 2).contiguous(), b, padding=1)
        x = F.leaky_relu(x, LEAKY_ALPHA) if self.spec["act"] == "leaky" else F.relu(x)
        s = (g / torch.sqrt(var + BN_EPS)).view(1, -1, 1, 1, 1)
        return (x - mu.view(1, -1, 1, 1, 1)) * s + be.view(1, -1, 1, 1, 1)

    def forward(self, x_ncxyz, return_intermediates=False):
        spec = self.spec
        pool = spec["pool"]
        x = x_ncxyz.to(self.dtype)
        inter = []
        skips = []
        i = 0
        for _ in spec["down"]:
            x = self._block(x, i); inter.append(x)
            x = self._block(x, i + 1); inter.append(x)
            i += 2
            skips.append(x)
            x = F.max_pool3d(x, kernel_size=pool, stride=pool)
        for _ in spec["up"]:
            x = self._block(x, i); inter.append(x)
            x = self._block(x, i + 1); inter.append(x)
            i += 2
            x = F.interpolate(x, scale_factor=tuple(float(p) for p in pool), mode="nearest")
            x = torch.cat([x, skips.pop()], dim=1)          # concatenate([UpSampling3D(im_2), horiz])
        x = self._block(x, i); inter.append(x)
        x = self._block(x, i + 1); inter.append(x)
        k, b = self.w[-2], self.w[-1]
        x = torch.sigmoid(F.conv3d(x, k.permute(4, 3, 0, 1, 2).contiguous(), b))
        if return_intermediates:
            return x, inter
        return x

    def predict(self, tiles_bxyzc):
        """(B,x,y,z,1) ndarray -> (B,x,y,z,1) float32 ndarray, like keras Model.predict."""
        t = torch.from_numpy(np.ascontiguousarray(tiles_bxyzc)).permute(0, 4, 1, 2, 3)
        with torch.no_grad():
            y = self.forward(t)
        return y.permute(0, 2, 3, 4, 1).to(torch.float32).numpy()


# --------------------------------------------------------------------------------------------------
# tiling (unet3d.py:203-279)
# --------------------------------------------------------------------------------------------------
def padded_size(size_i, centre_i):
    n = int(math.ceil(size_i * 1.0 / centre_i))
    return n * centre_i, n


def tile_grid(shape_xyz, model_in_xyz, shrink):
    """Centre sizes, tile counts and pad widths of unet3_prediction (unet3d.py:221-233)."""
    centre = tuple(model_in_xyz[i] - 2 * shrink[i] for i in range(3))
    padded, num = zip(*[padded_size(shape_xyz[i], centre[i]) for i in range(3)])
    before = tuple(shrink)
    after = tuple(shrink[i] + padded[i] - shape_xyz[i] for i in range(3))
    return centre, num, padded, before, after


def unet3_prediction(img_1xyz1, model, shrink=(24, 24, 2)):
    """Restatement of unet3d.py:203-256 (one model.predict per tile, batch 1)."""
    x, y, z = img_1xyz1.shape[1:4]
    tin = model.input_shape[1:4]
    centre, num, padded, before, after = tile_grid((x, y, z), model.output_shape[1:4], shrink)
    pad = np.pad(img_1xyz1[0, :, :, :, 0], tuple(zip(before, after)), "reflect")
    out = np.zeros((1,) + tuple(padded) + (1,), dtype="float32")
    for i, j, k in itertools.product(range(num[0]), range(num[1]), range(num[2])):
        o = (i * centre[0], j * centre[1], k * centre[2])
        tile = pad[o[0]:o[0] + tin[0], o[1]:o[1] + tin[1], o[2]:o[2] + tin[2]][None, ..., None]
        pred = model.predict(tile)
        out[0, o[0]:o[0] + centre[0], o[1]:o[1] + centre[1], o[2]:o[2] + centre[2], 0] = \
            pred[0, before[0]:before[0] + centre[0], before[1]:before[1] + centre[1],
                 before[2]:before[2] + centre[2], 0]
    return out[:, :x, :y, :z, :]


# --------------------------------------------------------------------------------------------------
# LCN normalisation (preprocess.py:117-188)
# --------------------------------------------------------------------------------------------------
def _box_sum_keras_fp32(vol_xyz, filter_size):
    """Keras Conv3D(1, filter_size, ones kernel, zero bias, padding='same') on a float32 copy of the
    input (keras casts predict() inputs to float32).  preprocess.py:117-133."""
    t = torch.from_numpy(np.asarray(vol_xyz, dtype=np.float32))[None, None]
    k = torch.ones((1, 1) + tuple(filter_size), dtype=torch.float32)
    pad = tuple(s // 2 for s in filter_size)
    with torch.no_grad():
        return F.conv3d(t, k, padding=pad)[0, 0].numpy()


def lcn(img3d, noise_level=5, filter_size=(27, 27, 1)):
    """preprocess.py:136-167 with the NumPy dtype promotions of the reference kept: img3d float64,
    avg/std float32, result float64."""
    volume = filter_size[0] * filter_size[1] * filter_size[2]
    img3d = np.asarray(img3d)
    avg = _box_sum_keras_fp32(img3d, filter_size) / volume
    diff_sqr = np.square(img3d - avg)
    std = np.sqrt(_box_sum_keras_fp32(diff_sqr, filter_size) / volume)
    return np.divide(img3d - avg, std + np.float32(noise_level))


def normalize_image(image, noise_level):
    """preprocess.py:170-188."""
    image_norm = image - np.median(image)
    image_norm[image_norm < 0] = 0
    return lcn(image_norm, noise_level, filter_size=(27, 27, 1))
