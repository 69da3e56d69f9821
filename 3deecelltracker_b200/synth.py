"""Seeded synthetic inputs for tests and bench.py: uint16 stacks of Gaussian blobs, point sets
with affine motion + dropped / added cells.  Pure NumPy, host side; nothing here is on the timed path."""
import numpy as np


def blob_centres(shape_xyz, k, seed, margin=8):
    rng = np.random.default_rng(seed)
    lo = np.array([margin, margin, min(margin, shape_xyz[2] // 4)], dtype=np.float64)
    hi = np.array(shape_xyz, dtype=np.float64) - lo
    return rng.uniform(lo, hi, (k, 3))


def blob_stack(shape_xyz, centres, seed, z_xy_ratio=1.0, sigma_xy=4.0, background=None):
    """uint16 (x,y,z): background N(100,10^2) clipped at 0 + Gaussian blobs of amplitude U(300,3000).
    `background`: a ready-made float32 noise field (time-lapses reuse one field instead of drawing 9 M normals per
    volume); the blob amplitudes then come first from the generator, so they are the same for every volume."""
    rng = np.random.default_rng(seed)
    x, y, z = shape_xyz
    if background is None:
        img = rng.normal(100.0, 10.0, shape_xyz).astype(np.float32)
    else:
        img = np.array(background, dtype=np.float32, copy=True)
    sigma_z = max(1.0, sigma_xy / z_xy_ratio)
    amp = rng.uniform(300, 3000, len(centres))
    rx, rz = int(3 * sigma_xy) + 1, int(3 * sigma_z) + 1
    for (cx, cy, cz), a in zip(centres, amp):
        x0, x1 = max(int(cx) - rx, 0), min(int(cx) + rx + 1, x)
        y0, y1 = max(int(cy) - rx, 0), min(int(cy) + rx + 1, y)
        z0, z1 = max(int(cz) - rz, 0), min(int(cz) + rz + 1, z)
        gx = np.exp(-0.5 * ((np.arange(x0, x1) - cx) / sigma_xy) ** 2)
        gy = np.exp(-0.5 * ((np.arange(y0, y1) - cy) / sigma_xy) ** 2)
        gz = np.exp(-0.5 * ((np.arange(z0, z1) - cz) / sigma_z) ** 2)
        img[x0:x1, y0:y1, z0:z1] += (a * gx[:, None, None] * gy[None, :, None] * gz[None, None, :]).astype(np.float32)
    return np.clip(img, 0, 65535).astype(np.uint16)


def move_points(points, seed, affine_level=0.05, noise=0.002, drop=0.05, add=0.05):
    """Affine-perturbed copy with dropped / added points (semantics of ffn.py:29-54 in normalised coordinates)."""
    rng = np.random.default_rng(seed)
    mean = points.mean(axis=0)
    c = points - mean
    scale = np.abs(c).max()
    a = np.eye(3) + (rng.random((3, 3)) - 0.5) * affine_level
    t = (c / scale) @ a + (rng.random(c.shape) - 0.5) * 4 * noise
    t = t * scale + mean
    t = t[rng.random(len(t)) > drop]
    extra = rng.uniform(points.min(axis=0), points.max(axis=0), (int(add * len(points)), 3))
    t = np.concatenate([t, extra], axis=0)
    return t[rng.permutation(len(t))]


def random_points(n, seed, extent=(512.0, 512.0, 320.0)):
    rng = np.random.default_rng(seed)
    return rng.uniform(0, 1, (n, 3)) * np.asarray(extent)


def unet_weights(variant="a", seed=0):
    """Seeded random-init U-Net weights in Keras order (He-normal kernels, non-trivial BatchNorm statistics) for
    benchmarks: throughput does not depend on the weight values."""
    import math
    from .unet3d import _SPECS, _conv_layers
    rng = np.random.default_rng(seed)
    spec = _SPECS[variant]
    ws = []
    for cin, cout in _conv_layers(spec):
        ws.append((rng.standard_normal((3, 3, 3, cin, cout)) * math.sqrt(2.0 / (27 * cin))).astype(np.float32))
        ws.append((rng.standard_normal(cout) * 0.05).astype(np.float32))
        ws.append(rng.uniform(0.5, 1.5, cout).astype(np.float32))
        ws.append((rng.standard_normal(cout) * 0.1).astype(np.float32))
        ws.append((rng.standard_normal(cout) * 0.1).astype(np.float32))
        ws.append(rng.uniform(0.5, 1.5, cout).astype(np.float32))
    c = spec["out"][1]
    ws.append((rng.standard_normal((1, 1, 1, c, 1)) * math.sqrt(2.0 / c)).astype(np.float32))
    ws.append((rng.standard_normal(1) * 0.05).astype(np.float32))
    return ws


def detector_unet_weights(seed=0, gain=8.0, bias=-4.0, mix=0.05):
    """unet3_a weights that DETECT bright blobs, for end-to-end runs without trained weights (the reference's
    unet3_pretrained.h5 is a download, README.md:67-69): output channel 0 of d0a, d0b, o_m2 (reading the level-0 skip
    half of its concat input) and o_m1 carries the normalised intensity (3x3x3 box means in d0a and d0b, which
    suppress the single-voxel background noise, then centre taps) with identity
    BatchNorm, every other kernel keeps its seeded random-init value, and the 1x1x1 head is
    sigmoid(gain * ch0 + mix * (random combination of the other 7 channels) + bias).  The probability map is therefore
    a steep monotone function of the LCN-normalised intensity, perturbed by the full random network -- every layer
    contributes to every output value, and the cell / background decision is the blob detector's."""
    from .unet3d import _SPECS, _conv_layers
    ws = unet_weights("a", seed)
    layers = _conv_layers(_SPECS["a"])
    src_channel = {0: 0, 1: 0, len(layers) - 2: 16, len(layers) - 1: 0}
    for i, src in src_channel.items():
        k = ws[6 * i]
        k[..., 0] = 0.0
        if i <= 1:
            k[:, :, :, src, 0] = 1.0 / 27.0
        else:
            k[1, 1, 1, src, 0] = 1.0
        ws[6 * i + 1][0] = 0.0                       # conv bias
        ws[6 * i + 2][0] = 1.0                       # gamma
        ws[6 * i + 3][0] = 0.0                       # beta
        ws[6 * i + 4][0] = 0.0                       # moving mean
        ws[6 * i + 5][0] = 1.0                       # moving variance
    head = ws[-2]
    head[0, 0, 0, 1:, 0] *= mix
    head[0, 0, 0, 0, 0] = gain
    ws[-1][0] = bias
    return ws


def ffn_weights(seed=0):
    """Seeded random-init FFN weights in Keras order (Glorot-uniform kernels, non-trivial BatchNorm statistics)."""
    import math
    rng = np.random.default_rng(seed)

    def glorot(fi, fo):
        lim = math.sqrt(6.0 / (fi + fo))
        return rng.uniform(-lim, lim, (fi, fo)).astype(np.float32)

    def bn(c):
        return [rng.uniform(0.5, 1.5, c).astype(np.float32), (rng.standard_normal(c) * 0.1).astype(np.float32),
                (rng.standard_normal(c) * 0.1).astype(np.float32), rng.uniform(0.5, 1.5, c).astype(np.float32)]

    return [glorot(61, 512)] + bn(512) + [glorot(1024, 512)] + bn(512) + \
           [glorot(512, 1), (rng.standard_normal(1) * 0.05).astype(np.float32)]
