"""CPU: a second, independent restatement of the Keras layers (plain NumPy, written from the Keras documentation) against
the torch oracle of the U-Net graph (oracle/unet.py).

TensorFlow cannot be installed here, so the Keras half of the oracle cannot be pinned to the reference itself (DESIGN.md
section 4).  What this test does pin is the oracle's READING of the layer semantics the reference relies on
(unet3d.py:84-200): Conv3D(3, 'same') is a zero-padded cross-correlation with kernel layout (kx, ky, kz, Cin, Cout);
LeakyReLU() has alpha 0.3; BatchNormalization() at inference is gamma (x - mean) / sqrt(var + 1e-3) + beta;
MaxPooling3D / UpSampling3D(size) are non-overlapping max windows / element repetition; concatenate([up, skip]) puts the
up-sampled channels first; the head is a 1x1x1 convolution with a sigmoid.  Two restatements written independently and
agreeing to fp64 round-off make a misreading of a default unlikely."""
import numpy as np
import pytest
import torch

from oracle import unet as ounet


def conv3_same(x, k, b):
    """x (X,Y,Z,Cin), k (3,3,3,Cin,Cout) -> (X,Y,Z,Cout): zero 'same' padding, no kernel flip."""
    X, Y, Z, _ = x.shape
    p = np.pad(x, ((1, 1), (1, 1), (1, 1), (0, 0)))
    out = np.zeros((X, Y, Z, k.shape[4]))
    for dx in range(3):
        for dy in range(3):
            for dz in range(3):
                out += np.einsum("xyzc,co->xyzo", p[dx:dx + X, dy:dy + Y, dz:dz + Z], k[dx, dy, dz])
    return out + b


def block(x, w6, act):
    k, b, gamma, beta, mean, var = (np.asarray(a, np.float64) for a in w6)
    y = conv3_same(x, k, b)
    y = np.where(y > 0, y, 0.3 * y) if act == "leaky" else np.maximum(y, 0.0)
    return gamma * (y - mean) / np.sqrt(var + 1e-3) + beta


def max_pool(x, pool):
    X, Y, Z, C = x.shape
    px, py, pz = pool
    return x.reshape(X // px, px, Y // py, py, Z // pz, pz, C).max(axis=(1, 3, 5))


def up_sample(x, size):
    for axis, s in enumerate(size):
        x = np.repeat(x, s, axis=axis)
    return x


def unet_numpy(x, ws, spec):
    i, skips = 0, []
    for _ in spec["down"]:
        x = block(x, ws[6 * i:6 * i + 6], spec["act"]); i += 1
        x = block(x, ws[6 * i:6 * i + 6], spec["act"]); i += 1
        skips.append(x)
        x = max_pool(x, spec["pool"])
    for _ in spec["up"]:
        x = block(x, ws[6 * i:6 * i + 6], spec["act"]); i += 1
        x = block(x, ws[6 * i:6 * i + 6], spec["act"]); i += 1
        x = np.concatenate([up_sample(x, spec["pool"]), skips.pop()], axis=3)
    x = block(x, ws[6 * i:6 * i + 6], spec["act"]); i += 1
    x = block(x, ws[6 * i:6 * i + 6], spec["act"]); i += 1
    k, b = np.asarray(ws[-2], np.float64), np.asarray(ws[-1], np.float64)
    return 1.0 / (1.0 + np.exp(-(np.einsum("xyzc,co->xyzo", x, k[0, 0, 0]) + b)))


@pytest.mark.parametrize("variant,shape", [("a", (8, 16, 3)), ("b", (8, 4, 2)), ("c", (8, 8, 16))])
def test_numpy_layers_agree_with_the_torch_oracle(variant, shape):
    spec = ounet.unet_spec(variant)
    ws = ounet.random_weights(variant, seed=7)
    x = np.random.default_rng(1).normal(0, 1, shape + (1,))
    oracle = ounet.UNetOracle(variant, ws, dtype=torch.float64)
    with torch.no_grad():
        want = oracle.forward(torch.from_numpy(x).permute(3, 0, 1, 2)[None]).numpy()[0].transpose(1, 2, 3, 0)
    got = unet_numpy(x, ws, spec)
    assert got.shape == want.shape == shape + (1,)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-12)
