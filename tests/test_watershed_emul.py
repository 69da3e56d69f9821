"""CPU: the watershed pass pipeline (csrc/watershed_core.cuh), instantiated sequentially by tests/emul/ws_host.cpp,
against the oracle (oracle/watershed.py: SciPy's EDT / Gaussian for real + scikit-image restated) -- bit-exact label
images, centres, min_size and cell_num.  The CUDA instantiation of the same passes is checked on the GPU box by
tests/test_gpu_watershed.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import watershed as ows


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("ws_emul") / "libws_emul.so")
    src = os.path.join(ROOT, "tests", "emul", "ws_host.cpp")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, src], check=True)
    lib = C.CDLL(out)

    def weights(sigma):
        radius = int(4.0 * sigma + 0.5)
        x = np.arange(-radius, radius + 1)
        phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
        phi = phi / phi.sum()
        return np.ascontiguousarray(phi[radius:])

    def run(prob, ratio, method, min_size, cell_num, max_cells=4096):
        prob = np.ascontiguousarray(prob, np.float32)
        x, y, z = prob.shape
        lab = np.zeros(prob.shape, np.int32)
        cen = np.zeros((2, max_cells, 3))
        sc = np.zeros(4, np.int32)
        wxy, wz = weights(2.0), weights(0.3)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = lib.ws_emul_segment(vp(prob), x, y, z, C.c_double(ratio), 0 if method == "min_size" else 1, min_size,
                                 cell_num, vp(wxy), vp(wz), vp(lab), vp(cen), max_cells, vp(sc))
        assert rc == 0
        assert np.array_equal(cen[1, :sc[0]], cen[0, :sc[0]] * np.array([1.0, 1.0, ratio]))
        return lab, cen[0, :sc[0]], int(sc[1]), int(sc[2])
    return run


def shapes_volume(rng, shape, n):
    """Probability map made of boxes (flat plateaus: exact ties in the distance map) and balls, some touching."""
    p = np.zeros(shape, np.float32)
    X, Y, Z = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    for _ in range(n):
        c = rng.uniform([5, 5, 0], [shape[0] - 5, shape[1] - 5, shape[2]])
        if rng.random() < 0.5:
            h = rng.integers(3, 10, 3)
            m = (np.abs(X - int(c[0])) <= h[0]) & (np.abs(Y - int(c[1])) <= h[1]) & (np.abs(Z - int(c[2])) <= max(1, h[2] // 3))
        else:
            r = rng.uniform(4, 9)
            m = ((X - c[0]) ** 2 + (Y - c[1]) ** 2 + ((Z - c[2]) * 3) ** 2) <= r * r
        p[m] = rng.uniform(0.6, 1.0)
    return p


@pytest.mark.parametrize("trial", range(6))
def test_pass_pipeline_matches_oracle(emul, trial):
    rng = np.random.default_rng(100 + trial)
    shape = (int(rng.integers(40, 80)), int(rng.integers(40, 80)), int(rng.integers(2, 10)))
    prob = shapes_volume(rng, shape, int(rng.integers(3, 12)))
    ratio = [9.2, 1.0, 2.5, 3.0][trial % 4]
    method = "cell_num" if trial % 3 == 0 else "min_size"
    min_size, cell_num = int(rng.integers(0, 40)), int(rng.integers(1, 3))
    seg, cen, ms, cn = ows.segment(prob, ratio, method, min_size, cell_num)
    lab, cen2, ms2, cn2 = emul(prob, ratio, method, min_size, cell_num)
    assert np.array_equal(seg, lab)
    assert (ms, cn) == (ms2, cn2)
    assert cen.shape == cen2.shape and np.array_equal(cen, cen2)


def test_empty_and_full_foreground(emul):
    prob = np.zeros((20, 24, 3), np.float32)
    seg, cen, ms, cn = ows.segment(prob, 2.0, "min_size", 5, 0)
    lab, cen2, ms2, cn2 = emul(prob, 2.0, "min_size", 5, 0)
    assert seg.max() == 0 and lab.max() == 0 and len(cen2) == 0 and (ms, cn) == (ms2, cn2)
    prob[4:17, 5:19, :] = 0.9                                # one slab through every slice
    seg, cen, ms, cn = ows.segment(prob, 2.0, "min_size", 5, 0)
    lab, cen2, ms2, cn2 = emul(prob, 2.0, "min_size", 5, 0)
    assert np.array_equal(seg, lab) and np.array_equal(cen, cen2) and (ms, cn) == (ms2, cn2)


def test_oracle_pieces():
    """Small known answers for the scikit-image restatements, evaluated by hand from the definitions."""
    # find_boundaries(mode='outer', connectivity=1): background voxels 4-adjacent to a label, plus labelled voxels that
    # differ from a 4-neighbour AND see another label in their 3x3 neighbourhood (the 1 | 5 contact column)
    labels = np.array([[0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
                       [0, 0, 0, 0, 0, 5, 5, 5, 0, 0],
                       [0, 0, 1, 1, 1, 5, 5, 5, 0, 0],
                       [0, 0, 1, 1, 1, 5, 5, 5, 0, 0],
                       [0, 0, 1, 1, 1, 5, 5, 5, 0, 0],
                       [0, 0, 0, 0, 0, 5, 5, 5, 0, 0],
                       [0, 0, 0, 0, 0, 0, 0, 0, 0, 0]], dtype=np.uint8).astype(np.int32)
    want = np.array([[0, 0, 0, 0, 0, 1, 1, 1, 0, 0],
                     [0, 0, 1, 1, 1, 1, 0, 0, 1, 0],
                     [0, 1, 0, 0, 1, 1, 0, 0, 1, 0],
                     [0, 1, 0, 0, 1, 1, 0, 0, 1, 0],
                     [0, 1, 0, 0, 1, 1, 0, 0, 1, 0],
                     [0, 0, 1, 1, 1, 1, 0, 0, 1, 0],
                     [0, 0, 0, 0, 0, 1, 1, 1, 0, 0]], dtype=bool)
    assert np.array_equal(ows.find_boundaries_outer(labels, 1), want)
    # relabel_sequential: surviving labels keep their order
    assert list(ows.relabel_sequential(np.array([1, 1, 5, 5, 8, 99, 42]))) == [1, 1, 2, 2, 3, 5, 4]
    # watershed: two seeds on a 1-D ramp meet at the ridge; ties go to the earlier push
    img = np.array([[3.0, 2.0, 1.0, 2.0, 3.0, 2.0, 1.0, 2.0, 3.0]])
    markers = np.zeros((1, 9), np.int64)
    markers[0, 2], markers[0, 6] = 1, 2
    out = ows.watershed(img, markers, np.ones((1, 9), bool))
    assert list(out[0]) == [1, 1, 1, 1, 1, 2, 2, 2, 2]
