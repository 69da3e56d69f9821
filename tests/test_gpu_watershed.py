"""GPU parity: the watershed + centroid stage (ct_watershed_segment through the C ABI) against the oracle
(oracle/watershed.py: SciPy's distance_transform_edt / gaussian_filter executed for real, scikit-image restated).

Bar: BIT-EXACT label images (north_star: "bit-exact integer label maps after watershed"), bit-exact centres
(integer coordinate sums divided once in fp64), equal min_size / cell_num."""
import importlib

import numpy as np
import pytest
import torch

from conftest import load_pkg
from oracle import unet as ounet
from oracle import watershed as ows
from test_watershed_emul import shapes_volume

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m():
    load_pkg()
    names = ("watershed", "synth", "unet3d", "preprocess", "tracker", "ffn")
    return {n: importlib.import_module("3deecelltracker_b200." + n) for n in names}


def check(m, prob, ratio, method, min_size, cell_num):
    seg, cen, ms, cn = ows.segment(prob, ratio, method, min_size, cell_num)
    lab, cen2, ms2, cn2 = m["watershed"].segment(prob, ratio, method, min_size, cell_num)
    assert lab.dtype == np.int32 and lab.shape == prob.shape
    assert np.array_equal(seg, lab), f"{int((seg != lab).sum())} voxels differ"
    assert (ms, cn) == (ms2, cn2)
    assert cen.shape == cen2.shape and np.array_equal(cen, cen2)
    return seg


@pytest.mark.parametrize("trial", range(8))
def test_watershed_matches_oracle_on_shapes(m, trial):
    """Boxes (exact ties in the distance map) and balls, touching and isolated; every sampling ratio / method."""
    rng = np.random.default_rng(200 + trial)
    shape = (int(rng.integers(40, 120)), int(rng.integers(40, 120)), int(rng.integers(1, 14)))
    prob = shapes_volume(rng, shape, int(rng.integers(3, 25)))
    ratio = [9.2, 1.0, 2.5, 3.0][trial % 4]
    method = "cell_num" if trial % 3 == 0 else "min_size"
    check(m, prob, ratio, method, int(rng.integers(0, 40)), int(rng.integers(1, 3)))


def test_watershed_edge_cases(m):
    prob = np.zeros((24, 20, 3), np.float32)
    lab, cen, ms, cn = m["watershed"].segment(prob, 2.0, "min_size", 5, 0)           # no foreground at all
    assert lab.max() == 0 and cen.shape == (0, 3)
    prob[4:20, 5:17, :] = 0.9
    check(m, prob, 2.0, "min_size", 5, 0)                                              # one slab through every slice
    prob[:] = 0.9
    prob[0, 0, 0] = 0.1                                                                # almost everything foreground
    check(m, prob, 1.0, "min_size", 0, 0)
    with pytest.raises(ValueError):
        m["watershed"].segment(prob, 1.0, "nonsense", 0, 0)                            # watershed.py:99-100


@pytest.mark.parametrize("config", ["config1", "config2"])
def test_watershed_on_named_configs(m, config):
    """BASELINE configs[1] / configs[2]: probability map of the GPU U-Net (blob-detector weights) on the synthetic
    stack -> label image on the GPU vs the oracle applied to the SAME probability map: bit-identical labels, centres."""
    shape, cells, ratio = {"config1": ((512, 512, 35), 164, 9.2), "config2": ((160, 160, 16), 113, 9.2)}[config]
    synth, u, pre = m["synth"], m["unet3d"], m["preprocess"]
    raw = synth.blob_stack(shape, synth.blob_centres(shape, cells, 1234), 1234, z_xy_ratio=ratio)
    model = u.UNet3("a", weights=synth.detector_unet_weights(0), tiles_per_batch=38)
    norm = pre.normalize_image_device(pre._raw_to_device(raw), 20)
    prob_dev = model.prediction_device(norm, (24, 24, 2))
    seg_dev = m["watershed"].segment_device(prob_dev, ratio, "min_size", 40, 0)
    n, ms, cn = seg_dev.host_scalars()
    prob = prob_dev.cpu().numpy()
    seg, cen, ms_o, cn_o = ows.segment(prob, ratio, "min_size", 40, 0)
    assert 0.5 * cells <= n <= 1.5 * cells, f"{n} cells segmented from {cells} blobs"
    assert (n, ms, cn) == (int(seg.max()), ms_o, cn_o)
    assert np.array_equal(seg_dev.labels.cpu().numpy(), seg)
    assert np.array_equal(seg_dev.centres_host(), cen)
    # and end to end from the ORACLE's U-Net output: the masks are identical (test_gpu_lcn_unet.py), so are the labels
    want_norm = ounet.normalize_image(raw.copy(), 20).astype(np.float32)
    if config == "config2":
        prob_o = ounet.unet3_prediction(want_norm[None, ..., None], ounet.UNetOracle("a", synth.detector_unet_weights(0)),
                                        (24, 24, 2))[0, ..., 0]
        band = np.abs(prob_o.astype(np.float64) - 0.5) <= 0.5e-4
        if not band.any():
            seg_o, _, _, _ = ows.segment(prob_o, ratio, "min_size", 40, 0)
            assert np.array_equal(seg_o, seg)


def test_repeat_runs_are_deterministic(m):
    """Atomics and concurrent floods must not leak into the result: same labels on every run."""
    rng = np.random.default_rng(9)
    prob = torch.from_numpy(shapes_volume(rng, (150, 140, 12), 40)).cuda()
    first = m["watershed"].segment_device(prob, 9.2, "min_size", 10, 0)
    a, c = first.labels.clone(), first.centres_host()
    for _ in range(5):
        again = m["watershed"].segment_device(prob, 9.2, "min_size", 10, 0)
        assert torch.equal(again.labels, a) and np.array_equal(again.centres_host(), c)
