"""CPU: file formats of the reference (io_formats.py): Keras .h5 weight traversal (with an h5py stand-in -- h5py is
not installed here), float16 U-Net cache, TIFF label sequences, coords .npy, folder layout."""
import importlib
import os

import numpy as np
import pytest

from conftest import load_pkg


@pytest.fixture(scope="module")
def io():
    load_pkg()
    return importlib.import_module("3deecelltracker_b200.io_formats")


class FakeH5(dict):
    """Mapping with .attrs, like h5py.File / h5py.Group."""

    def __init__(self, items=(), **attrs):
        super().__init__(items)
        self.attrs = attrs


def _keras_layer(name, weights):
    """Keras writes weight_names like b'conv3d/kernel:0' and nests the datasets under a group of the layer's name."""
    inner = FakeH5({k: v for k, v in weights})
    return FakeH5({name: inner}, weight_names=[f"{name}/{k}".encode() for k, _ in weights])


def test_keras_h5_traversal_follows_get_weights_order(io):
    rng = np.random.default_rng(0)
    k1, b1 = rng.normal(size=(3, 3, 3, 1, 8)), rng.normal(size=8)
    gamma, beta, mean, var = (rng.normal(size=8) for _ in range(4))
    k2, b2 = rng.normal(size=(1, 1, 1, 8, 1)), rng.normal(size=1)
    layers = {
        "input_1": FakeH5(weight_names=[]),
        "conv3d": _keras_layer("conv3d", [("kernel:0", k1), ("bias:0", b1)]),
        "leaky_re_lu": FakeH5(weight_names=[]),
        "batch_normalization": _keras_layer("batch_normalization", [("gamma:0", gamma), ("beta:0", beta),
                                                                    ("moving_mean:0", mean), ("moving_variance:0", var)]),
        "conv3d_1": _keras_layer("conv3d_1", [("kernel:0", k2), ("bias:0", b2)]),
    }
    names = [n.encode() for n in layers]
    bare = FakeH5(layers, layer_names=names)                        # model.save_weights layout
    full = FakeH5({"model_weights": FakeH5(layers, layer_names=names)}, keras_version=b"2.11.0")   # model.save layout
    want = [k1, b1, gamma, beta, mean, var, k2, b2]
    for f in (bare, full):
        got = io.keras_h5_weight_list(f)
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert g.dtype == np.float32 and np.array_equal(g, w.astype(np.float32))


def test_h5_without_h5py_raises_with_instruction(io, tmp_path):
    try:
        import h5py  # noqa: F401
        pytest.skip("h5py is installed")
    except ImportError:
        pass
    with pytest.raises(ImportError, match="io_formats"):
        io.load_weight_file(str(tmp_path / "unet3_pretrained.h5"))


def test_npz_weight_container_round_trip(io, tmp_path):
    ws = [np.arange(6, dtype=np.float32).reshape(2, 3), np.ones(4, np.float32)]
    np.savez(tmp_path / "w.npz", *ws)
    got = io.load_weight_file(str(tmp_path / "w.npz"))
    assert all(np.array_equal(a, b) for a, b in zip(got, ws))


def test_unet_cache_is_float16_like_the_reference(io, tmp_path):
    prob = np.random.default_rng(1).random((1, 9, 8, 3, 1)).astype(np.float32)
    assert io.load_unet_cache(str(tmp_path), 7) is None
    io.save_unet_cache(str(tmp_path), 7, prob)
    assert os.path.isfile(tmp_path / "t000007.npy")
    back = io.load_unet_cache(str(tmp_path), 7)
    assert back.dtype == np.float16 and back.shape == prob.shape
    assert np.array_equal(back, prob.astype(np.float16))


def test_label_tiff_sequences_and_coords(io, tmp_path):
    labels = (np.random.default_rng(2).integers(0, 300, (12, 10, 4))).astype(np.int32)
    io.save_tracked_labels(tmp_path, labels, t=3, use_8_bit=False)
    back = io.read_tiff_stack(os.path.join(str(tmp_path), "track_results", "labels", "track_results_t%06i_z%04i.tif"), 3, (1, 5))
    assert back.dtype == np.uint16 and np.array_equal(back, labels.astype(np.uint16))
    os.makedirs(tmp_path / "res", exist_ok=True)
    io.save_img3ts(range(0, 4), labels % 200, os.path.join(str(tmp_path), "res", "track_results_t%06i_z%04i.tif"), t=1, use_8_bit=True)
    back8 = io.read_tiff_stack(os.path.join(str(tmp_path), "res", "track_results_t%06i_z%04i.tif"), 1, (1, 5))
    assert back8.dtype == np.uint8 and np.array_equal(back8, (labels % 200).astype(np.uint8))
    io.save_automatic_segmentation(labels % 200, str(tmp_path), use_8_bit=True)
    assert os.path.isfile(tmp_path / "auto_vol1" / "auto_vol1_z0004.tif")
    coords = np.random.default_rng(3).random((5, 3))
    io.save_coords_real(tmp_path, coords, 12)
    assert np.array_equal(np.load(tmp_path / "track_results" / "coords_real" / "coords000012.npy"), coords)


def test_paths_layout(io, tmp_path):
    p = io.Paths(str(tmp_path), "img_t%06d_z%04i.tif", "unet3_pretrained.h5", "ffn_pretrained.h5")
    p.make_folders(adjacent=False, ensemble=20)
    for sub in ("data", "auto_vol1", "manual_vol1", "track_information", "models", "unet_cache", "anim",
                "track_results_EnsembleDstrbtMode", os.path.join("models", "unet_weights")):
        assert os.path.isdir(tmp_path / sub), sub
    assert io.get_tracking_path(False, 0, "x").endswith("track_results_SingleMode/")
    assert io.get_tracking_path(True, 5, "x").endswith("track_results_EnsembleAdjctMode/")
