"""File formats of the reference, so that a folder prepared for / written by 3DeeCellTracker works unchanged:

  * Keras `.h5` models / weight files (tracker.py:579, :1121; trackerlite.py:57-63) -> lists of arrays in
    `model.get_weights()` order.  Reading needs `h5py` (not installed in the build image; the loader raises with the
    conversion instruction when it is missing).  `convert_keras_h5` / `python -m 3deecelltracker_b200.io_formats`
    turns a `.h5` into the `.npz` container `UNet3.load_weights` / `FFN.load_weights` read everywhere.
  * the U-Net probability cache `unet_cache/t%06i.npy` in float16 (tracker.py:652-669);
  * label images as 2-D TIFF sequences `track_results_t%06i_z%04i.tif`, `auto_vol1_z%04i.tif` (tracker.py:145-190,
    coord_image_transformer.py:561-583) and real coordinates `coords%06d.npy` (coord_image_transformer.py:267,512);
  * the folder layout of `Paths.make_folders` (tracker.py:734-752).
Host-side glue only: nothing here is on the timed path.
"""
import os

import numpy as np


# ---------------------------------------------------------------------------------------------- Keras .h5 weights
def _decode(name):
    return name.decode("utf-8") if isinstance(name, bytes) else str(name)


def keras_h5_weight_list(h5file):
    """Arrays of a Keras HDF5 file in `model.get_weights()` order.  `h5file`: an open h5py.File (or any mapping with
    the same `.attrs` / item protocol -- the traversal is tested with a stand-in).  Handles both layouts Keras writes:
    a full model saved with `model.save(path)` (weights under the group `model_weights`) and a bare
    `model.save_weights(path)` file.  Order: `layer_names` attribute of the group, then each layer's `weight_names`."""
    root = h5file["model_weights"] if "model_weights" in h5file else h5file
    out = []
    for layer in root.attrs["layer_names"]:
        g = root[_decode(layer)]
        for wname in g.attrs["weight_names"]:
            node = g
            for part in _decode(wname).split("/"):
                node = node[part]
            out.append(np.asarray(node, dtype=np.float32))
    return out


def load_keras_h5(path):
    try:
        import h5py
    except ImportError as e:
        raise ImportError(f"{path}: reading Keras .h5 files needs h5py, which is not installed here.  Convert the file "
                          "once on a machine that has it: `python -m 3deecelltracker_b200.io_formats model.h5 "
                          "model.npz`, then pass the .npz") from e
    with h5py.File(path, "r") as f:
        return keras_h5_weight_list(f)


def load_weight_file(path):
    """`.npz` written by save_weights / convert_keras_h5, or a Keras `.h5`."""
    if str(path).lower().endswith((".h5", ".hdf5", ".keras")):
        return load_keras_h5(path)
    with np.load(path) as f:
        return [f[f"arr_{i}"] for i in range(len(f.files))]


def convert_keras_h5(src_h5, dst_npz):
    ws = load_keras_h5(src_h5)
    np.savez(dst_npz, *ws)
    return len(ws)


# ---------------------------------------------------------------------------------------------- folders
def _make_folder(path_i, print_=False):
    os.makedirs(path_i, exist_ok=True)                     # tracker.py:64-80
    if print_:
        print(os.path.relpath(path_i, start=os.getcwd()))
    return path_i


def get_tracking_path(adjacent, ensemble, folder_path):
    """tracker.py:83-110."""
    if not ensemble:
        return os.path.join(folder_path, "track_results_SingleMode/")
    if not adjacent:
        return os.path.join(folder_path, "track_results_EnsembleDstrbtMode/")
    return os.path.join(folder_path, "track_results_EnsembleAdjctMode/")


class Paths:
    """tracker.py:687-752: same attributes, same folder names."""

    def __init__(self, folder_path, image_name, unet_model_file, ffn_model_file):
        self.folder = folder_path
        self.models = self.unet_cache = self.raw_image = None
        self.auto_segmentation_vol1 = self.manual_segmentation_vol1 = None
        self.unet_weights = self.track_results = self.track_information = self.anim = None
        self.image_name = image_name
        self.unet_model_file = unet_model_file
        self.ffn_model_file = ffn_model_file

    def make_folders(self, adjacent, ensemble):
        f = self.folder
        self.raw_image = _make_folder(os.path.join(f, "data/"))
        self.auto_segmentation_vol1 = _make_folder(os.path.join(f, "auto_vol1/"))
        self.manual_segmentation_vol1 = _make_folder(os.path.join(f, "manual_vol1/"))
        self.track_information = _make_folder(os.path.join(f, "track_information/"))
        self.models = _make_folder(os.path.join(f, "models/"))
        self.unet_cache = _make_folder(os.path.join(f, "unet_cache/"))
        self.track_results = _make_folder(get_tracking_path(adjacent, ensemble, f))
        self.anim = _make_folder(os.path.join(f, "anim/"))
        self.unet_weights = _make_folder(os.path.join(self.models, "unet_weights/"))


# ---------------------------------------------------------------------------------------------- U-Net cache
def unet_cache_file(cache_dir, vol):
    return os.path.join(cache_dir, "t%06i.npy" % vol)


def save_unet_cache(cache_dir, vol, image_cell_bg):
    """tracker.py:668: the (1, x, y, z, 1) probability map as float16."""
    np.save(unet_cache_file(cache_dir, vol), np.array(image_cell_bg, dtype="float16"))


def load_unet_cache(cache_dir, vol):
    """tracker.py:656-660: the cached map (float16) or None."""
    try:
        return np.load(unet_cache_file(cache_dir, vol), allow_pickle=True)
    except OSError:
        return None


# ---------------------------------------------------------------------------------------------- label images / coordinates
def save_img3ts(z_range, img, path, t, use_8_bit=True):
    """tracker.py:168-190: slices `z_range` of a (x, y, z) volume as path % (t, 1..)."""
    from PIL import Image
    dtype = np.uint8 if use_8_bit else np.uint16
    for i, z in enumerate(z_range):
        Image.fromarray(np.asarray(img[:, :, z]).astype(dtype)).save(path % (t, i + 1))


def save_automatic_segmentation(labels_xyz, folder_path, use_8_bit):
    """tracker.py:145-165."""
    from PIL import Image
    os.makedirs(os.path.join(folder_path, "auto_vol1"), exist_ok=True)
    dtype = np.uint8 if use_8_bit else np.uint16
    for z in range(1, labels_xyz.shape[2] + 1):
        Image.fromarray(np.asarray(labels_xyz[:, :, z - 1]).astype(dtype)).save(
            os.path.join(folder_path, "auto_vol1", "auto_vol1_z%04i.tif" % z))


def save_tracked_labels(results_folder, labels_xyz, t, use_8_bit):
    """coord_image_transformer.py:561-583 (LZW-compressed TIFF slices under track_results/labels)."""
    from PIL import Image
    path = os.path.join(str(results_folder), "track_results", "labels")
    os.makedirs(path, exist_ok=True)
    dtype = np.uint8 if use_8_bit else np.uint16
    for z in range(1, labels_xyz.shape[2] + 1):
        with Image.fromarray(np.asarray(labels_xyz[:, :, z - 1]).astype(dtype)) as img:
            img.save(os.path.join(path, "track_results_t%06i_z%04i.tif" % (t, z)), compression="tiff_lzw")


def save_coords_real(results_folder, coords_real, t):
    """coord_image_transformer.py:267,512."""
    path = os.path.join(str(results_folder), "track_results", "coords_real")
    os.makedirs(path, exist_ok=True)
    np.save(os.path.join(path, "coords%06d.npy" % t), np.asarray(coords_real))


def read_tiff_stack(pattern_path, t, z_range):
    """tracker.py:113-142 with PIL instead of tifffile: slices path % (t, z) stacked into (x, y, z)."""
    from PIL import Image
    return np.array([np.array(Image.open(pattern_path % (t, z))) for z in range(z_range[0], z_range[1])]).transpose((1, 2, 0))


if __name__ == "__main__":
    import sys
    if len(sys.argv) != 3:
        raise SystemExit("usage: python -m 3deecelltracker_b200.io_formats <keras model or weights .h5> <out.npz>")
    print(f"wrote {convert_keras_h5(sys.argv[1], sys.argv[2])} arrays to {sys.argv[2]}")
