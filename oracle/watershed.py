"""CPU oracle for the watershed + centroid stage between the two hot paths.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package never imports it.

Reference: CellTracker/watershed.py:16-108 (watershed_2d, watershed_3d), CellTracker/tracker.py:671-684
(Tracker._watershed), :646-648 (centre of mass), called once per volume from Tracker._segment (:605-650).

Parity status: PARTLY PINNED.
  * The SciPy calls of the reference (`distance_transform_edt`, `gaussian_filter`; scipy is installed here and on the
    GPU box) are executed for real: EDT and smoothing are pinned to SciPy itself, bit for bit.
  * scikit-image is NOT installable here and the reference calls an API (`peak_local_max(..., indices=False)`,
    watershed.py:42,92) that scikit-image removed in 0.20 although pyproject.toml pins ^0.22 -- no released version
    satisfies both.  `peak_local_max`, `morphology.label`, `segmentation.watershed`, `find_boundaries`,
    `remove_small_objects` and `relabel_sequential` are therefore RESTATED from the scikit-image algorithm text
    (the last release that accepts the reference's call, 0.17/0.18-style semantics), with every tie-break made
    explicit below.  PARITY UNPINNED for these six functions: nothing external checks them.

Explicit choices (where the scikit-image text leaves order to the implementation):
  * peak_local_max: a voxel is a peak iff it equals the maximum of the (2 d + 1)^n box around it (zero padded) and is
    > image.min(); `exclude_border` clears that many voxels at both ends of every axis; no greedy spacing pass
    (pre-0.18 behaviour, which is why the reference merges plateau voxels with `morphology.label`).
  * label: full connectivity (8 / 26), components numbered by their first voxel in C (raster) order.
  * watershed (connectivity 1, no compactness, no watershed line): priority flood.  Queue order = (value, age, index)
    lexicographic; markers enter with age 0, every later push takes the next value of ONE global counter; neighbours
    are visited in C order of their offsets (x-1, y-1, [z-1, z+1,] y+1, x+1); a voxel is labelled when it is pushed.
    (scikit-image orders by (value, age) and leaves equal marker entries to its heap layout; `index` replaces that.)
  * find_boundaries(mode='outer', background=0): grey dilation != grey erosion over the connectivity footprint,
    restricted to background voxels and to voxels whose full 3^n neighbourhood holds two different non-zero labels;
    neighbourhoods are clipped at the image border.
"""
import ctypes
import heapq
import os

import numpy as np
from scipy import ndimage as ndi

_FLOOD_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libws_flood.so")
_flood_c = None


def _load_flood():
    """The C restatement of the flood (oracle/ws_flood.c, built by __graft_entry__.build()); None when not built."""
    global _flood_c
    if _flood_c is None and os.path.isfile(_FLOOD_SO):
        lib = ctypes.CDLL(_FLOOD_SO)
        lib.ws_flood.restype = ctypes.c_int
        lib.ws_flood.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        _flood_c = lib
    return _flood_c


# --------------------------------------------------------------------------------------------------
# scikit-image restatements
# --------------------------------------------------------------------------------------------------
def peak_local_max_mask(image, min_distance, exclude_border):
    """skimage.feature.peak_local_max(image, min_distance, exclude_border, indices=False), pre-0.18 semantics.
    exclude_border: True -> min_distance voxels on every axis, 0 / False -> none (watershed.py:42,92)."""
    image = np.asarray(image)
    out = np.zeros(image.shape, dtype=bool)
    if image.size == 0 or np.all(image == image.flat[0]):
        return out
    size = 2 * min_distance + 1
    image_max = ndi.maximum_filter(image, size=size, mode="constant")
    out = (image == image_max) & (image > image.min())
    border = min_distance if exclude_border is True else int(exclude_border)
    if border > 0:
        for ax in range(image.ndim):
            sl = [slice(None)] * image.ndim
            sl[ax] = slice(None, border)
            out[tuple(sl)] = False
            sl[ax] = slice(-border, None)
            out[tuple(sl)] = False
    return out


def label_full(mask):
    """skimage.morphology.label(mask): full connectivity, raster-order numbering."""
    lab, n = ndi.label(mask, structure=np.ones((3,) * mask.ndim, dtype=bool))
    return lab.astype(np.int64), n


def _neighbour_offsets(shape):
    """Raveled offsets of the connectivity-1 neighbours in C order of their coordinate offsets."""
    nd = len(shape)
    strides = [int(np.prod(shape[a + 1:])) for a in range(nd)]
    offs = []
    for ax in range(nd):
        offs.append((tuple(-1 if a == ax else 0 for a in range(nd)), -strides[ax], ax, -1))
    for ax in reversed(range(nd)):
        offs.append((tuple(1 if a == ax else 0 for a in range(nd)), strides[ax], ax, 1))
    offs.sort(key=lambda t: t[0])
    return [(o[1], o[2], o[3]) for o in offs]


def watershed(image, markers, mask, force_python=False):
    """skimage.segmentation.watershed(image, markers, mask=mask) (connectivity 1): priority flood.  Runs the C
    restatement (oracle/ws_flood.c, same order) when it has been built, else the pure-Python loop below."""
    lib = None if force_python else _load_flood()
    if lib is not None and np.asarray(image).ndim in (2, 3):
        img = np.ascontiguousarray(image, dtype=np.float64)
        msk = np.ascontiguousarray(mask, dtype=np.uint8)
        out = np.ascontiguousarray(np.asarray(markers, dtype=np.int64) * (msk != 0))
        shp = np.asarray(img.shape, dtype=np.int64)
        rc = lib.ws_flood(img.ctypes.data, out.ctypes.data, msk.ctypes.data, img.ndim, shp.ctypes.data)
        assert rc == 0
        return out
    image = np.asarray(image, dtype=np.float64)
    shape = image.shape
    mask_f = np.asarray(mask, dtype=bool).ravel()
    out = (np.asarray(markers, dtype=np.int64) * np.asarray(mask, dtype=bool)).ravel().copy()
    val = image.ravel()
    offs = _neighbour_offsets(shape)
    coords = np.unravel_index(np.arange(val.size), shape) if val.size < (1 << 24) else None
    heap = [(val[i], 0, int(i)) for i in np.flatnonzero(out)]
    heapq.heapify(heap)
    age = 0
    dims = shape
    while heap:
        _, _, idx = heapq.heappop(heap)
        if coords is not None:
            pos = [int(c[idx]) for c in coords]
        else:
            pos = list(np.unravel_index(idx, shape))
        lab = out[idx]
        for off, ax, step in offs:
            p = pos[ax] + step
            if p < 0 or p >= dims[ax]:
                continue
            nb = idx + off
            if not mask_f[nb] or out[nb]:
                continue
            age += 1
            out[nb] = lab
            heapq.heappush(heap, (val[nb], age, nb))
    return out.reshape(shape)


def _extreme_filter(a, footprint, take_max):
    """Grey dilation / erosion over a 3^n footprint with the neighbourhood clipped at the border (scikit-image pads
    by reflection, which for a 3^n window only repeats voxels already inside it).  Written with shifted views: SciPy's
    rank filters go through double and overflow on the int64 maximum used for the inverted background."""
    pad = np.pad(a, 1, mode="edge")
    out = a.copy()
    for off in np.argwhere(footprint):
        sl = tuple(slice(int(o), int(o) + n) for o, n in zip(off, a.shape))
        out = np.maximum(out, pad[sl]) if take_max else np.minimum(out, pad[sl])
    return out


def find_boundaries_outer(labels, connectivity):
    """skimage.segmentation.find_boundaries(labels, connectivity, mode='outer', background=0)."""
    labels = np.asarray(labels)
    nd = labels.ndim
    fp = ndi.generate_binary_structure(nd, connectivity)
    full = ndi.generate_binary_structure(nd, nd)
    big = np.iinfo(labels.dtype).max
    boundaries = _extreme_filter(labels, fp, True) != _extreme_filter(labels, fp, False)
    background = labels == 0
    inverted = labels.copy()
    inverted[background] = big
    adjacent = (_extreme_filter(labels, full, True) != _extreme_filter(inverted, full, False)) & ~background
    return boundaries & (background | adjacent)


def remove_small_objects(labels, min_size):
    """skimage.morphology.remove_small_objects on an integer label image: labels with fewer voxels are zeroed."""
    out = labels.copy()
    if min_size == 0:
        return out
    sizes = np.bincount(out.ravel())
    out[(sizes < min_size)[out]] = 0
    return out


def relabel_sequential(labels):
    """skimage.segmentation.relabel_sequential(labels)[0]: surviving labels -> 1..n, order preserved."""
    uniq = np.unique(labels)
    uniq = uniq[uniq != 0]
    fw = np.zeros(int(labels.max()) + 1, dtype=labels.dtype)
    fw[uniq] = np.arange(1, len(uniq) + 1)
    return fw[labels]


# --------------------------------------------------------------------------------------------------
# the reference's two functions and their caller
# --------------------------------------------------------------------------------------------------
def watershed_2d(image_pred, z_range, min_distance=7):
    """watershed.py:16-52."""
    boundary = np.zeros(image_pred.shape, dtype=bool)
    for z in range(z_range):
        bn_image = image_pred[:, :, z] > 0.5
        dist = ndi.distance_transform_edt(bn_image, sampling=[1, 1])
        dist_smooth = ndi.gaussian_filter(dist, 2, mode="constant")
        local_maxi = peak_local_max_mask(dist_smooth, min_distance, True)
        markers, _ = label_full(local_maxi)
        labels_ws = watershed(-dist_smooth, markers, bn_image)
        boundary[:, :, z] = find_boundaries_outer(labels_ws, 2)
    bn_output = image_pred > 0.5
    bn_output[boundary] = 0
    return bn_output, boundary


def watershed_3d(image_watershed2d, samplingrate, method, min_size, cell_num, min_distance):
    """watershed.py:55-108 up to `labels_clear` (the only output Tracker._watershed keeps)."""
    dist = ndi.distance_transform_edt(image_watershed2d, sampling=samplingrate)
    dist_smooth = ndi.gaussian_filter(dist, (2, 2, 0.3), mode="constant")
    local_maxi = peak_local_max_mask(dist_smooth, min_distance, 0)
    markers, _ = label_full(local_maxi)
    labels_ws = watershed(-dist_smooth, markers, image_watershed2d)
    counts = np.bincount(labels_ws.ravel())
    if method == "min_size":
        cell_num = int(np.sum(np.sort(counts) >= min_size) - 1)
    elif method == "cell_num":
        min_size = int(np.sort(counts)[-cell_num - 1])
    else:
        raise ValueError("The method parameter should be either min_size or cell_num")
    labels_clear = remove_small_objects(labels_ws, min_size)
    return labels_clear, min_size, cell_num


def segment(image_cell_bg_xyz, z_xy_ratio, method="min_size", min_size=0, cell_num=0):
    """Tracker._watershed (tracker.py:671-684) + the centre-of-mass lines of Tracker._segment (:646-648).
    Returns (segmentation_auto int (x,y,z), centres (n,3) float64 in voxel units, min_size, cell_num)."""
    z_siz = image_cell_bg_xyz.shape[2]
    wo_border, _ = watershed_2d(image_cell_bg_xyz, z_range=z_siz, min_distance=7)
    labels_clear, min_size, cell_num = watershed_3d(wo_border, [1, 1, z_xy_ratio], method, min_size, cell_num, 3)
    seg = relabel_sequential(labels_clear)
    n = int(seg.max())
    if n == 0:
        return seg, np.zeros((0, 3)), min_size, cell_num
    centres = np.asarray(ndi.center_of_mass(seg > 0, seg, range(1, n + 1)), dtype=np.float64)
    return seg, centres, min_size, cell_num
