"""GPU parity of the spatially decomposed segmentation (config 3 path) on ONE device.

Every rank of the decomposition is played in turn by the same GPU: the rank's owned raw block is cut from the whole
stack, the all-reduce of the select histogram is an explicit sum over the ranks' states, and the halo boxes are cut
from the whole stack exactly as `exchange_halo` delivers them (that function itself is tested over gloo in
test_spatial_gloo.py and over NCCL by scripts/spatial_run.py).  Bar: the assembled probability volume is
BIT-IDENTICAL to the single-call `_normalize_image` + `unet3_prediction` result, which in turn is held to the
oracle elsewhere (test_gpu_lcn_unet.py)."""
import importlib

import numpy as np
import pytest
import torch

from conftest import load_pkg
from oracle import unet as ounet

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m():
    load_pkg()
    names = ("preprocess", "unet3d", "synth", "spatial", "_lib", "_device")
    return {n: importlib.import_module("3deecelltracker_b200." + n) for n in names}


def _emulated_median(m, blocks, total):
    """Lock-step select with the all-reduce replaced by an explicit sum of the ranks' histograms."""
    lib, L, sptr = m["_lib"].lib(), m["_lib"], m["_device"].stream_ptr
    dtype = m["preprocess"]._DTYPES[blocks[0].dtype]
    words = (lib.ct_select_state_bytes() + 3) // 4 + 64
    off = lib.ct_select_hist_offset() // 4
    states = [torch.zeros(words, dtype=torch.int32, device="cuda") for _ in blocks]
    for st in states:
        L.check(lib.ct_select_begin(st.data_ptr(), total, sptr()))
    for p in range(lib.ct_select_passes(dtype)):
        for st, b in zip(states, blocks):
            L.check(lib.ct_select_hist(b.data_ptr(), dtype, b.numel(), st.data_ptr(), p, sptr()))
        summed = torch.stack([st[off:off + 512] for st in states]).sum(0).to(torch.int32)
        for st in states:
            st[off:off + 512] = summed
            L.check(lib.ct_select_scan(st.data_ptr(), dtype, p, sptr()))
    meds = []
    for st in states:
        med = torch.empty(1, dtype=torch.float64, device="cuda")
        L.check(lib.ct_select_finish(st.data_ptr(), dtype, med.data_ptr(), sptr()))
        meds.append(med)
    return meds


@pytest.mark.parametrize("shape,grid,dtype", [((200, 150, 16), (2, 1, 1), np.uint16), ((230, 240, 30), (2, 2, 2), np.uint16),
                                              ((120, 130, 14), (1, 2, 2), np.float32)])
def test_decomposed_segmentation_is_bit_identical(m, shape, grid, dtype):
    pre, u, synth, sp = m["preprocess"], m["unet3d"], m["synth"], m["spatial"]
    raw = synth.blob_stack(shape, synth.blob_centres(shape, 30, 7), 7)
    if dtype == np.float32:
        raw = raw.astype(np.float32) * 0.37 - 11.0
    model = u.UNet3("a", weights=ounet.random_weights("a", seed=2), tiles_per_batch=4)
    raw_dev = pre._raw_to_device(raw)
    norm_full = pre.normalize_image_device(raw_dev, 20)
    want = model.prediction_device(norm_full, (24, 24, 2))

    plan = sp.SpatialPlan(shape, grid)
    owned = []
    for r in range(plan.world):
        lo, hi = plan.owned_box(r)
        owned.append(raw_dev[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]].contiguous())
    meds = _emulated_median(m, owned, raw_dev.numel())
    assert all(float(md[0]) == float(np.median(raw)) for md in meds)

    got = torch.full(shape, -1.0, dtype=torch.float32, device="cuda")
    for r in range(plan.world):
        if not plan.has_tiles(r):
            continue
        need, out = plan.raw_box(r), plan.out_box(r)
        ext = raw_dev[need[0][0]:need[1][0], need[0][1]:need[1][1], need[0][2]:need[1][2]].contiguous()
        norm = pre.normalize_block_device(ext, 20, meds[r])
        nb = plan.norm_box(r)                       # where the block LCN must equal the whole-volume LCN
        sl = tuple(slice(l, h) for l, h in zip(*nb))
        sl_loc = tuple(slice(l - o, h - o) for l, h, o in zip(nb[0], nb[1], need[0]))
        assert torch.equal(norm[sl_loc], norm_full[sl])
        tlo, thi = plan.tile_box(r)
        prob = model.prediction_block_device(norm, need[0], shape, (24, 24, 2), tlo, thi, out[0],
                                             tuple(h - l for l, h in zip(*out)))
        got[out[0][0]:out[1][0], out[0][1]:out[1][1], out[0][2]:out[1][2]] = prob
    assert torch.equal(got, want)


def test_block_prediction_rejects_a_box_that_misses_tile_reads(m):
    u = m["unet3d"]
    model = u.UNet3("a", weights=ounet.random_weights("a", seed=2), tiles_per_batch=2)
    block = torch.zeros((100, 100, 16), dtype=torch.float32, device="cuda")
    with pytest.raises(m["_lib"].Ct3dError, match="outside the input box"):
        model.prediction_block_device(block, (0, 0, 0), (300, 300, 16), (24, 24, 2), (0, 0, 0), (1, 1, 1), (0, 0, 0),
                                      (112, 112, 12))
    with pytest.raises(m["_lib"].Ct3dError, match="tile range"):
        model.prediction_block_device(block, (0, 0, 0), (100, 100, 16), (24, 24, 2), (0, 0, 0), (2, 1, 1), (0, 0, 0),
                                      (100, 100, 12))
