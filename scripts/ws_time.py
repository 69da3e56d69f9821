"""Timing of the watershed stage on config 1 (512 x 512 x 35, 164 blobs, blob-detector U-Net weights)."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
m = {n: importlib.import_module("3deecelltracker_b200." + n) for n in ("watershed", "synth", "unet3d", "preprocess")}
shape, cells, ratio = (512, 512, 35), 164, 9.2
raw = m["synth"].blob_stack(shape, m["synth"].blob_centres(shape, cells, 1234), 1234, z_xy_ratio=ratio)
model = m["unet3d"].UNet3("a", weights=m["synth"].detector_unet_weights(0), tiles_per_batch=38)
norm = m["preprocess"].normalize_image_device(m["preprocess"]._raw_to_device(raw), 20)
prob = model.prediction_device(norm, (24, 24, 2))
for _ in range(3):
    seg = m["watershed"].segment_device(prob, ratio, "min_size", 40, 0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    seg = m["watershed"].segment_device(prob, ratio, "min_size", 40, 0)
e1.record(); torch.cuda.synchronize()
print("watershed stage: %.3f ms per volume, %s cells, fg voxels %d" % (e0.elapsed_time(e1) / 10, seg.host_scalars(), int((prob > 0.5).sum())))
