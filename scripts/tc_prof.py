"""One U-Net batch (15 tiles) per engine, for ncu launch lists / captures (run under gpurun + ncu)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
u = importlib.import_module("3deecelltracker_b200.unet3d")
synth = importlib.import_module("3deecelltracker_b200.synth")
eng = sys.argv[1] if len(sys.argv) > 1 else "tcgen05"
ntile = int(sys.argv[2]) if len(sys.argv) > 2 else 15
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
model = u.UNet3("a", weights=synth.unet_weights("a", 0), tiles_per_batch=ntile, engine=eng)
tiles = torch.from_numpy(np.random.default_rng(0).normal(0, 1, (ntile, 160, 160, 16)).astype(np.float32)).cuda()
for _ in range(reps):
    model.predict_device(tiles)
torch.cuda.synchronize()
