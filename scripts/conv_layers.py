"""Per-block timing of the 14 convolution blocks of unet3_a at the bench's batch size (38 tiles), per engine, from the
library's own CUDA-event profiler (tag 1 = convolution kernels only; layout conversion of the test harness excluded).
usage: conv_layers.py [tiles] [engine ...]"""
import ctypes as C, importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
u = importlib.import_module("3deecelltracker_b200.unet3d")
synth = importlib.import_module("3deecelltracker_b200.synth")
lib = importlib.import_module("3deecelltracker_b200._lib").lib()
tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 38
engines = sys.argv[2:] or ["tcgen05_classic", "tcgen05"]
model = u.UNet3("a", weights=synth.unet_weights("a", 0), tiles_per_batch=tiles)
layers = u._conv_layers(u._SPECS["a"])
names = ["d0a", "d0b", "d1a", "d1b", "d2a", "d2b", "u2a", "u2b", "u1a", "u1b", "u0a", "u0b", "o_m2", "o_m1"]
res = [160, 160, 80, 80, 40, 40, 20, 20, 40, 40, 80, 80, 160, 160]
gmac = [0.088, 1.416, 0.708, 1.416, 0.708, 1.416, 0.708, 0.708, 2.831, 0.708, 2.831, 0.708, 2.831, 0.708]
rng = np.random.default_rng(0)
tot = {e: 0.0 for e in engines}
print("| block | " + " | ".join(f"{e} ms (TF/s)" for e in engines) + " |")
for li, ((cin, cout), xy) in enumerate(zip(layers, res)):
    if li == 0:
        continue
    x = torch.from_numpy(rng.normal(0, 1, (tiles, xy, xy, 16, cin)).astype(np.float32)).cuda()
    cells = []
    for e in engines:
        for _ in range(2):
            model.conv_block_device(li, x, e)
        torch.cuda.synchronize()
        lib.ct_profile_enable(1)
        ms, cnt = C.c_double(), C.c_ulonglong()
        lib.ct_profile_read(1, C.byref(ms), C.byref(cnt), 1)
        for _ in range(5):
            model.conv_block_device(li, x, e)
        torch.cuda.synchronize()
        lib.ct_profile_read(1, C.byref(ms), C.byref(cnt), 1)
        lib.ct_profile_enable(0)
        t = ms.value / max(cnt.value, 1)
        tot[e] += t
        cells.append(f"{t:.3f} ({2 * gmac[li] * tiles / t:.0f})")
    print(f"| {names[li]} {cin}>{cout} @{xy} | " + " | ".join(cells) + " |")
    del x
print("| sum | " + " | ".join(f"{tot[e]:.3f}" for e in engines) + " |")
