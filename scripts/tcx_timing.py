"""Wait-cycle breakdown of the x-stacked conv kernel (debug build with -DTX_TIMING; block 0's role warps).
Run on the GPU box: builds csrc with TX_TIMING into a scratch library, runs single conv blocks at U-Net sizes."""
import ctypes as C, importlib, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
csrc = os.path.join(ROOT, "3deecelltracker_b200", "csrc")
subprocess.run(["make", "-C", csrc, "-j", "8", "NVCCFLAGS_EXTRA=-DTX_TIMING", "BUILD=build_timing", "TARGET=../libct3d.so"], check=True,
               stdout=subprocess.DEVNULL)
u = importlib.import_module("3deecelltracker_b200.unet3d")
synth = importlib.import_module("3deecelltracker_b200.synth")
L = importlib.import_module("3deecelltracker_b200._lib")
lib = L.lib()
fn = C.CDLL(L.LIB_PATH).ct_debug_tcx_timers
model = u.UNet3("a", weights=synth.unet_weights("a", 0), tiles_per_batch=38)
layers = u._conv_layers(u._SPECS["a"])
sizes = {1: 160, 2: 80, 3: 80, 4: 40, 9: 40, 10: 80, 11: 80, 12: 160, 13: 160}
names = {1: "d0b", 2: "d1a", 3: "d1b", 4: "d2a", 9: "u1b", 10: "u0a", 11: "u0b", 12: "o_m2", 13: "o_m1"}
rng = np.random.default_rng(0)
for li, xy in sizes.items():
    cin, cout = layers[li]
    x = torch.from_numpy(rng.normal(0, 1, (38, xy, xy, 16, cin)).astype(np.float32)).cuda()
    for _ in range(2):
        model.conv_block_device(li, x, "tcgen05")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = (C.c_ulonglong * 16)()
    fn(out)
    t = list(out)
    tot = max(t[2], 1)
    print(f"{names[li]:5s} {cin:3d}>{cout:2d} @{xy}: issuer total {t[2]:8d} clk | wait conv {t[0]/tot:5.1%} wait acc_empty {t[1]/tot:5.1%} | "
          f"converter wait full {t[3]/max(t[4],1):5.1%} | drain wait acc_full {t[5]/max(t[7],1):5.1%} epilogue {t[6]/max(t[7],1):5.1%}")
