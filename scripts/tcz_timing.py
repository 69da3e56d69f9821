"""Wait-cycle breakdown of the plane-walk conv kernel (debug build with -DTZ_TIMING; block 0's role warps).
Run on the GPU box: builds csrc with TZ_TIMING into a scratch library, runs single conv blocks at U-Net sizes."""
import ctypes as C, importlib, os, subprocess, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
csrc = os.path.join(ROOT, "3deecelltracker_b200", "csrc")
extra = "-DTZ_TIMING " + " ".join(sys.argv[1:])            # e.g. -DTZ_EXP_NOSTORE / -DTZ_EXP_NOEPI (results are then wrong: timing only)
subprocess.run(["make", "-C", csrc, "-j", "8", "NVCCFLAGS_EXTRA=" + extra, "BUILD=build_timing", "TARGET=../libct3d_timing.so"], check=True,
               stdout=subprocess.DEVNULL)
print("build flags:", extra)
os.environ["CT3D_LIB"] = os.path.join(ROOT, "3deecelltracker_b200", "libct3d_timing.so")
u = importlib.import_module("3deecelltracker_b200.unet3d")
synth = importlib.import_module("3deecelltracker_b200.synth")
L = importlib.import_module("3deecelltracker_b200._lib")
lib = L.lib()
fn = C.CDLL(os.environ["CT3D_LIB"]).ct_debug_tcz_timers
model = u.UNet3("a", weights=synth.unet_weights("a", 0), tiles_per_batch=38)
layers = u._conv_layers(u._SPECS["a"])
sizes = {2: 80, 11: 80, 12: 160, 13: 160}
names = {2: "d1a", 11: "u0b", 12: "o_m2", 13: "o_m1"}
rng = np.random.default_rng(0)
for li, xy in sizes.items():
    cin, cout = layers[li]
    x = torch.from_numpy(rng.normal(0, 1, (38, xy, xy, 16, cin)).astype(np.float32)).cuda()
    for _ in range(2):
        model.conv_block_device(li, x, "planewalk_split")
    torch.cuda.synchronize()
    out = (C.c_ulonglong * 16)()
    fn(out)
    t = list(out)
    tot, dt = max(t[2], 1), max(t[8], 1)
    print(f"{names[li]:5s} {cin:3d}>{cout:2d} @{xy}: issuer total {t[2]:8d} clk | wait stage {t[0]/tot:5.1%} wait acc_empty {t[1]/tot:5.1%} | "
          f"producer wait empty {t[3]/tot:5.1%} | drain total {t[8]:8d} wait acc_full {t[5]/dt:5.1%} tmem ld {t[6]/dt:5.1%} finish {t[7]/dt:5.1%}")
