"""Plane-walk conv kernel (unet_tcz.cu) bring-up: every block it takes, through ct_unet_conv_block (engines planewalk_split /
planewalk_split_src) against fp64 torch, plus the whole network (engine auto vs tcgen05).  Prints errors, asserts nothing."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import unet as ounet
u = importlib.import_module("3deecelltracker_b200.unet3d")
from test_gpu_lcn_unet import _block_reference

ws = ounet.random_weights("a", seed=3)
model = u.UNet3("a", weights=ws, tiles_per_batch=2)
layers = u._conv_layers(u._SPECS["a"])
only = [int(a) for a in sys.argv[1:]] or range(1, 14)
for layer in only:
    cin, cout = layers[layer]
    if cin % 8 or cout > 32:
        continue
    rng = np.random.default_rng(300 + layer)
    for b, (x, y, z), mag in [(1, (8, 16, 16), 1.0), (2, (11, 21, 16), 3e4), (1, (20, 40, 16), 2e-5), (1, (37, 24, 16), 1.0)]:
        xin = (rng.normal(0, 1, (b, x, y, z, cin)) * mag).astype(np.float32)
        ref = _block_reference(ws, layer, xin)
        scale = np.abs(ref).max()
        dev = torch.from_numpy(xin).cuda()
        for engine in ("planewalk_split", "planewalk_split_src", "tcgen05_split"):
            try:
                got = model.conv_block_device(layer, dev, engine).cpu().numpy().astype(np.float64)
                torch.cuda.synchronize()
            except Exception as e:
                print(f"layer {layer} {cin}->{cout} {engine} {x}x{y}x{z}: EXC {e}")
                raise
            d = np.abs(got - ref) / scale
            msg = f"layer {layer} {cin}->{cout} {engine:15s} {b}x{x}x{y}x{z} x{mag:g}: max {d.max():.2e}"
            if d.max() > 2e-5:
                bad = np.argwhere(d > 2e-5)
                msg += f"  BAD {len(bad)}/{d.size} first {bad[0].tolist()} per-x {np.unique(bad[:,1]).tolist()[:12]} per-y {np.unique(bad[:,2]).tolist()[:12]} per-z {np.unique(bad[:,3]).tolist()} per-c {np.unique(bad[:,4]).tolist()}"
            print(msg, flush=True)
if len(sys.argv) > 1:
    sys.exit(0)
# whole network
from test_gpu_lcn_unet import assert_prob_close
model = u.UNet3("a", weights=ws, tiles_per_batch=4)
rng = np.random.default_rng(5)
tiles = rng.normal(0, 1, (3, 160, 160, 16)).astype(np.float32)
dev = torch.from_numpy(tiles).cuda()
outs = {}
for eng in ("tcgen05", "auto"):
    model.set_engine(eng)
    outs[eng] = model.predict_device(dev).cpu().numpy()
    torch.cuda.synchronize()
d = np.abs(outs["auto"] - outs["tcgen05"])
print("network auto vs tcgen05: max abs diff", d.max(), "max rel", (d / np.maximum(outs["tcgen05"], 1e-12)).max())
want = ounet.UNetOracle("a", ws)
w32 = want.predict(tiles[..., None]) if hasattr(want, "predict") else None
if w32 is not None:
    w32 = np.asarray(w32).reshape(outs["auto"].shape)
    print("auto vs oracle fp32: max rel", (np.abs(outs["auto"] - w32) / np.maximum(w32, 1e-12)).max())
# timing: 38 tiles, engine auto vs tcgen05
synth = importlib.import_module("3deecelltracker_b200.synth")
for eng in ("tcgen05", "auto"):
    m = u.UNet3("a", weights=synth.unet_weights("a", 0), tiles_per_batch=38, engine=eng)
    t38 = torch.from_numpy(np.random.default_rng(0).normal(0, 1, (38, 160, 160, 16)).astype(np.float32)).cuda()
    for _ in range(2):
        m.predict_device(t38)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m.predict_device(t38)
    e1.record(); torch.cuda.synchronize()
    print(f"engine {eng}: {e0.elapsed_time(e1) / 5:.3f} ms per 38-tile batch", flush=True)
