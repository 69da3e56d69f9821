"""GPU debug: tcgen05 conv block vs CUDA-core conv block vs torch fp64, layer by layer (run under gpurun)."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
u = importlib.import_module("3deecelltracker_b200.unet3d")
from oracle import unet as ounet

ws = ounet.random_weights("a", seed=3)
model = u.UNet3("a", weights=ws, tiles_per_batch=2)
layers = u._conv_layers(u._SPECS["a"])
rng = np.random.default_rng(0)
cases = [(1, (8, 16, 8)), (1, (12, 40, 16)), (2, (20, 20, 16))]
only = [int(a) for a in sys.argv[1:]] or range(len(layers))
for li in only:
    cin, cout = layers[li]
    for b, (x, y, z) in cases:
        xin = torch.from_numpy(rng.normal(0, 1, (b, x, y, z, cin)).astype(np.float32)).cuda()
        w = torch.from_numpy(ws[6 * li]).double().permute(4, 3, 0, 1, 2)
        bias, gamma, beta, mean, var = (torch.from_numpy(ws[6 * li + k]).double() for k in range(1, 6))
        ref = torch.nn.functional.conv3d(xin.cpu().double().permute(0, 4, 1, 2, 3), w, bias, padding=1)
        ref = torch.nn.functional.leaky_relu(ref, 0.3)
        ref = (ref - mean[None, :, None, None, None]) / torch.sqrt(var[None, :, None, None, None] + 1e-3) * gamma[None, :, None, None, None] + beta[None, :, None, None, None]
        ref = ref.permute(0, 2, 3, 4, 1).numpy()
        out = {}
        for eng in ("direct", "tcgen05"):
            try:
                o = model.conv_block_device(li, xin, eng)
                torch.cuda.synchronize()
                out[eng] = o.cpu().numpy().astype(np.float64)
            except Exception as e:
                print(f"layer {li} ({cin}->{cout}) {eng} FAILED: {e}")
                raise
        sc = np.abs(ref).max()
        print(f"layer {li:2d} {cin:3d}->{cout:2d} b={b} {x}x{y}x{z}: direct err {np.abs(out['direct']-ref).max()/sc:.2e}  "
              f"tc err {np.abs(out['tcgen05']-ref).max()/sc:.2e}", flush=True)
print("block checks done")
# whole network, both engines
ntile = 15
model = u.UNet3("a", weights=ws, tiles_per_batch=ntile)
tiles = torch.from_numpy(rng.normal(0, 1, (ntile, 160, 160, 16)).astype(np.float32)).cuda()
res = {}
for eng in ("direct", "tcgen05"):
    model.set_engine(eng)
    for _ in range(2):
        res[eng] = model.predict_device(tiles)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        model.predict_device(tiles)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{eng}: {ntile} tiles {ms:.3f} ms -> {ntile * 35.573e9 / ms / 1e9:.1f} TFLOP/s, {ms / ntile * 75:.2f} ms per 75-tile volume")
a_, b_ = res["direct"].double().cpu().numpy(), res["tcgen05"].double().cpu().numpy()
d = np.abs(a_ - b_) / np.maximum(np.abs(a_), 1e-30)
print("full net direct vs tc: max rel", d.max(), "q99.99", np.quantile(d, 0.9999))
want64 = ounet.UNetOracle("a", ws, dtype=torch.float64)
with torch.no_grad():
    y64 = want64.forward(tiles[:1].cpu().double()[:, None]).numpy()[0, 0]
for eng in ("direct", "tcgen05"):
    r = np.abs(res[eng][0].double().cpu().numpy() - y64) / np.maximum(np.abs(y64), 1e-30)
    print(eng, "vs fp64 oracle: max rel", r.max(), "q99.99", np.quantile(r, 0.9999))
