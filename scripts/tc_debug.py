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
tiles = rng.normal(0, 1, (2, 160, 160, 16, 1)).astype(np.float32)
res = {}
for eng in ("direct", "tcgen05"):
    model.set_engine(eng)
    model.predict(tiles)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res[eng] = model.predict(tiles)
    torch.cuda.synchronize(); print(eng, "2 tiles", time.perf_counter() - t0, "s")
d = np.abs(res["direct"].astype(np.float64) - res["tcgen05"]) / np.maximum(np.abs(res["direct"]), 1e-30)
print("full net direct vs tc: max rel", d.max(), "q99.99", np.quantile(d, 0.9999))
