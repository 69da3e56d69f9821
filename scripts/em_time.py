"""PR-GLS EM timings: microseconds per EM iteration and per pr_gls_quick call, single problems and
ensemble batches (one persistent CTA per problem), CUDA events on the launching stream.  Run on the GPU box."""
import importlib, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
tr, synth = bench.mod("track"), bench.mod("synth")
ffn = bench.mod("ffn").FFN(synth.ffn_weights(0))


def case(n, batch, iters, beta, lam, reps=5):
    probs = []
    for b in range(batch):
        ref = synth.random_points(n, 10 + b)
        tgt = synth.move_points(ref, 100 + b)
        r, t = torch.from_numpy(ref).cuda(), torch.from_numpy(tgt).cuda()
        probs.append((r, t, ffn.match_device(r, t, 20)))
    def run():
        return tr.run_em([tr.EmProblem(r, t, c) for r, t, c in probs], tr.MODE_TRACK, beta, lam, iters + 1, 1e8, 0.5)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"n": n, "batch": batch, "iterations": iters, "ms_per_launch": round(ms, 3),
            "us_per_iteration": round(ms * 1e3 / iters, 1), "us_per_iteration_per_problem": round(ms * 1e3 / iters / batch, 2)}


rows = [case(164, 1, 19, 300, 0.1), case(164, 20, 19, 300, 0.1), case(113, 1, 9, 1000, 1e-5), case(113, 20, 9, 1000, 1e-5),
        case(164, 148, 19, 300, 0.1), case(64, 1, 19, 300, 0.1), case(300, 1, 19, 300, 0.1), case(512, 1, 4, 300, 0.1, reps=2),
        case(1024, 1, 4, 300, 0.1, reps=2), case(2048, 1, 4, 300, 0.1, reps=2), case(4096, 1, 2, 300, 0.1, reps=1)]
hbm = bench.peaks()["hbm"]
for r in rows:
    n = r["n"]
    m = int(n * 1.0)
    # algorithmic bytes per iteration (fp64) = 8 (M N [prior] + 2 N^2 [G for a, G for C G]) + 24 (N + M); LU 2/3 N^3 + 6 N^2 flop
    r["algorithmic_mb_per_iteration"] = round((8 * (m * n + 2 * n * n) + 24 * (n + m)) / 1e6, 2)
    r["hbm_fraction"] = round(r["algorithmic_mb_per_iteration"] * 1e6 / (r["us_per_iteration"] * 1e-6) / 1e9 / hbm / r["batch"] * r["batch"], 4)
    r["lu_gflop"] = round((2 / 3 * n ** 3 + 6 * n * n) / 1e9, 3)
    r["lu_tflops_if_all_time_were_lu"] = round(r["lu_gflop"] / (r["us_per_iteration"] * 1e-6) / 1e3, 3)
    print(json.dumps(r))
