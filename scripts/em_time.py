import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
tr = bench.mod("track"); ffn = bench.mod("ffn").FFN(bench.mod("synth").ffn_weights(0))
real0, real_t = bench.make_points()
ref = torch.from_numpy(real0).cuda(); tgt = torch.from_numpy(real_t).cuda()
corr = ffn.match_device(ref, tgt, 20)
for _ in range(2):
    p = tr.run_em([tr.EmProblem(ref, tgt, corr)], tr.MODE_TRACK, 300, 0.1, 20, 1e8, 0.5)[0]
    torch.cuda.synchronize()
