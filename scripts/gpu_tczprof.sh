cd $GRAFT_REPO_ROOT
cat > /tmp/one.py <<'PY'
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.environ["GRAFT_REPO_ROOT"])
u = importlib.import_module("3deecelltracker_b200.unet3d")
synth = importlib.import_module("3deecelltracker_b200.synth")
model = u.UNet3("a", weights=synth.unet_weights("a", 0), tiles_per_batch=38)
for li, xy in ((2, 80), (13, 160)):
    cin, cout = u._conv_layers(u._SPECS["a"])[li]
    x = torch.from_numpy(np.random.default_rng(0).normal(0, 1, (38, xy, xy, 16, cin)).astype(np.float32)).cuda()
    for _ in range(2):
        model.conv_block_device(li, x, "planewalk_split")
    torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:conv3_tcz" -s 1 -c 1 -o gpurun_out/prof_tcz_d1a python /tmp/one.py > gpurun_out/tczprof.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:conv3_tcz" -s 3 -c 1 -o gpurun_out/prof_tcz_om1 python /tmp/one.py >> gpurun_out/tczprof.log 2>&1
tail -2 gpurun_out/tczprof.log
nvidia-smi --query-gpu=name,temperature.gpu --format=csv
