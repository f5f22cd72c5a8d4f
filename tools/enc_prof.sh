python tools/encode_cloud_profile.py 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_encoder_v10.csv python - <<'PY' > gpurun_out/enc_ncu.log 2>&1
import sys, torch, numpy as np
sys.path.insert(0, '.')
import ppsurf_b200
from ppsurf_b200 import synthetic
dev = torch.device('cuda:0')
model = ppsurf_b200.PPSurfModel(256, ['imp_surf_sign'], 3, 2, 64, 0.0, False, 'bench', 'results', 0.05, 'p', 256, 10, 10000, 129, 50, 50000, 10, 8)
net = model.network
net.load_state_dict(synthetic.make_state_dict(net, 42))
model = model.to(dev)
pts = torch.from_numpy(synthetic.synthetic_cloud(20000, 42).T[None].copy()).to(dev)
net.sampling_seed = 42
model.encode_cloud(pts, generator=torch.Generator().manual_seed(42), batch_passes=16)
torch.cuda.synchronize()
PY
tail -2 gpurun_out/enc_ncu.log
