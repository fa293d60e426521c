"""Kernel-only timing of the C2 bench launch (256 pairs resident in HBM), one JSON line; dev aid."""
import json, os, sys
import numpy as np
sys.path.insert(0, ".")
import torch
from unsynchronized_stereo_vision_proj325_b200 import _abi, api, synth
ctx = api.Context(0)
st = torch.cuda.current_stream().cuda_stream
n, W, H = 256, 640, 480
kw = dict(tmpl_w=16, tmpl_h=16, cost="sad")
if len(sys.argv) > 1:
    kw["search_max"] = int(sys.argv[1])
p = _abi.make_params(**kw)
f = _abi.FrameDesc(W, H, 1, W, W * H)
nx, ny, ev = api.grid_dims(f, p)
left, right = synth.make_pairs(n, W, H, 1, shift=37, noise_sigma=2.0, seed=325)
dl, dr = torch.from_numpy(np.ascontiguousarray(left)).cuda(), torch.from_numpy(np.ascontiguousarray(right)).cuda()
o_d = torch.empty(n * nx * ny, dtype=torch.int16, device="cuda"); o_c = torch.empty(n * nx * ny, dtype=torch.int16, device="cuda")
out = _abi.Outputs(); out.disparity_u16, out.raw_cost_u16 = o_d.data_ptr(), o_c.data_ptr()
for _ in range(3):
    ctx.match_dense_device(dl.data_ptr(), dr.data_ptr(), f, n, p, out, st)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10):
    ctx.match_dense_device(dl.data_ptr(), dr.data_ptr(), f, n, p, out, st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
chk = int(o_d.cpu().numpy().view(np.uint16)[::997].astype(np.int64).sum()) ^ int(o_c.cpu().numpy().view(np.uint16)[::991].astype(np.int64).sum())
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("USV_DEV")}, "ms": ms, "pairs_per_s": n / ms * 1e3, "T_evals_per_s": n * ev / ms / 1e9, "checksum": chk}))
