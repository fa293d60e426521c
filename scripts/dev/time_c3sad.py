"""Kernel-only timing of colour SAD on the C3 geometry (1280x720x3, 32x32, D = 256), 32 pairs resident; dev aid."""
import json, os, sys
import numpy as np
sys.path.insert(0, ".")
import torch
from unsynchronized_stereo_vision_proj325_b200 import _abi, api, synth
ctx = api.Context(0)
st = torch.cuda.current_stream().cuda_stream
n, W, H = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 1280, 720
p = _abi.make_params(tmpl_w=32, tmpl_h=32, cost="sad", search_max=255)
f = _abi.FrameDesc(W, H, 3, W * 3, W * H * 3)
nx, ny, ev = api.grid_dims(f, p)
left, right = synth.make_pairs(4, W, H, 3, shift=60, noise_sigma=3.0, seed=33)
left, right = np.concatenate([left] * (n // 4)), np.concatenate([right] * (n // 4))
dl, dr = torch.from_numpy(np.ascontiguousarray(left)).cuda(), torch.from_numpy(np.ascontiguousarray(right)).cuda()
o_d = torch.empty(n * nx * ny, dtype=torch.int16, device="cuda"); o_c = torch.empty(n * nx * ny, dtype=torch.int32, device="cuda")
out = _abi.Outputs(); out.disparity_u16, out.raw_cost = o_d.data_ptr(), o_c.data_ptr()
for _ in range(2):
    ctx.match_dense_device(dl.data_ptr(), dr.data_ptr(), f, n, p, out, st)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(5):
    ctx.match_dense_device(dl.data_ptr(), dr.data_ptr(), f, n, p, out, st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
chk = int(o_d.cpu().numpy().view(np.uint16)[::997].astype(np.int64).sum()) ^ int(o_c.cpu().numpy()[::991].astype(np.int64).sum())
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("USV_DEV")}, "pairs": n, "ms": ms, "pairs_per_s": n / ms * 1e3, "T_evals_per_s": n * ev / ms / 1e9, "checksum": chk, "kernel": ctx.last_kernel}))
