"""Dev helper for ncu: C3 (1280x720 colour, 32x32 ZNCC, D = 256) on the sliding correlation kernel, 4 pairs."""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from unsynchronized_stereo_vision_proj325_b200 import _abi, api, synth
n = 4
ctx = api.Context(0)
ctx.corr_kernel(sys.argv[1] if len(sys.argv) > 1 else "auto")
left, right = synth.make_pairs(n, 1280, 720, 3, shift=37, noise_sigma=2.0)
dl, dr = torch.from_numpy(np.ascontiguousarray(left)).cuda(), torch.from_numpy(np.ascontiguousarray(right)).cuda()
f = _abi.frame_desc_for(left)
p = _abi.make_params(tmpl_w=32, tmpl_h=32, cost="zncc", search_max=255)
nx, ny, ev = api.grid_dims(f, p)
o = torch.empty(n * nx * ny, dtype=torch.int16, device="cuda")
out = _abi.Outputs(); out.disparity_u16 = o.data_ptr()
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    ctx.match_dense_device(dl.data_ptr(), dr.data_ptr(), f, n, p, out, st)
torch.cuda.synchronize()
print(ctx.last_kernel)
