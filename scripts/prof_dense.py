"""Dev helper for ncu: one dense launch configuration, few launches."""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from unsynchronized_stereo_vision_proj325_b200 import _abi, api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dmax = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 20)
ctx = api.Context(0)
left, right = synth.make_pairs(n, 640, 480, 1, shift=37, noise_sigma=2.0)
dl = torch.from_numpy(np.ascontiguousarray(left)).cuda(); dr = torch.from_numpy(np.ascontiguousarray(right)).cuda()
f = _abi.FrameDesc(640, 480, 1, 640, 640 * 480)
p = _abi.make_params(tmpl_w=16, tmpl_h=16, cost="sad", search_max=dmax)
nx, ny, ev = api.grid_dims(f, p)
o_ri = torch.empty(n * nx * ny, dtype=torch.int32, device="cuda"); o_rc = torch.empty_like(o_ri)
out = _abi.Outputs(); out.right_index = o_ri.data_ptr(); out.raw_cost = o_rc.data_ptr()
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    ctx.match_dense_device(dl.data_ptr(), dr.data_ptr(), f, n, p, out, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); ctx.match_dense_device(dl.data_ptr(), dr.data_ptr(), f, n, p, out, st); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("pairs %d dmax %d: %.3f ms, %.3f T cand-evals/s" % (n, dmax, ms, n * ev / ms / 1e9))
