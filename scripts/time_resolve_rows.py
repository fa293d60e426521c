"""Device time of the resolved-disparity step (matching kernel + dense_resolve_rows_kernel) against the matching kernel alone, on
textured frames (the bench workload) and on flat frames (every window of a row claims the same x': the resolve kernel's worst case)."""
import json, sys
import numpy as np
sys.path.insert(0, ".")
import torch
from unsynchronized_stereo_vision_proj325_b200 import _abi, api, synth

ctx = api.Context(0)
st = torch.cuda.current_stream().cuda_stream
n, W, H = 256, 640, 480
p = _abi.make_params(tmpl_w=16, tmpl_h=16, cost="sad")
f = _abi.FrameDesc(W, H, 1, W, W * H)
nx, ny, ev = api.grid_dims(f, p)
for name in ("textured", "flat"):
    left, right = synth.make_pairs(n, W, H, 1, shift=37, noise_sigma=2.0, seed=325)
    if name == "flat":
        left[:] = 90; right[:] = 90
    dl, dr = torch.from_numpy(np.ascontiguousarray(left)).cuda(), torch.from_numpy(np.ascontiguousarray(right)).cuda()
    o_r = torch.empty(n * nx * ny, dtype=torch.int16, device="cuda")
    o_d = torch.empty(n * nx * ny, dtype=torch.int16, device="cuda")
    o_c = torch.empty(n * nx * ny, dtype=torch.int16, device="cuda")
    res = {}
    for what in ("kernel_only", "with_resolve", "resolve_scratch_only"):
        out = _abi.Outputs()
        if what != "resolve_scratch_only":
            out.disparity_u16, out.raw_cost_u16 = o_d.data_ptr(), o_c.data_ptr()
        if what != "kernel_only":
            out.resolved_disparity_u16 = o_r.data_ptr()
        for _ in range(3):
            ctx.match_dense_device(dl.data_ptr(), dr.data_ptr(), f, n, p, out, st)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            ctx.match_dense_device(dl.data_ptr(), dr.data_ptr(), f, n, p, out, st)
        e1.record(); torch.cuda.synchronize()
        res[what] = e0.elapsed_time(e1) / 10
    surv = float((o_r.cpu().numpy().view(np.uint16) != 0xFFFF).mean())
    print(json.dumps({"frames": name, "pairs": n, "ms": res, "resolve_ms": res["with_resolve"] - res["kernel_only"], "survivor_fraction": surv}))
