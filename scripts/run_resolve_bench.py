"""ResolveMatchList over the per-window winners of a dense sweep, on the device.
Times usv_resolve_match_list_device (CUDA events) on the winners of one 640x480 pair (290 625 records) and, beside
it, the CPU restatement of the reference's O(M*T) loop on a bounded prefix of the same list."""
import ctypes as C
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from unsynchronized_stereo_vision_proj325_b200 import _abi, api, synth
from oracle import oracle

ctx = api.Context(0)
left, right = synth.make_pairs(1, 640, 480, 1, shift=37, noise_sigma=2.0, seed=325)
p = _abi.make_params(tmpl_w=16, tmpl_h=16, cost="sad")
win = ctx.match_dense(left, right, p, mask=_abi.OUT_MATCHES)["matches"][0]
n = len(win)
d_in = torch.from_numpy(win.view(np.uint8).reshape(n, 16).copy()).cuda()
d_out = torch.empty_like(d_in)
d_n = torch.zeros(1, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
L = api.lib()
def run():
    rc = L.usv_resolve_match_list_device(ctx._h, C.c_void_p(d_in.data_ptr()), C.c_int64(n), C.c_int32(1), C.c_void_p(d_out.data_ptr()),
                                         C.c_int64(n), C.c_void_p(d_n.data_ptr()), C.c_void_p(st))
    assert rc == 0, ctx.last_error
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
reps = 20
e0.record()
for _ in range(reps): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
n_out = int(d_n.item())
got = d_out.cpu().numpy().reshape(-1)[:n_out * 16].view(_abi.MATCH_DTYPE)
# CPU: the reference's loop restated (oracle), on a prefix it finishes in seconds
kept = win[win["RightIndex"] != _abi.NO_MATCH]
m_cpu = min(len(kept), 30000)
t0 = time.perf_counter(); exp = oracle.resolve_match_list(kept[:m_cpu]); t_cpu = time.perf_counter() - t0
same_prefix = ctx.resolve_match_list(kept[:m_cpu]).tobytes() == exp.tobytes()
print(json.dumps({"what": "ResolveMatchList over dense winners (P/Main.cpp:432-477)", "records_in": n, "records_kept": int(len(kept)),
                  "records_out": n_out, "gpu_ms": ms, "gpu_records_per_s": n / ms * 1e3,
                  "cpu_prefix_records": m_cpu, "cpu_prefix_s": t_cpu, "cpu_records_per_s_on_prefix": m_cpu / t_cpu,
                  "cpu_note": "O(M*T) loop: cost grows quadratically, the full list would take (n/prefix)^2 times longer",
                  "gpu_equals_cpu_on_prefix": bool(same_prefix)}))
