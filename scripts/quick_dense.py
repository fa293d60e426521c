"""Dev helper: dense-kernel parity on small cases + a quick device-resident timing of C2."""
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, ".")
import torch
from unsynchronized_stereo_vision_proj325_b200 import _abi, api, synth
from oracle import oracle

ctx = api.Context(0)
ok = True
for (w, h, shift, kw) in [
    (640, 40, 37, dict(tmpl_w=16, tmpl_h=16, cost="sad")),
    (640, 36, 37, dict(tmpl_w=16, tmpl_h=16, cost="sad", search_max=127)),
    (300, 50, 11, dict(tmpl_w=32, tmpl_h=32, cost="sad", search_max=100)),
    (200, 30, -15, dict(tmpl_w=16, tmpl_h=16, cost="sad", camera_side=_abi.RIGHT_CAM, search_max=64)),
    (600, 24, 3, dict(tmpl_w=16, tmpl_h=16, cost="sad", search_min=-8, search_max=8)),
    (131, 29, 9, dict(tmpl_w=16, tmpl_h=5, cost="sad", search_min=1, search_max=33)),
    (640, 480, 37, dict(tmpl_w=16, tmpl_h=16, cost="sad", search_max=63)),
]:
    left, right = synth.make_pairs(2, w, h, 1, shift=shift, noise_sigma=2.0, seed=w + h)
    p = _abi.make_params(**kw)
    t0 = time.time(); got = ctx.match_dense(left, right, p); t1 = time.time()
    exp = oracle.match_dense(left, right, p)
    same = all(np.array_equal(got[k], exp[k]) for k in ("right_index", "raw_cost", "disparity_u16"))
    same &= got["matches"].tobytes() == exp["matches"].tobytes()
    print(w, h, kw, ctx.last_kernel, "OK" if same else "MISMATCH", "%.1f ms" % ((t1 - t0) * 1e3), flush=True)
    if not same:
        ok = False
        bad = np.argwhere(got["raw_cost"] != exp["raw_cost"])
        print("  first mismatches (pair, window):", bad[:5].tolist(), "n_bad", len(bad))
        for b in bad[:5]:
            print("   got", got["raw_cost"][tuple(b)], got["right_index"][tuple(b)], "exp", exp["raw_cost"][tuple(b)], exp["right_index"][tuple(b)])

# timing: C2, device resident
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
left, right = synth.make_pairs(n, 640, 480, 1, shift=37, noise_sigma=2.0)
dl = torch.from_numpy(np.ascontiguousarray(left)).cuda(); dr = torch.from_numpy(np.ascontiguousarray(right)).cuda()
f = _abi.FrameDesc(640, 480, 1, 640, 640 * 480)
for kw in (dict(), dict(search_max=127)):
    p = _abi.make_params(tmpl_w=16, tmpl_h=16, cost="sad", **kw)
    nx, ny, ev = api.grid_dims(f, p)
    o_ri = torch.empty(n * nx * ny, dtype=torch.int32, device="cuda"); o_rc = torch.empty_like(o_ri)
    o_d = torch.empty(n * nx * ny, dtype=torch.float32, device="cuda")
    out = _abi.Outputs(); out.right_index = o_ri.data_ptr(); out.raw_cost = o_rc.data_ptr(); out.distance_f32 = o_d.data_ptr()
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2): ctx.match_dense_device(dl.data_ptr(), dr.data_ptr(), f, n, p, out, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); reps = 3
    for _ in range(reps): ctx.match_dense_device(dl.data_ptr(), dr.data_ptr(), f, n, p, out, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("C2%s: %d pairs %.2f ms -> %.0f pairs/s, %.3f T cand-evals/s (%s)" % (kw, n, ms, n / ms * 1e3, n * ev / ms / 1e9, ctx.last_kernel), flush=True)
print("ALL OK" if ok else "FAILURES")
