"""Device-resident throughput of the BASELINE.json configurations other than the bench line (C2), one JSON line each:
C1 single template (latency), C2 with D = 128, C3 (1280x720 colour, 32x32, D = 256, ZNCC), C4 (1920x1080, 16x16 SAD, D = 256).
Inputs resident in HBM, CUDA events on the launching stream, 3 warm-up launches.

  python scripts/run_configs.py [--c3-pairs 1] [--c4-pairs 16]"""
import argparse, json, sys
import numpy as np
sys.path.insert(0, ".")
import torch
from unsynchronized_stereo_vision_proj325_b200 import _abi, api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--c3-pairs", type=int, default=1)
ap.add_argument("--c4-pairs", type=int, default=16)
ap.add_argument("--only", default="")
ap.add_argument("--corr-kernel", default="auto", help="auto | alu | mma | tcgen05 (Context.corr_kernel)")
a = ap.parse_args()
ctx = api.Context(0)
ctx.corr_kernel(a.corr_kernel)
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    if e0.elapsed_time(e1) > 1500.0:  # a launch of seconds: one (cold) sample is the measurement
        return e0.elapsed_time(e1)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def dense(name, n, w, h, c, reps, **kw):
    if a.only and a.only not in name:
        return
    left, right = synth.make_pairs(n, w, h, c, shift=37, noise_sigma=2.0, seed=325)
    dl, dr = torch.from_numpy(np.ascontiguousarray(left)).cuda(), torch.from_numpy(np.ascontiguousarray(right)).cuda()
    f = _abi.frame_desc_for(left)
    p = _abi.make_params(**kw)
    nx, ny, ev = api.grid_dims(f, p)
    o_d = torch.empty(n * nx * ny, dtype=torch.int16, device="cuda")
    o_f = torch.empty(n * nx * ny, dtype=torch.float32, device="cuda")
    out = _abi.Outputs(); out.disparity_u16 = o_d.data_ptr(); out.distance_f32 = o_f.data_ptr()
    ms = timed(lambda: ctx.match_dense_device(dl.data_ptr(), dr.data_ptr(), f, n, p, out, st), reps)
    d = o_d.cpu().numpy().view(np.uint16)
    print(json.dumps({"config": name, "pairs": n, "frame": [w, h, c], "params": kw, "windows_per_pair": nx * ny, "cand_evals_per_pair": ev,
                      "ms_per_launch": ms, "pairs_per_s": n / ms * 1e3, "cand_evals_per_s": n * ev / ms * 1e3, "kernel": ctx.last_kernel,
                      "mode_disparity": int(np.bincount(d[d != 0xFFFF]).argmax())}), flush=True)


# C1: one 16x16 template at (300, 200) of one 640x480 pair, full-row search (625 candidates): launch latency
if not a.only or "C1" in a.only:
    left, right = synth.make_pairs(1, 640, 480, 1, shift=37, noise_sigma=2.0, seed=325)
    dl, dr = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
    f = _abi.frame_desc_for(left)
    p = _abi.make_params(tmpl_w=16, tmpl_h=16, cost="sad")
    tx, ty = torch.tensor([300], dtype=torch.int32).cuda(), torch.tensor([200], dtype=torch.int32).cuda()
    o_m = torch.empty(16, dtype=torch.uint8, device="cuda"); o_dd = torch.empty(1, dtype=torch.float64, device="cuda")
    out = _abi.Outputs(); out.matches = o_m.data_ptr(); out.distance = o_dd.data_ptr()
    ms = timed(lambda: ctx.match_templates_device(dl.data_ptr(), dr.data_ptr(), f, 1, tx.data_ptr(), ty.data_ptr(), 1, p, out, st), 200)
    m = o_m.cpu().numpy().view(_abi.MATCH_DTYPE)[0]
    print(json.dumps({"config": "C1 single template", "us_per_launch": ms * 1e3, "cand_evals": 301, "kernel": ctx.last_kernel,
                      "right_index": int(m["RightIndex"]), "expected_right_index": 200 * 625 + 300 - 37, "distance_cm": float(o_dd.item())}), flush=True)

dense("C2 D=128", 256, 640, 480, 1, 5, tmpl_w=16, tmpl_h=16, cost="sad", search_max=127)
dense("C4 1920x1080 D=256", a.c4_pairs, 1920, 1080, 1, 3, tmpl_w=16, tmpl_h=16, cost="sad", search_max=255)
dense("C3geom 1280x720 gray 32x32 SAD D=256", 16, 1280, 720, 1, 3, tmpl_w=32, tmpl_h=32, cost="sad", search_max=255)
dense("C3 1280x720 colour 32x32 ZNCC D=256", a.c3_pairs, 1280, 720, 3, 1, tmpl_w=32, tmpl_h=32, cost="zncc", search_max=255)
dense("C2 geometry, ssd gray 16x16 full range", 64, 640, 480, 1, 3, tmpl_w=16, tmpl_h=16, cost="ssd")
dense("C2 geometry, ncc gray 16x16 full range", 64, 640, 480, 1, 3, tmpl_w=16, tmpl_h=16, cost="ncc")
dense("C2 geometry, zncc gray 16x16 full range", 64, 640, 480, 1, 3, tmpl_w=16, tmpl_h=16, cost="zncc")
dense("C3geom 1280x720 colour 32x32 NCC D=256", 16, 1280, 720, 3, 1, tmpl_w=32, tmpl_h=32, cost="ncc", search_max=255)
dense("C3geom 1280x720 colour 32x32 SSD D=256", 16, 1280, 720, 3, 1, tmpl_w=32, tmpl_h=32, cost="ssd", search_max=255)
dense("C3geom 1280x720 colour 32x32 SAD-colour D=256", 32, 1280, 720, 3, 3, tmpl_w=32, tmpl_h=32, cost="sad", search_max=255)
