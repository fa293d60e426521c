"""Randomised parity stress of the sliding-window kernel against the CPU oracle: random frame sizes (incl. widths that are
not multiples of 4 and frames narrower than one x-tile), one and three channels, template sizes, camera sides, disparity
ranges (negative, empty, beyond the frame), accept thresholds and batch sizes; every fourth case also checks the resolved
disparity map (the device-side ResolveMatchList) against the oracle's restatement of P/Main.cpp:432-477.
python scripts/stress_dense.py [n_cases] [seed]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from unsynchronized_stereo_vision_proj325_b200 import _abi, api, synth
from oracle import oracle

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ctx = api.Context(0)
bad = dense = resolved = 0
for t in range(n_cases):
    tw = int(rng.choice([8, 12, 16, 24, 32]))
    th = int(rng.choice([1, 3, 8, 15, 16, 17, 24, 32, 40, 64]))
    w = int(rng.integers(tw, 420))
    h = int(rng.integers(th, th + 70))
    n = int(rng.integers(1, 4))
    side = int(rng.integers(0, 2))
    lo = 0 if rng.random() < 0.3 else int(rng.integers(-40, 60))
    hi = lo + int(rng.integers(0, 300)) if rng.random() < 0.8 else 1 << 20
    thr = float(rng.choice([0.75, 0.05, 2.0, 0.0]))
    shift = int(rng.integers(-30, 60))
    c = 3 if rng.random() < 0.3 else 1
    left, right = synth.make_pairs(n, w, h, c, shift=shift if abs(shift) < w else 0, noise_sigma=float(rng.choice([0, 2.0])), seed=1000 + t)
    if rng.random() < 0.15:
        left[:] = 128; right[:] = 128           # flat frames: every candidate ties, the first one must win
    sx, sy = (int(rng.integers(1, 4)), int(rng.integers(1, 4))) if rng.random() < 0.1 else (1, 1)  # strided grids run on the direct-form kernel
    p = _abi.make_params(tmpl_w=tw, tmpl_h=th, cost="sad", search_min=lo, search_max=hi, camera_side=side, accept_threshold=thr,
                         distance_kind=int(rng.integers(0, 3)), stride_x=sx, stride_y=sy)
    with_resolve = t % 4 == 0
    got = ctx.match_dense(left, right, p, mask=api.ALL_OUTPUTS | (_abi.OUT_RESOLVED_DISPARITY_U16 if with_resolve else 0))
    dense += ctx.last_kernel == "dense_sad_argmin_kernel"
    exp = oracle.match_dense(left, right, p)
    ok = all(np.array_equal(got[k], exp[k]) for k in ("right_index", "raw_cost", "disparity_u16")) and got["matches"].tobytes() == exp["matches"].tobytes()
    fin = np.isfinite(exp["distance"])
    ok = ok and np.array_equal(np.isfinite(got["distance"]), fin) and np.allclose(got["distance"][fin], exp["distance"][fin], rtol=1e-12, atol=0)
    if ok and with_resolve:
        for k in range(n):
            win = got["matches"][k]
            out = oracle.resolve_match_list(win[win["RightIndex"] != _abi.NO_MATCH])
            alive = np.zeros(len(win), bool)
            alive[np.unique(out["LeftIndex"])] = True
            ok = ok and np.array_equal(got["resolved_disparity_u16"][k], np.where(alive, got["disparity_u16"][k], _abi.NO_DISPARITY).astype(np.uint16))
        resolved += 1
    if not ok:
        bad += 1
        print("MISMATCH case", t, dict(w=w, c=c, h=h, n=n, tw=tw, th=th, side=side, lo=lo, hi=hi, thr=thr, shift=shift), ctx.last_kernel, flush=True)
print("cases %d, on the sliding kernel %d, resolved maps checked %d, mismatches %d" % (n_cases, dense, resolved, bad))
sys.exit(1 if bad else 0)
