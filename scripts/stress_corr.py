"""Randomised parity stress of the sliding-window correlation kernel (NCC / ZNCC / SSD, gray and colour; SAD on colour frames) against the CPU oracle:
bit-exact indices, disparities, f64 scores and MatchValues.  python scripts/stress_corr.py [n_cases] [seed] [mma|all] [auto|alu|mma|tcgen05]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from unsynchronized_stereo_vision_proj325_b200 import _abi, api, synth
from oracle import oracle

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
mma_only = len(sys.argv) > 3 and sys.argv[3] == "mma"  # only the tensor-pipe kernel's coverage: 16 / 32-px templates, no SAD, wider frames
ctx = api.Context(0)
forced = sys.argv[4] if len(sys.argv) > 4 else "auto"  # auto | alu | mma | tcgen05: Context.corr_kernel (usv_set_option)
ctx.corr_kernel(forced)
bad = dense = 0
kernels = {}
for t in range(n_cases):
    tw = int(rng.choice([1, 5, 8, 9, 12, 16, 16, 17, 20, 24, 31, 32, 32] if mma_only else [8, 12, 16, 24, 32]))
    th = int(rng.choice([1, 4, 8, 13, 16, 24, 32] + ([70] if mma_only else [])))
    c = int(rng.choice([1, 3]))
    w = int(rng.integers(tw, 700 if mma_only and rng.random() < 0.3 else 200))
    h = int(rng.integers(th, th + 40))
    n = int(rng.integers(1, 3))
    side = int(rng.integers(0, 2))
    lo = 0 if rng.random() < 0.4 else int(rng.integers(-20, 40))
    hi = lo + int(rng.integers(0, 150)) if rng.random() < 0.8 else 1 << 20
    thr = float(rng.choice([0.75, 0.2, 2.5]))
    shift = int(rng.integers(-20, 40))
    left, right = synth.make_pairs(n, w, h, c, shift=shift if abs(shift) < w else 0, noise_sigma=float(rng.choice([0, 2.0])), seed=2000 + t)
    mode = rng.random()
    if mode < 0.1:
        left[:] = 90; right[:] = 90                      # flat: zero variance everywhere (score 0, all ties)
    elif mode < 0.25:
        left[:, : h // 2] = 200; right[:, :, : w // 2] = 17   # flat regions next to texture
    p = _abi.make_params(tmpl_w=tw, tmpl_h=th, cost=str(rng.choice(["ncc", "zncc", "ssd"] if mma_only else ["ncc", "zncc", "ssd", "sad"])), search_min=lo, search_max=hi, camera_side=side,
                         accept_threshold=thr, distance_kind=int(rng.integers(0, 3)))
    got = ctx.match_dense(left, right, p)
    dense += ctx.last_kernel in ("dense_corr_argmin_kernel", "dense_corr_mma_kernel", "dense_corr_umma_kernel", "dense_sad_argmin_kernel")  # gray SAD has its own kernel
    kernels[ctx.last_kernel] = kernels.get(ctx.last_kernel, 0) + 1
    exp = oracle.match_dense(left, right, p)
    ok = all(np.array_equal(got[k], exp[k]) for k in ("right_index", "raw_cost", "disparity_u16"))
    ok = ok and got["matches"].tobytes() == exp["matches"].tobytes() and got["score"].tobytes() == exp["score"].tobytes()
    # the same job without the score output: the tensor-pipe kernel then tracks only v = 1 - score (another instantiation)
    g2 = ctx.match_dense(left, right, p, mask=_abi.OUT_MATCHES | _abi.OUT_DISPARITY_U16 | _abi.OUT_DISTANCE)
    ok = ok and g2["matches"].tobytes() == exp["matches"].tobytes() and np.array_equal(g2["disparity_u16"], exp["disparity_u16"])
    ok = ok and g2["distance"].tobytes() == got["distance"].tobytes()
    if ok and t % 4 == 0:  # the resolved disparity map (device-side ResolveMatchList over f64 / integer values) against the oracle's restatement
        g3 = ctx.match_dense(left, right, p, mask=_abi.OUT_MATCHES | _abi.OUT_DISPARITY_U16 | _abi.OUT_RESOLVED_DISPARITY_U16)
        for k in range(n):
            win = g3["matches"][k]
            out = oracle.resolve_match_list(win[win["RightIndex"] != _abi.NO_MATCH])
            alive = np.zeros(len(win), bool)
            alive[np.unique(out["LeftIndex"])] = True
            ok = ok and np.array_equal(g3["resolved_disparity_u16"][k], np.where(alive, g3["disparity_u16"][k], _abi.NO_DISPARITY).astype(np.uint16))
    if not ok:
        bad += 1
        nb = int((got["right_index"] != exp["right_index"]).sum())
        ns = int((got["score"].view(np.uint64) != exp["score"].view(np.uint64)).sum())
        print("MISMATCH case", t, dict(w=w, h=h, c=c, n=n, tw=tw, th=th, side=side, lo=lo, hi=hi, kind=p.cost_kind, mode=round(mode, 2)), ctx.last_kernel,
              "index diffs", nb, "score diffs", ns, flush=True)
        if bad <= 3:
            ii = np.argwhere(got["right_index"] != exp["right_index"])[:4]
            for a in ii:
                print("   win", a.tolist(), "got", got["right_index"][tuple(a)], got["score"][tuple(a)], "exp", exp["right_index"][tuple(a)], exp["score"][tuple(a)])
print("cases %d, on the sliding kernels %d, mismatches %d, kernels %s" % (n_cases, dense, bad, kernels))
sys.exit(1 if bad else 0)
