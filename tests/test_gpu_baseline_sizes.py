"""GPU parity at BASELINE.json's sizes for the configurations the small sweeps only cover cropped:

  C3  1280x720x3, 32x32 templates, D = 256, ZNCC / NCC / colour SSD — the dense result of the full frame against
      the oracle's answer for several hundred single windows (indices AND the f64 score / MatchValue bytes), incl.
      the frame edges and both sides of every x-tile boundary of the two tensor-pipe kernels (32 windows for
      mma.sync, 128 for tcgen05), once per kernel (usv_set_option picks it, usv_last_kernel is asserted);
  C5  2 x 10 000 timestamps with jitter, phase offset and drops: the host pairing equals the brute-force oracle,
      all pairs are streamed through the pinned ring (usv_stream_submit_gather) and a sample of the streamed pairs
      equals the oracle's dense result for those frames.

Scan order P/Main.cpp:408-423, tie rule :451."""
import numpy as np
import pytest

from unsynchronized_stereo_vision_proj325_b200 import _abi, api, pipeline, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def c3_frames():
    return synth.make_pairs(1, 1280, 720, 3, shift=60, noise_sigma=3.0, seed=33)


def _c3_windows(nx, ny):
    rng = np.random.default_rng(5)
    tx, ty = list(rng.integers(0, nx, 220)), list(rng.integers(0, ny, 220))
    ys = [0, ny - 1, ny // 3, (2 * ny) // 3 + 1]
    for k, b in enumerate(range(32, nx, 32)):  # both sides of every 32-window boundary (every fourth one is a 128-window boundary)
        for x in (b - 1, b):
            tx.append(x); ty.append(ys[k % len(ys)])
    for x in (0, 1, 255, 256, 257, nx - 1):    # frame edges; the first windows whose whole 256-px range is inside the frame
        for y in (0, ny - 1, ny // 2):
            tx.append(x); ty.append(y)
    return np.asarray(tx, np.int32), np.asarray(ty, np.int32)


@pytest.mark.parametrize("kernel,name", [("mma", "dense_corr_mma_kernel"), ("tcgen05", "dense_corr_umma_kernel")])
@pytest.mark.parametrize("cost", ["zncc", "ncc", "ssd"])
def test_c3_full_frame_sampled_windows(ctx, oracle, c3_frames, cost, kernel, name):
    left, right = c3_frames
    p = _abi.make_params(tmpl_w=32, tmpl_h=32, cost=cost, search_max=255, accept_threshold=0.6)
    f = _abi.frame_desc_for(left)
    nx, ny, ev = api.grid_dims(f, p)
    assert (nx, ny) == (1249, 689)  # 860 561 windows
    mask = _abi.OUT_MATCHES | _abi.OUT_RIGHT_INDEX | _abi.OUT_RAW_COST | _abi.OUT_SCORE | _abi.OUT_DISPARITY_U16
    ctx.corr_kernel(kernel)
    try:
        got = ctx.match_dense(left, right, p, mask=mask)
        assert ctx.last_kernel == name
    finally:
        ctx.corr_kernel("auto")
    tx, ty = _c3_windows(nx, ny)
    assert len(tx) >= 300
    exp = oracle.match_templates(left, right, tx, ty, p, mask=mask)
    w_idx = ty.astype(np.int64) * nx + tx
    for k in ("right_index", "raw_cost", "disparity_u16"):
        assert np.array_equal(got[k][0][w_idx], exp[k][0]), k
    assert got["score"][0][w_idx].tobytes() == exp["score"][0].tobytes()
    gm, em = got["matches"][0][w_idx], exp["matches"][0]
    assert np.array_equal(gm["RightIndex"], em["RightIndex"]) and gm["MatchValue"].tobytes() == em["MatchValue"].tobytes()
    assert np.array_equal(gm["LeftIndex"], w_idx.astype(np.uint32))
    if cost != "ssd":  # the known shift is recovered where its counterpart is inside the frame (noise sigma 3 on 32x32x3 bytes)
        d = got["disparity_u16"][0].reshape(ny, nx)[:, 60:]
        assert (d == 60).mean() > 0.99


def test_c3_colour_sad_full_frame(ctx, oracle, c3_frames):
    left, right = c3_frames
    p = _abi.make_params(tmpl_w=32, tmpl_h=32, cost="sad", search_max=255)
    f = _abi.frame_desc_for(left)
    nx, ny, _ = api.grid_dims(f, p)
    mask = _abi.OUT_RIGHT_INDEX | _abi.OUT_RAW_COST | _abi.OUT_DISPARITY_U16
    got = ctx.match_dense(left, right, p, mask=mask)
    assert ctx.last_kernel in ("dense_corr_argmin_kernel", "dense_sad_colour_kernel", "dense_sad_argmin_kernel")
    tx, ty = _c3_windows(nx, ny)
    exp = oracle.match_templates(left, right, tx, ty, p, mask=mask)
    w_idx = ty.astype(np.int64) * nx + tx
    for k in ("right_index", "raw_cost", "disparity_u16"):
        assert np.array_equal(got[k][0][w_idx], exp[k][0]), k


def _brute_force_pairing(tl, tr, max_dt):
    """Every left frame takes the right frame with the smallest |dt| (earlier one on a tie); rejected above max_dt; a right
    frame is kept by its closest left frame (lowest index on a tie). O(n^2) in chunks — the definition, not the algorithm."""
    n = len(tl)
    pick = np.empty(n, np.int64)
    gap = np.empty(n, np.float64)
    for a in range(0, n, 512):
        d = np.abs(tl[a:a + 512, None] - tr[None, :])
        j = d.argmin(axis=1)  # first minimum = earlier frame on a tie
        pick[a:a + 512], gap[a:a + 512] = j, d[np.arange(len(j)), j]
    ok = gap <= max_dt
    best = {}
    for i in np.nonzero(ok)[0]:
        j = int(pick[i])
        if j not in best or gap[i] < gap[best[j]]:
            best[j] = int(i)
    pairs = sorted((i, j) for j, i in best.items())
    return np.asarray([p[0] for p in pairs], np.int32), np.asarray([p[1] for p in pairs], np.int32)


def test_c5_streams_at_size(ctx, oracle):
    n_frames, pool = 10000, 48
    tl, idl = synth.make_timestamps(n_frames, fps=30.0, jitter_sigma=0.002, phase=0.0, drop_prob=0.01, seed=1)
    tr, idr = synth.make_timestamps(n_frames, fps=30.0, jitter_sigma=0.002, phase=0.011, drop_prob=0.01, seed=2)
    li, ri, dt = pipeline.pair_streams(tl, tr, 1.0 / 60.0)
    bl, br = _brute_force_pairing(tl, tr, 1.0 / 60.0)
    assert np.array_equal(li, bl) and np.array_equal(ri, br)
    assert 9000 < len(li) <= min(len(tl), len(tr)) and np.abs(dt).max() <= 1.0 / 60.0

    w, h = 640, 480
    left, right = synth.make_pairs(pool, w, h, 1, shift=37, noise_sigma=2.0, seed=325)
    # the right store is rotated against the left one, so that a pair is (left[a], right[b]) with a != b in general:
    # what an unsynchronised pairing produces, and a wrong gather index cannot go unnoticed
    right = np.ascontiguousarray(np.roll(right, 5, axis=0))
    p = _abi.make_params(tmpl_w=16, tmpl_h=16, cost="sad", search_max=127)
    frame = _abi.FrameDesc(w, h, 1, w, w * h)
    mask = _abi.OUT_DISPARITY_U16 | _abi.OUT_RAW_COST_U16
    pps, ns = 32, 4
    st = ctx.stream(frame, p, pairs_per_slot=pps, n_slots=ns, mask=mask)
    fl_all, fr_all = (idl[li] % pool).astype(np.int32), (idr[ri] % pool).astype(np.int32)
    sample = set(np.random.default_rng(3).choice(len(li), 12, replace=False).tolist()) | {0, len(li) - 1}
    kept, pending, n_done = {}, [], 0
    ctx.host_register(left)
    ctx.host_register(right)
    try:
        def drain():
            nonlocal n_done
            slot, b0, cnt = pending.pop(0)
            st.wait(slot)
            n_done += cnt
            for k in range(cnt):
                if b0 + k in sample:
                    kept[b0 + k] = {n: st.slots[slot]["out"][n][k].copy() for n in ("disparity_u16", "raw_cost_u16")}
        for b0 in range(0, len(li), pps):
            slot = (b0 // pps) % ns
            if len(pending) == ns:
                drain()
            cnt = min(pps, len(li) - b0)
            st.submit_gather(slot, left, fl_all[b0:b0 + cnt], right, fr_all[b0:b0 + cnt])
            pending.append((slot, b0, cnt))
        while pending:
            drain()
    finally:
        st.close()
        ctx.host_unregister(left)
        ctx.host_unregister(right)
    assert n_done == len(li) and ctx.last_kernel == "dense_sad_argmin_kernel"
    assert sorted(kept) == sorted(sample)
    for k in sorted(kept):
        a, b = int(fl_all[k]), int(fr_all[k])
        exp = oracle.match_dense(left[a:a + 1], right[b:b + 1], p, mask=_abi.OUT_DISPARITY_U16 | _abi.OUT_RAW_COST)
        assert np.array_equal(kept[k]["disparity_u16"], exp["disparity_u16"][0]), k
        assert np.array_equal(kept[k]["raw_cost_u16"].astype(np.uint32), exp["raw_cost"][0]), k
