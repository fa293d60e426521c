"""Unsynchronised-streams driver: host nearest-timestamp pairing, sharding of the paired frames
over ranks (one process per GPU, no data-path collective) and streamed matching through the
pinned ring. Replaces the reference's two free-running CameraThread loops (P/Main.cpp:738-1309:
capture + timestamp :876-879, exchange of the other camera's data :1100-1113) for offline /
synthetic streams (BASELINE.json configs[4], "C5")."""
import numpy as np

from . import _abi, api


def shard_range(n_items, rank, world):
    """Contiguous block of `n_items` owned by `rank`: item p goes to rank floor(p * world / n_items)
    (SURVEY.md 8e). Returns (start, stop)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    start = -(-rank * n_items // world)
    stop = -(-(rank + 1) * n_items // world)
    return start, stop


def pair_streams(t_left, t_right, max_dt):
    """Nearest-timestamp pairing on the host (C-ABI usv_pair_nearest). Returns (left_idx, right_idx, dt)."""
    li, ri = api.pair_nearest(t_left, t_right, max_dt)
    dt = np.asarray(t_left)[li] - np.asarray(t_right)[ri]
    return li, ri, dt


def default_matcher(ctx, frame, params, pairs_per_slot, n_slots, mask):
    """The product matcher: a usv_stream over `ctx` (GPU only)."""
    return ctx.stream(frame, params, pairs_per_slot=pairs_per_slot, n_slots=n_slots, mask=mask)


def match_streams(frames_left, t_left, frames_right, t_right, params, max_dt=1.0 / 60.0, rank=0, world=1, ctx=None,
                  pairs_per_slot=16, n_slots=3, mask=_abi.OUT_DISPARITY_U16 | _abi.OUT_RAW_COST, stream_factory=default_matcher,
                  copy_threads=1, gather=False):
    """Pair two unsynchronised streams and match this rank's shard of the pairs.

    frames_*: [n, H, W(, C)] uint8 arrays (host) indexed by frame number; t_*: ascending timestamps.
    Returns dict(pair_left, pair_right, dt — for this rank's pairs — and one array per output in `mask`).
    `stream_factory(ctx, frame, params, pairs_per_slot, n_slots, mask)` must return an object with the
    usv_stream interface (slots / submit / wait / close); tests inject a CPU stub to exercise the host logic.
    gather=True: frames_* are the cameras' frame stores and the paired frames go store -> HBM directly
    (usv_stream_submit_gather; the stores are page-locked for the duration of the call) — no staging memcpy on the host.
    """
    li, ri, dt = pair_streams(t_left, t_right, max_dt)
    lo, hi = shard_range(len(li), rank, world)
    li, ri, dt = li[lo:hi], ri[lo:hi], dt[lo:hi]
    n = len(li)
    h, w = frames_left.shape[1:3]
    c = frames_left.shape[3] if frames_left.ndim == 4 else 1
    frame = _abi.FrameDesc(w, h, c, w * c, w * c * h)
    st = stream_factory(ctx, frame, params, pairs_per_slot, n_slots, mask)
    outs = {name: [] for name, bit, _ in _abi.OUTPUT_FIELDS if mask & bit}
    pending = []  # (slot, count) in submission order
    pool = None
    registered = []
    try:
        if gather and ctx is not None:
            for store in (frames_left, frames_right):
                if store.flags["C_CONTIGUOUS"] and store.nbytes:
                    try:
                        ctx.host_register(store)
                        registered.append(store)
                    except api.UsvError:
                        pass  # already page-locked by the caller (or not lockable): the copies still work, synchronously
        if copy_threads > 1 and not gather:  # the "capture" memcpy into the pinned ring is the host-side bound of the stream: spread it
            from concurrent.futures import ThreadPoolExecutor
            pool = ThreadPoolExecutor(copy_threads)

        def drain_one():
            slot, cnt = pending.pop(0)
            st.wait(slot)
            for k in outs:
                outs[k].append(st.slots[slot]["out"][k][:cnt].copy())

        for b0 in range(0, n, pairs_per_slot):
            slot = (b0 // pairs_per_slot) % n_slots
            if len(pending) == n_slots:
                drain_one()  # the oldest in-flight batch owns this slot
            cnt = min(pairs_per_slot, n - b0)
            if gather:
                st.submit_gather(slot, frames_left, li[b0:b0 + cnt], frames_right, ri[b0:b0 + cnt])
                pending.append((slot, cnt))
                continue
            # "capture": the paired frames land in the slot's pinned buffers
            sl, sr = st.slots[slot]["left"], st.slots[slot]["right"]
            if pool is None:
                sl[:cnt] = frames_left[li[b0:b0 + cnt]].reshape(cnt, h, w * c)
                sr[:cnt] = frames_right[ri[b0:b0 + cnt]].reshape(cnt, h, w * c)
            else:
                list(pool.map(lambda k: (np.copyto(sl[k], frames_left[li[b0 + k]].reshape(h, w * c)),
                                         np.copyto(sr[k], frames_right[ri[b0 + k]].reshape(h, w * c))), range(cnt)))
            st.submit(slot, cnt)
            pending.append((slot, cnt))
        while pending:
            drain_one()
    finally:  # an exception must not leak the page-locked stores, the pinned ring or the copy threads
        st.close()
        for store in registered:
            try:
                ctx.host_unregister(store)
            except api.UsvError:
                pass
        if pool is not None:
            pool.shutdown()
    res = {"pair_left": li, "pair_right": ri, "dt": dt}
    for k, v in outs.items():
        res[k] = np.concatenate(v, axis=0) if v else np.zeros((0, 0), dtype=dict((n_, d) for n_, _, d in _abi.OUTPUT_FIELDS)[k])
    return res


def gather_on_host(local, rank, world, group=None):
    """Host-side gather of per-rank result dicts onto rank 0 in rank (= pair) order. Uses
    torch.distributed only as plumbing (gloo or nccl-backed object gather); returns the merged dict on
    rank 0 and None elsewhere."""
    if world == 1:
        return local
    import torch.distributed as dist
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(local, bucket, dst=0, group=group)
    if rank != 0:
        return None
    return {k: np.concatenate([b[k] for b in bucket], axis=0) for k in local}


def extrapolated_distances(ctx, camera_side, this_frame, t_this, other_frames, t_other, templates, params):
    """The reference's "unsynchronised" step for pixel templates: the other camera has no frame at
    this camera's capture time, so the matched position of every template is tracked through the other camera's
    last three frames and extrapolated to `t_this` before the disparity is taken
    (MovingObjectDistanceCalculator, P/DistanceCalculator.cpp:53-84; call order of P/Main.cpp:1115-1143, :1238).

    this_frame: [H, W] uint8; other_frames: (older, old, cur) [H, W] uint8 with timestamps t_other (seconds,
    ascending); templates: [n, 2] int (x, y) window corners in this camera's frame.
    Every step runs on the GPU: three sparse block searches (usv_match_templates_*) and the distance kernel
    (usv_moving_object_distance). Returns dict(distance [n] f64 cm,
    nearest_distance [n] f64 (disparity against the newest frame only), tracks [3, n] f32 x', accepted [n] bool).
    """
    tpl = np.ascontiguousarray(np.asarray(templates, np.int32).reshape(-1, 2))
    n = len(tpl)
    half_w, half_h = params.tmpl_w / 2.0, params.tmpl_h / 2.0
    nxc = this_frame.shape[1] - params.tmpl_w + 1
    tracks, ok = np.zeros((3, n), np.float32), np.ones(n, bool)
    for k, fr in enumerate(other_frames):
        got = ctx.match_templates(this_frame[None], fr[None], tpl[:, 0], tpl[:, 1], params, mask=_abi.OUT_RIGHT_INDEX)
        ri = got["right_index"][0]
        ok &= ri != _abi.NO_MATCH
        tracks[k] = (ri % nxc).astype(np.float32)  # x' of the winner; y' = y (rectified pair)
    centre = lambda xs: np.stack([xs + half_w, tpl[:, 1] + half_h], 1).astype(np.float32)  # noqa: E731
    this_xy = centre(tpl[:, 0].astype(np.float32))
    older_xy, old_xy, cur_xy = (centre(tracks[k]) for k in range(3))
    # identity tracks: template i is object i in all three frames, so the index triples are (i, i, i) — what
    # IDMatcher (P/Main.cpp:483-499, usv_id_matcher) is meant to deliver; the reference's own join collapses every
    # triple to (old.RightIndex, 0, 0) through the comma operator at :492
    idx3 = np.repeat(np.arange(n, dtype=np.int32)[:, None], 3, axis=1)
    ns = lambda t: int(round(t * 1e9))  # noqa: E731
    dist = ctx.moving_object_distance(camera_side, ns(t_this), this_xy, cur_xy, old_xy, older_xy, idx3,
                                      ns(t_other[2]), ns(t_other[1]), ns(t_other[0]))
    dist = np.where(ok, dist, np.inf) if len(dist) == n else np.full(n, np.inf)
    disp_now = (tpl[:, 0] - tracks[2]).astype(np.int32) if camera_side == _abi.LEFT_CAM else (tracks[2] - tpl[:, 0]).astype(np.int32)
    nearest = np.where(ok, ctx.disparity_to_distance(np.maximum(disp_now, 0), _abi.DIST_POWERLAW), np.inf)
    return {"distance": dist, "nearest_distance": nearest, "tracks": tracks, "accepted": ok, "idx3": idx3}
