"""Host logic of the unsynchronised-streams driver: sharding, pairing, slot rotation and the
world_size-2 gather (gloo, CPU, with a stub matcher), plus the real thing on one GPU."""
import os
import socket

import numpy as np
import pytest

from unsynchronized_stereo_vision_proj325_b200 import _abi, api, pipeline, synth


def test_shard_range_partitions():
    for n in (0, 1, 7, 256, 4096, 9999):
        for world in (1, 2, 3, 4, 8):
            spans = [pipeline.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        pipeline.shard_range(10, 2, 2)


class StubStream:
    """CPU stand-in with the usv_stream interface: 'matches' by writing the mean of the left frame and
    the pair's row sum into the outputs, and checks the ring protocol (no slot reused while in flight)."""

    def __init__(self, ctx, frame, params, pairs_per_slot, n_slots, mask):
        self.pps, self.n_slots = pairs_per_slot, n_slots
        self.nwin = 3
        self.slots = [{"left": np.zeros((pairs_per_slot, frame.height, frame.width * frame.channels), np.uint8),
                       "right": np.zeros((pairs_per_slot, frame.height, frame.width * frame.channels), np.uint8),
                       "out": {"raw_cost": np.zeros((pairs_per_slot, self.nwin), np.uint32),
                               "disparity_u16": np.zeros((pairs_per_slot, self.nwin), np.uint16)}} for _ in range(n_slots)]
        self.in_flight = set()

    def submit(self, slot, n):
        assert slot not in self.in_flight, "slot resubmitted before wait"
        self.in_flight.add(slot)
        s = self.slots[slot]
        s["out"]["raw_cost"][:n, 0] = s["left"][:n].reshape(n, -1).sum(1)
        s["out"]["raw_cost"][:n, 1] = s["right"][:n].reshape(n, -1).sum(1)
        s["out"]["disparity_u16"][:n, 0] = s["left"][:n, 0, 0]

    def submit_gather(self, slot, left_store, idx_left, right_store, idx_right):
        n = len(idx_left)
        s = self.slots[slot]
        s["left"][:n] = left_store[idx_left].reshape(n, s["left"].shape[1], -1)
        s["right"][:n] = right_store[idx_right].reshape(n, s["right"].shape[1], -1)
        self.submit(slot, n)

    def wait(self, slot):
        self.in_flight.discard(slot)

    def close(self):
        assert not self.in_flight


def _streams(n=50, w=16, h=8):
    tl, idl = synth.make_timestamps(n, seed=1)
    tr, idr = synth.make_timestamps(n, phase=0.011, seed=2)
    rng = np.random.default_rng(5)
    fl = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    fr = rng.integers(0, 256, (n, h, w), dtype=np.uint8)
    return fl[idl], tl, fr[idr], tr


def _run_rank(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fl, tl, fr, tr = _streams()
    p = _abi.make_params(tmpl_w=4, tmpl_h=4)
    local = pipeline.match_streams(fl, tl, fr, tr, p, rank=rank, world=world, pairs_per_slot=4, n_slots=3, stream_factory=StubStream)
    merged = pipeline.gather_on_host(local, rank, world)
    if rank == 0:
        q.put({k: v.tolist() for k, v in merged.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_rank():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_run_rank, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    merged = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    fl, tl, fr, tr = _streams()
    single = pipeline.match_streams(fl, tl, fr, tr, _abi.make_params(tmpl_w=4, tmpl_h=4), pairs_per_slot=4, n_slots=3, stream_factory=StubStream)
    assert merged["pair_left"] == single["pair_left"].tolist() and merged["pair_right"] == single["pair_right"].tolist()
    assert merged["raw_cost"] == single["raw_cost"].tolist()
    # the stub's outputs identify the frames that were paired
    li, ri = np.array(merged["pair_left"]), np.array(merged["pair_right"])
    rc = np.array(merged["raw_cost"])
    assert np.array_equal(rc[:, 0], fl[li].reshape(len(li), -1).sum(1)) and np.array_equal(rc[:, 1], fr[ri].reshape(len(ri), -1).sum(1))
    assert len(li) > 40 and (np.abs(np.array(merged["dt"])) <= 1 / 60).all()


def test_gather_mode_equals_staged_mode():
    fl, tl, fr, tr = _streams()
    p = _abi.make_params(tmpl_w=4, tmpl_h=4)
    a = pipeline.match_streams(fl, tl, fr, tr, p, pairs_per_slot=4, n_slots=3, stream_factory=StubStream)
    b = pipeline.match_streams(fl, tl, fr, tr, p, pairs_per_slot=4, n_slots=3, stream_factory=StubStream, gather=True)
    assert all(np.array_equal(a[k], b[k]) for k in a)


@pytest.mark.gpu
@pytest.mark.parametrize("w", [128, 150])
def test_streams_gather_on_gpu_vs_oracle(oracle, w):
    """Frames go from the cameras' (page-locked) frame stores straight to HBM: same results as the staged ring, both for
    a store pitch equal to the device pitch (one copy per run of consecutive frames) and for one that needs 2-D copies."""
    n, h = 40, 40
    left, right = synth.make_pairs(n, w, h, 1, shift=9, noise_sigma=2.0, seed=4)
    tl, idl = synth.make_timestamps(n, drop_prob=0.1, seed=1)
    tr, idr = synth.make_timestamps(n, phase=0.011, drop_prob=0.1, seed=2)
    fl, fr = np.ascontiguousarray(left[idl]), np.ascontiguousarray(right[idr])
    p = _abi.make_params(tmpl_w=16, tmpl_h=16, cost="sad", search_max=31)
    ctx = api.Context(0)
    res = pipeline.match_streams(fl, tl, fr, tr, p, ctx=ctx, pairs_per_slot=8, n_slots=3, gather=True)
    li, ri = res["pair_left"], res["pair_right"]
    exp = oracle.match_dense(fl[li], fr[ri], p, mask=_abi.OUT_DISPARITY_U16 | _abi.OUT_RAW_COST)
    assert np.array_equal(res["raw_cost"], exp["raw_cost"]) and np.array_equal(res["disparity_u16"], exp["disparity_u16"])
    st = ctx.stream(_abi.frame_desc_for(fl), p, pairs_per_slot=4, n_slots=1)
    with pytest.raises(api.UsvError):
        st.submit_gather(0, fl, [0, len(fl)], fr, [0, 1])  # index outside the store
    st.close()
    ctx.close()


@pytest.mark.gpu
def test_streams_on_gpu_vs_oracle(oracle):
    """C5 in miniature: jittered timestamps with drops, nearest pairing, streamed matching."""
    n, w, h = 40, 160, 40
    left, right = synth.make_pairs(n, w, h, 1, shift=9, noise_sigma=2.0, seed=3)
    tl, idl = synth.make_timestamps(n, drop_prob=0.05, seed=1)
    tr, idr = synth.make_timestamps(n, phase=0.011, drop_prob=0.05, seed=2)
    fl, fr = np.ascontiguousarray(left[idl]), np.ascontiguousarray(right[idr])
    p = _abi.make_params(tmpl_w=16, tmpl_h=16, cost="sad", search_max=31)
    ctx = api.Context(0)
    res = pipeline.match_streams(fl, tl, fr, tr, p, ctx=ctx, pairs_per_slot=8, n_slots=3)
    li, ri = res["pair_left"], res["pair_right"]
    ol, orr = oracle.pair_nearest(tl, tr, 1 / 60)
    assert np.array_equal(li, ol) and np.array_equal(ri, orr)
    exp = oracle.match_dense(fl[li], fr[ri], p, mask=_abi.OUT_DISPARITY_U16 | _abi.OUT_RAW_COST)
    assert np.array_equal(res["raw_cost"], exp["raw_cost"]) and np.array_equal(res["disparity_u16"], exp["disparity_u16"])
    ctx.close()
