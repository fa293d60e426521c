#!/usr/bin/env python
"""bench.py — frame-pairs/s and candidate-evals/s of the stereo block-search hot path.

Workload (BASELINE.json configs[1], "C2" in SURVEY.md 8(d)): a batch of 256 synthetic
640x480 grayscale pairs per GPU, dense template sweep — every 16x16 window, stride 1,
full-row search range (x' in [0, x]), SAD cost, first-minimum selection, pinhole distance.
290 625 windows and 90 965 625 candidate evaluations per pair.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...     (one rank per GPU; no data-path collective)

`value`  : pairs/s with the frames already resident in HBM (CUDA events on the launching stream).
`e2e`    : pairs/s through the C-ABI streaming ring with HOST (pinned) buffers — H2D of every
           frame and D2H of every result inside the timed region.
`roofline`: the sliding-window kernel is integer-ALU bound; achieved = candidate evals/s x the
           ALU-pipe lane-ops one evaluation needs in this formulation, peak = the
           VABSDIFF4.U8.ACC issue rate measured on this GPU in this run.
`cpu_baseline` / `--impl reference`: the CPU oracle port (direct-form, OpenMP, all host cores)
           on a bounded sample of the same workload. The reference repo has no pixel path and
           cannot be built here (MSVC + OpenCV 3.0), so the port is the CPU arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, TW, TH = 640, 480, 16, 16
PAIRS_PER_GPU = 256
# ALU-pipe thread-instructions one candidate evaluation needs in the sliding-window formulation
# (DESIGN.md "roofline"): VABSDIFF4.U8.ACC for the row entering the window, one for the row
# leaving it, and one VIMNMX at twice the VABSDIFF4 issue rate (counted 0.5).
ALU_OPS_PER_EVAL = 2.5
OUT_BYTES_PER_WINDOW = 2 + 2 + 4  # disparity_u16 + raw_cost_u16 (lossless: 255*16*16 < 2^16) + distance_f32


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (recipe in B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [c.strip() for c in line.split(",")]))
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": sorted(reasons)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_sample(params, seconds=12.0, seed=325):
    """Time the CPU oracle (direct-form port, all host cores) on whole window rows of pairs of the
    workload (pair after pair) until about `seconds` of CPU work has been done."""
    from oracle import oracle
    from unsynchronized_stereo_vision_proj325_b200 import _abi, api, synth
    n_src = 4
    left, right = synth.make_pairs(n_src, W, H, 1, shift=37, noise_sigma=2.0, seed=seed)
    f = _abi.frame_desc_for(left)
    nx, ny, ev_pair = api.grid_dims(f, params)
    cores = host_cores()
    oracle.match_dense_rows(left[0], right[0], params, 0, min(ny, cores), threads=cores)  # warm-up / page-in
    rows_done, evals, t_used = 0, 0, 0.0
    chunk = min(ny, 8 * cores)
    while t_used < seconds:
        pair, r0 = (rows_done // ny) % n_src, rows_done % ny
        n = min(chunk, ny - r0)
        t0 = time.perf_counter()
        _, _, ev = oracle.match_dense_rows(left[pair], right[pair], params, r0, r0 + n, threads=cores)
        t_used += time.perf_counter() - t0
        evals += ev
        rows_done += n
    return {"evals": evals, "seconds": t_used, "rows": rows_done, "ny": ny, "cores": cores, "evals_per_pair": ev_pair,
            "pairs_per_s": evals / ev_pair / t_used, "evals_per_s": evals / t_used}


def run_reference(args, rank, world):
    """`--impl reference`: the CPU arm. Rank 0 alone works; other ranks exit 0."""
    if rank != 0:
        return
    from unsynchronized_stereo_vision_proj325_b200 import _abi
    params = _abi.make_params(tmpl_w=TW, tmpl_h=TH, cost="sad", distance_kind=_abi.DIST_PINHOLE)
    per_step = max(1.0, min(15.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_sample(params, seconds=per_step)
    t, ev, last = 0.0, 0, None
    for _ in range(args.steps):
        last = cpu_sample(params, seconds=per_step)
        t += last["seconds"]
        ev += last["evals"]
    pairs_s = ev / last["evals_per_pair"] / t
    sample = "%d window rows (%.2f pairs of 465 rows) of the same 640x480 workload per step, %.1f s of CPU work per step" % (
        last["rows"], last["rows"] / last["ny"], per_step)
    line = {
        "impl": "reference", "metric": "frame-pairs/sec", "value": pairs_s, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cand_evals_per_s": ev / t,
        "cpu_baseline": {"value": pairs_s, "unit": "pairs/s", "cores": last["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": pairs_s, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference repo has no pixel-level search and is MSVC/OpenCV-3.0 only (unbuildable here): the CPU arm is the "
                "oracle port (direct-form SAD, reference loop order and tie rule, gcc -O3 -march=native, OpenMP)",
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": "C2: %d synthetic 640x480 gray pairs per GPU, 16x16 SAD templates, stride 1, full-row range "
                        "(290625 windows, 90965625 candidate evals per pair), first-min + pinhole distance" % PAIRS_PER_GPU,
            "pairs_per_gpu": PAIRS_PER_GPU, "global_pairs": PAIRS_PER_GPU * n_gpus, "parallelism": "pairs sharded, no collective",
            "outputs": "disparity_u16 + raw_cost_u16 + distance_f32 per window (8 B)",
            "l2": "inputs (157 MB) + outputs (595 MB) per step exceed the 126 MB L2; no explicit flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU, help="pairs per GPU (dev only; the contract value is 256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from unsynchronized_stereo_vision_proj325_b200 import _abi, api, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # torch.distributed is measurement plumbing only (barrier + max over ranks); the data path has no collective
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.pairs
    params = _abi.make_params(tmpl_w=TW, tmpl_h=TH, cost="sad", distance_kind=_abi.DIST_PINHOLE)
    ctx = api.Context(local_rank)
    left, right = synth.make_pairs(n, W, H, 1, shift=37, noise_sigma=2.0, seed=325 + rank)
    frame = _abi.FrameDesc(W, H, 1, W, W * H)
    nx, ny, ev_pair = api.grid_dims(frame, params)
    n_win = nx * ny
    mask = _abi.OUT_DISPARITY_U16 | _abi.OUT_RAW_COST_U16 | _abi.OUT_DISTANCE_F32

    # ---------------- value: inputs resident in HBM, CUDA events on the launching stream -----------
    d_left = torch.from_numpy(np.ascontiguousarray(left)).cuda()
    d_right = torch.from_numpy(np.ascontiguousarray(right)).cuda()
    o_disp = torch.empty(n * n_win, dtype=torch.int16, device="cuda")
    o_cost = torch.empty(n * n_win, dtype=torch.int16, device="cuda")
    o_dist = torch.empty(n * n_win, dtype=torch.float32, device="cuda")
    d_out = _abi.Outputs()
    d_out.disparity_u16, d_out.raw_cost_u16, d_out.distance_f32 = o_disp.data_ptr(), o_cost.data_ptr(), o_dist.data_ptr()
    stream = torch.cuda.current_stream().cuda_stream

    def step_device():
        ctx.match_dense_device(d_left.data_ptr(), d_right.data_ptr(), frame, n, params, d_out, stream)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_clk0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    t_clk1 = time.time()
    launches = ctx.launch_count - launches0
    kernel_name = ctx.last_kernel
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = world * n * args.steps / (ms_total * 1e-3)
    evals_s = value * ev_pair

    # parity spot-check of the timed launches' output against the oracle (rank 0, one window row of pair 0)
    parity = None
    if rank == 0:
        try:
            from oracle import oracle
            ri, rc, _ = oracle.match_dense_rows(left[0], right[0], params, 100, 101, threads=host_cores())
            got_c = o_cost[100 * nx:101 * nx].cpu().numpy().view(np.uint16).astype(np.uint32)
            got_d = o_disp[100 * nx:101 * nx].cpu().numpy().view(np.uint16)
            exp_d = (np.arange(nx) - (ri.astype(np.int64) - 100 * nx)).astype(np.uint16)
            parity = bool(np.array_equal(got_c, rc) and np.array_equal(got_d, exp_d))
        except Exception as e:  # the oracle is optional at bench time
            parity = "unchecked: %s" % e

    # ---------------- roofline denominator: live VABSDIFF4 issue rate ---------------------------------
    peak_lane = ctx.probe_issue_rate(0, 25.0)
    achieved_lane = evals_s / world * ALU_OPS_PER_EVAL
    algo_bytes = 2 * W * H + OUT_BYTES_PER_WINDOW * n_win  # per pair
    hbm_peak = 6449.7
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        hbm_src = "measured"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback"
    traffic, warp_inst_pair = None, None
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "dense_traffic.json")))
        traffic = rec["dram_bytes_per_pair"] * n
        warp_inst_pair = rec.get("warp_instructions_per_pair")  # executed warp-instructions, ncu on the same launch
    except Exception:
        pass

    # ---------------- e2e: C-ABI streaming ring, pinned host buffers, H2D + kernels + D2H ------------
    pps = 32 if n % 32 == 0 else n
    n_slots = n // pps

    def measure_e2e(out_mask, check):
        # up to two batches in flight: the ring has two halves of n_slots slots; with depth 2, step k is submitted into
        # half k & 1 before step k - 1 is waited for, so the copies of one step overlap the kernels of its neighbours
        # (every step's H2D and D2H still run inside the timed region, which ends when the last results are on the host)
        st = ctx.stream(frame, params, pairs_per_slot=pps, n_slots=2 * n_slots, mask=out_mask)
        for s in range(2 * n_slots):  # the capture side writes frames straight into the pinned ring
            a = (s % n_slots) * pps
            st.slots[s]["left"][:] = left[a:a + pps]
            st.slots[s]["right"][:] = right[a:a + pps]

        def run_steps(k_steps, depth):
            for k in range(k_steps):
                for s in range(n_slots):
                    st.submit((k & 1) * n_slots + s)
                if depth == 1:
                    for s in range(n_slots):
                        st.wait((k & 1) * n_slots + s)
                elif k > 0:
                    for s in range(n_slots):
                        st.wait(((k - 1) & 1) * n_slots + s)
            if depth == 2:
                for s in range(n_slots):
                    st.wait(((k_steps - 1) & 1) * n_slots + s)

        # warm-up doubles as calibration of the pipeline depth: one GPU alone gains from two steps in flight, several
        # GPUs behind one host lose when copies in both directions overlap all the time (scripts/pcie_probe.py)
        t_cal = {}
        for depth in (1, 2):
            run_steps(1, depth)
            barrier()
            t0 = time.perf_counter()
            run_steps(max(4, args.warmup), depth)
            torch.cuda.synchronize()
            t_cal[depth] = max_over_ranks(time.perf_counter() - t0)
        depth = 1 if t_cal[1] <= 1.03 * t_cal[2] else 2  # two steps in flight only when clearly faster
        forced = os.environ.get("USV_BENCH_E2E_DEPTH")    # measurement aid: pin the depth for A/B runs
        if forced in ("1", "2"):
            depth = int(forced)
        barrier()
        t0 = time.perf_counter()
        run_steps(args.steps, depth)
        torch.cuda.synchronize()
        t = max_over_ranks(time.perf_counter() - t0)
        barrier()
        ok = None
        if rank == 0 and check:
            got = st.slots[0]["out"]["raw_cost_u16"][0, 100 * nx:101 * nx]
            ok = bool(np.array_equal(got, o_cost[100 * nx:101 * nx].cpu().numpy().view(np.uint16)))
        res = (world * n * args.steps / t, st.h2d_bytes_per_pair * n, st.d2h_bytes_per_pair * n, ok, depth, {k: round(v * 1e3, 2) for k, v in t_cal.items()})
        st.close()
        return res

    e2e_value, h2d, d2h, e2e_ok, e2e_depth, e2e_cal = measure_e2e(mask, True)
    t_clk1 = time.time()  # the clock samples cover both timed regions (device-resident and e2e)
    # the same step with the 4-byte result record (the distance is a function of the disparity: the host can look it
    # up in the W-entry table): what the copies back to the host cost. Reported beside e2e, not instead of it.
    c_value, c_h2d, c_d2h, _, c_depth, _ = measure_e2e(_abi.OUT_DISPARITY_U16 | _abi.OUT_RAW_COST_U16, False)

    clocks = None
    if rank == 0:
        sampler.stop()
        clocks = sampler.summary(t_clk0, t_clk1)

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        c = cpu_sample(params, seconds=12.0)
        cpu = {"value": c["pairs_per_s"], "unit": "pairs/s", "cores": c["cores"], "kind": "port",
               "cand_evals_per_s": c["evals_per_s"],
               "sample": "%d window rows (%.2f pairs of %d rows) of the same workload, %.1f s" % (c["rows"], c["rows"] / c["ny"], c["ny"], c["seconds"])}

    if rank == 0:
        line = {
            "metric": "frame-pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(world) if n == PAIRS_PER_GPU else dict(workload_config(world), pairs_per_gpu=n, global_pairs=n * world),
            "cand_evals_per_s": evals_s, "kernel": kernel_name, "gpu_launches": int(launches), "parity_vs_oracle": parity,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "cand_evals_per_s": e2e_value * ev_pair, "api": "usv_stream_submit/usv_stream_wait, %d slots x %d pairs per step, %d step(s) in flight (calibrated in the warm-up)" % (n_slots, pps, e2e_depth),
                    "matches_device_path": e2e_ok, "calibration_ms_by_depth": e2e_cal,
                    "compact_results": {"value": c_value, "unit": "pairs/s", "h2d_bytes_per_step": int(c_h2d), "d2h_bytes_per_step": int(c_d2h),
                                        "outputs": "disparity_u16 + raw_cost_u16 per window (4 B); distance left to a host table lookup"}},
            "roofline": {"bound": "alu", "achieved": achieved_lane / 1e12, "peak": peak_lane / 1e12, "unit": "Tlaneop/s",
                         "frac": achieved_lane / peak_lane, "traffic": traffic,
                         "alu_ops_per_eval": ALU_OPS_PER_EVAL, "peak_source": "VABSDIFF4.U8.ACC issue rate measured in this run (usv_probe_issue_rate)",
                         "direct_form_byteops_per_s": evals_s / world * TW * TH,
                         # how busy the SM's four issue ports are: executed warp-instructions (ncu count of this launch
                         # shape) per second against 4 per clock and SM at the sampled clock
                         "issue_slots": None if not (warp_inst_pair and clocks) else {
                             "achieved_gwarpinst_per_s": value / world * warp_inst_pair / 1e9,
                             "peak_gwarpinst_per_s": torch.cuda.get_device_properties(local_rank).multi_processor_count * 4 * clocks["sm_mhz"] * 1e6 / 1e9,
                             "frac": value / world * warp_inst_pair / (torch.cuda.get_device_properties(local_rank).multi_processor_count * 4 * clocks["sm_mhz"] * 1e6)}},
            "roofline_hbm": {"bound": "hbm", "achieved": value / world * algo_bytes / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": value / world * algo_bytes / 1e9 / hbm_peak, "peak_source": hbm_src,
                             "algorithmic_bytes_per_pair": algo_bytes},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
