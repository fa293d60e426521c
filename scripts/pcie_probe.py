"""Raw host<->device copy rates of this box with every rank copying at the same time: the ceiling the end-to-end
(`e2e`) numbers of bench.py run into at N > 1. Pinned host buffers, cudaMemcpyAsync on two streams per GPU.

  python scripts/pcie_probe.py                                   (one GPU)
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/pcie_probe.py"""
import json
import os
import time

import torch

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("gloo")
MB = 256
h_in = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
h_out = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
d_in = torch.empty(MB << 20, dtype=torch.uint8, device="cuda")
d_out = torch.zeros(MB << 20, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=8):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return reps * (MB << 20) / dt / 1e9  # GB/s per GPU and direction (slowest rank)


run(True, True, 2)
res = {"n_gpus": world, "buffer_mb": MB,
       "h2d_alone_gbps_per_gpu": run(True, False), "d2h_alone_gbps_per_gpu": run(False, True),
       "both_gbps_per_gpu_each_direction": run(True, True)}
res["d2h_alone_aggregate_gbps"] = res["d2h_alone_gbps_per_gpu"] * world
res["both_aggregate_gbps_each_direction"] = res["both_gbps_per_gpu_each_direction"] * world
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
