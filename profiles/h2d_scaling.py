#!/usr/bin/env python
"""What the platform gives N ranks that copy pinned host memory to their GPUs at the same time (the ceiling of `e2e` at N GPUs:
the engine's end-to-end path moves 12 B per atom over PCIe).  Launch like the bench:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/h2d_scaling.py
Prints one JSON line: per-rank and aggregate GB/s, alone (rank 0 only) and all ranks together."""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 1 << 30
host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
host.fill_(1)
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")


def run(seconds):
    torch.cuda.synchronize()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(4):
            dev.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        n += 4
    return n * nbytes / (time.perf_counter() - t0) / 1e9


run(0.5)
alone = torch.zeros(1, device="cuda")
if world > 1:
    dist.barrier()
if rank == 0:
    alone[0] = run(2.0)
if world > 1:
    dist.barrier()
together = torch.tensor([run(3.0)], device="cuda")
if world > 1:
    allv = [torch.zeros(1, device="cuda") for _ in range(world)]
    dist.all_gather(allv, together)
    dist.broadcast(alone, 0)
else:
    allv = [together]
if rank == 0:
    per = [float(v.item()) for v in allv]
    print(json.dumps({"n_gpus": world, "h2d_gbs_rank0_alone": float(alone.item()), "h2d_gbs_per_rank_together": per, "h2d_gbs_aggregate": sum(per),
                      "cpu_count": os.cpu_count()}))
if world > 1:
    dist.destroy_process_group()
