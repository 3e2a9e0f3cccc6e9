#!/usr/bin/env python
"""Aggregate pinned-host -> device copy rate of the box with N ranks copying at once (the ceiling of every host-buffer e2e
figure at N GPUs).  torchrun --nproc-per-node N scripts/h2d_ceiling.py [--affinity]  -> one JSON line on rank 0.
--affinity pins each rank to its own slice of the host cores before it allocates its pinned buffer (first-touch locality)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--affinity", action="store_true")
ap.add_argument("--mb", type=int, default=256)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if args.affinity:
    cpus = sorted(os.sched_getaffinity(0))
    per = max(1, len(cpus) // world)
    os.sched_setaffinity(0, set(cpus[local * per:(local + 1) * per]))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = args.mb << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device=dev)
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record()
for _ in range(reps):
    d.copy_(h, non_blocking=True)
e1.record()
torch.cuda.synchronize()
gbs = torch.tensor([n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9], dtype=torch.float64, device=dev)
all_ = [torch.zeros_like(gbs) for _ in range(world)]
if world > 1:
    dist.all_gather(all_, gbs)
else:
    all_ = [gbs]
if rank == 0:
    v = [float(x.item()) for x in all_]
    print(json.dumps({"world": world, "affinity": args.affinity, "host_cpus": len(os.sched_getaffinity(0)) if not args.affinity else None,
                      "per_rank_GBps": [round(x, 2) for x in v], "aggregate_GBps": round(sum(v), 2), "min_GBps": round(min(v), 2)}))
if world > 1:
    dist.destroy_process_group()
