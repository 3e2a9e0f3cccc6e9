#!/usr/bin/env python
"""BASELINE configs[3]: ONE fixed batch of 65 536 BG1 Zc=384 code blocks (4096 transport blocks, 16QAM R=0.6) with early
termination, sharded over the ranks at transport-block granularity (STRONG scaling: the batch does not grow with N).

    python scripts/exp_cfg3_multigpu.py --out profiles/r02_config3_multigpu.jsonl                      # N = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        scripts/exp_cfg3_multigpu.py --out ...

Payloads and channel noise are functions of the GLOBAL transport-block index only, so every N decodes exactly the same
batch: the reduced counters (NCCL all-reduce of the int64[8] vector) and the iteration histogram must be identical for all
N -- scripts/check_cfg3_counters.py asserts that over the lines of the output file.  Time = max over ranks of the CUDA-event
time of the timed passes (inputs resident in HBM)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from neoradium_b200 import dist as nd
from neoradium_b200.batch import TbBatchCodec
from neoradium_b200.modulation import awgn_llr

ap = argparse.ArgumentParser()
ap.add_argument("--tbs", type=int, default=4096)
ap.add_argument("--snr", type=float, default=9.0)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--no-es", action="store_true")
ap.add_argument("--es-from", type=int, default=1)
ap.add_argument("--out", default="")
args = ap.parse_args()
rank, world, local = nd.env_rank_world()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
C = 16
A, G = 8424 * C - 24, 14040 * C
lo, hi = nd.shard_range(args.tbs, rank, world)
n = hi - lo
codec = TbBatchCodec(1, "16QAM", A, G, precision="fp32", earlyStop=not args.no_es, earlyStopFrom=args.es_from, device=dev)
sym_per_tb = codec.sumE // codec.qm
gen = torch.Generator(device=dev)
CH = 256                      # generation chunk: payload + TX chain + channel per 256 transport blocks
llr = torch.empty((n, G), dtype=torch.float32, device=dev)
payload = torch.empty((n, A), dtype=torch.int8, device=dev)
for c0 in range(0, n, CH):
    c1 = min(n, c0 + CH)
    for t in range(c0, c1):   # payload bits depend on the global transport-block index only
        gen.manual_seed(1000003 * 17 + lo + t)
        payload[t] = torch.randint(0, 2, (A,), dtype=torch.int8, device=dev, generator=gen)
    llr[c0:c1] = awgn_llr(codec.encode(payload[c0:c1]), codec.qm, snr_db=args.snr, seed=99, offset=(lo + c0) * sym_per_tb)
out = codec.alloc_outputs(n)
for _ in range(2):
    codec.decode(llr, 8, out=out)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    codec.decode(llr, 8, out=out)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
counters = torch.zeros(len(nd.COUNTER_NAMES), dtype=torch.int64, device=dev)
codec.accumulate(out, counters, refPayload=payload)
hist = torch.bincount(out["iters"].reshape(-1).long(), minlength=9)[:9].to(torch.int64)
ms_all = [torch.zeros_like(ms) for _ in range(world)]
hist_all = [torch.zeros_like(hist) for _ in range(world)]
if world > 1:
    dist.all_gather(ms_all, ms)
    dist.all_gather(hist_all, hist)
    nd.reduce_counters(counters)          # the ONE collective of the path: an int64[8] all-reduce over NCCL
else:
    ms_all, hist_all = [ms], [hist]
if rank == 0:
    ms_max = max(float(x.item()) for x in ms_all)
    d = nd.counters_dict(counters)
    line = {"config": "BASELINE configs[3]: %d code blocks BG1 Zc=384 16QAM R=0.6, Es/N0 %.1f dB, early termination %s, 8 iterations max, fp32"
                      % (args.tbs * C, args.snr, "off" if args.no_es else "on (from iteration %d)" % args.es_from),
            "world": world, "scaling": "strong", "tbs_total": args.tbs, "ms_max_over_ranks": ms_max,
            "ms_per_rank": [float(x.item()) for x in ms_all], "gbps": args.tbs * A / ms_max / 1e6,
            "counters": {k: d[k] for k in nd.COUNTER_NAMES[:6]}, "mean_iterations": d["meanIterations"],
            "iteration_histogram_total": [int(v) for v in torch.stack(hist_all).sum(0).tolist()],
            "iteration_histogram_per_rank": [[int(v) for v in h.tolist()] for h in hist_all]}
    print(json.dumps(line))
    if args.out:
        with open(os.path.join(ROOT, args.out) if not os.path.isabs(args.out) else args.out, "a") as f:
            f.write(json.dumps(line) + "\n")
if world > 1:
    dist.destroy_process_group()
