#!/usr/bin/env python
"""Quick device-resident A/B timing of the fused RX chain (bench.py's workload): single stream and two streams, CUDA events.
   NRLDPC_LIB=/path/to/other/libnrldpc.so python scripts/ab_quick.py [--tbs 64] [--steps 40] [--es] [--snr 9.0]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neoradium_b200.batch import TbBatchCodec
from neoradium_b200.modulation import awgn_llr

ap = argparse.ArgumentParser()
ap.add_argument("--tbs", type=int, default=64)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--es", action="store_true")
ap.add_argument("--snr", type=float, default=9.0)
ap.add_argument("--rate", type=float, default=0.6)
ap.add_argument("--tag", default="")
ap.add_argument("--es-from", default="1")   # iteration number or "auto"
args = ap.parse_args()
C = 16
A = 8424 * C - 24
E = int(round(8424 / args.rate / 4)) * 4
G = E * C
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
mk = lambda own: TbBatchCodec(1, "16QAM", A, G, precision="fp32", device=dev, ownHandle=own, earlyStop=args.es, earlyStopFrom=(args.es_from if args.es_from == "auto" else int(args.es_from)))
codec = mk(False)
gen = torch.Generator(device=dev)
gen.manual_seed(1)
NB = 4 if args.tbs <= 256 else 2
pls, llrs = [], []
for b in range(NB):
    pl = torch.randint(0, 2, (args.tbs, A), dtype=torch.int8, device=dev, generator=gen)
    llrs.append(awgn_llr(codec.encode(pl), 4, snr_db=args.snr, seed=77 + b, offset=0))
    pls.append(pl)
out = codec.alloc_outputs(args.tbs)
codec.decode(llrs[0], 8, out=out)
torch.cuda.synchronize()
ok = int(out["tbOk"].sum().item())
err = int((out["tb"][:, :A] != pls[0]).sum().item())
for i in range(5):
    codec.decode(llrs[i % NB], 8, out=out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for i in range(args.steps):
    codec.decode(llrs[i % NB], 8, out=out)
e1.record()
torch.cuda.synchronize()
ms1 = e0.elapsed_time(e1) / args.steps
codecs = [mk(True) for _ in range(2)]
outs = [c.alloc_outputs(args.tbs) for c in codecs]
streams = [torch.cuda.Stream(dev) for _ in range(2)]
cur = torch.cuda.current_stream()


def run(n):
    for s in streams:
        s.wait_stream(cur)
    for i in range(n):
        with torch.cuda.stream(streams[i % 2]):
            codecs[i % 2].decode(llrs[i % NB], 8, out=outs[i % 2])
    for s in streams:
        cur.wait_stream(s)


run(6)
torch.cuda.synchronize()
e0.record()
run(args.steps)
e1.record()
torch.cuda.synchronize()
ms2 = e0.elapsed_time(e1) / args.steps
bits = args.tbs * A
print(json.dumps({"tag": args.tag, "lib": os.environ.get("NRLDPC_LIB", "default"), "blocks": args.tbs * C, "rate": args.rate, "es": args.es,
                  "snr": args.snr, "single_ms": round(ms1, 4), "single_gbps": round(bits / ms1 / 1e6, 3),
                  "two_stream_gbps": round(bits / ms2 / 1e6, 3), "tb_ok": ok, "tbs": args.tbs, "bit_err": err,
                  "mean_iters": float(out["iters"].float().mean().item())}))
