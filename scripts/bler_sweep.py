"""BLER-vs-SNR Monte-Carlo sweep (BASELINE configs[4], reduced block count by default) sharded over the GPUs of one box:
every rank generates, encodes, modulates, decodes and counts on its own device; one NCCL all-reduce of the int64[8]
counter vector per SNR point.  torchrun --nproc-per-node N scripts/bler_sweep.py [--tbs 2048] [--early-stop]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, torch.distributed as dist
from neoradium_b200 import dist as nd
from neoradium_b200.batch import TbBatchCodec

ap = argparse.ArgumentParser()
ap.add_argument("--tbs", type=int, default=2048)
ap.add_argument("--bg", type=int, default=1)
ap.add_argument("--mod", default="16QAM")
ap.add_argument("--A", type=int, default=8424 * 4 - 24)
ap.add_argument("--rate", type=float, default=0.6)
ap.add_argument("--iters", type=int, default=8)
ap.add_argument("--snrs", default="7.0,7.4,7.8,8.0,8.2,8.4,8.6,8.8,9.0,9.4")
ap.add_argument("--early-stop", action="store_true")
ap.add_argument("--batch-tbs", type=int, default=256)
ap.add_argument("--out", default="")
args = ap.parse_args()
rank, world, local = nd.env_rank_world()
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
qm = {'QPSK': 2, '16QAM': 4, '64QAM': 6, '256QAM': 8}[args.mod]
g = int(-(-args.A / args.rate // qm) * qm)
codec = TbBatchCodec(args.bg, args.mod, args.A, g, precision='fp32', earlyStop=args.early_stop)
from neoradium_b200.sweep import BlerSweep
sweep = BlerSweep(codec, numIter=args.iters, tbsPerPoint=args.tbs, batchTbs=args.batch_tbs, seed=1)
clock = [time.perf_counter()]


def report(d):
    torch.cuda.synchronize()
    now = time.perf_counter()
    d["seconds"] = now - clock[0]
    clock[0] = now
    if rank == 0:
        print("SNR %.1f dB  TBs %d  BLER %.4f  CB-BLER %.4f  BER %.2e  mean iters %.2f  %.2fs" % (
            d["snr_db"], d["txBlocks"], d["bler"], d["cbler"], d["bitErrors"] / (d["txBlocks"] * args.A), d["meanIterations"],
            d["seconds"]), flush=True)


res = sweep.run([float(x) for x in args.snrs.split(",")], on_point=report)
if rank == 0:
    line = json.dumps({"config": vars(args), "world": world, "C": codec.C, "Zc": codec.Zc, "K": codec.K, "F": codec.F,
                       "E": codec.lens[0], "points": res})
    print(line)
    if args.out:
        with open(args.out, "a") as f:
            f.write(line + "\n")
if world > 1:
    dist.destroy_process_group()
