#!/usr/bin/env python
"""Playground/HARQ/Harq.ipynb (cells 3-7) with the UNMODIFIED reference HarqEntity / Modem / random on top of the
neoradium_b200 drop-in encoder/decoder, on the GPU box ("HARQ must keep working unmodified on top of the
drop-in").  The reference modules come from baseline/_ref (pip-installed copy, git-ignored) or /root/reference.

    python scripts/run_harq_notebook.py [--transmissions 1000] [--ref-transmissions 12] [--out profiles/r02_harq_notebook.json]

Prints / stores the notebook's statistics (raw line 127 of the notebook: txBlocks per try [504 496 0 0], rxBlocks
[0 496 0 0] at Eb/N0 = 3 dB) and the per-transmission latency, next to the same loop on the reference's own NumPy
LdpcEncoder/LdpcDecoder for a few transmissions (CPU)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import ref_loader


def run(LdpcEncoder, HarqEntity, Modem, random, toLinear, n, harqType="IR", ebNoDb=3.0, seed=123, modulation="16QAM",
        codeRate=490 / 1024, txLayers=1, tbSize=10000, numProc=16):
    enc = LdpcEncoder(baseGraphNo=1, modulation=modulation, txLayers=txLayers, targetRate=codeRate)
    harq = HarqEntity(enc, harqType, numProc)
    snrDb = ebNoDb + 10 * np.log10(enc.qm * codeRate)
    noiseStd = np.sqrt(1 / toLinear(snrDb))
    rangen = random.getGenerator(seed)
    bitgen = random.getGenerator(seed + 1)
    modem = Modem(modulation)
    sizes = harq.numCW * [tbSize]
    harq.reset()
    lat = []
    for t in range(n):
        txBlocks = [bitgen.bits(sizes[c]) if harq.needNewData[c] else None for c in range(harq.numCW)]
        t0 = time.perf_counter()
        rm = harq.getRateMatchedCodeBlocks(txBlocks)
        t1 = time.perf_counter()
        llrs = []
        for c in range(harq.numCW):
            y = modem.modulate(rm[c])
            y = y + rangen.awgn(y.shape, noiseStd)
            llrs += [modem.getLLRsFromSymbols(y, noiseStd ** 2)]
        t2 = time.perf_counter()
        dec, errs = harq.decodeLLRs(llrs, sizes)
        t3 = time.perf_counter()
        lat.append((t1 - t0, t3 - t2, 0 if txBlocks[0] is not None else 1))
        harq.goNext()
    lat = np.array(lat)
    first, re = lat[lat[:, 2] == 0], lat[lat[:, 2] == 1]
    med = lambda a, k: float(np.median(a[:, k]) * 1e3) if len(a) else None
    return {"transmissions": n, "txBlocks_per_try": [int(v) for v in harq.txBlocks], "rxBlocks_per_try": [int(v) for v in harq.rxBlocks],
            "numTimeouts": int(harq.numTimeouts), "throughput_pct": float(harq.throughput), "bler_pct": float(harq.bler),
            "meanTries": float(harq.meanTries),
            "latency_ms": {"tx_first": med(first, 0), "tx_retransmission": med(re, 0), "rx_first": med(first, 1),
                           "rx_retransmission": med(re, 1)}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--transmissions", type=int, default=1000)
    ap.add_argument("--ref-transmissions", type=int, default=12)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    if not ref_loader.reference_available():
        print(json.dumps({"unavailable": "no reference copy (baseline/_ref or /root/reference)"}))
        return
    harq, modulation, rnd, utils, rldpc = ref_loader.load_reference("harq", "modulation", "random", "utils", "ldpc")
    res = {"reference_root": ref_loader.REFERENCE_ROOT, "notebook": "Playground/HARQ/Harq.ipynb cells 3-7, Eb/N0 = 3 dB, 16QAM, R = 490/1024, IR, 16 processes"}
    import torch
    if torch.cuda.is_available():
        from neoradium_b200 import LdpcEncoder
        run(LdpcEncoder, harq.HarqEntity, modulation.Modem, rnd.random, utils.toLinear, 8)     # warm-up (JIT-free, but first-call allocations)
        t0 = time.perf_counter()
        res["neoradium_b200"] = run(LdpcEncoder, harq.HarqEntity, modulation.Modem, rnd.random, utils.toLinear, args.transmissions)
        res["neoradium_b200"]["wall_s"] = time.perf_counter() - t0
        os.environ["NRLDPC_NO_MANAGED"] = "1"       # the same with explicit H2D / D2H copies of the HARQ buffers (round-1 path)
        res["neoradium_b200_no_managed"] = run(LdpcEncoder, harq.HarqEntity, modulation.Modem, rnd.random, utils.toLinear, min(200, args.transmissions))
        del os.environ["NRLDPC_NO_MANAGED"]
        # BASELINE configs[2]'s large codeword (256QAM, 4 layers, R = 0.75, A = 176 208: C = 21, Zc = 384; 4.2 MB float64 soft
        # buffer per HarqCW) below its waterfall, so that most blocks need retransmissions: per-retransmission latency with the
        # HARQ buffers resident on the device (ManagedArray) and with explicit H2D / D2H copies (round-1 path)
        big = dict(modulation="256QAM", codeRate=0.75, txLayers=4, tbSize=176208, numProc=4, ebNoDb=9.5)
        run(LdpcEncoder, harq.HarqEntity, modulation.Modem, rnd.random, utils.toLinear, 8, **big)
        res["large_slot_managed"] = run(LdpcEncoder, harq.HarqEntity, modulation.Modem, rnd.random, utils.toLinear, 48, **big)
        os.environ["NRLDPC_NO_MANAGED"] = "1"
        res["large_slot_no_managed"] = run(LdpcEncoder, harq.HarqEntity, modulation.Modem, rnd.random, utils.toLinear, 48, **big)
        del os.environ["NRLDPC_NO_MANAGED"]
    if args.ref_transmissions > 0:
        t0 = time.perf_counter()
        res["reference_numpy"] = run(rldpc.LdpcEncoder, harq.HarqEntity, modulation.Modem, rnd.random, utils.toLinear, args.ref_transmissions)
        res["reference_numpy"]["wall_s"] = time.perf_counter() - t0
    print(json.dumps(res, indent=1))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
