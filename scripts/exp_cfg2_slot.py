#!/usr/bin/env python
"""BASELINE configs[2] at full size: the 51-RB 256QAM slot (codeword 1: BG1, 4 layers, R=0.75, A=176 208, C=21; codeword 2:
BG2, 2 layers, R=0.3, A=37 896, C=10; both Zc=384) with device-resident HARQ soft buffers, decoded (a) by ONE grouped library
call (nrldpc_decode_tb_groups: descriptor array, groups run concurrently) and (b) by one call per group (the round-1 Python
loop).  Reports the latency of one slot and the throughput of batches of slots.  -> gpurun_out/exp_cfg2_slot.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neoradium_b200.batch import TbBatchCodec, decode_groups
from neoradium_b200.modulation import awgn_llr

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
cw = [dict(bg=1, A=176208, nl=4, g=235008, snr=21.5), dict(bg=2, A=37896, nl=2, g=127296, snr=9.5)]
res = {"config": "BASELINE configs[2]: 256QAM slot, BG1 C=21 (A=176208, 4 layers, R=0.75) + BG2 C=10 (A=37896, 2 layers, R=0.3), Zc=384, "
                 "fp32, 8 iterations, device-resident soft buffers (HARQ combine in the fused load)", "points": []}
gen = torch.Generator(device=dev)
gen.manual_seed(3)
for slots in (1, 8, 64):
    codecs = [TbBatchCodec(c["bg"], '256QAM', c["A"], c["g"], txLayers=c["nl"], rv=0, precision='fp32', device=dev) for c in cw]
    llr, pl, soft = [], [], []
    for c, k in zip(cw, codecs):
        p = torch.randint(0, 2, (slots, c["A"]), dtype=torch.int8, device=dev, generator=gen)
        llr.append(awgn_llr(k.encode(p), 8, snr_db=c["snr"], seed=5, offset=0))
        pl.append(p)
        soft.append(torch.zeros((slots * k.C, k.ncb - k.F), dtype=torch.float32, device=dev))
    outs = [k.alloc_outputs(slots) for k in codecs]
    info_bits = slots * sum(c["A"] for c in cw)

    def grouped():
        for s in soft:
            s.zero_()
        decode_groups(codecs, llr, 8, outs=outs, softBuffers=soft)

    def looped():
        for s in soft:
            s.zero_()
        for k, x, o, s in zip(codecs, llr, outs, soft):
            k.decode(x, 8, out=o, softBuffer=s)

    point = {"slots": slots, "code_blocks": slots * 31}
    for name, fn in (("grouped_call", grouped), ("one_call_per_group", looped)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        ok = all(bool(o["tbOk"].all().item()) for o in outs) and all(torch.equal(o["tb"][:, :c["A"]], p) for o, c, p in zip(outs, cw, pl))
        reps = 50 if slots <= 8 else 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / reps * 1e3
        ms = e0.elapsed_time(e1) / reps
        # latency of ONE isolated call: host submit -> results visible (synchronised each time)
        lat = []
        for _ in range(20):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            lat.append((time.perf_counter() - t0) * 1e3)
        lat.sort()
        point[name] = {"device_ms": ms, "wall_ms": wall, "isolated_call_latency_ms_median": lat[len(lat) // 2],
                       "gbps": info_bits / ms / 1e6, "decoded_ok": ok}
    res["points"].append(point)
    print(json.dumps(point), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "exp_cfg2_slot.json"), "w"), indent=1)
