"""Fused RX chain (rate recovery + decode + CRC) throughput vs code rate: BG1 Zc=384, 16QAM, 1024 code blocks (64 TBs x 16),
8 fp32 iterations, device-resident LLRs from the fused QAM/AWGN kernel.  Prints one JSON object.
Env: RATES=0.4,0.48,... ; library knobs (NRLDPC_NO_SPLIT, NRLDPC_DEC_OCC, NRLDPC_NO_STAGE) pass through for A/B runs."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from neoradium_b200.batch import TbBatchCodec
from neoradium_b200.modulation import awgn_llr

A, numTb = 8424 * 16 - 24, 64
res = {}
for rate in [float(v) for v in os.environ.get("RATES", "0.4,0.45,0.48,0.5,0.54,0.56,0.6,0.75,0.9").split(",")]:
    G = int(round(A / rate / 64)) * 64
    codec = TbBatchCodec(1, '16QAM', A, G, precision='fp32')
    gen = torch.Generator(device='cuda'); gen.manual_seed(7)
    pl = torch.randint(0, 2, (numTb, A), dtype=torch.int8, device='cuda', generator=gen)
    snr = 9.0 + 12.0 * (rate - 0.6)          # roughly tracks the waterfall so that the blocks decode
    llrs = [awgn_llr(codec.encode(pl), 4, snr_db=snr, seed=11 + b) for b in range(4)]   # rotate inputs (> L2)
    out = codec.alloc_outputs(numTb)
    for b in range(4): codec.decode(llrs[b], 8, out=out)
    torch.cuda.synchronize()
    ok = int(out["tbOk"].sum().item())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 40
    e0.record()
    for i in range(reps): codec.decode(llrs[i % 4], 8, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res["R=%.2f" % rate] = {"G": G, "E": G // 16, "ms_per_1024_blocks": ms, "info_gbps": numTb * A / ms / 1e6, "tb_ok": ok, "snr_db": snr}
    print("R=%.2f  E=%5d  %.4f ms  %6.2f Gbit/s  tbOk %d/%d" % (rate, G // 16, ms, numTb * A / ms / 1e6, ok, numTb), flush=True)
json.dump(res, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'gpurun_out', os.environ.get("OUT", "exp_rate.json")), "w"), indent=1)
