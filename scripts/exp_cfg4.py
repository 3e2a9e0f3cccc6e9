"""BASELINE configs[3]: 65 536 BG1 Zc=384 code blocks on one GPU, with and without early termination."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neoradium_b200.batch import TbBatchCodec
from neoradium_b200.modulation import awgn_llr
C = 16
A, G = 8424 * C - 24, 14040 * C
dev = torch.device("cuda", 0)
tbs = int(os.environ.get("TBS", "4096"))
gen = torch.Generator(device=dev); gen.manual_seed(5)
res = {}
for snr in (10.5, 9.0, 8.6):
    codec0 = TbBatchCodec(1, "16QAM", A, G, precision="fp32", device=dev)
    pl = torch.randint(0, 2, (tbs, A), dtype=torch.int8, device=dev, generator=gen)
    llr = awgn_llr(codec0.encode(pl), 4, snr_db=snr, seed=77, offset=0)
    for es in (False, True):
        codec = TbBatchCodec(1, "16QAM", A, G, precision="fp32", earlyStop=es, device=dev, earlyStopFrom=int(os.environ.get("ES_FROM", "1")))   # ES_FROM=k: test the syndrome from iteration k on
        out = codec.alloc_outputs(tbs)
        for _ in range(2):
            codec.decode(llr, 8, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            codec.decode(llr, 8, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        it = out["iters"].float().mean().item()
        ok = int(out["tbOk"].sum().item())
        be = int((out["tb"][:, :A] != pl).sum().item())
        print("snr %.1f early_stop=%s: %.3f ms  %.2f Gbit/s  mean iters %.2f  tbOk %d/%d bitErr %d" % (snr, es, ms, tbs * A / ms / 1e6, it, ok, tbs, be), flush=True)
        res["snr%.1f_es%d" % (snr, es)] = dict(ms=ms, gbps=tbs * A / ms / 1e6, mean_iters=it, tb_ok=ok, bit_err=be)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "exp_cfg4.json"), "w"))
