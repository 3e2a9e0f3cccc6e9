import sys, os, faulthandler
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import nr_oracle as O
from neoradium_b200 import LdpcDecoder, LdpcEncoder
faulthandler.dump_traceback_later(40, exit=True)
bg, A, mod, rate, nl, rv = 1, 3000, '64QAM', 0.5, 1, 2
class Harq:
    def __init__(self, rv=0): self.rv, self.decBuffer = rv, None
rng = np.random.default_rng(A * 7 + rv)
enc = LdpcEncoder(bg, mod, nl, 0, rate)
tb = rng.integers(0, 2, A).astype(np.int8)
g = int(np.ceil(A / rate))
coded = enc.encode(enc.doSegmentation(enc.appendCrc(tb, '24A')))
rm = enc.rateMatch(coded, g, True, rv)
sigma = 0.7
llr = (2 * ((1 - 2.0 * rm) + sigma * rng.standard_normal(len(rm))) / sigma ** 2).astype(np.float32).astype(np.float64)
dec = LdpcDecoder(bg, mod, nl, 0, precision='fp32')
h = Harq(rv)
rr = dec.recoverRate(llr, A, h)
print("Z", dec.liftingSize, "rr", rr.shape, flush=True)
for nit in (0, 1, 6):
    print("decode beliefs nit", nit, flush=True)
    bel = dec.decode(rr, nit, False, True)
    torch.cuda.synchronize()
    print("  ok", bel.shape, flush=True)
print("decode bits", flush=True)
bits = dec.decode(rr, 6); torch.cuda.synchronize(); print("  ok", flush=True)
print("fused softbuf", flush=True)
ftb, fcb, ftbok = dec.decodeLLRs(llr.astype(np.float32), A, 6, harq=Harq(rv)); torch.cuda.synchronize(); print("  ok", ftbok, flush=True)
print("fused no softbuf", flush=True)
ftb, fcb, ftbok = dec.decodeLLRs(llr.astype(np.float32), A, 6); torch.cuda.synchronize(); print("  ok", ftbok, flush=True)
# more multi-block static cases (bounded by the watchdog above): several blocks per CTA with a partly filled last group, C > 1
# (CRC24B + in-kernel TB CRC), soft buffers, LBRM, fp16 input
faulthandler.cancel_dump_traceback_later()
faulthandler.dump_traceback_later(60, exit=True)
from neoradium_b200.batch import TbBatchCodec
import nr_oracle_c as OC
for (bg2, A2, mod2, g2, numTb, nref, rv2) in [(2, 500, 'QPSK', 1668, 23, 0, 0), (1, 600, '16QAM', 1200, 7, 0, 1), (2, 24, 'QPSK', 100, 130, 0, 0),
                                              (1, 9000, '16QAM', 18000, 5, 0, 0), (2, 100, 'QPSK', 600, 9, 400, 3), (1, 8424 * 2 - 24 - 4000, '16QAM', 25000, 3, 0, 2)]:
    codec = TbBatchCodec(bg2, mod2, A2, g2, 1, nref, rv2, 'fp32')
    print("codec", bg2, A2, "Z", codec.Zc, "C", codec.C, "numTb", numTb, flush=True)
    pl = codec.random_payload(numTb, 3)
    rmb = codec.encode(pl)
    from neoradium_b200.modulation import awgn_llr
    x = awgn_llr(rmb, codec.qm, snr_db=6.0, seed=1)
    out = codec.decode(x, 6); torch.cuda.synchronize()
    soft = torch.zeros((numTb * codec.C, codec.ncb - codec.F), dtype=torch.float32, device='cuda')
    out2 = codec.decode(x, 6, softBuffer=soft); torch.cuda.synchronize()
    assert torch.equal(out['tb'], out2['tb']) and torch.equal(out['tbOk'], out2['tbOk'])
    os.environ["NRLDPC_NO_STATIC_MB"] = "1"
    ref = TbBatchCodec(bg2, mod2, A2, g2, 1, nref, rv2, 'fp32', ownHandle=True).decode(x, 6); torch.cuda.synchronize()
    del os.environ["NRLDPC_NO_STATIC_MB"]
    for k in ('tb', 'cbOk', 'tbOk'):
        assert torch.equal(out[k], ref[k]), k
    out3 = codec.decode(x.half(), 6); torch.cuda.synchronize()
    print("   ok  tbOk %d/%d" % (int(out['tbOk'].sum()), numTb), flush=True)
print("ALL MB CASES OK", flush=True)
