"""Per-call latency of the drop-in API on single transport blocks (host NumPy arrays in and out, as the notebooks use it):
BASELINE configs[0] (BG2, QPSK, R~0.3, one code block) and the PDSCH notebook case (BG1, 16QAM, 2 layers, R=490/1024,
4 code blocks, numIter 5 and 20).  The SAME script times either implementation:
    python scripts/bench_dropin_calls.py            -> neoradium_b200 on cuda:0 (run under gpurun)
    python scripts/bench_dropin_calls.py reference  -> the unmodified reference on the host CPU (container only: it needs
                                                       /root/reference; oracle/ref_loader.py is the import shim)
Prints one JSON object; medians over REPS calls after one warm-up call."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
which = sys.argv[1] if len(sys.argv) > 1 else "ours"
if which == "reference":
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from ref_loader import load_reference
    ldpc = load_reference("ldpc")
    LdpcEncoder, LdpcDecoder = ldpc.LdpcEncoder, ldpc.LdpcDecoder
    REPS = 3
else:
    from neoradium_b200 import LdpcEncoder, LdpcDecoder
    REPS = 30


def med(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        r = fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3, r


cases = [("configs[0]: BG2 QPSK R=0.3, A=3000, G=10000", 2, 'QPSK', 1, 0.3, 3000, 10000, (5,)),
         ("PDSCH notebook: BG1 16QAM 2 layers R=490/1024, A=30000", 1, '16QAM', 2, 490 / 1024, 30000, 62692, (5, 20))]
out = {"impl": which, "cases": {}}
rng = np.random.default_rng(3)
for name, bg, mod, nl, rate, A, G, iters in cases:
    enc = LdpcEncoder(baseGraphNo=bg, modulation=mod, txLayers=nl, targetRate=rate)
    dec = enc.getDecoder()
    tb = rng.integers(0, 2, A).astype(np.int8)
    res = {}
    res["getRateMatchedCodeBlocks_ms"], rm = med(lambda: enc.getRateMatchedCodeBlocks(tb, G), REPS)
    llr = (1.0 - 2.0 * np.asarray(rm, np.float64)) * 4 + rng.standard_normal(len(rm)) * 1.6
    res["recoverRate_ms"], rr = med(lambda: dec.recoverRate(llr, A), REPS)
    for it in iters:
        res["decode_numIter%d_ms" % it], bits = med(lambda: dec.decode(rr, numIter=it), REPS)
    res["checkCrcAndMerge_ms"], (tbo, crc) = med(lambda: dec.checkCrcAndMerge(bits), REPS)
    res["checkCrc_24A_ms"], ok = med(lambda: dec.checkCrc(tbo, '24A'), REPS)
    res["payload_recovered"] = bool(np.array_equal(np.asarray(tbo)[:A], tb))
    res["C"], res["Zc"] = int(enc.numCodeBlocks), int(enc.liftingSize)
    out["cases"][name] = res
    print(name, res, file=sys.stderr, flush=True)
print(json.dumps(out))
