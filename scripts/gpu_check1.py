"""First GPU bring-up check: every kernel against the CPU oracle on a spread of (BG, Zc) -- run under gpurun."""
import sys, os, time, traceback
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'oracle'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
import torch
import nr_oracle as O, nr_oracle_c as OC
from neoradium_b200 import LdpcEncoder, LdpcDecoder, ChanCodeBase

rng = np.random.default_rng(7)
fails = 0
def check(name, cond, extra=""):
    global fails
    if not cond: fails += 1
    print(("PASS " if cond else "FAIL ") + name + " " + str(extra), flush=True)

def guarded(fn, *a, **k):
    global fails
    try:
        return fn(*a, **k)
    except Exception:
        fails += 1
        traceback.print_exc()
        print("FAIL exception in", fn.__name__, a, flush=True)

def crc_tests():
    for poly in O.CRC_POLYS:
        for shape in [(37,), (3, 200), (2, 8448), (1, 100000)]:
            b = rng.integers(0, 2, shape).astype(np.int8)
            check("crc %s %s" % (poly, shape), np.array_equal(ChanCodeBase.getCrc(b, poly), O.crc_remainder(b, poly)))
            a = ChanCodeBase.appendCrc(b, poly)
            check("appendCrc %s %s" % (poly, shape), np.array_equal(a, O.crc_attach(b, poly)))
            ok = ChanCodeBase.checkCrc(a, poly)
            check("checkCrc %s %s" % (poly, shape), np.all(ok))
            a2 = a.copy(); a2[..., 5] ^= 1
            check("checkCrc-bad %s %s" % (poly, shape), not np.any(ChanCodeBase.checkCrc(a2, poly)))

def chain(bg, A, mod, rate, nl=1, rv=0, sigma=0.7, numIter=5, prec='fp64'):
    tag = "bg%d A%d %s R%.2f nl%d rv%d %s" % (bg, A, mod, rate, nl, rv, prec)
    enc = LdpcEncoder(bg, mod, nl, 0, rate)
    tb = rng.integers(0, 2, A).astype(np.int8)
    g = int(np.ceil(A / rate))
    tbc = enc.appendCrc(tb, '24A')
    cbs = enc.doSegmentation(tbc)
    ocbs, p = O.segment(O.crc_attach(tb, '24A'), bg)
    check(tag + " segment", np.array_equal(cbs, ocbs), (p['C'], p['Zc'], p['F']))
    coded = enc.encode(cbs)
    ocoded = O.encode(ocbs, bg, p['Zc'], p['iLS'])
    check(tag + " encode", np.array_equal(coded, ocoded))
    full = enc.encode(cbs, puncture=False)
    check(tag + " parity", all(enc.isValidCodedBlock(f) for f in full) and not enc.isValidCodedBlock(1 - full[0]) or p['Zc'] < 4)
    rm = enc.rateMatch(coded, g, True, rv)
    orm = O.rate_match(ocoded, bg, p['Zc'], p['K'], p['F'], g, enc.qm, nl, 0, rv)
    check(tag + " rateMatch", np.array_equal(rm, orm))
    if rv == 0:
        check(tag + " chain", np.array_equal(enc.getRateMatchedCodeBlocks(tb, g), orm))
    llr = (1 - 2.0 * orm) + sigma * rng.standard_normal(len(orm)); llr = 2 * llr / sigma ** 2
    llr = llr.astype(np.float32).astype(np.float64)
    dec = LdpcDecoder(bg, mod, nl, 0, precision=prec)
    class H: pass
    h = H(); h.rv = rv; h.decBuffer = None
    rr = dec.recoverRate(llr, A, h)
    orr, obuf, p2 = O.rate_recover(llr, A, bg, enc.qm, nl, 0, rv)
    check(tag + " recoverRate", np.array_equal(rr, orr) and np.array_equal(h.decBuffer, obuf))
    # second transmission combined
    rr2 = dec.recoverRate(llr * 0.5, A, h)
    orr2, obuf2, _ = O.rate_recover(llr * 0.5, A, bg, enc.qm, nl, 0, rv, soft_buffer=obuf)
    check(tag + " recoverRate+harq", np.array_equal(rr2, orr2) and np.array_equal(h.decBuffer, obuf2))
    dt = np.float64 if prec == 'fp64' else np.float32
    bel = dec.decode(rr, numIter, False, True)
    obel = OC.decode_beliefs(orr, bg, p['Zc'], p['iLS'], numIter, dt).astype(np.float64)
    check(tag + " decode beliefs(all)", np.array_equal(bel, obel), np.abs(bel - obel).max())
    bits = dec.decode(rr, numIter)
    obits = (obel[:, :p['K']] < 0).astype(np.int8)
    check(tag + " decode bits", np.array_equal(bits, obits))
    tbm, ok = dec.checkCrcAndMerge(bits)
    otbm, ook = O.check_crc_and_merge(obits, p['K'], p['F'], p['C'])
    check(tag + " merge", np.array_equal(tbm, otbm) and list(ok) == list(ook), list(ok))
    # fused
    x = llr if prec == 'fp64' else llr.astype(np.float32)
    ftb, fcb, ftbok = dec.decodeLLRs(x, A, numIter)
    check(tag + " fused", np.array_equal(ftb, otbm[:A]) and list(fcb) == list(ook) and bool(ftbok) == bool(O.crc_check(otbm, '24A')),
          (list(fcb), ftbok))

t0 = time.time()
guarded(crc_tests)
cases = [
    (1, 10000, 'QPSK', 449 / 1024, 1, 0), (2, 3000, 'QPSK', 0.3, 1, 0), (1, 8400 * 4, '16QAM', 0.6, 1, 0),
    (1, 3000, '64QAM', 0.5, 1, 2), (2, 300, 'QPSK', 0.25, 1, 3), (1, 1200, '16QAM', 0.4, 2, 0), (2, 100, 'BPSK', 0.2, 1, 0),
    (2, 40, 'QPSK', 0.2, 1, 1), (1, 500, '256QAM', 0.7, 1, 0), (1, 20000, '256QAM', 0.8, 4, 1), (2, 3800, '1024QAM', 0.5, 1, 0),
    (1, 8424 * 3 - 24, '16QAM', 1 / 3, 1, 0), (2, 9000, 'QPSK', 0.2, 1, 0),
]
for prec in ('fp64', 'fp32'):
    for cs in cases:
        guarded(chain, *cs, prec=prec, numIter=6)
print("TOTAL FAILS", fails, "time %.1f" % (time.time() - t0))
sys.exit(1 if fails else 0)
