"""Generate tests/golden/*.npz (run in the build container only; needs /root/reference).

  matlab_ldpc.npz   the seven MATLAB 5G-Toolbox vectors the reference ships for this path
                    (Playground/CompareWithMatlab/LDPC/MatlabFiles/*.mat; asserted bit-exact by the reference's own
                    LDPC-Matlab.ipynb) plus the Polar notebook's CRC24C vector (Polar/MatlabFiles/msg.mat, msgcrc.mat)
  ref_cases.npz     inputs and outputs of the UNMODIFIED reference (float64, as shipped) on seeded inputs for a spread
                    of configurations the reference's own tests do not pin: BG2, Zc=384, the iLS=4 `880` entry,
                    rv != 0 with F > 0, E > Ncb wrap, LBRM nRef, C == 1 vs C > 1, HARQ combining, every CRC polynomial,
                    outputBelief.
The fixtures are small (< 1.5 MB) and travel with the repository; the reference does not.
"""
import os
import sys

import numpy as np
import scipy.io

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import REFERENCE_ROOT, load_reference  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")

# (name, bg, A, modulation, rate, layers, nRef, rv sequence, sigma, numIter)
CASES = [
    ("bg1_z240_matlabcfg", 1, 10000, "QPSK", 449 / 1024, 1, 0, [0], 0.75, 5),
    ("bg2_z320_cfg1", 2, 3000, "QPSK", 0.3, 1, 0, [0], 1.05, 8),
    ("bg1_z384_c2", 1, 8424 * 2 - 24, "16QAM", 0.6, 1, 0, [0], 0.62, 8),
    ("bg2_z384_c2", 2, 3816 * 2 - 24, "QPSK", 0.3, 1, 0, [0], 1.0, 6),
    ("bg1_z144_ils4_880", 1, 3000, "64QAM", 0.5, 1, 0, [0, 2], 0.8, 5),
    ("bg1_z36_ils4_880", 1, 700, "QPSK", 0.4, 1, 0, [0], 0.9, 5),
    ("bg2_z44_rv3_fill", 2, 300, "QPSK", 0.25, 1, 0, [3, 0], 1.0, 5),
    ("bg1_z56_2layers", 1, 1200, "16QAM", 0.4, 2, 0, [0, 1], 0.8, 4),
    ("bg2_z22_bpsk_wrap", 2, 100, "BPSK", 0.1, 1, 0, [0], 1.2, 5),      # E > Ncb: wrap-around accumulation
    ("bg1_z208_256qam_hi", 1, 4500, "256QAM", 0.85, 1, 0, [0, 2, 3, 1], 0.45, 6),
    ("bg1_z112_lbrm", 1, 2400, "QPSK", 0.5, 1, 7392, [0, 2], 0.8, 5),   # nRef = N (LBRM limit equal to N)
    ("bg2_z6_tiny", 2, 20, "QPSK", 0.3, 1, 0, [0], 0.9, 5),
]


def matlab_vectors():
    p = os.path.join(REFERENCE_ROOT, "Playground", "CompareWithMatlab", "LDPC", "MatlabFiles")
    d = {}
    for name, key in [("in", "in"), ("cbsIn", "cbsIn"), ("enc", "enc"), ("chIn", "chIn"), ("raterec", "raterec"),
                      ("decBits", "decBits"), ("decBlk", "decBlk")]:
        d[name] = scipy.io.loadmat(os.path.join(p, name + ".mat"))[key]
    pp = os.path.join(REFERENCE_ROOT, "Playground", "CompareWithMatlab", "Polar", "MatlabFiles")
    d["polar_msg"] = scipy.io.loadmat(os.path.join(pp, "msg.mat"))["msg"]
    d["polar_msgcrc"] = scipy.io.loadmat(os.path.join(pp, "msgcrc.mat"))["msgcrc"]
    np.savez_compressed(os.path.join(OUT, "matlab_ldpc.npz"), **d)
    print("matlab_ldpc.npz", {k: v.shape for k, v in d.items()})


class _Harq:
    def __init__(self):
        self.rv, self.decBuffer = 0, None


def reference_cases():
    ldpc, cc = load_reference("ldpc", "chancodebase")
    out = {}
    rng = np.random.default_rng(20261017)
    names = []
    for (name, bg, A, mod, rate, nl, nref, rvs, sigma, nit) in CASES:
        enc = ldpc.LdpcEncoder(bg, mod, nl, nref, rate)
        tb = rng.integers(0, 2, A).astype(np.int8)
        g = int(np.ceil(A / rate))
        tbc = enc.appendCrc(tb, "24A")
        cbs = enc.doSegmentation(tbc)
        coded = enc.encode(cbs)
        dec = ldpc.LdpcDecoder(bg, mod, nl, nref)
        h = _Harq()
        pre = "%s/" % name
        out[pre + "meta"] = np.array([bg, A, enc.qm, nl, nref, g, nit, enc.numCodeBlocks, enc.liftingSize,
                                      enc.setIndex, enc.codeBlockSize, enc.numFillerBits], np.int64)
        out[pre + "tb"] = tb
        out[pre + "cbs"] = cbs.astype(np.int8)
        out[pre + "coded"] = coded.astype(np.int8)
        out[pre + "rvs"] = np.array(rvs, np.int64)
        for t, rv in enumerate(rvs):
            rm = enc.rateMatch(coded, g, True, rv)
            llr = (1 - 2.0 * rm) + sigma * rng.standard_normal(len(rm))
            llr = (2 * llr / sigma ** 2).astype(np.float32)     # fp32 LLRs, widened to float64 for the reference
            if t == 0 and len(llr) > 40:
                llr[7] = 0.0
                llr[11] = -0.0
            h.rv = rv
            rr = dec.recoverRate(llr.astype(np.float64), A, h)
            bel = dec.decode(rr, nit, False, True)
            bits = dec.decode(rr, nit)
            tbm, ok = dec.checkCrcAndMerge(bits)
            tbok = dec.checkCrc(tbm, "24A")
            out[pre + "rm%d" % t] = rm.astype(np.int8)
            out[pre + "llr%d" % t] = llr
            out[pre + "decbuf%d" % t] = h.decBuffer.copy()
            out[pre + "bel%d" % t] = bel
            out[pre + "merged%d" % t] = tbm.astype(np.int8)
            out[pre + "cbok%d" % t] = np.asarray(ok, bool)
            out[pre + "tbok%d" % t] = np.array(bool(tbok))
            print(name, "rv", rv, "C", enc.numCodeBlocks, "Zc", enc.liftingSize, "F", enc.numFillerBits, "E",
                  dec.getRateMatchedCbLens(len(llr), enc.numCodeBlocks)[:2], "cbok", list(ok), "tbok", bool(tbok))
        names.append(name)
    # CRC known answers for all six polynomials, 1-D and 2-D, incl. the probe of SURVEY 8a (0000 0001 -> poly bits)
    for poly in ("6", "11", "16", "24A", "24B", "24C"):
        b1 = rng.integers(0, 2, 333).astype(np.int8)
        b2 = rng.integers(0, 2, (3, 64)).astype(np.int8)
        one = np.array([0, 0, 0, 0, 0, 0, 0, 1], np.int8)
        out["crc/%s/in1" % poly], out["crc/%s/out1" % poly] = b1, cc.ChanCodeBase.getCrc(b1, poly).astype(np.int8)
        out["crc/%s/in2" % poly], out["crc/%s/out2" % poly] = b2, cc.ChanCodeBase.getCrc(b2, poly).astype(np.int8)
        out["crc/%s/one" % poly] = cc.ChanCodeBase.getCrc(one, poly).astype(np.int8)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "ref_cases.npz"), **out)
    print("ref_cases.npz written:", os.path.getsize(os.path.join(OUT, "ref_cases.npz")) // 1024, "KiB")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    matlab_vectors()
    reference_cases()
