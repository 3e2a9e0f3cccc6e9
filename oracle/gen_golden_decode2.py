"""Generate tests/golden/decode2_cases.npz (build container only; needs /root/reference): inputs and float64 outputs
of the UNMODIFIED reference's LdpcDecoder.decode2 (ldpc.py:1421-1492) on seeded inputs, with stopOnGoodParity False
(the pure recursion) and True (the reference's first-row-only stop test), for a few small lifting sizes and alphas."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "decode2_cases.npz")
# (name, bg, A, modulation, rate, sigma, maxIter, alpha)
CASES = [("bg2_z6", 2, 20, "QPSK", 0.3, 0.9, 4, 0.75), ("bg2_z16", 2, 120, "QPSK", 0.25, 1.0, 5, 0.8),
         ("bg1_z10", 1, 150, "QPSK", 0.4, 0.85, 3, 0.75), ("bg2_z13_zeros", 2, 100, "QPSK", 0.3, 0.0, 3, 0.7)]


def main():
    ldpc = load_reference("ldpc")
    d, names = {}, []
    for name, bg, A, mod, rate, sigma, nit, alpha in CASES:
        rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
        rng = np.random.default_rng(sum(map(ord, name)))
        enc = ldpc.LdpcEncoder(bg, mod, 1, 0, rate)
        tb = rng.integers(0, 2, A).astype(np.int8)
        g = int(np.ceil(A / rate))
        rm = enc.getRateMatchedCodeBlocks(tb, g)
        llr = (1 - 2.0 * rm) * 2 + sigma * 2 * rng.standard_normal(len(rm))
        if sigma == 0.0:
            llr[::3] = 0.0      # exact zeros: the min1 == 0 / min2 == 0 branches (ldpc.py:1478-1483)
            llr[5:40] = 0.0
        llr = llr.astype(np.float32).astype(np.float64)
        dec = enc.getDecoder()
        rr = dec.recoverRate(llr, A)
        bel = dec.decode2(rr, nit, False, True, alpha, False)
        dec2 = enc.getDecoder()
        rr2 = dec2.recoverRate(llr, A)
        bel_stop = dec2.decode2(rr2, nit + 4, False, True, alpha, True)
        d[name + "/rr"], d[name + "/bel"], d[name + "/bel_stop"] = rr, bel, bel_stop
        d[name + "/meta"] = np.array([bg, A, enc.liftingSize, enc.setIndex, nit, enc.codeBlockSize], np.int64)
        d[name + "/alpha"] = np.float64(alpha)
        names.append(name)
        print(name, rr.shape, "Zc", enc.liftingSize, "bit errors after decode2:", int(np.sum((bel[:, :enc.codeBlockSize] < 0).reshape(-1)[:A] != tb)))
    d["names"] = np.array(names)
    np.savez_compressed(OUT, **d)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
