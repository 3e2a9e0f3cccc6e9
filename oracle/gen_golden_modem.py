"""Generate tests/golden/modem_cases.npz from the UNMODIFIED reference modem (run in the build container only; needs
/root/reference): per modulation the constellation, a seeded bit stream, its symbols, noisy symbols and the max-log
LLRs of Modem.getLLRsFromSymbols (float64)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_loader import load_reference  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "modem_cases.npz")


def main():
    modulation = load_reference("modulation")
    rng = np.random.default_rng(20261017)
    d = {}
    for name, snr_db in [("BPSK", 0.0), ("QPSK", 3.0), ("16QAM", 9.0), ("64QAM", 15.0), ("256QAM", 21.0), ("1024QAM", 27.0)]:
        m = modulation.Modem(name)
        qm = m.qm
        bits = rng.integers(0, 2, (3, 40 * qm)).astype(np.int8)
        sym = m.modulate(bits)
        n0 = 10.0 ** (-snr_db / 10.0)
        noisy = sym + (rng.standard_normal(sym.shape) + 1j * rng.standard_normal(sym.shape)) * np.sqrt(n0 / 2)
        noisy[0, :4] *= 3.0          # far outside the constellation
        noisy[0, 4] = sym[0, 4]      # noise-free
        llr = m.getLLRsFromSymbols(noisy, n0)
        d[name + "_constellation"] = m.constellation
        d[name + "_bits"] = bits
        d[name + "_symbols"] = sym
        d[name + "_noisy"] = noisy
        d[name + "_n0"] = np.float64(n0)
        d[name + "_llr"] = llr
        d[name + "_hard"] = m.demodulate(noisy, n0)
    utils = load_reference("utils")
    for k, (c_init, n) in enumerate([(1, 5), (0x12345, 12), (20001 * (1 << 15) + 1 * (1 << 14) + 17, 13), (777, 43), (778, 44),
                                     (0x7FFFFFFF, 1000), (0, 200), (4660 * (1 << 15) + 42, 70000)]):
        d["gold%d_cinit_n" % k] = np.int64([c_init, n])
        d["gold%d_bits" % k] = np.packbits(np.uint8(utils.goldSequence(c_init, n)))
    np.savez_compressed(OUT, **d)
    print("modem_cases.npz", {k: v.shape for k, v in d.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
