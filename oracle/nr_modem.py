"""CPU restatement of the reference's modem for the LDPC loops (test infrastructure; see oracle/nr_oracle.py header):
Modem.__init__ constellation (neoradium/modulation.py:60-74), modulate (:127-157), getLLRsFromSymbols (:159-204).
Pinned against the unmodified reference by tests/golden/modem_cases.npz (oracle/gen_golden_modem.py) and, when
/root/reference exists, live in tests/test_oracle_vs_reference.py."""
import numpy as np

QM = {'BPSK': 1, 'QPSK': 2, '16QAM': 4, '64QAM': 6, '256QAM': 8, '1024QAM': 10}
_SCALE = {1: 2, 2: 2, 4: 10, 6: 42, 8: 170, 10: 682}


def constellation(qm):
    """complex128 [2^qm]; label = integer with b0 as MSB (modulation.py:64-75)."""
    scale = 1 / np.sqrt(_SCALE[qm])
    pts = []
    for value in range(1 << qm):
        b = [(value >> (qm - 1 - i)) & 1 for i in range(qm)]
        real, img = 1, 1
        for q in range(2, qm, 2):
            real = (1 << (q // 2)) - (1 - 2 * b[qm - q]) * real
            img = (1 << (q // 2)) - (1 - 2 * b[qm + 1 - q]) * img
        real *= 1 - 2 * b[0]
        img *= 1 - 2 * b[min(1, qm - 1)]
        pts.append(scale * (real + 1j * img))
    return np.array(pts)


def modulate(bits, qm):
    """modulation.py:136-139"""
    b = np.asarray(bits)
    idx = (np.uint16(b).reshape((-1, qm)) * [[1 << (qm - i - 1) for i in range(qm)]]).sum(1)
    return constellation(qm)[idx].reshape(b.shape[:-1] + (b.shape[-1] // qm,))


def llrs_maxlog(symbols, qm, noise_var):
    """modulation.py:190-204 with useMax=True: full 2-D search, |.| then square, exactly as the reference."""
    symbols = np.asarray(symbols)
    con = constellation(qm)
    all_bin = np.int8([[(i >> (qm - 1 - k)) & 1 for k in range(qm)] for i in range(1 << qm)])
    c = np.int16([np.stack([np.where(all_bin[:, i] == bit)[0] for i in range(qm)], axis=1) for bit in [0, 1]])
    d = np.abs(symbols[..., None] - con)
    exponents = (-d ** 2 / noise_var)[..., c]
    lls = exponents.max(-2)
    llrs = lls[..., 0, :] - lls[..., 1, :]
    return llrs.reshape(llrs.shape[:-2] + (-1,))
