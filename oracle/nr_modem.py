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


def gold_sequence(c_init, num_bits):
    """goldSequence of neoradium/utils.py:70-94: both 31-bit LFSRs of TS 38.211 5.2.1 advanced one 31-bit word per
    step; word 51 yields c(0..11) from its 12 top bits, every later word 31 bits LSB first."""
    x1, x2 = 0x42054D21, int(c_init)
    for _ in range(51):
        x2 ^= (x2 >> 3) ^ (x2 >> 2) ^ (x2 >> 1)
        x2 ^= ((x2 << 28) ^ (x2 << 29) ^ (x2 << 30)) & 0x7FFFFFFF
    c = x1 ^ x2
    bits = [(c >> i) & 1 for i in range(19, 31)]
    remaining = num_bits - 12
    while remaining > 0:
        x1 ^= x1 >> 3
        x1 ^= (x1 << 28) & 0x7FFFFFFF
        x2 ^= (x2 >> 3) ^ (x2 >> 2) ^ (x2 >> 1)
        x2 ^= ((x2 << 28) ^ (x2 << 29) ^ (x2 << 30)) & 0x7FFFFFFF
        c = x1 ^ x2
        bits += [(c >> i) & 1 for i in range(31)]
        remaining -= 31
    return bits[:num_bits]


def gold_sequence_38211(c_init, num_bits):
    """Independent check: the bit-serial definition of TS 38.211 5.2.1 (Nc = 1600)."""
    n_total = 1600 + num_bits
    x1 = [0] * (n_total + 31)
    x2 = [0] * (n_total + 31)
    x1[0] = 1
    for i in range(31):
        x2[i] = (c_init >> i) & 1
    for n in range(n_total):
        x1[n + 31] = (x1[n + 3] + x1[n]) & 1
        x2[n + 31] = (x2[n + 3] + x2[n + 2] + x2[n + 1] + x2[n]) & 1
    return [(x1[n + 1600] + x2[n + 1600]) & 1 for n in range(num_bits)]
