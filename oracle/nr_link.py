"""Synthetic link for workload generation on the CPU side (test infrastructure): Gray QAM of TS 38.211 5.1
(neoradium/modulation.py:60-74), complex AWGN at a given Es/N0, max-log LLRs (modulation.py:190-204; positive => 0).
Used by the CPU baseline / reference arm of bench.py and by tests to synthesise the decoder's input."""
import numpy as np

_SCALE = {1: 2, 2: 2, 4: 10, 6: 42, 8: 170, 10: 682}


def _pam(bb):
    half = bb.shape[1]
    a = np.ones(bb.shape[0])
    for q in range(half - 1, 0, -1):
        a = (1 << (half - q)) - (1 - 2 * bb[:, q]) * a
    return (1 - 2 * bb[:, 0]) * a


def qam_awgn_llr(bits, qm, snr_db, rng, dtype=np.float32):
    assert qm in (2, 4, 6, 8, 10)
    half = qm // 2
    b = np.asarray(bits).reshape(-1, qm).astype(np.float64)
    scale = 1.0 / np.sqrt(_SCALE[qm])
    n0 = 10.0 ** (-snr_db / 10.0)
    re, im = _pam(b[:, 0::2]) * scale, _pam(b[:, 1::2]) * scale
    noise = rng.standard_normal((2, len(re))) * np.sqrt(n0 / 2)
    lv = np.arange(1 << half)
    lb = ((lv[:, None] >> np.arange(half - 1, -1, -1)[None, :]) & 1).astype(np.float64)
    levels = _pam(lb) * scale
    out = np.empty((len(re), qm))
    for comp, y in ((0, re + noise[0]), (1, im + noise[1])):
        d2 = (y[:, None] - levels[None, :]) ** 2
        for q in range(half):
            m0 = np.where(lb[None, :, q] == 0, d2, np.inf).min(1)
            m1 = np.where(lb[None, :, q] == 1, d2, np.inf).min(1)
            out[:, 2 * q + comp] = (m1 - m0) / n0
    return out.reshape(np.asarray(bits).shape).astype(dtype)
