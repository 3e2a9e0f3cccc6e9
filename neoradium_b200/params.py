"""Host-side integer parameter derivation for the NR LDPC chain (no arithmetic on payload data happens here).

Mirrors LdpcBase.initialize (neoradium/ldpc.py:859-892), LdpcBase.getRateMatchedCbLens (ldpc.py:846-856) and the
k0 / Ncb / filler bookkeeping of rateMatch / recoverRate (ldpc.py:1135-1145, 1365-1395).  These are a handful of
scalar integer operations per transport block; they stay in Python exactly as in the reference and feed the
`nrldpc_tb_config` struct of the C-ABI.
"""
import math

LIFTING_SETS = (   # TS 38.212 Table 5.3.2-1 (ldpc.py:657-666)
    (2, 4, 8, 16, 32, 64, 128, 256), (3, 6, 12, 24, 48, 96, 192, 384), (5, 10, 20, 40, 80, 160, 320),
    (7, 14, 28, 56, 112, 224), (9, 18, 36, 72, 144, 288), (11, 22, 44, 88, 176, 352), (13, 26, 52, 104, 208),
    (15, 30, 60, 120, 240))
MOD_ORDER = {'BPSK': 1, 'QPSK': 2, '16QAM': 4, '64QAM': 6, '256QAM': 8, '1024QAM': 10}   # ldpc.py:743
K0_NUM = {1: (0, 17, 33, 56), 2: (0, 13, 25, 43)}                                          # ldpc.py:1145


def bg_dims(bg):
    """(rows, columns, systematic columns) -- ldpc.py:780"""
    return (46, 68, 22) if bg == 1 else (42, 52, 10)


def segmentation_params(bg, tb_size_with_crc):
    """(C, Zc, iLS, K) for a transport block of B bits including its CRC (ldpc.py:859-892)."""
    B = int(tb_size_with_crc)
    kcb = 8448 if bg == 1 else 3840
    if B <= kcb:
        C, total = 1, B
    else:
        C = int(math.ceil(B / (kcb - 24)))
        total = B + C * 24
    k_prime = total / C          # kept fractional like the reference (:874)
    if bg == 1:
        kb = 22
    elif B > 640:
        kb = 10
    elif B > 560:
        kb = 9
    elif B > 192:
        kb = 8
    else:
        kb = 6
    zc, ils = 10000, -1
    for i, zs in enumerate(LIFTING_SETS):
        for z in zs:
            if kb * z >= k_prime and z < zc:
                zc, ils = z, i
    return C, zc, ils, (22 if bg == 1 else 10) * zc


def rate_matched_cb_lens(g, c, tx_layers, qm):
    """E_r of every code block (ldpc.py:846-856) as a list of ints."""
    f = tx_layers * qm
    g_base = int(math.ceil(g / f))
    lens = [(g_base // c) * f] * c
    for r in range(c - g_base % c, c):
        lens[r] += f
    return lens


def k0_start(bg, rv, ncb, n, zc):
    """Start index in the filler-less circular buffer (ldpc.py:1145, :1395)."""
    return (K0_NUM[bg][rv] * ncb // n) * zc
