"""CPU ORACLE (test infrastructure, not product code) for the 5G NR LDPC hot path of NeoRadium v0.4.0.

This is a NumPy restatement of the reference's algorithm for the path named by BASELINE.json:north_star:
  neoradium/chancodebase.py  (CRC)                      -> crc_remainder / crc_check / crc_attach
  neoradium/ldpc.py          (segmentation .. merge)    -> derive_params / segment / encode / rate_match /
                                                           rate_recover / decode / check_crc_and_merge
Every function cites the reference lines it follows.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline / reference arm may import this module; the product path (neoradium_b200/) never does and fails loudly when
its CUDA library is missing.

PARITY PIN: this oracle is pinned against (1) the seven MATLAB 5G-Toolbox golden vectors the reference ships in
Playground/CompareWithMatlab/LDPC/MatlabFiles (committed as tests/golden/matlab_ldpc.npz) and (2) outputs of the
unmodified reference run in the build container on seeded inputs (tests/golden/ref_*.npz, made by
oracle/gen_golden.py; tests/test_oracle_vs_reference.py re-runs the comparison live whenever /root/reference exists).
At dtype=float64 `decode` is bit-identical to the reference's beliefs; at dtype=float32 it is the "reference min-sum
evaluated in fp32" that north_star names as the LLR oracle.
"""
import numpy as np

from nr_tables import EDGES, LIFTING_SETS

LARGE_LLR = 1e20  # chancodebase.py:52

# chancodebase.py:37-44 -- generator polynomials, MSB first, including the leading 1
CRC_POLYS = {
    "6": 0x61, "11": 0xE21, "16": 0x11021, "24A": 0x1864CFB, "24B": 0x1800063, "24C": 0x1B2B117,
}
MOD_ORDER = {"BPSK": 1, "QPSK": 2, "16QAM": 4, "64QAM": 6, "256QAM": 8, "1024QAM": 10}  # ldpc.py:743


# ---------------------------------------------------------------------------------------------------------------------
# a1  CRC  (chancodebase.py:59-63, 83-128, 132-157, 161-189)
# ---------------------------------------------------------------------------------------------------------------------
def crc_len(poly):
    return 24 if poly[:2] == "24" else int(poly)


def crc_remainder(bits, poly):
    """Remainder of bits(x) * x^c divided by g(x): MSB-first long division, zero initial state, no reflection and
    no final XOR (chancodebase.py:120-128).  `bits` is [L] or [m, L], one value per bit.  Returns int8 [c] / [m, c]
    (the reference returns int64 because it concatenates a Python list; values are identical)."""
    bits = np.asarray(bits)
    flat = bits.ndim == 1
    b = (bits[None, :] if flat else bits).astype(np.int64) & 1
    m, n = b.shape
    c = crc_len(poly)
    g = CRC_POLYS[poly] & ((1 << c) - 1)          # drop the leading 1
    top = 1 << (c - 1)
    mask = (1 << c) - 1
    reg = np.zeros(m, dtype=np.int64)
    for d in range(n):                             # one division step per message bit, all streams at once
        fb = ((reg & top) != 0).astype(np.int64) ^ b[:, d]
        reg = ((reg << 1) & mask) ^ (fb * g)
    out = ((reg[:, None] >> np.arange(c - 1, -1, -1)[None, :]) & 1).astype(np.int8)
    return out[0] if flat else out


def crc_check(bits, poly):
    """True where the remainder of the whole stream (data || crc) is zero (chancodebase.py:157)."""
    return np.count_nonzero(crc_remainder(bits, poly), -1) == 0


def crc_attach(bits, poly):
    """chancodebase.py:189"""
    bits = np.asarray(bits)
    return np.append(bits, crc_remainder(bits, poly).astype(bits.dtype if bits.dtype.kind in "iu" else np.int8),
                     axis=-1)


# ---------------------------------------------------------------------------------------------------------------------
# a2/a3  graph and parameter derivation  (ldpc.py:775-789, 846-856, 859-892)
# ---------------------------------------------------------------------------------------------------------------------
def bg_dims(bg):
    """(rows P, cols n, systematic cols k)  -- ldpc.py:780"""
    return (46, 68, 22) if bg == 1 else (42, 52, 10)


def base_graph(bg, zc, ils):
    """int16 [P, n]; -1 = no edge, else V % Zc (ldpc.py:781-788)."""
    P, n, _ = bg_dims(bg)
    h = np.full((P, n), -1, dtype=np.int16)
    for row, col, vals in EDGES[bg]:
        h[row, col] = vals[ils] % zc
    return h


def set_index_of(zc):
    for i, s in enumerate(LIFTING_SETS):
        if zc in s:
            return i
    raise ValueError("illegal lifting size %r" % (zc,))


def derive_params(bg, tb_size_with_crc):
    """ldpc.py:859-892.  Input B (transport block length INCLUDING its 24-bit CRC).
    Returns dict(C, Zc, iLS, K)."""
    B = int(tb_size_with_crc)
    kcb = 8448 if bg == 1 else 3840
    if B <= kcb:
        C, total = 1, B
    else:
        C = int(np.ceil(B / (kcb - 24)))
        total = B + 24 * C
    k_prime = total / C                                    # may be fractional, as in the reference (:874)
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
    for i, s in enumerate(LIFTING_SETS):
        for z in s:
            if kb * z >= k_prime and z < zc:
                zc, ils = z, i
    K = (22 if bg == 1 else 10) * zc
    return dict(C=C, Zc=zc, iLS=ils, K=K)


def rm_cb_lens(g, c, n_layers, qm):
    """E_r per code block (ldpc.py:846-856)."""
    f = n_layers * qm
    g_base = int(np.ceil(g / f))
    lens = np.zeros(c, dtype=np.int64)
    lens[c - g_base % c:] = f
    lens += (g_base // c) * f
    return lens


def k0_start(bg, rv, ncb, n, zc):
    """ldpc.py:1145 / :1395 -- applied to the filler-less circular buffer (reference quirk)."""
    num = ([0, 17, 33, 56] if bg == 1 else [0, 13, 25, 43])[rv]
    return (num * ncb // n) * zc


# ---------------------------------------------------------------------------------------------------------------------
# a5  segmentation  (ldpc.py:1011-1030)
# ---------------------------------------------------------------------------------------------------------------------
def segment(tb_with_crc, bg):
    """[B] -> ([C, K] int8, params incl. F).  Zero pad at the END of the TB, CRC24B per CB when C>1, F zero fillers."""
    tb = np.asarray(tb_with_crc)
    p = derive_params(bg, len(tb))
    C, K = p["C"], p["K"]
    per_cb = int(np.ceil(len(tb) / C))
    pad = per_cb * C - len(tb)
    cbs = np.concatenate([tb.astype(np.int8), np.zeros(pad, np.int8)]).reshape(C, per_cb)
    if C > 1:
        cbs = np.concatenate([cbs, crc_remainder(cbs, "24B")], axis=1)
    F = K - cbs.shape[1]
    cbs = np.concatenate([cbs, np.zeros((C, F), np.int8)], axis=1)
    p["F"] = F
    return cbs, p


# ---------------------------------------------------------------------------------------------------------------------
# a6  encoder  (ldpc.py:1057-1090)
# ---------------------------------------------------------------------------------------------------------------------
def _rot(x, s):
    """Left-rotate the last axis: out[..., i] = x[..., (i+s) mod Z]   (ldpc.py:792-810, scalar-shift case)."""
    return np.roll(x, -int(s), axis=-1)


def encode(code_blocks, bg, zc, ils, puncture=True):
    """[C, K] bits -> [C, N] (or [C, N+2Z]).  Double-diagonal core for the first four parity blocks, then one
    XOR-accumulate per extension row."""
    cbs = np.asarray(code_blocks).astype(np.int8)
    P, n, k = bg_dims(bg)
    C, kk = cbs.shape
    assert kk == k * zc
    h = base_graph(bg, zc, ils)
    blk = np.zeros((C, n, zc), np.int8)
    blk[:, :k] = cbs.reshape(C, k, zc)
    lam = np.zeros((C, 4, zc), np.int8)                      # lam_i = XOR_j rot(s_j, h[i,j]) over systematic cols
    for i in range(4):
        for j in range(k):
            if h[i, j] >= 0:
                lam[:, i] ^= _rot(blk[:, j], h[i, j])
    b = h[2, k] if h[1, k] == -1 else h[1, k]                # ldpc.py:1068
    blk[:, k] = _rot(lam[:, 0] ^ lam[:, 1] ^ lam[:, 2] ^ lam[:, 3], zc - b)
    for i in range(3):                                        # ldpc.py:1077-1080
        acc = lam[:, i].copy()
        for t in range(i + 1):
            if h[i, k + t] >= 0:
                acc ^= _rot(blk[:, k + t], h[i, k + t])
        blk[:, k + i + 1] = acc
    for r in range(4, P):                                     # ldpc.py:1083-1084
        acc = np.zeros((C, zc), np.int8)
        for j in range(k + 4):
            if h[r, j] >= 0:
                acc ^= _rot(blk[:, j], h[r, j])
        blk[:, k + r] = acc
    out = blk.reshape(C, n * zc)
    return out[:, 2 * zc:] if puncture else out


def parity_ok(coded_full, bg, zc, ils):
    """All P*Zc parity checks of the un-punctured coded block(s) [.., n*Zc] (a correct `isValidCodedBlock`;
    the reference's ldpc.py:841-843 returns after the first base-graph row)."""
    x = np.asarray(coded_full).astype(np.int8)
    P, n, _ = bg_dims(bg)
    x = x.reshape(x.shape[:-1] + (n, zc))
    h = base_graph(bg, zc, ils)
    ok = np.ones(x.shape[:-2], bool)
    for i in range(P):
        acc = np.zeros(x.shape[:-2] + (zc,), np.int8)
        for j in np.nonzero(h[i] >= 0)[0]:
            acc ^= _rot(x[..., j, :], h[i, j])
        ok &= ~acc.any(-1)
    return ok


# ---------------------------------------------------------------------------------------------------------------------
# a7  rate matching  (ldpc.py:1128-1159)
# ---------------------------------------------------------------------------------------------------------------------
def rate_match(coded, bg, zc, K, F, g, qm, n_layers=1, n_ref=0, rv=0, concat=True):
    coded = np.asarray(coded)
    C, N = coded.shape
    assert N in (66 * zc, 50 * zc)
    if rv not in (0, 1, 2, 3):
        raise ValueError("rv")
    ncb = N if n_ref == 0 else min(N, n_ref)
    sys_len = K - 2 * zc
    circ = np.concatenate([coded[:, :sys_len - F], coded[:, sys_len:ncb]], axis=1)   # fillers are NOT in the buffer
    L = circ.shape[1]
    start = k0_start(bg, rv, ncb, N, zc)
    lens = rm_cb_lens(g, C, n_layers, qm)
    out = []
    for r in range(C):
        e = int(lens[r])
        sel = circ[r, (start + np.arange(e)) % L]
        out.append(sel.reshape(qm, e // qm).T.reshape(-1))       # bit interleaver, ldpc.py:1155
    return np.concatenate(out) if concat else out


# ---------------------------------------------------------------------------------------------------------------------
# a9  rate recovery with soft combining  (ldpc.py:1365-1418)
# ---------------------------------------------------------------------------------------------------------------------
def rate_recover(llrs, tb_size, bg, qm, n_layers=1, n_ref=0, rv=0, soft_buffer=None, dtype=np.float64):
    """[G] LLRs -> ([C, N] dtype, soft buffer [C, Ncb-F] dtype, params).  `soft_buffer` (HARQ decBuffer) is combined
    into and returned (a fresh array here; the reference mutates it in place).  Accumulation order = ascending
    position in the de-interleaved stream, wrap by wrap (ldpc.py:1407-1410)."""
    llrs = np.asarray(llrs)
    p = derive_params(bg, tb_size + 24)
    C, zc, K = p["C"], p["Zc"], p["K"]
    per_cb = int(np.ceil((tb_size + 24) / C)) + (24 if C > 1 else 0)
    F = K - per_cb
    N = (66 if bg == 1 else 50) * zc
    ncb = N if n_ref == 0 else min(N, n_ref)
    L = ncb - F
    buf = np.zeros((C, L), dtype) if soft_buffer is None else np.array(soft_buffer, dtype=dtype, copy=True)
    assert buf.shape == (C, L)
    sys_len = K - F - 2 * zc
    g = len(llrs)
    lens = rm_cb_lens(g, C, n_layers, qm)
    offs = np.concatenate([[0], np.cumsum(lens)])
    start = k0_start(bg, rv, ncb, N, zc)
    for r in range(C):
        e = int(lens[r])
        x = llrs[offs[r]:offs[r + 1]].astype(dtype)
        if len(x) < e:
            x = np.concatenate([x, np.zeros(e - len(x), dtype)])
        x = x.reshape(e // qm, qm).T.reshape(-1)                  # de-interleave, ldpc.py:1405
        pos = (start + np.arange(e)) % L
        for s in range(0, e, L):                                  # one pass per wrap => true accumulate
            buf[r, pos[s:s + L]] += x[s:s + L]
    out = np.concatenate([buf[:, :sys_len], np.full((C, F), LARGE_LLR, dtype), buf[:, sys_len:]], axis=1)
    if ncb < N:
        pass  # the reference returns [C, Ncb] in the LBRM case (no zero extension); we follow it
    p["F"] = F
    return out, buf, p


# ---------------------------------------------------------------------------------------------------------------------
# a10  layered normalised min-sum  (ldpc.py:1535-1581)
# ---------------------------------------------------------------------------------------------------------------------
def decode(rx, bg, zc, ils, num_iter=5, only_info=True, output_belief=False, dtype=np.float64, K=None,
           num_rows=None):
    """[C, N] LLRs -> bits int8 / beliefs `dtype`.

    Per check (layer i, lifted index m), edges j in ascending column order, p_j = (m + s_j) mod Z:
        t_j = r[j,p_j] - msg_i[j];  a_j = |t_j|;  sg_j = -1 if t_j < 0 else +1          (:1550-1556)
        j*  = first argmin a_j;  min1 = a_j*                                               (:1559-1561)
        min2 = min( min_{j != j*} a_j, |t_j* + 100000| )                                   (:1563-1564, the quirk)
        new_j = ((min2 if j == j* else min1) * sg_j * prod(sg)) * 0.75                     (:1567-1573)
        r[j,p_j] = t_j + new_j                                                             (:1576)
    All arithmetic in `dtype` (the reference is float64).  `num_rows` (default all P) restricts the schedule to the
    first rows -- only used by tests that prove row skipping exact."""
    dtype = np.dtype(dtype).type
    P, n, k = bg_dims(bg)
    rx = np.asarray(rx)
    C = rx.shape[0]
    x = np.clip(rx.astype(dtype), dtype(-1e10), dtype(1e10))                     # :1536
    ncols_in = x.shape[1] // zc
    r = np.zeros((C, n, zc), dtype)
    r[:, 2:2 + ncols_in] = x.reshape(C, ncols_in, zc)                           # :1538
    h = base_graph(bg, zc, ils)
    assert r.shape[1] == h.shape[1]
    rows = range(P if num_rows is None else num_rows)
    m = np.arange(zc)
    layers = []
    for i in rows:
        cols = np.nonzero(h[i] >= 0)[0]
        pos = (m[None, :] + h[i, cols].astype(np.int64)[:, None]) % zc          # [q, Z]
        layers.append((cols, pos, np.zeros((C, len(cols), zc), dtype)))
    q_idx_cache = {}
    for _ in range(num_iter):
        for li, (cols, pos, msg) in enumerate(layers):
            t = r[:, cols[:, None], pos] - msg                                   # [C, q, Z]
            a = np.abs(t)
            neg = t < 0
            jstar = np.argmin(a, axis=1)                                         # first minimum
            js = jstar[:, None, :]
            min1 = np.take_along_axis(a, js, 1)
            bumped = np.abs(np.take_along_axis(t, js, 1) + dtype(100000))
            a2 = a.copy()
            np.put_along_axis(a2, js, bumped, 1)
            min2 = a2.min(axis=1, keepdims=True)
            q = len(cols)
            if q not in q_idx_cache:
                q_idx_cache[q] = np.arange(q)[None, :, None]
            mag = np.where(q_idx_cache[q] == js, min2, min1)
            sgn = np.where(neg, dtype(-1), dtype(1))
            par = np.where((neg.sum(1, keepdims=True) & 1) == 1, dtype(-1), dtype(1))
            new = (mag * (sgn * par)) * dtype(0.75)
            r[:, cols[:, None], pos] = t + new
            layers[li] = (cols, pos, new)
    out = r.reshape(C, n * zc)
    if only_info:
        out = out[:, :(k * zc if K is None else K)]
    return out if output_belief else (out < 0).astype(np.int8)


# ---------------------------------------------------------------------------------------------------------------------
# a12  decode2: the row-by-row verification decoder  (ldpc.py:1421-1492)
# ---------------------------------------------------------------------------------------------------------------------
def decode2(rx, bg, zc, ils, max_iter=6, only_info=True, output_belief=False, alpha=0.75, stop_on_good_parity=False,
            stop_rule="all", K=None, beta=0.0):
    """[C, N] LLRs -> bits int8 / float64 beliefs, one lifted parity-check row at a time, exactly as the reference walks
    them (row = i*Z + m, edges in ascending column order, column index col*Z + (m + V) % Z, ldpc.py:1434-1445):
        t = rx[cols] - rr;  a = |t|;  j* = first argmin;  min1 = a[j*];  min2 = min_{j != j*} a_j      (:1462-1470)
        min1 > 0 : rr_j = ((prod(sign t) * sign t_j) * (min2 if j == j* else min1)) * alpha            (:1472-1477)
        min1 == 0 < min2 : rr_j* = prod(1 - 2 (t < 0)) * min2 * alpha, other rr_j = 0                  (:1478-1481)
        both zero : rr = 0                                                                           (:1482-1483)
        rx[cols] = t + rr                                                                            (:1485)
    `beta` (extension, the reference has none): offset min-sum, |message| = max(alpha * min - beta, 0).
    `stop_rule`: "all" = stop after an iteration whose hard decisions satisfy every check (what the CUDA path does);
    "first_row" = what the reference actually tests (isValidCodedBlock returns after the first base-graph row,
    ldpc.py:841-843).  float64 like the reference."""
    P, n, k = bg_dims(bg)
    h = base_graph(bg, zc, ils)
    rx = np.asarray(rx, np.float64)
    C = rx.shape[0]
    rows = []
    for i in range(P):
        cols = np.nonzero(h[i] >= 0)[0]
        for m in range(zc):
            rows.append(cols * zc + (m + h[i, cols].astype(np.int64)) % zc)
    out = []
    for c in range(C):
        r = np.clip(np.concatenate([np.zeros(2 * zc), rx[c]]), -1e10, 1e10)                      # :1452-1456
        rr = [np.zeros(len(ix)) for ix in rows]
        for _ in range(max_iter):
            for ri, ix in enumerate(rows):
                t = r[ix] - rr[ri]
                a = np.abs(t)
                js = int(np.argmin(a))
                min1 = a[js]
                a2 = a.copy()
                a2[js] = 1e10
                min2 = a2.min()
                if min1 > 0:
                    sg = np.sign(t)
                    s1 = np.prod(sg) * sg
                    new = s1 * min1
                    new[js] = s1[js] * min2
                    new = new * alpha
                    if beta:    # offset min-sum (extension, not in the reference): |message| = max(alpha * min - beta, 0)
                        new = np.sign(new) * np.maximum(np.abs(new) - beta, 0.0)
                elif min2 > 0:
                    new = np.zeros_like(t)
                    new[js] = np.prod(1 - 2 * (t < 0)) * min2
                    new = new * alpha
                    if beta:
                        new = np.sign(new) * np.maximum(np.abs(new) - beta, 0.0)
                else:
                    new = np.zeros_like(t)
                rr[ri] = new
                r[ix] = t + new
            if stop_on_good_parity:
                hard = (r < 0).astype(np.int8)
                if stop_rule == "first_row":
                    hb = hard.reshape(n, zc)
                    acc = np.zeros(zc, np.int64)
                    for j in np.nonzero(h[0] >= 0)[0]:
                        acc += np.roll(hb[j], -int(h[0, j]))
                    good = not (acc % 2).any()
                else:
                    good = parity_ok(hard, bg, zc, ils)
                if good:
                    break
        out.append(r)
    out = np.array(out)
    if only_info:
        out = out[:, :(k * zc if K is None else K)]
    return out if output_belief else (out < 0).astype(np.int8)


# ---------------------------------------------------------------------------------------------------------------------
# a11  CRC check + merge  (ldpc.py:1610-1619)
# ---------------------------------------------------------------------------------------------------------------------
def check_crc_and_merge(decoded, K, F, C):
    d = np.asarray(decoded)[:, :K - F]
    if C == 1:
        flat = d.reshape(-1)
        return flat, [bool(crc_check(flat, "24A"))]
    return d[:, :-24].reshape(-1), crc_check(d, "24B")


# ---------------------------------------------------------------------------------------------------------------------
# chains (ldpc.py:1200-1204 and the documented RX usage :1234-1251)
# ---------------------------------------------------------------------------------------------------------------------
def tx_chain(tb_bits, bg, g, qm, n_layers=1, n_ref=0, rv=0):
    tb = crc_attach(np.asarray(tb_bits).astype(np.int8), "24A")
    cbs, p = segment(tb, bg)
    coded = encode(cbs, bg, p["Zc"], p["iLS"])
    return rate_match(coded, bg, p["Zc"], p["K"], p["F"], g, qm, n_layers, n_ref, rv), p


def rx_chain(llrs, tb_size, bg, qm, num_iter, n_layers=1, n_ref=0, rv=0, soft_buffer=None, dtype=np.float64):
    rr, buf, p = rate_recover(llrs, tb_size, bg, qm, n_layers, n_ref, rv, soft_buffer, dtype)
    bits = decode(rr, bg, p["Zc"], p["iLS"], num_iter, dtype=dtype)
    tb, cb_ok = check_crc_and_merge(bits, p["K"], p["F"], p["C"])
    tb_ok = bool(crc_check(tb, "24A"))
    return tb[:-24], cb_ok, tb_ok, buf, p
