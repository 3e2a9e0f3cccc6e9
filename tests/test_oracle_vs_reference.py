"""CPU, build container only: the oracle against the UNMODIFIED reference imported live from /root/reference
(skipped automatically on the GPU box, where the reference does not exist)."""
import numpy as np
import pytest

import nr_oracle as O
from ref_loader import load_reference, reference_available

pytestmark = pytest.mark.needs_reference


class _Harq:
    rv, decBuffer = 0, None


@pytest.mark.parametrize("bg,A,mod,rate,nl,rv", [
    (1, 5000, "QPSK", 0.5, 1, 0), (2, 2000, "16QAM", 0.35, 1, 1), (1, 9000, "64QAM", 0.7, 2, 2), (2, 640, "QPSK", 0.2, 1, 3),
    (2, 150, "QPSK", 0.3, 1, 0), (1, 300, "256QAM", 0.6, 1, 0), (2, 561, "QPSK", 0.4, 1, 0), (2, 193, "QPSK", 0.5, 1, 2)])
def test_live_chain(bg, A, mod, rate, nl, rv):
    ldpc = load_reference("ldpc")
    rng = np.random.default_rng(A + rv)
    enc = ldpc.LdpcEncoder(bg, mod, nl, 0, rate)
    tb = rng.integers(0, 2, A).astype(np.int8)
    g = int(np.ceil(A / rate))
    cbs = enc.doSegmentation(enc.appendCrc(tb, "24A"))
    coded = enc.encode(cbs)
    rm = enc.rateMatch(coded, g, True, rv)
    orm, p = O.tx_chain(tb, bg, g, enc.qm, nl, 0, rv)
    assert np.array_equal(rm, orm)
    llr = ((1 - 2.0 * rm) * 2 + 1.6 * rng.standard_normal(len(rm))).astype(np.float32).astype(np.float64)
    dec = ldpc.LdpcDecoder(bg, mod, nl, 0)
    h = _Harq()
    h.rv = rv
    rr = dec.recoverRate(llr, A, h)
    orr, obuf, _ = O.rate_recover(llr, A, bg, enc.qm, nl, 0, rv)
    assert np.array_equal(rr, orr) and np.array_equal(h.decBuffer, obuf)
    assert np.array_equal(dec.decode(rr, 4, False, True), O.decode(orr, bg, p["Zc"], p["iLS"], 4, False, True))
    bits = dec.decode(rr, 4)
    tbm, ok = dec.checkCrcAndMerge(bits)
    otbm, ook = O.check_crc_and_merge(O.decode(orr, bg, p["Zc"], p["iLS"], 4), p["K"], p["F"], p["C"])
    assert np.array_equal(tbm, otbm) and list(ok) == list(ook)


def test_live_params_sweep():
    ldpc = load_reference("ldpc")
    for bg in (1, 2):
        e = ldpc.LdpcEncoder(bg)
        for B in list(range(25, 1000, 7)) + list(range(1000, 40000, 997)) + [3840, 3841, 8448, 8449, 100000]:
            e.initialize(B)
            p = O.derive_params(bg, B)
            assert (p["C"], p["Zc"], p["iLS"], p["K"]) == (e.numCodeBlocks, e.liftingSize, e.setIndex, e.codeBlockSize)


def test_live_base_graphs():
    ldpc = load_reference("ldpc")
    for bg in (1, 2):
        for ils, zs in enumerate(ldpc.liftingSizeSets):
            for z in zs:
                e = ldpc.LdpcEncoder(bg)
                e.setIndex, e.liftingSize = ils, z
                assert np.array_equal(e.baseGraph, O.base_graph(bg, z, ils))
