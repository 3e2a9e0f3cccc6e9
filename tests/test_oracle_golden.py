"""CPU: the oracle against every golden vector the reference holds for this path (MATLAB 5G Toolbox files of
Playground/CompareWithMatlab) and against the committed outputs of the unmodified reference (ref_cases.npz)."""
import os
import numpy as np
import pytest

import nr_oracle as O
import nr_oracle_c as OC


def test_matlab_golden_flow(matlab):
    """Replays Playground/CompareWithMatlab/LDPC/LDPC-Matlab.ipynb (asserts at raw lines 113-326) with the oracle."""
    in_bits = matlab["in"].reshape(-1).astype(np.int8)
    tbc = O.crc_attach(in_bits, "24A")
    cbs, p = O.segment(tbc, 1)
    assert (p["C"], p["Zc"], p["iLS"], p["F"], p["K"]) == (2, 240, 7, 244, 5280)
    cbs_m = cbs.copy()
    cbs_m[:, p["K"] - p["F"]:] = -1                                   # MATLAB marks fillers with -1
    assert np.array_equal(cbs_m, matlab["cbsIn"].T)
    coded = O.encode(cbs, 1, 240, 7)
    coded_m = coded.copy()
    fs = p["K"] - p["F"] - 2 * 240
    coded_m[:, fs:fs + p["F"]] = -1
    assert np.array_equal(coded_m, matlab["enc"].T)
    full = O.encode(cbs, 1, 240, 7, puncture=False)
    assert O.parity_ok(full, 1, 240, 7).all()
    assert O.parity_ok(np.zeros(68 * 240, np.int8), 1, 240, 7) and not O.parity_ok(np.ones(68 * 240, np.int8), 1, 240, 7)
    g = int(np.ceil((len(tbc) - 24) / (449 / 1024)))
    rm = O.rate_match(coded, 1, 240, p["K"], p["F"], g, 2)
    assert np.array_equal(rm, matlab["chIn"].reshape(-1))
    ch = 1 - 2.0 * rm
    rr, _, p2 = O.rate_recover(ch, len(in_bits), 1, 2)
    rr_m = matlab["raterec"].T.copy()
    rr_m[rr_m == np.inf] = O.LARGE_LLR
    assert np.array_equal(rr, rr_m)
    bits = O.decode(rr, 1, 240, 7, 5)
    assert np.array_equal(bits, matlab["decBits"].T)
    tb, ok = O.check_crc_and_merge(bits, p["K"], p["F"], p["C"])
    assert all(ok) and O.crc_check(tb, "24A")
    assert np.array_equal(tb, matlab["decBlk"].reshape(-1))
    assert np.array_equal(tb[:-24], in_bits)


def test_matlab_crc24c(matlab):
    """Playground/CompareWithMatlab/Polar/PolarMatlab.ipynb raw line 125: the only golden vector for CRC24C."""
    msg = matlab["polar_msg"].reshape(-1).astype(np.int8)
    assert np.array_equal(O.crc_attach(msg, "24C"), matlab["polar_msgcrc"].reshape(-1))


def test_crc_known_answers(ref_cases):
    _, crc = ref_cases
    for poly in O.CRC_POLYS:
        assert np.array_equal(O.crc_remainder(crc[poly + "/in1"], poly), crc[poly + "/out1"])
        assert np.array_equal(O.crc_remainder(crc[poly + "/in2"], poly), crc[poly + "/out2"])
        # CRC of 0000 0001 = the low bits of the generator (SURVEY 8a probe)
        one = O.crc_remainder(np.array([0, 0, 0, 0, 0, 0, 0, 1], np.int8), poly)
        assert np.array_equal(one, crc[poly + "/one"])
        c = O.crc_len(poly)
        assert int("".join(map(str, one)), 2) == O.CRC_POLYS[poly] & ((1 << c) - 1)
        assert OC.crc(crc[poly + "/in1"], poly) == int("".join(map(str, crc[poly + "/out1"])), 2)


def _case_names():
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_cases.npz"))
    return [str(n) for n in z["names"]]


@pytest.mark.parametrize("name", _case_names())
def test_oracle_vs_reference_fixture(ref_cases, name):
    """Every stage of the oracle against the committed outputs of the unmodified reference, incl. HARQ combining."""
    cases, _ = ref_cases
    d = cases[name]
    bg, A, qm, nl, nref, g = d["bg"], d["A"], d["qm"], d["nl"], d["nref"], d["g"]
    p = O.derive_params(bg, A + 24)
    assert (p["C"], p["Zc"], p["iLS"], p["K"]) == (d["C"], d["Zc"], d["iLS"], d["K"])
    cbs, p = O.segment(O.crc_attach(d["tb"], "24A"), bg)
    assert p["F"] == d["F"] and np.array_equal(cbs, d["cbs"])
    coded = O.encode(cbs, bg, p["Zc"], p["iLS"])
    assert np.array_equal(coded, d["coded"])
    buf = None
    for t, rv in enumerate(d["rvs"]):
        rv = int(rv)
        rm = O.rate_match(coded, bg, p["Zc"], p["K"], p["F"], g, qm, nl, nref, rv)
        assert np.array_equal(rm, d["rm%d" % t])
        llr = d["llr%d" % t].astype(np.float64)
        rr, buf, _ = O.rate_recover(llr, A, bg, qm, nl, nref, rv, soft_buffer=buf)
        assert np.array_equal(buf, d["decbuf%d" % t])
        bel = O.decode(rr, bg, p["Zc"], p["iLS"], d["nit"], False, True)
        assert np.array_equal(bel, d["bel%d" % t])                      # float64 beliefs bit for bit
        assert np.array_equal(OC.decode_beliefs(rr, bg, p["Zc"], p["iLS"], d["nit"], np.float64), d["bel%d" % t])
        bits = (bel[:, :p["K"]] < 0).astype(np.int8)
        tb, ok = O.check_crc_and_merge(bits, p["K"], p["F"], p["C"])
        assert np.array_equal(tb, d["merged%d" % t]) and list(ok) == list(d["cbok%d" % t])
        assert bool(O.crc_check(tb, "24A")) == bool(d["tbok%d" % t])


def test_c_oracle_matches_numpy_oracle_fp32_fp64():
    rng = np.random.default_rng(11)
    for bg, zc, ils in [(1, 384, 1), (2, 52, 6), (1, 18, 4), (2, 384, 1), (1, 15, 7)]:
        _, _, k = O.bg_dims(bg)
        cw = O.encode(rng.integers(0, 2, (2, k * zc)).astype(np.int8), bg, zc, ils)
        llr = (1 - 2.0 * cw) * 3 + 2.5 * rng.standard_normal(cw.shape)
        llr[:, -5 * zc:] = 0
        llr[0, :3] = 1e20
        llr[1, 5] = -0.0
        for dt in (np.float64, np.float32):
            a = O.decode(llr.astype(dt), bg, zc, ils, 4, False, True, dtype=dt)
            b = OC.decode_beliefs(llr.astype(dt), bg, zc, ils, 4, dt)
            assert a.dtype == dt and np.array_equal(a, b)


def test_row_skipping_is_exact():
    """Extension rows whose parity LLRs are all zero never change another column: scheduling only the rows up to the
    last non-zero extension column gives bit-identical posteriors on every column those rows cover."""
    rng = np.random.default_rng(5)
    bg, zc, ils = 1, 96, 1
    cw = O.encode(rng.integers(0, 2, (2, 22 * zc)).astype(np.int8), bg, zc, ils)
    llr = ((1 - 2.0 * cw) * 4 + 2.0 * rng.standard_normal(cw.shape)).astype(np.float32)
    keep = 30 * zc + 17                      # LLRs present for punctured-frame positions < keep
    llr[:, keep:] = 0
    rows = (keep - 1) // zc + 2 - 22 + 1     # row owning the last non-zero extension column
    full = OC.decode_beliefs(llr, bg, zc, ils, 6, np.float32)
    part = OC.decode_beliefs(llr, bg, zc, ils, 6, np.float32, num_rows=rows)
    ncols = 22 + rows
    assert np.array_equal(full[:, :ncols * zc], part[:, :ncols * zc])


# ----------------------------------------------------------------------------------------------------------------------
# modem oracle pinned by outputs of the unmodified reference Modem (oracle/gen_golden_modem.py)
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mod", ["BPSK", "QPSK", "16QAM", "64QAM", "256QAM", "1024QAM"])
def test_modem_oracle_matches_reference_outputs(mod):
    import nr_modem
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "modem_cases.npz"))
    qm = nr_modem.QM[mod]
    assert np.array_equal(nr_modem.constellation(qm), g[mod + "_constellation"])
    assert np.array_equal(nr_modem.modulate(g[mod + "_bits"], qm), g[mod + "_symbols"])
    llr = nr_modem.llrs_maxlog(g[mod + "_noisy"], qm, float(g[mod + "_n0"]))
    assert np.array_equal(llr, g[mod + "_llr"])
    assert np.array_equal(np.int8((llr <= 0) * 1), g[mod + "_hard"])


def test_gold_sequence_oracle_matches_reference_and_38211_definition():
    import nr_modem
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "modem_cases.npz"))
    for k in range(8):
        c_init, n = (int(v) for v in g["gold%d_cinit_n" % k])
        ref = np.unpackbits(g["gold%d_bits" % k])[:n]
        assert np.array_equal(np.array(nr_modem.gold_sequence(c_init, n)), ref)
        if n <= 1000:   # the bit-serial definition of TS 38.211 5.2.1 is slow in Python
            assert np.array_equal(np.array(nr_modem.gold_sequence_38211(c_init, n)), ref)


# ----------------------------------------------------------------------------------------------------------------------
# decode2: oracle restatement pinned by outputs of the unmodified reference
# (oracle/gen_golden_decode2.py)
# ----------------------------------------------------------------------------------------------------------------------
def _decode2_cases():
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "decode2_cases.npz"))
    return g, [str(n) for n in g["names"]]


@pytest.mark.parametrize("name", _decode2_cases()[1])
def test_decode2_oracle_matches_reference_outputs(name):
    g, _ = _decode2_cases()
    bg, A, zc, ils, nit, K = (int(v) for v in g[name + "/meta"])
    alpha = float(g[name + "/alpha"])
    rr = g[name + "/rr"]
    assert np.array_equal(O.decode2(rr, bg, zc, ils, nit, False, True, alpha, False), g[name + "/bel"])
    # the reference's stop test only looks at the first base-graph row (ldpc.py:841-843)
    assert np.array_equal(O.decode2(rr, bg, zc, ils, nit + 4, False, True, alpha, True, "first_row"), g[name + "/bel_stop"])
    # decode2 without a stop is the layered schedule of decode() with the true second minimum: on inputs where the
    # "+100000" quirk cannot bite (all |LLR| far below 5e4) and alpha = 0.75 both give the same beliefs
    if alpha == 0.75 and np.abs(rr).max() < 1e3:
        assert np.array_equal(g[name + "/bel"], O.decode(rr, bg, zc, ils, nit, False, True))
