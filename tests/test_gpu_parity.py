"""GPU parity tests (pytest -m gpu): the CUDA path, called through the Python drop-in classes and the C-ABI, against
the CPU oracle on the same seeded inputs and against the committed golden fixtures.  Bit-exact for bits, indices and
float64 beliefs; fp32 beliefs are compared bit-for-bit against the fp32 evaluation of the reference algorithm (the
1e-4 tolerance of north_star is therefore slack, asserted as exact equality)."""
import ctypes

import os

import numpy as np
import pytest
import torch

import nr_link
import nr_oracle as O
import nr_oracle_c as OC
from conftest import MOD_NAME
from neoradium_b200 import ChanCodeBase, LdpcDecoder, LdpcEncoder, _dev, _native
from neoradium_b200.batch import TbBatchCodec, qam_awgn_llr

pytestmark = pytest.mark.gpu


class Harq:
    def __init__(self, rv=0):
        self.rv, self.decBuffer = rv, None


# ----------------------------------------------------------------------------------------------------------------------
def test_matlab_golden_flow_on_gpu(matlab):
    """Playground/CompareWithMatlab/LDPC/LDPC-Matlab.ipynb replayed on the CUDA path (all seven .mat comparisons)."""
    in_bits = matlab["in"].reshape(-1)
    enc = LdpcEncoder(baseGraphNo=1, modulation='QPSK', txLayers=1, nRef=0, targetRate=449 / 1024)
    tbc = enc.appendCrc(in_bits, '24A')
    assert tbc.shape == (10024,)
    cbs = enc.doSegmentation(tbc)
    assert (enc.liftingSize, enc.setIndex, enc.numFillerBits) == (240, 7, 244)
    fs = enc.codeBlockSize - enc.numFillerBits
    cbs_m = cbs.copy()
    cbs_m[:, fs:] = -1
    assert np.abs(cbs_m - matlab["cbsIn"].T).sum() == 0
    full = enc.encode(cbs, puncture=False)
    assert (enc.isValidCodedBlock(full[0]), enc.isValidCodedBlock(full[1]), enc.isValidCodedBlock(np.zeros(68 * 240)),
            enc.isValidCodedBlock(np.ones(68 * 240))) == (True, True, True, False)
    coded = enc.encode(cbs)
    coded_m = coded.copy()
    coded_m[:, fs - 480:fs - 480 + 244] = -1
    assert np.abs(coded_m - matlab["enc"].T).sum() == 0
    rm = enc.rateMatch(coded)
    assert np.abs(rm - matlab["chIn"].reshape(-1)).sum() == 0
    assert np.abs(enc.getRateMatchedCodeBlocks(in_bits) - matlab["chIn"].reshape(-1)).sum() == 0
    ch = 1 - 2.0 * rm
    dec = LdpcDecoder(baseGraphNo=1, modulation='QPSK', txLayers=1, nRef=0)
    rr = dec.recoverRate(ch, len(in_bits))
    rr_m = matlab["raterec"].T.copy()
    rr_m[rr_m == np.inf] = LdpcDecoder.LARGE_LLR
    assert rr.dtype == np.float64 and np.abs(rr - rr_m).sum() == 0
    bits = dec.decode(rr)
    assert bits.dtype == np.int8 and np.abs(bits - matlab["decBits"].T).sum() == 0
    tb, ok = dec.checkCrcAndMerge(bits)
    assert list(ok) == [True, True] and np.abs(tb - matlab["decBlk"].reshape(-1)).sum() == 0
    assert dec.checkCrc(tb, '24A')
    assert np.abs(tb[:-24] - in_bits).sum() == 0
    # fused chain, both precisions
    for prec in ("fp64", "fp32"):
        d2 = LdpcDecoder(1, 'QPSK', precision=prec)
        out, cbok, tbok = d2.decodeLLRs(ch, len(in_bits), 5)
        assert np.array_equal(out, in_bits) and all(cbok) and tbok


def test_crc24c_matlab_vector(matlab):
    msg = matlab["polar_msg"].reshape(-1).astype(np.int8)
    assert np.array_equal(ChanCodeBase.appendCrc(msg, '24C'), matlab["polar_msgcrc"].reshape(-1))


# ----------------------------------------------------------------------------------------------------------------------
def _fixture_names():
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_cases.npz"))
    return [str(n) for n in z["names"]]


@pytest.mark.parametrize("name", _fixture_names())
def test_reference_fixtures_on_gpu(ref_cases, name):
    """Outputs of the UNMODIFIED reference (committed fixtures): every stage bit-exact, float64 beliefs bit-exact,
    HARQ soft buffers bit-exact over the rv sequence."""
    cases, _ = ref_cases
    d = cases[name]
    bg, A, nl, nref, g = d["bg"], d["A"], d["nl"], d["nref"], d["g"]
    enc = LdpcEncoder(bg, MOD_NAME[d["qm"]], nl, nref, 0.5)
    cbs = enc.doSegmentation(enc.appendCrc(d["tb"], '24A'))
    assert (enc.numCodeBlocks, enc.liftingSize, enc.setIndex, enc.codeBlockSize, enc.numFillerBits) == \
        (d["C"], d["Zc"], d["iLS"], d["K"], d["F"])
    assert np.array_equal(cbs, d["cbs"])
    coded = enc.encode(cbs)
    assert np.array_equal(coded, d["coded"])
    dec = LdpcDecoder(bg, MOD_NAME[d["qm"]], nl, nref)          # fp64 = the reference's arithmetic
    dec32 = LdpcDecoder(bg, MOD_NAME[d["qm"]], nl, nref, precision='fp32')
    dec32.initialize(A + 24)
    h, hf = Harq(), Harq()
    for t, rv in enumerate(d["rvs"]):
        rv = int(rv)
        assert np.array_equal(enc.rateMatch(coded, g, True, rv), d["rm%d" % t])
        llr = d["llr%d" % t].astype(np.float64)
        h.rv = hf.rv = rv
        rr = dec.recoverRate(llr, A, h)
        assert np.array_equal(h.decBuffer, d["decbuf%d" % t])
        bel = dec.decode(rr, d["nit"], False, True)
        assert np.array_equal(bel, d["bel%d" % t])
        bits = dec.decode(rr, d["nit"])
        tb, ok = dec.checkCrcAndMerge(bits)
        assert np.array_equal(tb, d["merged%d" % t]) and list(ok) == list(d["cbok%d" % t])
        assert bool(dec.checkCrc(tb, '24A')) == bool(d["tbok%d" % t])
        # fused chain with HARQ combining (fp64): same bits, same soft buffer
        out, cbok, tbok = dec.decodeLLRs(llr, A, d["nit"], harq=hf)
        assert np.array_equal(out, d["merged%d" % t][:A]) and list(cbok) == list(d["cbok%d" % t])
        assert bool(tbok) == bool(d["tbok%d" % t]) and np.array_equal(hf.decBuffer, d["decbuf%d" % t])
        # fp32 kernel == reference algorithm evaluated in fp32 (bit for bit)
        bel32 = dec32.decode(rr, d["nit"], False, True)
        obel32 = OC.decode_beliefs(rr.astype(np.float32), bg, d["Zc"], d["iLS"], d["nit"], np.float32)
        assert np.array_equal(bel32, obel32.astype(np.float64))


def test_crc_all_polynomials(ref_cases):
    _, crc = ref_cases
    rng = np.random.default_rng(3)
    for poly in O.CRC_POLYS:
        assert np.array_equal(ChanCodeBase.getCrc(crc[poly + "/in1"], poly), crc[poly + "/out1"])
        assert np.array_equal(ChanCodeBase.getCrc(crc[poly + "/in2"], poly), crc[poly + "/out2"])
        assert np.array_equal(ChanCodeBase.getCrc(np.array([0, 0, 0, 0, 0, 0, 0, 1]), poly), crc[poly + "/one"])
        for shape in [(1,), (5,), (255,), (256,), (257,), (4, 1000), (2, 8448), (1, 300001)]:
            b = rng.integers(0, 2, shape).astype(np.int8)
            got = ChanCodeBase.getCrc(b, poly)
            assert got.dtype == np.int64 and np.array_equal(got, O.crc_remainder(b, poly))
            a = ChanCodeBase.appendCrc(b, poly)
            assert np.array_equal(a, O.crc_attach(b, poly)) and np.all(ChanCodeBase.checkCrc(a, poly))
            a[..., 0] ^= 1
            assert not np.any(ChanCodeBase.checkCrc(a, poly))
    assert isinstance(ChanCodeBase.checkCrc(np.zeros(30, np.int8), '24A'), (bool, np.bool_))
    assert ChanCodeBase.checkCrc(np.zeros((3, 30), np.int8), '16').shape == (3,)
    with pytest.raises(KeyError):
        ChanCodeBase.getCrc(np.zeros(8, np.int8), '32')


# ----------------------------------------------------------------------------------------------------------------------
CHAIN_CASES = [
    (1, 10000, 'QPSK', 449 / 1024, 1, 0), (2, 3000, 'QPSK', 0.3, 1, 0), (1, 8400 * 4, '16QAM', 0.6, 1, 0),
    (1, 3000, '64QAM', 0.5, 1, 2), (2, 300, 'QPSK', 0.25, 1, 3), (1, 1200, '16QAM', 0.4, 2, 0), (2, 100, 'BPSK', 0.2, 1, 0),
    (2, 40, 'QPSK', 0.2, 1, 1), (1, 500, '256QAM', 0.7, 1, 0), (1, 20000, '256QAM', 0.8, 4, 1), (2, 3800, '1024QAM', 0.5, 1, 0),
    (1, 8424 * 3 - 24, '16QAM', 1 / 3, 1, 0), (2, 9000, 'QPSK', 0.2, 1, 0), (2, 12, 'QPSK', 0.2, 1, 0), (1, 30, 'QPSK', 0.34, 1, 0),
    # 22-33 scheduled rows at Zc = 384 / 256: the split (Tensor Memory + shared planes) static decoder kernels
    (1, 8424 * 2 - 24, 'QPSK', 0.45, 1, 0), (2, 3816, '16QAM', 0.3, 1, 0), (1, 5608, '64QAM', 0.42, 1, 0),
]


@pytest.mark.parametrize("prec", ["fp64", "fp32"])
@pytest.mark.parametrize("bg,A,mod,rate,nl,rv", CHAIN_CASES)
def test_chain_vs_oracle(bg, A, mod, rate, nl, rv, prec):
    rng = np.random.default_rng(A * 7 + rv)
    enc = LdpcEncoder(bg, mod, nl, 0, rate)
    tb = rng.integers(0, 2, A).astype(np.int8)
    g = int(np.ceil(A / rate))
    cbs = enc.doSegmentation(enc.appendCrc(tb, '24A'))
    ocbs, p = O.segment(O.crc_attach(tb, '24A'), bg)
    assert np.array_equal(cbs, ocbs)
    coded = enc.encode(cbs)
    ocoded = O.encode(ocbs, bg, p['Zc'], p['iLS'])
    assert np.array_equal(coded, ocoded)
    rm = enc.rateMatch(coded, g, True, rv)
    orm = O.rate_match(ocoded, bg, p['Zc'], p['K'], p['F'], g, enc.qm, nl, 0, rv)
    assert np.array_equal(rm, orm)
    lst = enc.rateMatch(coded, g, False, rv)
    assert isinstance(lst, list) and np.array_equal(np.concatenate(lst), orm)
    with pytest.raises(ValueError):
        enc.rateMatch(coded, g, True, 4)
    sigma = 0.7
    llr = (2 * ((1 - 2.0 * orm) + sigma * rng.standard_normal(len(orm))) / sigma ** 2).astype(np.float32).astype(np.float64)
    dt = np.float64 if prec == 'fp64' else np.float32
    dec = LdpcDecoder(bg, mod, nl, 0, precision=prec)
    h = Harq(rv)
    rr = dec.recoverRate(llr, A, h)
    orr, obuf, _ = O.rate_recover(llr, A, bg, enc.qm, nl, 0, rv)
    assert np.array_equal(rr, orr) and np.array_equal(h.decBuffer, obuf)
    rr2 = dec.recoverRate(llr * 0.5, A, h)                       # soft combining of a second transmission
    orr2, obuf2, _ = O.rate_recover(llr * 0.5, A, bg, enc.qm, nl, 0, rv, soft_buffer=obuf)
    assert np.array_equal(rr2, orr2) and np.array_equal(h.decBuffer, obuf2)
    for nit in (0, 1, 6):
        bel = dec.decode(rr, nit, False, True)
        obel = OC.decode_beliefs(orr, bg, p['Zc'], p['iLS'], nit, dt).astype(np.float64)
        assert bel.dtype == np.float64 and np.array_equal(bel, obel)
    bits = dec.decode(rr, 6)
    obits = (obel[:, :p['K']] < 0).astype(np.int8)
    assert np.array_equal(bits, obits)
    assert np.array_equal(dec.decode(rr, 6, True, True), obel[:, :p['K']])
    tbm, ok = dec.checkCrcAndMerge(bits)
    otbm, ook = O.check_crc_and_merge(obits, p['K'], p['F'], p['C'])
    assert np.array_equal(tbm, otbm) and list(ok) == list(ook)
    x = llr if prec == 'fp64' else llr.astype(np.float32)
    ftb, fcb, ftbok = dec.decodeLLRs(x, A, 6, harq=Harq(rv))
    assert np.array_equal(ftb, otbm[:A]) and list(fcb) == list(ook) and bool(ftbok) == bool(O.crc_check(otbm, '24A'))
    # truncated input: the missing tail counts as zero LLRs (ldpc.py:1402-1403)
    cut = len(llr) - min(7, len(llr) // 3)
    assert np.array_equal(dec.recoverRate(llr[:cut], A, Harq(rv)), O.rate_recover(llr[:cut], A, bg, enc.qm, nl, 0, rv)[0])


def test_decode_special_values():
    """+-0.0, huge / infinite LLRs (clipped to 1e10), all-zero input, the +100000 quirk regime."""
    rng = np.random.default_rng(9)
    bg, zc, ils = 2, 40, 2
    cw = O.encode(rng.integers(0, 2, (3, 10 * zc)).astype(np.int8), bg, zc, ils)
    llr = (1 - 2.0 * cw) * 3 + 2.0 * rng.standard_normal(cw.shape)
    llr[0, ::7] = -0.0
    llr[0, 1::11] = 0.0
    llr[1] = (1 - 2.0 * cw[1]) * 1e20            # noise-free "certain" LLRs: every |t| > 1e5 => the quirk decides min2
    llr[1, 5] = np.inf
    llr[1, 6] = -np.inf
    llr[2] = 0.0
    for prec, dt in (("fp64", np.float64), ("fp32", np.float32)):
        dec = LdpcDecoder(bg, 'QPSK', precision=prec)
        dec.liftingSize, dec.setIndex, dec.codeBlockSize = zc, ils, 10 * zc
        for nit in (1, 3):
            got = dec.decode(llr, nit, False, True)
            want = OC.decode_beliefs(llr.astype(dt), bg, zc, ils, nit, dt).astype(np.float64)
            assert np.array_equal(got, want)


def test_row_skipping_flag_is_exact():
    """nrldpc_decode with and without NRLDPC_DEC_ALL_ROWS gives identical bits and beliefs on ALL columns."""
    rng = np.random.default_rng(21)
    for bg, zc, keep_cols in [(1, 384, 37), (2, 64, 20), (1, 48, 66), (1, 30, 22)]:
        _, n, k = O.bg_dims(bg)
        ils = O.set_index_of(zc)
        C = 5
        cw = O.encode(rng.integers(0, 2, (C, k * zc)).astype(np.int8), bg, zc, ils)
        llr = ((1 - 2.0 * cw) * 3 + 2.2 * rng.standard_normal(cw.shape)).astype(np.float32)
        llr[:, keep_cols * zc - zc // 3:] = 0
        x = torch.from_numpy(llr).cuda()
        res = []
        for flags in (0, _native.DEC_ALL_ROWS):
            bel = torch.empty((C, n * zc), dtype=torch.float32, device='cuda')
            bits = torch.empty((C, n * zc), dtype=torch.int8, device='cuda')
            _native.check(_native.lib().nrldpc_decode(_dev.handle(), bg, zc, _native.F32, _native.F32, _dev.ptr(x), C,
                                                      (n - 2) * zc, n - 2, 5, flags, n, _dev.ptr(bits), _dev.ptr(bel),
                                                      None, _dev.stream_ptr()))
            res.append((bel.cpu().numpy(), bits.cpu().numpy()))
        want = OC.decode_beliefs(llr, bg, zc, ils, 5, np.float32)
        assert np.array_equal(res[0][0], want) and np.array_equal(res[1][0], want)
        assert np.array_equal(res[0][1], (want < 0).astype(np.int8)) and np.array_equal(res[1][1], res[0][1])


def test_decoder_kernel_variants_agree(monkeypatch):
    """Every static fp32 kernel variant a configuration can be routed to (compile-time edge table for Zc=384, with / without the
    early-termination code, split vs single all-TMEM CTA vs tiered state, forced residency) returns the oracle's beliefs bit for
    bit.  The library reads these knobs at every launch."""
    rng = np.random.default_rng(33)
    knobs = [{}, {"NRLDPC_NO_SPECZ": "1"}, {"NRLDPC_ES_CODE": "1"}, {"NRLDPC_NO_SPLIT": "1"}, {"NRLDPC_NO_SPLIT": "1", "NRLDPC_NO_SPECZ": "1"},
             {"NRLDPC_NO_SPECZ": "1", "NRLDPC_ES_CODE": "1"}]
    # (the second line: low code rates on CTAs of at most 8 warps -- three tiered CTAs per SM -- and on multi-block CTAs)
    for bg, zc, rows in [(1, 384, 17), (1, 384, 26), (2, 384, 26), (2, 384, 42), (1, 320, 24), (1, 384, 40), (1, 384, 46),
                         (1, 256, 42), (1, 256, 46), (1, 240, 42), (1, 240, 38), (1, 224, 46), (1, 208, 42), (2, 240, 42), (1, 128, 46), (1, 60, 44)]:
        _, n, k = O.bg_dims(bg)
        ils = O.set_index_of(zc)
        C = 3
        cw = O.encode(rng.integers(0, 2, (C, k * zc)).astype(np.int8), bg, zc, ils)
        llr = ((1 - 2.0 * cw) * 3 + 2.4 * rng.standard_normal(cw.shape)).astype(np.float32)
        llr[:, (k + rows - 2) * zc:] = 0           # `rows` scheduled rows
        want = OC.decode_beliefs(llr, bg, zc, ils, 4, np.float32)
        x = torch.from_numpy(llr).cuda()
        for kn in knobs:
            with monkeypatch.context() as mp:
                for name, val in kn.items():
                    mp.setenv(name, val)
                bel = torch.empty((C, n * zc), dtype=torch.float32, device='cuda')
                bits = torch.empty((C, n * zc), dtype=torch.int8, device='cuda')
                _native.check(_native.lib().nrldpc_decode(_dev.handle(), bg, zc, _native.F32, _native.F32, _dev.ptr(x), C,
                                                          (n - 2) * zc, n - 2, 4, 0, n, _dev.ptr(bits), _dev.ptr(bel),
                                                          None, _dev.stream_ptr()))
                assert np.array_equal(bel.cpu().numpy(), want), (bg, zc, rows, kn)
                assert np.array_equal(bits.cpu().numpy(), (want < 0).astype(np.int8)), (bg, zc, rows, kn)


def test_early_stop_parity_protocol():
    """Early termination is an extension: the kernel reports the per-block iteration count n, and the block's output
    must equal the oracle run with numIter = n; a stopped block satisfies every parity check."""
    rng = np.random.default_rng(33)
    bg, A, g = 1, 8424 * 2 - 24, 14040 * 2
    codec = TbBatchCodec(bg, '16QAM', A, g, precision='fp32', earlyStop=True)
    numTb = 6
    pl = torch.from_numpy(rng.integers(0, 2, (numTb, A)).astype(np.int8)).cuda()
    rm = codec.encode(pl)
    gen = torch.Generator(device='cuda')
    gen.manual_seed(5)
    llr = qam_awgn_llr(rm, 4, 8.3, generator=gen)
    out = codec.decode(llr, 12)
    iters = out["iters"].cpu().numpy().reshape(-1)
    assert iters.min() >= 1 and iters.max() <= 12 and len(set(iters.tolist())) > 1
    llr_h = llr.cpu().numpy()
    tb_h = out["tb"].cpu().numpy()
    for t in range(numTb):
        rr, _, p = O.rate_recover(llr_h[t], A, bg, 4, dtype=np.float32)
        for r in range(2):
            n_it = int(iters[t * 2 + r])
            bel = OC.decode_beliefs(rr[r:r + 1], bg, 384, 1, n_it, np.float32)
            hard = (bel < 0).astype(np.int8)
            assert np.array_equal(tb_h[t, r * codec.per:(r + 1) * codec.per], hard[0, :codec.per])
            if n_it < 12:
                assert O.parity_ok(hard[0], bg, 384, 1)
                if n_it > 1:   # and it had NOT converged one iteration earlier
                    prev = (OC.decode_beliefs(rr[r:r + 1], bg, 384, 1, n_it - 1, np.float32) < 0).astype(np.int8)
                    assert not O.parity_ok(prev[0], bg, 384, 1)
    # earlyStopFrom = k: the syndrome is tested from iteration k on -- blocks that converged earlier report k, the others are unchanged,
    # and every output still equals the fixed-iteration decode with the reported count (static and generic kernels)
    for kw, zc_case in (({}, None), ({"precision": "fp32"}, "generic")):
        k = int(np.median(iters))
        c2 = TbBatchCodec(bg, '16QAM', A, g, precision='fp32', earlyStop=True, earlyStopFrom=k)
        if zc_case == "generic":
            os.environ["NRLDPC_NO_STATIC_ROWS"] = "1"
            c2 = TbBatchCodec(bg, '16QAM', A, g, precision='fp32', earlyStop=True, earlyStopFrom=k, ownHandle=True)
        try:
            out2 = c2.decode(llr, 12)
            it2 = out2["iters"].cpu().numpy().reshape(-1)
        finally:
            os.environ.pop("NRLDPC_NO_STATIC_ROWS", None)
        assert np.array_equal(it2, np.maximum(iters, k))
        tb2 = out2["tb"].cpu().numpy()
        for t in range(numTb):
            rr, _, p = O.rate_recover(llr_h[t], A, bg, 4, dtype=np.float32)
            for r in range(2):
                hard = (OC.decode_beliefs(rr[r:r + 1], bg, 384, 1, int(it2[t * 2 + r]), np.float32) < 0).astype(np.int8)
                assert np.array_equal(tb2[t, r * codec.per:(r + 1) * codec.per], hard[0, :codec.per])
    # earlyStopFrom = 'auto' (NRLDPC_DEC_ES_AUTO): the first launch on a handle tests from iteration 1, every later one from
    # (smallest iteration count of the previous launch - 1); static and generic kernels, same protocol as above
    for generic in (False, True):
        if generic:
            os.environ["NRLDPC_NO_STATIC_ROWS"] = "1"
        try:
            c3 = TbBatchCodec(bg, '16QAM', A, g, precision='fp32', earlyStop=True, earlyStopFrom='auto', ownHandle=True)
            o1 = c3.decode(llr, 12)
            i1 = o1["iters"].cpu().numpy().reshape(-1)
            tb1 = o1["tb"].cpu().numpy()
            o2 = c3.decode(llr, 12)
            i2 = o2["iters"].cpu().numpy().reshape(-1)
            tb3 = o2["tb"].cpu().numpy()
            o3 = c3.decode(llr, 12)
            i3 = o3["iters"].cpu().numpy().reshape(-1)
        finally:
            os.environ.pop("NRLDPC_NO_STATIC_ROWS", None)
        assert np.array_equal(i1, iters) and np.array_equal(tb1, tb_h)
        k = max(1, int(iters.min()) - 1)
        assert np.array_equal(i2, np.maximum(iters, k)) and np.array_equal(i3, i2)
        for t in range(numTb):
            rr, _, p = O.rate_recover(llr_h[t], A, bg, 4, dtype=np.float32)
            for r in range(2):
                hard = (OC.decode_beliefs(rr[r:r + 1], bg, 384, 1, int(i2[t * 2 + r]), np.float32) < 0).astype(np.int8)
                assert np.array_equal(tb3[t, r * codec.per:(r + 1) * codec.per], hard[0, :codec.per])


@pytest.mark.parametrize("bg,A,mod,rate,numTb", [(2, 500, 'QPSK', 0.3, 23), (1, 600, '16QAM', 0.5, 40), (2, 24, 'QPSK', 0.25, 130),
                                                  (1, 20000, '64QAM', 0.75, 5), (1, 2000, 'QPSK', 0.4, 9),
                                                  # static kernels, 22-33 scheduled rows (split state: TMEM + shared planes)
                                                  (1, 8424 * 2 - 24, '16QAM', 0.48, 5), (2, 3816, 'QPSK', 0.3, 7),
                                                  (1, 5608, '16QAM', 0.45, 6)])
def test_batched_codec_small_z_multi_cb_per_cta(bg, A, mod, rate, numTb):
    """Batches of equally configured TBs: several code blocks share a CTA at small Zc, last group partially filled."""
    rng = np.random.default_rng(A)
    qm = {'QPSK': 2, '16QAM': 4, '64QAM': 6}[mod]
    g = int(np.ceil(A / rate / qm)) * qm
    codec = TbBatchCodec(bg, mod, A, g, precision='fp32')
    pl = rng.integers(0, 2, (numTb, A)).astype(np.int8)
    rm = codec.encode(torch.from_numpy(pl).cuda()).cpu().numpy()
    llr = np.empty(rm.shape, np.float32)
    for t in range(numTb):
        orm, p = O.tx_chain(pl[t], bg, g, qm)
        assert np.array_equal(rm[t, :len(orm)], orm)
        llr[t] = nr_link.qam_awgn_llr(orm, qm, 3.0 + 10 * np.log10(rate * qm), rng)
    out = codec.decode(torch.from_numpy(llr).cuda(), 5)
    tb, cbok, tbok = out["tb"].cpu().numpy(), out["cbOk"].cpu().numpy(), out["tbOk"].cpu().numpy()
    for t in range(numTb):
        rr, _, p = O.rate_recover(llr[t], A, bg, qm, dtype=np.float32)
        hard = (OC.decode_beliefs(rr, bg, p["Zc"], p["iLS"], 5, np.float32)[:, :p["K"]] < 0).astype(np.int8)
        otb, ocb = O.check_crc_and_merge(hard, p["K"], p["F"], p["C"])
        assert np.array_equal(tb[t], otb) and list(cbok[t].astype(bool)) == list(ocb)
        assert bool(tbok[t]) == bool(O.crc_check(otb, '24A'))
    counters = torch.zeros(8, dtype=torch.int64, device='cuda')
    codec.accumulate(out, counters, refPayload=torch.from_numpy(pl).cuda())
    c = counters.cpu().numpy()
    assert c[0] == numTb * codec.C and c[2] == numTb and c[1] == (cbok == 0).sum() and c[3] == (tbok == 0).sum()
    assert c[4] == (tb[:, :A] != pl).sum() and c[5] == 5 * numTb * codec.C


def test_full_size_config2_properties():
    """BASELINE configs[1] at full size (1024 code blocks, BG1 Zc=384, 16QAM, R=0.6): size-independent properties
    (valid code words, encoder linearity, encode -> noise-free recover -> decode round trip, CRC consistency) plus a
    64-block sample through the oracle bit for bit."""
    A, g, numTb = 8424 * 16 - 24, 14040 * 16, 64
    codec = TbBatchCodec(1, '16QAM', A, g, precision='fp32')
    gen = torch.Generator(device='cuda')
    gen.manual_seed(20261017)
    pa = torch.randint(0, 2, (numTb, A), dtype=torch.int8, device='cuda', generator=gen)
    pb = torch.randint(0, 2, (numTb, A), dtype=torch.int8, device='cuda', generator=gen)
    L, h, s = _native.lib(), _dev.handle(), _dev.stream_ptr()

    def coded_full(pl):
        tb = torch.empty((numTb, A + 24), dtype=torch.int8, device='cuda')
        _native.check(L.nrldpc_crc_attach(h, _dev.ptr(pl), numTb, A, A, 3, _dev.ptr(tb), s))
        cbs = torch.empty((numTb * 16, 8448), dtype=torch.int8, device='cuda')
        _native.check(L.nrldpc_segment(h, codec.cfg, _dev.ptr(tb), numTb, A + 24, A + 24, _dev.ptr(cbs), s))
        full = torch.empty((numTb * 16, 68 * 384), dtype=torch.int8, device='cuda')
        _native.check(L.nrldpc_encode(h, 1, 384, _dev.ptr(cbs), numTb * 16, _dev.ptr(full), 0, s))
        return cbs, full
    cbs_a, full_a = coded_full(pa)
    ok = torch.empty((numTb * 16,), dtype=torch.uint8, device='cuda')
    _native.check(L.nrldpc_parity_check(h, 1, 384, _dev.ptr(full_a), numTb * 16, _dev.ptr(ok), s))
    assert int(ok.sum()) == numTb * 16                                  # every encoded block is a code word
    assert torch.equal(full_a[:, :8448], cbs_a)                         # systematic
    # linearity over GF(2): enc(a ^ b) == enc(a) ^ enc(b) on raw code blocks
    cbs_b = torch.randint(0, 2, cbs_a.shape, dtype=torch.int8, device='cuda', generator=gen)
    ea, eb, eab = (torch.empty((numTb * 16, 66 * 384), dtype=torch.int8, device='cuda') for _ in range(3))
    for src, dst in ((cbs_a, ea), (cbs_b, eb), (cbs_a ^ cbs_b, eab)):
        _native.check(L.nrldpc_encode(h, 1, 384, _dev.ptr(src.contiguous()), numTb * 16, _dev.ptr(dst), 1, s))
    assert torch.equal(ea ^ eb, eab)
    # round trip without noise and with noise at 9 dB
    rm = codec.encode(pa)
    for llr in ((1 - 2 * rm.to(torch.float32)) * 4, qam_awgn_llr(rm, 4, 9.0, generator=gen)):
        out = codec.decode(llr.contiguous(), 8)
        assert torch.equal(out["tb"][:, :A], pa) and int(out["tbOk"].sum()) == numTb and int(out["cbOk"].sum()) == numTb * 16
    # oracle sample: first 4 TBs = 64 code blocks, fp32 bit for bit
    llr_h = llr[:4].cpu().numpy()
    tb_h = out["tb"][:4].cpu().numpy()
    for t in range(4):
        rr, _, p = O.rate_recover(llr_h[t], A, 1, 4, dtype=np.float32)
        hard = (OC.decode_beliefs(rr, 1, 384, 1, 8, np.float32)[:, :8448] < 0).astype(np.int8)
        otb, ocb = O.check_crc_and_merge(hard, 8448, 0, 16)
        assert np.array_equal(tb_h[t], otb) and all(ocb)


def test_wraparound_and_lbrm_recover():
    """E > Ncb (repetition: several LLRs accumulate into one buffer position, in stream order) and nRef-limited buffers."""
    rng = np.random.default_rng(2)
    for bg, A, mod, g, nref in [(2, 100, 'BPSK', 5000, 0), (2, 60, 'QPSK', 7000, 0), (1, 2400, 'QPSK', 4800, 7392), (1, 2400, 'QPSK', 4800, 5000),
                                 (2, 1000, '16QAM', 40000, 0)]:
        qm = O.MOD_ORDER[mod]
        llr = rng.standard_normal(g) * 3
        for prec_rv in (0, 2):
            dec = LdpcDecoder(bg, mod, 1, nref)
            h = Harq(prec_rv)
            got = dec.recoverRate(llr, A, h)
            want, buf, _ = O.rate_recover(llr, A, bg, qm, 1, nref, prec_rv)
            assert np.array_equal(got, want) and np.array_equal(h.decBuffer, buf)


def test_lbrm_small_bg2_start_beyond_buffer():
    """LBRM + small BG2 block (kb = 6: F >= 4 Zc) + rv 3: the start k0 of the reads lies BEYOND the filler-less circular
    buffer (k0 >= Ncb - F), which the reference wraps with (arange + start) % cirBufSize (ldpc.py:1148, 1407).  TX, the
    stand-alone rate recovery, and the fused decode (with and without a soft buffer, fp32 and fp64) must all agree with
    the oracle."""
    rng = np.random.default_rng(31)

    def rx_chain_lbrm(llr, A, bg, qm, nit, nref, rv, soft, dt):
        # O.rx_chain with the rate-recovered LLRs zero-extended to N (the reference's decode asserts on an LBRM-shortened
        # array, ldpc.py:1539; the fused chain treats the columns beyond Ncb as not received, i.e. LLR 0)
        rr, buf, p = O.rate_recover(llr, A, bg, qm, 1, nref, rv, soft, dt)
        N = (66 if bg == 1 else 50) * p["Zc"]
        rr = np.concatenate([rr, np.zeros((rr.shape[0], N - rr.shape[1]), dt)], axis=1)
        bits = O.decode(rr, bg, p["Zc"], p["iLS"], nit, dtype=dt)
        tbm, cb_ok = O.check_crc_and_merge(bits, p["K"], p["F"], p["C"])
        return tbm[:-24], cb_ok, bool(O.crc_check(tbm, "24A")), buf, p

    for A, nref, g in [(100, 400, 600), (150, 300, 420), (60, 300, 1000), (100, 600, 300)]:
        bg, mod, qm = 2, 'QPSK', 2
        p0 = O.derive_params(bg, A + 24)
        zc, K = p0["Zc"], p0["K"]
        F = K - (A + 24)
        ncb = min(50 * zc, nref)
        assert O.k0_start(bg, 3, ncb, 50 * zc, zc) >= ncb - F          # the case under test
        enc = LdpcEncoder(bg, mod, 1, nref, 0.3)
        tb = rng.integers(0, 2, A).astype(np.int8)
        coded = enc.encode(enc.doSegmentation(enc.appendCrc(tb, '24A')))
        ocbs, p = O.segment(O.crc_attach(tb, '24A'), bg)
        ocoded = O.encode(ocbs, bg, p['Zc'], p['iLS'])
        for prec, dt in (('fp64', np.float64), ('fp32', np.float32)):
            dec = LdpcDecoder(bg, mod, 1, nref, precision=prec)
            h1, h2, obuf = Harq(0), Harq(0), None
            for rv in (0, 3, 2, 3):
                rm = enc.rateMatch(coded, g, True, rv)
                assert np.array_equal(rm, O.rate_match(ocoded, bg, zc, K, F, g, qm, 1, nref, rv)), (A, nref, rv)
                llr = nr_link.qam_awgn_llr(rm.astype(np.int8), qm, -1.0, rng, np.float32)
                h1.rv = h2.rv = rv
                want_rr, obuf64, _ = O.rate_recover(llr, A, bg, qm, 1, nref, rv, None if h1.decBuffer is None else h1.decBuffer)
                got_rr = dec.recoverRate(llr.astype(np.float64), A, h1)
                assert np.array_equal(got_rr, want_rr) and np.array_equal(h1.decBuffer, obuf64), (A, nref, rv)
                # fused chain without history (closed-form row skipping) and with the soft buffer
                out, cbok, tbok = dec.decodeLLRs(llr if prec == 'fp32' else llr.astype(np.float64), A, 6,
                                                 harq=Harq(rv))
                otb, ocb, otbok, _, _ = rx_chain_lbrm(llr, A, bg, qm, 6, nref, rv, None, dt)
                assert np.array_equal(out, otb[:A]) and list(cbok) == list(ocb) and bool(tbok) == otbok, (A, nref, rv, prec)
                # device-resident batch codec: the fused load WITHOUT a soft buffer (closed-form row skipping)
                codec = TbBatchCodec(bg, mod, A, g, 1, nref, rv, prec)
                x = torch.from_numpy(np.stack([llr, llr])).cuda()
                res = codec.decode(x if prec == 'fp32' else x.double(), 6)
                torch.cuda.synchronize()
                assert np.array_equal(res['tb'][1, :A].cpu().numpy(), otb[:A]) and bool(res['tbOk'][1].item()) == otbok
                out, cbok, tbok = dec.decodeLLRs(llr if prec == 'fp32' else llr.astype(np.float64), A, 6, harq=h2)
                otb, ocb, otbok, obuf, _ = rx_chain_lbrm(llr, A, bg, qm, 6, nref, rv, obuf, dt)
                assert np.array_equal(out, otb[:A]) and bool(tbok) == otbok, (A, nref, rv, prec)
                assert np.array_equal(h2.decBuffer, obuf.astype(np.float64)), (A, nref, rv, prec)


def test_rate_match_repetition_and_lbrm():
    """rateMatch with E > Ncb - F (the circular buffer is sent more than once), LBRM buffers and every rv, C == 1 and
    C > 1: the staged scatter kernel's generic path (chunks that straddle interleaver rows / repeated buffers)."""
    rng = np.random.default_rng(12)
    for bg, A, mod, g, nref, nl in [(2, 100, 'BPSK', 5000, 0, 1), (2, 60, 'QPSK', 7000, 0, 1), (1, 2400, 'QPSK', 4800, 5000, 1),
                                    (2, 1000, '16QAM', 40000, 0, 1), (1, 9000, '64QAM', 60000, 0, 2), (2, 4000, '1024QAM', 30000, 0, 1),
                                    (1, 20000, '256QAM', 90000, 12000, 1)]:
        qm = O.MOD_ORDER[mod]
        enc = LdpcEncoder(bg, mod, nl, nref, 0.5)
        tb = rng.integers(0, 2, A).astype(np.int8)
        cbs = enc.doSegmentation(enc.appendCrc(tb, '24A'))
        ocbs, p = O.segment(O.crc_attach(tb, '24A'), bg)
        coded = enc.encode(cbs)
        ocoded = O.encode(ocbs, bg, p['Zc'], p['iLS'])
        assert np.array_equal(coded, ocoded)
        for rv in range(4):
            assert np.array_equal(enc.rateMatch(coded, g, True, rv),
                                  O.rate_match(ocoded, bg, p['Zc'], p['K'], p['F'], g, qm, nl, nref, rv)), (bg, A, mod, rv)


def test_error_behaviour_on_gpu():
    dec = LdpcDecoder(1, 'QPSK')
    h = Harq()
    h.decBuffer = np.zeros((3, 7))
    with pytest.raises(AssertionError):
        dec.recoverRate(np.zeros(22808), 10000, h)
    enc = LdpcEncoder(1, 'QPSK')
    enc.initialize(10024)
    with pytest.raises(AssertionError):
        enc.encode(np.zeros((2, 100), np.int8))
    p = ctypes.c_void_p()
    assert _native.lib().nrldpc_create(99, ctypes.byref(p)) == _native.ERR_ARG
    with pytest.raises(ValueError):
        _native.check(_native.lib().nrldpc_decode(_dev.handle(), 1, 100, 0, 0, None, 1, 100, 66, 1, 0, 22, None, None, None, None))


@pytest.mark.parametrize("bg,zc", [(1, 384), (1, 240), (2, 48), (2, 352), (1, 36), (2, 13), (1, 208)])
def test_parity_check_detects_single_bit_errors(bg, zc):
    """nrldpc_parity_check (bit-packed kernel for Zc % 16 == 0, byte-wise otherwise) against the oracle's full-row check:
    valid code words pass, a single flipped bit anywhere (core, punctured or extension column) fails."""
    rng = np.random.default_rng(100 * bg + zc)
    ils = O.set_index_of(zc)
    k, ncols = (22, 68) if bg == 1 else (10, 52)
    n_cb = 12
    cbs = rng.integers(0, 2, (n_cb, k * zc)).astype(np.int8)
    L, h, s = _native.lib(), _dev.handle(), _dev.stream_ptr()
    d_cbs = torch.from_numpy(cbs).cuda()
    full = torch.empty((n_cb, ncols * zc), dtype=torch.int8, device='cuda')
    _native.check(L.nrldpc_encode(h, bg, zc, _dev.ptr(d_cbs), n_cb, _dev.ptr(full), 0, s))
    host = full.cpu().numpy()
    assert np.array_equal(host, O.encode(cbs, bg, zc, ils, puncture=False))
    flipped = host.copy()
    pos = [0, zc - 1, zc, k * zc + 5, (k + 4) * zc - 1, (k + 4) * zc, ncols * zc - 1, (ncols - 7) * zc + zc // 2]
    for i, p in enumerate(pos):
        flipped[i, p] ^= 1                      # blocks 0..7 carry one error each, 8..11 stay valid
    ok = torch.empty((n_cb,), dtype=torch.uint8, device='cuda')
    d_fl = torch.from_numpy(flipped).cuda()
    _native.check(L.nrldpc_parity_check(h, bg, zc, _dev.ptr(d_fl), n_cb, _dev.ptr(ok), s))
    got = ok.cpu().numpy().astype(bool)
    want = np.array([O.parity_ok(flipped[i], bg, zc, ils) for i in range(n_cb)])
    assert list(want) == [False] * len(pos) + [True] * (n_cb - len(pos))
    assert list(got) == list(want)


def test_mixed_slot_256qam_harq_ir_device_soft_buffers():
    """BASELINE configs[2]: a slot carrying two transport blocks of different base graph / lifting size (256QAM, BG1
    4 layers R=0.75: C=21; BG2 2 layers R=0.3: C=10, both Zc=384 in the survey's 51-RB allocation, shrunk 4x here so
    that the oracle stays fast), incremental-redundancy HARQ with the rv sequence [0, 2, 3, 1] of harq.py:376, soft
    combining in DEVICE-resident buffers (the reference's HarqCW.decBuffer, harq.py:121,169) that never visit the
    host between transmissions.  Every transmission is checked against the oracle fed with the same LLRs and history:
    merged bits, CRC flags and the combined soft buffer bit for bit (fp32)."""
    rng = np.random.default_rng(2026)
    groups = [dict(bg=1, A=44040, nl=4, rate=0.75, snr=16.5), dict(bg=2, A=9480, nl=2, rate=0.3, snr=8.0)]
    qm, rvs, n_iter = 8, [0, 2, 3, 1], 6
    for gcfg in groups:
        bg, A, nl = gcfg["bg"], gcfg["A"], gcfg["nl"]
        g = int(np.ceil(A / gcfg["rate"] / (qm * nl))) * qm * nl
        codecs = {rv: TbBatchCodec(bg, '256QAM', A, g, txLayers=nl, rv=rv, precision='fp32') for rv in set(rvs)}
        c0 = codecs[0]
        numTb = 2
        pl = rng.integers(0, 2, (numTb, A)).astype(np.int8)
        d_pl = torch.from_numpy(pl).cuda()
        soft = torch.zeros((numTb * c0.C, c0.ncb - c0.F), dtype=torch.float32, device='cuda')   # decBuffer, on the device
        obuf = [None] * numTb
        first_ok = None
        for k, rv in enumerate(rvs):
            codec = codecs[rv]
            rm = codec.encode(d_pl)
            rm_h = rm.cpu().numpy()
            llr = np.empty(rm_h.shape, np.float32)
            for t in range(numTb):
                assert np.array_equal(rm_h[t], O.tx_chain(pl[t], bg, g, qm, nl, 0, rv)[0])
                llr[t] = nr_link.qam_awgn_llr(rm_h[t], qm, gcfg["snr"], rng)
            out = codec.decode(torch.from_numpy(llr).cuda(), n_iter, softBuffer=soft)
            tb, cbok, tbok = out["tb"].cpu().numpy(), out["cbOk"].cpu().numpy(), out["tbOk"].cpu().numpy()
            soft_h = soft.cpu().numpy().reshape(numTb, c0.C, -1)
            for t in range(numTb):
                rr, obuf[t], p = O.rate_recover(llr[t], A, bg, qm, nl, 0, rv, soft_buffer=obuf[t], dtype=np.float32)
                assert np.array_equal(soft_h[t], obuf[t]), "soft buffer differs after transmission %d" % k
                hard = (OC.decode_beliefs(rr, bg, p["Zc"], p["iLS"], n_iter, np.float32)[:, :p["K"]] < 0).astype(np.int8)
                otb, ocb = O.check_crc_and_merge(hard, p["K"], p["F"], p["C"])
                assert np.array_equal(tb[t], otb) and list(cbok[t].astype(bool)) == list(ocb)
                assert bool(tbok[t]) == bool(O.crc_check(otb, '24A'))
            if k == 0:
                first_ok = tbok.copy()
        assert not first_ok.all(), "the first transmission was meant to fail (operating point below the waterfall)"
        assert tbok.all() and np.array_equal(tb[:, :A], pl), "soft combining of four redundancy versions must decode"


def test_config2_full_size_grouped_call():
    """BASELINE configs[2] at FULL size through ONE library call (nrldpc_decode_tb_groups): the 51-RB 256QAM slot of SURVEY 8,
    codeword 1 = BG1, 4 layers, R = 0.75, A = 176 208 (C = 21, Zc = 384, G = 235 008), codeword 2 = BG2, 2 layers, R = 0.3,
    A = 37 896 (C = 10, Zc = 384, G = 127 296), incremental redundancy over rv = [0, 2, 3, 1] with DEVICE-resident soft buffers.
    Every transmission: merged bits, CRC flags and soft buffers equal the C oracle's on the same LLRs and history (fp32), and
    the grouped call equals two separate calls."""
    from neoradium_b200.batch import decode_groups
    rng = np.random.default_rng(51)
    cw = [dict(bg=1, A=176208, nl=4, g=235008, snr=17.2), dict(bg=2, A=37896, nl=2, g=127296, snr=7.6)]
    qm, rvs, n_iter, numTb = 8, [0, 2, 3, 1], 6, 1
    codecs = [{rv: TbBatchCodec(c["bg"], '256QAM', c["A"], c["g"], txLayers=c["nl"], rv=rv, precision='fp32') for rv in set(rvs)}
              for c in cw]
    assert (codecs[0][0].C, codecs[0][0].Zc, codecs[1][0].C, codecs[1][0].Zc) == (21, 384, 10, 384)
    pl = [rng.integers(0, 2, (numTb, c["A"])).astype(np.int8) for c in cw]
    soft = [torch.zeros((numTb * k[0].C, k[0].ncb - k[0].F), dtype=torch.float32, device='cuda') for k in codecs]
    soft2 = [torch.zeros_like(x) for x in soft]
    obuf = [[None] * numTb for _ in cw]
    tbok_hist = []
    for k, rv in enumerate(rvs):
        llr = []
        for i, c in enumerate(cw):
            rm_h = codecs[i][rv].encode(torch.from_numpy(pl[i]).cuda()).cpu().numpy()
            x = np.empty(rm_h.shape, np.float32)
            for t in range(numTb):
                assert np.array_equal(rm_h[t], O.tx_chain(pl[i][t], c["bg"], c["g"], qm, c["nl"], 0, rv)[0])
                x[t] = nr_link.qam_awgn_llr(rm_h[t], qm, c["snr"], rng)
            llr.append(x)
        d_llr = [torch.from_numpy(x).cuda() for x in llr]
        outs = decode_groups([codecs[0][rv], codecs[1][rv]], d_llr, n_iter, softBuffers=soft)
        sep = [codecs[i][rv].decode(d_llr[i], n_iter, softBuffer=soft2[i]) for i in range(2)]
        torch.cuda.synchronize()
        for i, c in enumerate(cw):
            for key in ("tb", "cbOk", "tbOk", "iters"):
                assert torch.equal(outs[i][key], sep[i][key]), (k, i, key)
            assert torch.equal(soft[i], soft2[i])
            tb, cbok, tbok = (outs[i][key].cpu().numpy() for key in ("tb", "cbOk", "tbOk"))
            C = codecs[i][0].C
            soft_h = soft[i].cpu().numpy().reshape(numTb, C, -1)
            for t in range(numTb):
                rr, obuf[i][t], p = O.rate_recover(llr[i][t], c["A"], c["bg"], qm, c["nl"], 0, rv, soft_buffer=obuf[i][t], dtype=np.float32)
                assert np.array_equal(soft_h[t], obuf[i][t]), "soft buffer differs after transmission %d" % k
                hard = (OC.decode_beliefs(rr, c["bg"], p["Zc"], p["iLS"], n_iter, np.float32)[:, :p["K"]] < 0).astype(np.int8)
                otb, ocb = O.check_crc_and_merge(hard, p["K"], p["F"], p["C"])
                assert np.array_equal(tb[t], otb) and list(cbok[t].astype(bool)) == list(ocb)
                assert bool(tbok[t]) == bool(O.crc_check(otb, '24A'))
        tbok_hist.append([bool(o["tbOk"].all().item()) for o in outs])
    assert tbok_hist[0] != [True, True], "the first transmission was meant to fail for at least one codeword"
    assert tbok_hist[-1] == [True, True], "soft combining of four redundancy versions must decode both codewords"
    for i in range(2):
        assert np.array_equal(outs[i]["tb"].cpu().numpy()[:, :cw[i]["A"]], pl[i])


@pytest.mark.timeout(180)
def test_multi_block_static_kernels_equal_generic_kernels(monkeypatch):
    """Statically scheduled kernels with SEVERAL code blocks per CTA (every lifting size that is no multiple of 32 or is
    below 224): padding threads (Zc = 8, 11, 22, 30: up to 21 of 32), a partly filled last group, C > 1 (CRC24B + the in-kernel
    transport-block CRC), soft buffers (the load path must not depend on the per-thread soft-buffer pointer), LBRM with
    k0 beyond the buffer, IEEE-half input.  Bit-identical to the generic (run-time row dispatch) kernels, which the rest of
    the suite pins to the oracle."""
    from neoradium_b200.modulation import awgn_llr
    cases = [(2, 500, 'QPSK', 1668, 23, 0, 0), (1, 600, '16QAM', 1200, 7, 0, 1), (2, 24, 'QPSK', 100, 130, 0, 0),
             (1, 9000, '16QAM', 18000, 5, 0, 0), (2, 100, 'QPSK', 600, 9, 400, 3), (2, 40, 'QPSK', 200, 3, 0, 1),
             (1, 3000, '64QAM', 6000, 2, 0, 2)]
    for bg, A, mod, g, numTb, nref, rv in cases:
        codec = TbBatchCodec(bg, mod, A, g, 1, nref, rv, 'fp32')
        x = awgn_llr(codec.encode(codec.random_payload(numTb, 3)), codec.qm, snr_db=2.0 * codec.qm, seed=1)
        out = codec.decode(x, 6)
        soft = torch.zeros((numTb * codec.C, codec.ncb - codec.F), dtype=torch.float32, device='cuda')
        out_s = codec.decode(x, 6, softBuffer=soft)
        out_h = codec.decode(x.half().float(), 6)
        out_h16 = codec.decode(x.half(), 6)
        torch.cuda.synchronize()
        monkeypatch.setenv("NRLDPC_NO_STATIC_MB", "1")
        ref_codec = TbBatchCodec(bg, mod, A, g, 1, nref, rv, 'fp32', ownHandle=True)
        soft_r = torch.zeros_like(soft)
        ref, ref_s = ref_codec.decode(x, 6), ref_codec.decode(x, 6, softBuffer=soft_r)
        torch.cuda.synchronize()
        monkeypatch.delenv("NRLDPC_NO_STATIC_MB")
        for k in ('tb', 'cbOk', 'tbOk', 'iters'):
            assert torch.equal(out[k], ref[k]) and torch.equal(out_s[k], ref_s[k]) and torch.equal(out_h[k], out_h16[k]), (bg, A, k)
        assert torch.equal(soft, soft_r)


def test_async_host_batches_and_concurrent_codecs():
    """Two host batches in flight (decodeLLRsAsync, alternating slots) and two codecs with private handles on two streams
    give exactly the results of the blocking single-stream calls, which equal the oracle's (low SNR: some blocks fail)."""
    bg, A, mod, qm, numTb = 1, 8424 * 3 - 24, '16QAM', 4, 6
    g = 14040 * 3
    rng = np.random.default_rng(77)
    batches, pls = [], []
    for b in range(3):
        llr = np.empty((numTb, g), np.float32)
        pl = rng.integers(0, 2, (numTb, A)).astype(np.int8)
        for t in range(numTb):
            orm, _ = O.tx_chain(pl[t], bg, g, qm)
            llr[t] = nr_link.qam_awgn_llr(orm, qm, 7.6 if t % 2 else 9.0, rng)
        batches.append(llr)
        pls.append(pl)
    dec = LdpcDecoder(bg, mod, 1, 0, precision='fp32')
    ref = [dec.decodeLLRs(x, A, 8) for x in batches]
    ref = [(a.copy(), b.copy(), c.copy()) for a, b, c in ref]
    for t in range(2):   # oracle on a sample of the first batch
        rr, _, p = O.rate_recover(batches[0][t], A, bg, qm, dtype=np.float32)
        hard = (OC.decode_beliefs(rr, bg, p["Zc"], p["iLS"], 8, np.float32)[:, :p["K"]] < 0).astype(np.int8)
        otb, ocb = O.check_crc_and_merge(hard, p["K"], p["F"], p["C"])
        assert np.array_equal(ref[0][0][t], otb[:A]) and list(ref[0][1][t]) == list(ocb)
    pend = [dec.decodeLLRsAsync(batches[i], A, 8, slot=i % 2) for i in range(2)]
    got = [pend[0].result()]
    got[0] = tuple(x.copy() for x in got[0])
    pend.append(dec.decodeLLRsAsync(batches[2], A, 8, slot=0))
    got += [tuple(x.copy() for x in p.result()) for p in pend[1:]]
    for r, o in zip(ref, got):
        assert all(np.array_equal(a, b) for a, b in zip(r, o))
    with pytest.raises(ValueError):
        dec.decodeLLRsAsync(batches[0][0], A, 8)
    # two device-resident codecs on two streams
    codecs = [TbBatchCodec(bg, mod, A, g, precision='fp32', ownHandle=True) for _ in range(2)]
    streams = [torch.cuda.Stream() for _ in range(2)]
    dl = [torch.from_numpy(x).cuda() for x in batches]
    torch.cuda.synchronize()
    outs = []
    for rep in range(3):
        for i in range(3):
            with torch.cuda.stream(streams[i % 2]):
                o = codecs[i % 2].decode(dl[i], 8)
            if rep == 2:
                outs.append(o)
    torch.cuda.synchronize()
    for r, o in zip(ref, outs):
        assert np.array_equal(r[0], o["tb"][:, :A].cpu().numpy()) and np.array_equal(r[1], o["cbOk"].cpu().numpy().astype(bool))
        assert np.array_equal(r[2], o["tbOk"].cpu().numpy().astype(bool))


def _decode2_cases():
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "decode2_cases.npz"))
    return g, [str(n) for n in g["names"]]


@pytest.mark.parametrize("name", _decode2_cases()[1])
def test_decode2_vs_reference_outputs_and_oracle(name):
    """LdpcDecoder.decode2 (ldpc.py:1421-1492) in float64: beliefs equal the unmodified reference's (committed fixture,
    no stop); with stopOnGoodParity the block stops after the first iteration whose hard decisions satisfy every check,
    which is what the oracle restatement does with stop_rule='all' (the reference tests the first row only)."""
    g, _ = _decode2_cases()
    bg, A, zc, ils, nit, K = (int(v) for v in g[name + "/meta"])
    alpha = float(g[name + "/alpha"])
    rr = g[name + "/rr"]
    dec = LdpcDecoder(bg, 'QPSK', 1, 0)
    dec.initialize(A + 24)
    assert (dec.liftingSize, dec.setIndex, dec.codeBlockSize) == (zc, ils, K)
    bel = dec.decode2(rr, nit, False, True, alpha, False)
    assert np.array_equal(bel, g[name + "/bel"])
    assert list(dec.lastIterations) == [nit] * rr.shape[0]
    assert np.array_equal(dec.decode2(rr, nit, True, False, alpha, False), (g[name + "/bel"][:, :K] < 0).astype(np.int8))
    bel_s = dec.decode2(rr, nit + 4, False, True, alpha, True)
    assert np.array_equal(bel_s, O.decode2(rr, bg, zc, ils, nit + 4, False, True, alpha, True, "all"))
    n_it = int(dec.lastIterations[0])
    assert np.array_equal(bel_s, O.decode2(rr, bg, zc, ils, n_it, False, True, alpha, False))
    if n_it < nit + 4:
        assert O.parity_ok((bel_s[0] < 0).astype(np.int8), bg, zc, ils)
    # offset min-sum (extension): |message| = max(alpha * min - beta, 0), against the oracle's restatement of the same rule
    bel_o = dec.decode2(rr, nit, False, True, alpha, False, beta=0.15)
    assert np.array_equal(bel_o, O.decode2(rr, bg, zc, ils, nit, False, True, alpha, False, beta=0.15))
    assert not np.array_equal(bel_o, bel) or not np.abs(rr).any()
    # compatibility switch: the reference's own stop test looks at the first base-graph row only (ldpc.py:841-843)
    bel_f = dec.decode2(rr, nit + 4, False, True, alpha, True, firstRowOnly=True)
    assert np.array_equal(bel_f, O.decode2(rr, bg, zc, ils, nit + 4, False, True, alpha, True, "first_row"))
    assert int(dec.lastIterations[0]) <= n_it
    # fp32 arithmetic: same hard decisions on these well-conditioned inputs
    d32 = LdpcDecoder(bg, 'QPSK', 1, 0, precision='fp32')
    d32.initialize(A + 24)
    if np.abs(rr).max() > 0 and name != "bg2_z13_zeros":
        assert np.array_equal(d32.decode2(rr, nit, True, False, alpha, False), (g[name + "/bel"][:, :K] < 0).astype(np.int8))


@pytest.mark.parametrize("bg", [1, 2])
def test_every_lifting_size(bg):
    """All 51 lifting sizes of TS 38.212 Table 5.3.2-1 on both base graphs (north_star: "every lifting size Zc up to
    384"): encoder bit-exact and a code word, fp32 beliefs after 3 iterations bit-exact vs the C oracle (every Zc picks its
    own kernel variant: several blocks per CTA below 193, dynamic-row and static one-block kernels above), fp64 on a
    subset, parity check accepts the code word and rejects a flipped bit."""
    import nr_tables
    rng = np.random.default_rng(100 + bg)
    P, n, k = O.bg_dims(bg)
    L, h, s = _native.lib(), _dev.handle(), _dev.stream_ptr()
    for ils, zs in enumerate(nr_tables.LIFTING_SETS):
        for zc in zs:
            C = 3
            cbs = rng.integers(0, 2, (C, k * zc)).astype(np.int8)
            dcbs = torch.from_numpy(cbs).cuda()
            full = torch.empty((C, n * zc), dtype=torch.int8, device='cuda')
            _native.check(L.nrldpc_encode(h, bg, zc, _dev.ptr(dcbs), C, _dev.ptr(full), 0, s))
            ofull = O.encode(cbs, bg, zc, ils, puncture=False)
            assert np.array_equal(full.cpu().numpy(), ofull), (bg, zc)
            ok = torch.empty((C,), dtype=torch.uint8, device='cuda')
            bad = full.clone()
            bad[1, int(rng.integers(0, n * zc))] ^= 1
            for src, exp in ((full, [1, 1, 1]), (bad, [1, 0, 1])):
                _native.check(L.nrldpc_parity_check(h, bg, zc, _dev.ptr(src), C, _dev.ptr(ok), s))
                assert ok.cpu().tolist() == exp, (bg, zc)
            coded = ofull[:, 2 * zc:]
            llr = ((1 - 2.0 * coded) * 1.5 + 1.3 * rng.standard_normal(coded.shape)).astype(np.float32)
            llr[:, (n - 2) * zc * 2 // 3:] = 0                      # unsent tail: exercises the exact row skipping
            x = torch.from_numpy(llr).cuda()
            for cdt, tdt, odt in ((_native.F32, torch.float32, np.float32),) + (((_native.F64, torch.float64, np.float64),) if zc in (2, 9, 52, 208, 384) else ()):
                bel = torch.empty((C, n * zc), dtype=tdt, device='cuda')
                _native.check(L.nrldpc_decode(h, bg, zc, _native.F32, cdt, _dev.ptr(x), C, (n - 2) * zc, n - 2, 3, 0, n, None,
                                              _dev.ptr(bel), None, s))
                obel = OC.decode_beliefs(llr.astype(odt), bg, zc, ils, 3, odt)
                assert np.array_equal(bel.cpu().numpy(), obel), (bg, zc, cdt)


def test_float16_llr_input_is_exact():
    """NRLDPC_F16 input (extension): half LLRs are widened exactly on load, so every output equals the one obtained
    from the same values passed as float32 -- staged (TMA) path at Zc=384, generic path at small Zc, soft-buffer path,
    host-batch path -- and therefore the oracle's."""
    rng = np.random.default_rng(16)
    for bg, A, mod, qm, g, numTb in ((1, 8424 * 3 - 24, '16QAM', 4, 14040 * 3, 5), (2, 500, 'QPSK', 2, 1668, 21),
                                     (1, 8424 * 2 - 24 - 7, '64QAM', 6, 14040 * 2 - 6, 3)):
        llr = np.empty((numTb, g), np.float32)
        for t in range(numTb):
            orm, _ = O.tx_chain(rng.integers(0, 2, A).astype(np.int8), bg, g, qm)
            llr[t] = nr_link.qam_awgn_llr(orm, qm, 3.0 + 10 * np.log10(A / g * qm) + (2.5 if qm > 2 else 0), rng)
        h16 = llr.astype(np.float16)
        as32 = h16.astype(np.float32)
        codec = TbBatchCodec(bg, mod, A, g, precision='fp32')
        o32 = codec.decode(torch.from_numpy(as32).cuda(), 6)
        o16 = codec.decode(torch.from_numpy(h16).cuda(), 6)
        for k in ('tb', 'cbOk', 'tbOk', 'iters'):
            assert torch.equal(o32[k], o16[k]), (bg, k)
        rr, _, p = O.rate_recover(as32[0], A, bg, qm, dtype=np.float32)
        hard = (OC.decode_beliefs(rr, bg, p["Zc"], p["iLS"], 6, np.float32)[:, :p["K"]] < 0).astype(np.int8)
        otb, _ = O.check_crc_and_merge(hard, p["K"], p["F"], p["C"])
        assert np.array_equal(o16['tb'][0].cpu().numpy(), otb)
        # soft-buffer (HARQ) path
        sb32 = torch.zeros((numTb * codec.C, codec.ncb - codec.F), dtype=torch.float32, device='cuda')
        sb16 = torch.zeros_like(sb32)
        a = codec.decode(torch.from_numpy(as32).cuda(), 4, softBuffer=sb32)
        b = codec.decode(torch.from_numpy(h16).cuda(), 4, softBuffer=sb16)
        assert torch.equal(sb32, sb16) and torch.equal(a['tb'], b['tb'])
        # host batch through the drop-in class
        dec = LdpcDecoder(bg, mod, 1, 0, precision='fp32')
        r32 = dec.decodeLLRs(as32, A, 6)
        r16 = dec.decodeLLRs(h16, A, 6)
        assert all(np.array_equal(x, y) for x, y in zip(r32, r16))
        one = dec.decodeLLRs(h16[0], A, 6)
        assert np.array_equal(one[0], r16[0][0])
    with pytest.raises(ValueError):   # decode() of rate-recovered blocks takes float32 / float64
        _native.check(_native.lib().nrldpc_decode(_dev.handle(), 1, 384, _native.F16, _native.F32, None, 1, 66 * 384, 66, 1, 0, 22,
                                                  None, None, None, _dev.stream_ptr()))
