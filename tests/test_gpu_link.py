"""Parity of the on-device link kernels (csrc/linksim.cu) with the reference modem: `Modem.modulate` bit
for bit, `Modem.getLLRsFromSymbols` within the stated tolerance (per-axis vs 2-D search: same value in exact
arithmetic), and the fused modulate -> AWGN -> LLR generator through its noise statistics, its reproducibility and the
oracle demapper applied to the same noise."""
import os

import numpy as np
import pytest
import torch

import nr_modem
from neoradium_b200.modulation import Modem, awgn_llr

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "modem_cases.npz"))
MODS = ["BPSK", "QPSK", "16QAM", "64QAM", "256QAM", "1024QAM"]


@pytest.mark.parametrize("mod", MODS)
def test_modem_against_reference_golden(mod):
    m = Modem(mod)
    assert np.array_equal(m.constellation, GOLD[mod + "_constellation"])            # bit-exact complex128
    sym = m.modulate(GOLD[mod + "_bits"])
    assert sym.dtype == np.complex128 and np.array_equal(sym, GOLD[mod + "_symbols"])
    assert np.array_equal(m.modulate(GOLD[mod + "_bits"][1]), GOLD[mod + "_symbols"][1])   # 1-D form
    n0, ref = float(GOLD[mod + "_n0"]), GOLD[mod + "_llr"]
    llr = m.getLLRsFromSymbols(GOLD[mod + "_noisy"], n0)
    assert llr.shape == ref.shape and llr.dtype == np.float64
    # tolerance: 1e-9 of the LLR scale (distances up to ~|y|^2 / N0 cancel in the difference)
    scale = np.maximum(1.0, np.abs(ref)) + np.abs(GOLD[mod + "_noisy"]).max() ** 2 / n0
    assert np.all(np.abs(llr - ref) <= 1e-9 * scale)
    hard = m.demodulate(GOLD[mod + "_noisy"], n0)
    sure = np.abs(ref) > 1e-6
    assert np.array_equal(hard[sure], GOLD[mod + "_hard"][sure])
    with pytest.raises(ValueError):
        m.modulate(np.zeros(m.qm + 1, np.int8)) if m.qm > 1 else (_ for _ in ()).throw(ValueError())


def test_modem_matches_oracle_on_random_symbols():
    rng = np.random.default_rng(3)
    for mod in MODS:
        qm = nr_modem.QM[mod]
        y = (rng.standard_normal(500) + 1j * rng.standard_normal(500)) * 0.9
        n0 = 0.07
        ref = nr_modem.llrs_maxlog(y, qm, n0)
        got = Modem(mod).getLLRsFromSymbols(y, n0)
        assert np.all(np.abs(got - ref) <= 1e-9 * (np.maximum(1.0, np.abs(ref)) + 20 / n0))


def _noise_from_qpsk(n_sym, n0, seed, offset=0):
    """The generator's noise per symbol, recovered exactly from a QPSK run (LLR = 4 a y / N0 per axis, linear in y)."""
    bits = torch.zeros(n_sym * 2, dtype=torch.int8, device='cuda')
    llr = awgn_llr(bits, 2, noise_var=n0, seed=seed, offset=offset).cpu().numpy().astype(np.float64).reshape(-1, 2)
    a = 1 / np.sqrt(2)
    y = llr * n0 / (4 * a)
    return (y[:, 0] - a) + 1j * (y[:, 1] - a)


def test_awgn_generator_statistics_and_reproducibility():
    n, n0 = 1 << 20, 0.25
    z = _noise_from_qpsk(n, n0, seed=11)
    for comp in (z.real, z.imag):
        assert abs(comp.mean()) < 4 * np.sqrt(n0 / 2 / n)
        assert abs(comp.var() / (n0 / 2) - 1) < 0.01
        k = ((comp / np.sqrt(n0 / 2)) ** 4).mean()
        assert abs(k - 3) < 0.05                                  # Gaussian kurtosis
        assert (np.abs(comp) > 4.5 * np.sqrt(n0 / 2)).sum() in range(0, 30)   # expected 7
    assert abs(np.corrcoef(z.real, z.imag)[0, 1]) < 0.01
    assert abs(np.corrcoef(z.real[:-1], z.real[1:])[0, 1]) < 0.01
    # same seed -> same noise; another seed -> different; offset shifts the stream (independent of the launch geometry)
    assert np.array_equal(z[:4096], _noise_from_qpsk(4096, n0, seed=11))
    assert not np.array_equal(z[:4096], _noise_from_qpsk(4096, n0, seed=12))
    assert np.array_equal(z[1000:3000], _noise_from_qpsk(2000, n0, seed=11, offset=1000))


@pytest.mark.parametrize("mod,snr", [("BPSK", 1.0), ("QPSK", 4.0), ("16QAM", 10.0), ("64QAM", 16.0), ("256QAM", 22.0), ("1024QAM", 28.0)])
def test_fused_awgn_llr_equals_oracle_on_the_same_noise(mod, snr):
    """The noise of symbol k depends on (seed, offset + k) only, not on the modulation: recover it from a QPSK run, apply
    it to the oracle's symbols and demap with the oracle -- the fused fp32 kernel must agree within fp32 accuracy."""
    qm = nr_modem.QM[mod]
    n_sym = 3001                                   # odd: exercises the partial last group of the vector stores
    rng = np.random.default_rng(qm)
    bits = rng.integers(0, 2, n_sym * qm).astype(np.int8)
    n0 = 10.0 ** (-snr / 10)
    got = awgn_llr(torch.from_numpy(bits).cuda(), qm, snr_db=snr, seed=5, offset=77).cpu().numpy().astype(np.float64)
    z = _noise_from_qpsk(n_sym, n0, seed=5, offset=77)
    y = nr_modem.modulate(bits, qm) + z
    ref = nr_modem.llrs_maxlog(y, qm, n0)
    tol = 2e-4 * (np.maximum(1.0, np.abs(ref)) + 4.0 / n0 * 0.05)
    assert np.all(np.abs(got - ref) <= tol), float(np.max(np.abs(got - ref) / tol))
    # decisions agree wherever the LLR is not within the tolerance of zero
    sure = np.abs(ref) > 2 * tol
    assert np.array_equal((got <= 0)[sure], (ref <= 0)[sure])


def test_gold_sequence_and_scrambling_bit_exact():
    """goldSequence / scrambleBits / scrambleLLRs (utils.py:70-94, pdsch.py:603-616) against the reference's stored
    sequences and the oracle: lengths around the 12-bit head and the 31-bit words, and sequences long enough for many
    threads (each thread jumps to its first LFSR word and walks 32 words)."""
    from neoradium_b200.scrambling import goldSequence, scramble_, scrambleBits, scrambleLLRs
    for k in range(8):
        c_init, n = (int(v) for v in GOLD["gold%d_cinit_n" % k])
        ref = np.unpackbits(GOLD["gold%d_bits" % k])[:n]
        got = goldSequence(c_init, n)
        assert isinstance(got, list) and np.array_equal(np.array(got), ref)
    rng = np.random.default_rng(9)
    for c_init, n in [(20001 * (1 << 15) + 17, 1), (5, 11), (6, 12 + 31 * 32), (7, 12 + 31 * 32 + 1), (99, 224640), (12345678, 1000003)]:
        c = np.array(nr_modem.gold_sequence(c_init, n), dtype=np.int8)
        assert np.array_equal(np.array(goldSequence(c_init, n), dtype=np.int8), c)
        bits = rng.integers(0, 2, n).astype(np.int8)
        sb = scrambleBits(c_init, bits)
        assert sb.dtype == np.int8 and np.array_equal(sb, bits ^ c)
        llr = rng.standard_normal(n)
        sl = scrambleLLRs(c_init, llr)
        assert sl.dtype == np.float64 and np.array_equal(sl, llr * (1 - 2 * np.float64(c)))
        d32 = torch.from_numpy(llr.astype(np.float32)).cuda()
        assert np.array_equal(scramble_(c_init, d32).cpu().numpy(), llr.astype(np.float32) * (1 - 2 * np.float32(c)))
    assert goldSequence(1, 0) == []


def test_symbol_input_host_pipeline_equals_llr_input():
    """LdpcDecoder.decodeSymbols / decodeSymbolsAsync (complex64 equalised symbols across PCIe, demapper + fused chain on the
    device) returns exactly what decodeLLRs returns for the fp32 max-log LLRs of the same symbols, and those LLRs agree with
    the reference demapper formula (oracle, float64) to a few ulp of float32."""
    import nr_modem
    from neoradium_b200 import LdpcDecoder, LdpcEncoder
    rng = np.random.default_rng(77)
    bg, mod, qm, A, numTb = 1, '16QAM', 4, 8424 * 2 - 24, 6
    G = 14040 * 2
    enc = LdpcEncoder(bg, mod, 1, 0, 0.6)
    n0 = 10 ** (-8.4 / 10)                                            # waterfall: some blocks fail, most pass
    sym = np.empty((numTb, G // qm), np.complex64)
    pl = rng.integers(0, 2, (numTb, A)).astype(np.int8)
    for t in range(numTb):
        x = nr_modem.modulate(enc.getRateMatchedCodeBlocks(pl[t], G).astype(np.int8), qm)
        sym[t] = (x + (rng.standard_normal(x.shape) + 1j * rng.standard_normal(x.shape)) * np.sqrt(n0 / 2)).astype(np.complex64)
    dec = LdpcDecoder(bg, mod, 1, 0, precision='fp32')
    tb_s, cb_s, ok_s = dec.decodeSymbols(sym, n0, A, 8)
    llr64 = np.stack([nr_modem.llrs_maxlog(sym[t].astype(np.complex128), qm, n0) for t in range(numTb)])
    llr32 = llr64.astype(np.float32)
    tb_l, cb_l, ok_l = dec.decodeLLRs(llr32, A, 8)
    assert np.array_equal(cb_s, cb_l) and np.array_equal(ok_s, ok_l)
    good = np.asarray(ok_l, bool)
    assert good.any() and np.array_equal(tb_s[good], tb_l[good]) and np.array_equal(tb_s[good], pl[good])
    p0 = dec.decodeSymbolsAsync(sym[:3], n0, A, 8, slot=0)
    p1 = dec.decodeSymbolsAsync(sym[3:], n0, A, 8, slot=1)
    a, b = p0.result(), p1.result()
    assert np.array_equal(np.concatenate([a[0], b[0]]), tb_s) and np.array_equal(np.concatenate([a[2], b[2]]), ok_s)


def test_fused_symbol_input_equals_demapper_plus_decoder():
    """nrldpc_decode_tb_symbols (TbBatchCodec.decode_symbols): the demapper inside the decoder's load phase -- and its
    demap-first fall-back for configurations without a fused form -- gives exactly the outputs of nrldpc_demap_maxlog (fp32 LLRs)
    followed by nrldpc_decode_tb: one block per CTA at several lifting sizes / modulations / redundancy versions, filler bits,
    a pitched symbol array, a partly present codeword (fewer symbols than G'/qm), and the fall-back cases (several blocks per
    CTA, repetition E > Ncb, BPSK)."""
    import torch
    from neoradium_b200 import _dev, _native
    from neoradium_b200.batch import TbBatchCodec
    from neoradium_b200.modulation import awgn_llr
    L, h = _native.lib(), _dev.handle()
    #        bg  mod      A      G      numTb rv  pad  cut
    cases = [(1, '16QAM', 8424 * 3 - 24, 14040 * 3, 5, 0, 0, 0), (1, '64QAM', 20000, 30000, 4, 2, 6, 0), (1, 'QPSK', 7000, 11000, 3, 3, 0, 0),
             (1, '256QAM', 8000, 12000, 3, 0, 0, 0), (2, 'QPSK', 3000, 9000, 4, 1, 2, 0), (1, '16QAM', 8424 * 2 - 24, 14040 * 2, 3, 0, 0, 500),
             (2, 'QPSK', 500, 1668, 7, 0, 0, 0), (1, '16QAM', 600, 1200, 5, 0, 0, 0), (2, 'QPSK', 100, 2000, 3, 0, 0, 0), (1, 'BPSK', 640, 1280, 2, 0, 0, 0)]
    for bg, mod, A, G, numTb, rv, pad, cut in cases:
        qm = {'BPSK': 1, 'QPSK': 2, '16QAM': 4, '64QAM': 6, '256QAM': 8}[mod]
        G = (G // qm) * qm
        codec = TbBatchCodec(bg, mod, A, G, 1, 0, rv, 'fp32', ownHandle=True)
        rm = codec.encode(codec.random_payload(numTb, 11))
        nsym = G // qm
        # symbols = hard constellation points + noise, built on the device from the modulator and torch's generator
        x = torch.empty((numTb, nsym, 2), dtype=torch.float32, device='cuda')
        _native.check(L.nrldpc_modulate(h, qm, _dev.ptr(rm), numTb * nsym, _native.F32, _dev.ptr(x), _dev.stream_ptr()))
        gen = torch.Generator(device='cuda'); gen.manual_seed(A)
        n0 = 10 ** (-{1: 4.0, 2: 6.0, 4: 12.0, 6: 18.0, 8: 24.0}[qm] / 10)   # Es/N0 at which most transport blocks decode
        y = x + torch.randn(x.shape, device='cuda', generator=gen) * float(np.sqrt(n0 / 2))
        buf = torch.zeros((numTb, nsym + pad, 2), dtype=torch.float32, device='cuda')
        buf[:, :nsym] = y
        sym = torch.view_as_complex(buf)[:, :nsym - cut]                      # pitch nsym + pad, nsym - cut symbols present
        llr = torch.empty((numTb, (nsym - cut) * qm), dtype=torch.float32, device='cuda')
        yc = y[:, :nsym - cut].contiguous()
        _native.check(L.nrldpc_demap_maxlog(h, qm, _native.F32, _dev.ptr(yc), numTb * (nsym - cut), float(n0), _native.F32, _dev.ptr(llr),
                                            _dev.stream_ptr()))
        ref = codec.decode(llr, 6)
        out = codec.decode_symbols(sym, n0, 6)
        torch.cuda.synchronize()
        for k in ('tb', 'cbOk', 'tbOk', 'iters'):
            assert torch.equal(out[k], ref[k]), (bg, mod, A, k)
        assert cut or int(ref['tbOk'].sum()) > 0, (bg, mod, A)
