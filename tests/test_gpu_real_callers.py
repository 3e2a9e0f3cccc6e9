"""The reference's own callers, UNMODIFIED, on top of the drop-in classes on the GPU:
``harq.HarqEntity`` (harq.py:376, 634-667) around ``neoradium_b200.LdpcEncoder`` and ``snrhelper.SnrScheduler``
(snrhelper.py:161-234) around the sweep runner.  The reference modules come from the pip-installed copy under baseline/_ref
(git-ignored, travels with the snapshot) or /root/reference; without either the tests are skipped."""
import os
import sys

import numpy as np
import pytest

import ref_loader

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_loader.reference_available(), reason="no reference copy (baseline/_ref or /root/reference)")]

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))


def _harq_loop(LdpcEncoder, mods, n, harqType, ebNoDb, nref=0, A=10000, record=None):
    harq, modulation, rnd, utils = mods
    modName, codeRate = "16QAM", 490 / 1024
    enc = LdpcEncoder(baseGraphNo=1, modulation=modName, txLayers=1, nRef=nref, targetRate=codeRate)
    ent = harq.HarqEntity(enc, harqType, 4)
    noiseStd = np.sqrt(1 / utils.toLinear(ebNoDb + 10 * np.log10(enc.qm * codeRate)))
    rangen, bitgen = rnd.random.getGenerator(123), rnd.random.getGenerator(7)
    modem = modulation.Modem(modName)
    out = []
    for t in range(n):
        txBlocks = [bitgen.bits(A) if ent.needNewData[0] else None]
        rm = ent.getRateMatchedCodeBlocks(txBlocks)
        y = modem.modulate(rm[0])
        y = y + rangen.awgn(y.shape, noiseStd)
        llr = modem.getLLRsFromSymbols(y, noiseStd ** 2)
        dec, errs = ent.decodeLLRs([llr], [A])
        out.append((np.array(rm[0]).copy(), np.array(dec[0]).copy(), int(errs[0])))
        if record is not None:
            record.append(ent)
        ent.goNext()
    return ent, out


@pytest.mark.parametrize("harqType,ebNoDb", [("IR", 3.0), ("CC", 2.2)])
def test_unmodified_harq_entity_on_dropin_equals_reference_codec(harqType, ebNoDb):
    """The same HARQ loop (Harq.ipynb cells 3-7, four processes) twice: the reference's NumPy codec and the drop-in codec,
    both driven by the unmodified HarqEntity, same seeds.  Rate-matched bits, decoded blocks, per-try statistics and the
    block-error flags must agree transmission by transmission (float64 precision: the reference's arithmetic)."""
    mods = ref_loader.load_reference("harq", "modulation", "random", "utils")
    rldpc = ref_loader.load_reference("ldpc")
    from neoradium_b200 import LdpcEncoder
    n = 16
    ent_ref, out_ref = _harq_loop(rldpc.LdpcEncoder, mods, n, harqType, ebNoDb)
    ent_gpu, out_gpu = _harq_loop(LdpcEncoder, mods, n, harqType, ebNoDb)
    for t, ((rm_r, dec_r, e_r), (rm_g, dec_g, e_g)) in enumerate(zip(out_ref, out_gpu)):
        assert np.array_equal(rm_r, rm_g), t
        assert e_r == e_g, t
        assert np.array_equal(dec_r, dec_g), t
    for k in ("txBlocks", "rxBlocks", "txBits", "rxBits"):
        assert np.array_equal(getattr(ent_ref, k), getattr(ent_gpu, k)), k
    assert ent_ref.numTimeouts == ent_gpu.numTimeouts
    assert ent_gpu.txBlocks[1] > 0, "the point is meant to need retransmissions"


def test_harq_notebook_statistic_and_device_resident_buffers(monkeypatch):
    """Harq.ipynb raw line 127: at Eb/N0 = 3 dB every first transmission fails and every second one succeeds
    (txBlocks per try [504 496 0 0], rxBlocks [0 496 0 0] for 1000 transmissions; here 96 with 16 processes).  The HARQ
    buffers the unmodified HarqCW holds are ManagedArrays: ndarrays whose pages stay on the device between transmissions."""
    import run_harq_notebook as nb
    from neoradium_b200 import LdpcEncoder, _dev, _managed
    harq, modulation, rnd, utils = ref_loader.load_reference("harq", "modulation", "random", "utils")
    res = nb.run(LdpcEncoder, harq.HarqEntity, modulation.Modem, rnd.random, utils.toLinear, 96)
    assert res["txBlocks_per_try"] == [48, 48, 0, 0] and res["rxBlocks_per_try"] == [0, 48, 0, 0], res
    assert res["numTimeouts"] == 0 and abs(res["bler_pct"] - 50.0) < 1e-9
    if not _dev.managed_ok():
        pytest.skip("no concurrent managed access on this device")
    monkeypatch.setenv("NRLDPC_MANAGED_MIN", "0")     # (the 10 000-bit block of the notebook is below the default size threshold)
    enc = LdpcEncoder(baseGraphNo=1, modulation="16QAM", txLayers=1, targetRate=490 / 1024)
    ent = harq.HarqEntity(enc, "IR", 2)
    rng = np.random.default_rng(0)
    rm = ent.getRateMatchedCodeBlocks([rng.integers(0, 2, 10000).astype(np.int8)])
    cw = ent.curProcess.cws[0]
    assert _managed.root_of(cw.encBuffer, np.int8) is not None and cw.encBuffer.shape == (2, 66 * enc.liftingSize)
    llr = (1.0 - 2.0 * rm[0]) * 0.3 + rng.standard_normal(len(rm[0]))      # a transmission that fails
    dec, errs = ent.decodeLLRs([llr], [10000])
    if errs[0] > 0:
        assert _managed.root_of(cw.decBuffer, np.float64) is not None
        first = np.array(cw.decBuffer)                                    # host read: pages migrate, values are exact
        ent.goNext(); ent.goNext()                                          # back to the same process (2 processes)
        rm2 = ent.getRateMatchedCodeBlocks([None])
        llr2 = (1.0 - 2.0 * rm2[0]) * 0.3 + rng.standard_normal(len(rm2[0]))
        buf_id = id(cw.decBuffer)
        ent.decodeLLRs([llr2], [10000])
        if cw.decBuffer is not None:                                         # (reset to None when the block got through)
            assert id(cw.decBuffer) == buf_id and not np.array_equal(first, cw.decBuffer)


def test_unmodified_snr_scheduler_drives_the_sweep():
    """snrhelper.SnrScheduler (adaptive SNR iterator, snrhelper.py:161-234) fed by BlerSweep.run: it must bracket the
    waterfall of the BG2 QPSK R=0.3 code and walk it from the 100 % to the 0 % BLER end, every point on the device."""
    snrhelper = ref_loader.load_reference("snrhelper")
    from neoradium_b200.batch import TbBatchCodec
    from neoradium_b200.sweep import BlerSweep
    codec = TbBatchCodec(2, 'QPSK', 500, 1668, precision='fp32')
    sched = snrhelper.SnrScheduler(snr0=-1.0, step=0.5)
    pts = BlerSweep(codec, numIter=6, tbsPerPoint=256, batchTbs=128, seed=5).run(sched)
    snrs = [p["snr_db"] for p in pts]
    assert len(pts) >= 4 and len(set(snrs)) == len(snrs)
    by_snr = sorted(pts, key=lambda p: p["snr_db"])
    assert by_snr[0]["bler"] == 1.0 and by_snr[-1]["bler"] == 0.0
    assert all(a["bler"] >= b["bler"] - 0.05 for a, b in zip(by_snr, by_snr[1:]))   # monotone up to sampling noise


def test_polar_crc_reuse():
    """The Polar codec inherits the CRC API from ChanCodeBase (polar.py:117, 235, 524, 973-977).  With the three classmethods
    patched to the CUDA ones the reference's Polar encoder / SCL decoder produce exactly what they
    produce with their own CRC: segmentation with appendCrc (1-D blocks), list decoding with checkCrc on the 2-D candidate
    matrix, for the DCI ('24C'), UCI ('11' / '6') configurations."""
    polar, rnd = ref_loader.load_reference("polar", "random")
    from neoradium_b200 import ChanCodeBase as GpuCrc
    rng = np.random.default_rng(5)

    class GpuPolarEncoder(polar.PolarEncoder):
        appendCrc = GpuCrc.appendCrc
        checkCrc = GpuCrc.checkCrc
        getCrc = GpuCrc.getCrc

    class GpuPolarDecoder(polar.PolarDecoder):
        appendCrc = GpuCrc.appendCrc
        checkCrc = GpuCrc.checkCrc
        getCrc = GpuCrc.getCrc

    for dataType, payload, e in (("DCI", 40, 216), ("UCI", 30, 120), ("UCI", 16, 40)):
        msg = rng.integers(0, 2, payload).astype(np.int8)
        outs = []
        for Enc, Dec in ((polar.PolarEncoder, polar.PolarDecoder), (GpuPolarEncoder, GpuPolarDecoder)):
            enc = Enc(payload, e, dataType)
            cbs = enc.doSegmentation(msg)
            rm = enc.rateMatch(enc.encode(cbs))
            llr = (1.0 - 2.0 * rm) * 4.0 + np.random.default_rng(9).standard_normal(rm.shape) * 1.2     # [C, E]
            dec = Dec(payload, e, dataType)
            decoded, crcErrors = dec.decode(dec.recoverRate(llr))
            outs.append((np.asarray(cbs), np.asarray(rm), np.asarray(decoded), crcErrors))
        for a, b in zip(outs[0], outs[1]):
            assert np.array_equal(a, b), dataType
        assert np.array_equal(outs[1][2], msg) and outs[1][3] == 0
