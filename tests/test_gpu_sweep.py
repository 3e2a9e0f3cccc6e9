"""GPU tests of the sweep runner (SURVEY 8f row 3 / BASELINE configs[4]): the counters of a BLER point equal what the
CPU oracle finds on exactly the same payloads and channel realisations; the runner drives an adaptive SNR scheduler
through the reference's iterator + setData protocol (neoradium/snrhelper.py:14-254)."""
import numpy as np
import pytest
import torch

import nr_oracle as O
import nr_oracle_c as OC
from neoradium_b200 import dist as nd
from neoradium_b200.batch import TbBatchCodec
from neoradium_b200.modulation import awgn_llr
from neoradium_b200.sweep import BlerSweep

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bg,mod,qm,A,rate,snr,tbs,batch,nit", [(2, 'QPSK', 2, 500, 0.3, -0.6, 48, 16, 6),
                                                                 (1, '16QAM', 4, 8424 * 2 - 24, 0.6, 8.38, 12, 5, 8)])
def test_bler_point_counters_equal_oracle(bg, mod, qm, A, rate, snr, tbs, batch, nit):
    g = int(-(-A / rate // qm) * qm)
    codec = TbBatchCodec(bg, mod, A, g, precision='fp32')
    seed = 4242
    d = nd.bler_point(codec, tbs, snr, nit, seed=seed, batch_tbs=batch)
    # the same payloads and noise, regenerated batch by batch exactly as bler_point does, through the oracle
    sym = codec.sumE // qm
    tb_fail = cb_fail = bit_err = 0
    done = 0
    while done < tbs:
        n = min(batch, tbs - done)
        pl = codec.random_payload(n, seed=seed ^ 0x5bd1e995, firstTb=done)
        llr = awgn_llr(codec.encode(pl), qm, snr_db=snr, seed=seed, offset=done * sym).cpu().numpy()
        plh = pl.cpu().numpy()
        for t in range(n):
            rr, _, p = O.rate_recover(llr[t], A, bg, qm, dtype=np.float32)
            hard = (OC.decode_beliefs(rr, bg, p["Zc"], p["iLS"], nit, np.float32)[:, :p["K"]] < 0).astype(np.int8)
            otb, ocb = O.check_crc_and_merge(hard, p["K"], p["F"], p["C"])
            cb_fail += int(np.sum(~np.asarray(ocb, bool)))
            tb_fail += int(not O.crc_check(otb, '24A'))
            bit_err += int(np.sum(otb[:A] != plh[t]))
        done += n
    assert d["txBlocks"] == tbs and d["codeBlocks"] == tbs * codec.C and d["sumIterations"] == nit * tbs * codec.C
    assert (d["tbCrcFail"], d["cbCrcFail"], d["bitErrors"]) == (tb_fail, cb_fail, bit_err)
    assert 0 < tb_fail < tbs, "the point is meant to sit in the waterfall (some blocks fail, some pass)"


def test_sweep_runner_drives_an_adaptive_scheduler():
    class Sched:   # iterator + setData, as SnrScheduler: walks down in SNR until the BLER exceeds 50 %
        def __init__(self):
            self.snr, self.metrics, self.stop = 1.0, [], False

        def __iter__(self):
            return self

        def __next__(self):
            if self.stop or len(self.metrics) >= 6:
                raise StopIteration
            return self.snr

        def setData(self, metric, *others):
            self.metrics.append((self.snr, metric) + others)
            self.stop = metric > 50.0
            self.snr -= 1.5

    codec = TbBatchCodec(2, 'QPSK', 500, 1668, precision='fp32')
    sw = BlerSweep(codec, numIter=6, tbsPerPoint=64, batchTbs=32, seed=3)
    sch = Sched()
    pts = sw.run(sch)
    assert len(pts) == len(sch.metrics) >= 2
    assert [p["snr_db"] for p in pts] == [m[0] for m in sch.metrics]
    assert all(abs(100.0 * p["bler"] - m[1]) < 1e-12 for p, m in zip(pts, sch.metrics))
    assert pts[0]["bler"] < 0.5 < pts[-1]["bler"]            # stopped by the scheduler's own criterion
    again = BlerSweep(codec, numIter=6, tbsPerPoint=64, batchTbs=32, seed=3).point(pts[-1]["snr_db"])
    for k in ("tbCrcFail", "cbCrcFail", "bitErrors"):            # a point is a pure function of (seed, SNR, sizes)
        assert again[k] == pts[-1][k]


def test_payload_generator_and_batch_independence():
    """nrldpc_random_bits: the payload of a transport block is a function of (seed, global block index) only -- any split of
    a range into calls gives the same bits (unaligned heads / tails included), bits are balanced, seeds differ; and a BLER
    point counts the same whatever the batch size (what makes N-GPU counters equal 1-GPU counters)."""
    codec = TbBatchCodec(2, 'QPSK', 501, 1670, precision='fp32')      # odd A: calls start at unaligned stream offsets
    whole = codec.random_payload(37, seed=11)
    parts = torch.cat([codec.random_payload(5, 11, 0), codec.random_payload(1, 11, 5), codec.random_payload(31, 11, 6)])
    assert torch.equal(whole, parts)
    assert set(whole.unique().tolist()) == {0, 1} and abs(whole.float().mean().item() - 0.5) < 0.02
    assert not torch.equal(whole, codec.random_payload(37, seed=12))
    a = nd.bler_point(codec, 96, -0.8, 6, seed=9, batch_tbs=96)
    b = nd.bler_point(codec, 96, -0.8, 6, seed=9, batch_tbs=17)
    for k in ("tbCrcFail", "cbCrcFail", "bitErrors", "sumIterations"):
        assert a[k] == b[k], k
    assert 0 < a["tbCrcFail"] < 96
