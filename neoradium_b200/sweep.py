"""BLER-vs-SNR sweep runner: whole SNR points run on the device(s) -- payload, TX chain, fused
QAM + AWGN + LLR, fused RX chain, counters -- with one all-reduce of the counter vector per point (dist.bler_point).

The SNR values come from any iterable.  NeoRadium's adaptive ``SnrScheduler`` (neoradium/snrhelper.py:14-254) works
unmodified: it is an iterator that expects ``setData(metric, *others)`` once per step, which ``run`` provides with the
BLER in percent (the convention of PDSCH-BLER.ipynb) -- every rank feeds it the same reduced value, so all ranks walk
the same SNR sequence."""
from . import dist as _dist


class BlerSweep:
    def __init__(self, codec, numIter=8, tbsPerPoint=2048, batchTbs=256, seed=0, group=None):
        self.codec, self.numIter, self.tbsPerPoint, self.batchTbs = codec, int(numIter), int(tbsPerPoint), int(batchTbs)
        self.seed, self.group = int(seed), group
        self.points = []

    def point(self, snr_db):
        """Counters of one SNR point (dict, identical on every rank); the point's seed mixes in the SNR."""
        d = _dist.bler_point(self.codec, self.tbsPerPoint, float(snr_db), self.numIter,
                             seed=self.seed * 7919 + int(round(float(snr_db) * 1000)) + (1 << 20),
                             batch_tbs=self.batchTbs, group=self.group)
        d["snr_db"] = float(snr_db)
        return d

    def run(self, snrs, on_point=None):
        """Iterate `snrs` (list / generator / SnrScheduler); returns the list of per-point dicts."""
        adaptive = hasattr(snrs, "setData")
        for snr in snrs:
            d = self.point(snr)
            self.points.append(d)
            if on_point is not None:
                on_point(d)
            if adaptive:
                snrs.setData(100.0 * d["bler"], d["bitErrors"] / max(1, d["txBlocks"] * self.codec.A))
        return self.points
