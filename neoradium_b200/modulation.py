"""Drop-in mirror of ``neoradium.modulation.Modem`` for the three methods the LDPC loops use (modulation.py:127-235):
``modulate``, ``getLLRsFromSymbols(useMax=True)`` and ``demodulate``.  SURVEY.md 8f row 1: the modem is the producer
of the decoder's LLRs in every BLER notebook; here it runs in the CUDA kernels of csrc/linksim.cu behind the C-ABI
(`nrldpc_modulate`, `nrldpc_demap_maxlog`, `nrldpc_awgn_llr`).  Host NumPy arrays in and out, like the reference;
``awgn_llr`` is the device-resident fused form (modulate -> AWGN -> LLR) used by sweeps and benchmarks.

The exact log-sum-exp demapper (``useMax=False``) is not built: none of the LDPC paths of the reference use it.
"""
import numpy as np
import torch

from . import _dev, _native
from .params import MOD_ORDER


class Modem:
    mod2qm = dict(MOD_ORDER)   # modulation.py:30-31

    def __init__(self, modulation='QPSK'):
        if modulation not in self.mod2qm:
            raise ValueError("Unsupported modulation '%s'!" % modulation)
        self.modulation = modulation
        self.qm = self.mod2qm[modulation]
        self._constellation = None

    @property
    def constellation(self):
        """complex128 [2^qm] lookup table (modulation.py:75), produced by the device kernel from all labels."""
        if self._constellation is None:
            qm = self.qm
            labels = np.arange(1 << qm)
            bits = ((labels[:, None] >> np.arange(qm - 1, -1, -1)[None, :]) & 1).astype(np.int8).reshape(-1)
            self._constellation = self.modulate(bits)
        return self._constellation

    def __repr__(self):
        return self.print(getStr=True)

    def print(self, indent=0, title=None, getStr=False):
        rep = "\n" if indent == 0 else ""
        rep += indent * ' ' + ("Modem Properties:" if title is None else title) + "\n"
        rep += indent * ' ' + "  Modulation Type ...........: %s\n" % self.modulation
        rep += indent * ' ' + "  Qm ........................: %d\n" % self.qm
        rep += indent * ' ' + "  Num constellation points ..: %d\n" % (1 << self.qm)
        if getStr:
            return rep
        print(rep)

    # ------------------------------------------------------------------------------------------------------------------
    def modulate(self, bitstreams):
        """bits 1-D [n*qm] or 2-D [m, n*qm] -> complex128 symbols [n] / [m, n] (modulation.py:127-157)."""
        b = np.asarray(bitstreams)
        if b.shape[-1] % self.qm:
            raise ValueError("The length of 'bitstream' (%d) must be a multiple of 'qm' (%d)!" % (b.shape[-1], self.qm))
        d = _dev.to_dev(b.reshape(-1), torch.int8)
        n = d.numel() // self.qm
        out = torch.empty((n, 2), dtype=torch.float64, device=d.device)
        if n:
            _native.check(_native.lib().nrldpc_modulate(_dev.handle(), self.qm, _dev.ptr(d), n, _native.F64, _dev.ptr(out),
                                                        _dev.stream_ptr()))
        sym = _dev.to_host(out).view(np.complex128).reshape(b.shape[:-1] + (b.shape[-1] // self.qm,))
        return sym

    def getLLRsFromSymbols(self, symbols, noiseVar, useMax=True):
        """complex symbols [n] / [m, n] -> float64 LLRs [n*qm] / [m, n*qm], positive => 0 (modulation.py:159-204)."""
        if not useMax:
            raise NotImplementedError("only the max-log demapper (useMax=True) is built; see neoradium_b200/modulation.py")
        s = np.ascontiguousarray(np.asarray(symbols, dtype=np.complex128))
        d = _dev.to_dev(s.view(np.float64).reshape(-1, 2), torch.float64)
        n = d.shape[0]
        out = torch.empty((n * self.qm,), dtype=torch.float64, device=d.device)
        if n:
            _native.check(_native.lib().nrldpc_demap_maxlog(_dev.handle(), self.qm, _native.F64, _dev.ptr(d), n, float(noiseVar),
                                                            _native.F64, _dev.ptr(out), _dev.stream_ptr()))
        return _dev.to_host(out).reshape(s.shape[:-1] + (s.shape[-1] * self.qm,))

    def demodulate(self, symbols, noiseVar, useMax=True):
        """Hard decisions of the LLRs (modulation.py:206-235)."""
        return np.int8((self.getLLRsFromSymbols(symbols, noiseVar, useMax) <= 0) * 1)


def awgn_llr(bits, qm, snr_db=None, noise_var=None, seed=0, offset=0, out=None):
    """Device-resident fused link: int8 CUDA tensor of bits [..., n*qm] -> fp32 LLRs of the same shape after unit-energy
    Gray QAM + complex AWGN (Es/N0 = snr_db, or the given noise variance) + max-log demapping.  Symbol k of the call
    uses noise counter offset + k under `seed`, so sharded / batched sweeps are reproducible."""
    assert bits.is_cuda and bits.dtype == torch.int8 and bits.is_contiguous() and bits.numel() % qm == 0
    if noise_var is None:
        noise_var = 10.0 ** (-float(snr_db) / 10.0)
    if out is None:
        out = torch.empty(bits.shape, dtype=torch.float32, device=bits.device)
    assert out.is_contiguous() and out.dtype == torch.float32 and out.numel() == bits.numel()
    _native.check(_native.lib().nrldpc_awgn_llr(_native.handle(bits.device.index), qm, _dev.ptr(bits), bits.numel() // qm,
                                                float(noise_var), int(seed) & (2 ** 64 - 1), int(offset), _dev.ptr(out),
                                                _dev.stream_ptr()))
    return out
