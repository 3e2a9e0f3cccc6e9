"""``ChanCodeBase`` -- drop-in for neoradium/chancodebase.py:49-189 on top of libnrldpc (CUDA, sm_100a).

Same classmethods, argument conventions, shapes and dtypes as the reference; the long division runs on the GPU
(csrc/crc.cu).  No CPU fallback: without a CUDA device every call raises.
"""
import numpy as np
import torch

from . import _dev, _native

# chancodebase.py:37-44 (kept for API compatibility; the kernels carry the same polynomials, csrc/nrldpc_internal.cuh)
strToPoly = {
    '6':   [1, 1, 0, 0, 0, 0, 1],
    '11':  [1, 1, 1, 0, 0, 0, 1, 0, 0, 0, 0, 1],
    '16':  [1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1],
    '24A': [int(c) for c in "1100001100100110011111011"],
    '24B': [int(c) for c in "1100000000000000001100011"],
    '24C': [int(c) for c in "1101100101011000100010111"],
}


def _poly_id(poly):
    if not isinstance(poly, str) or poly not in _native.CRC_IDS:
        raise KeyError(poly)        # the reference fails with KeyError on strToPoly[poly]
    return _native.CRC_IDS[poly]


class ChanCodeBase:
    """Base class of the channel-coding classes (CRC attach / check), chancodebase.py:49."""
    LARGE_LLR = 1e20    # chancodebase.py:52

    def __init__(self):
        pass

    @classmethod
    def getCrcLen(cls, poly):                                   # chancodebase.py:59-63
        if poly[:2] == "24":
            return 24
        return int(poly)

    @classmethod
    def _crc_device(cls, bits, poly):
        """bits: 1-D/2-D host array -> (device int8 [m, L], m, L, flat)"""
        bits = np.asarray(bits)
        flat = bits.ndim == 1
        b2 = bits[None, :] if flat else bits
        if b2.ndim != 2:
            raise ValueError("'bits' must be a 1-D or 2-D array")
        d = _dev.to_dev(b2, torch.int8)
        return d, b2.shape[0], b2.shape[1], flat, bits.dtype

    @classmethod
    def getCrc(cls, bits, poly):
        """CRC bits of one ([L] -> [c]) or several ([N, L] -> [N, c]) bit streams (chancodebase.py:83-128).
        The result dtype follows the reference (the input dtype promoted with int64)."""
        pid = _poly_id(poly)
        d, m, n, flat, in_dtype = cls._crc_device(bits, poly)
        c = cls.getCrcLen(poly)
        out = torch.empty((m, c), dtype=torch.int8, device=d.device)
        if n > 0:
            _native.check(_native.lib().nrldpc_crc(_dev.handle(), _dev.ptr(d), m, n, n, pid, _dev.ptr(out), None,
                                                   _dev.stream_ptr()))
        else:
            out.zero_()
        res = _dev.to_host(out).astype(np.result_type(in_dtype, np.int64))
        return res[0] if flat else res

    getCrcOld = getCrc                                          # chancodebase.py:67-79 (same values)

    @classmethod
    def checkCrc(cls, bits, poly):
        """True where the stream (data followed by its CRC) divides evenly (chancodebase.py:132-157)."""
        pid = _poly_id(poly)
        d, m, n, flat, _ = cls._crc_device(bits, poly)
        ok = torch.empty((m,), dtype=torch.uint8, device=d.device)
        if n > 0:
            _native.check(_native.lib().nrldpc_crc_check(_dev.handle(), _dev.ptr(d), m, n, n, pid, _dev.ptr(ok),
                                                         _dev.stream_ptr()))
        else:
            ok.fill_(1)
        res = _dev.to_host(ok).astype(np.bool_)
        return res[0] if flat else res

    @classmethod
    def appendCrc(cls, bits, poly):
        """The stream(s) followed by their CRC (chancodebase.py:161-189)."""
        pid = _poly_id(poly)
        d, m, n, flat, in_dtype = cls._crc_device(bits, poly)
        c = cls.getCrcLen(poly)
        out = torch.empty((m, n + c), dtype=torch.int8, device=d.device)
        _native.check(_native.lib().nrldpc_crc_attach(_dev.handle(), _dev.ptr(d), m, n, n, pid, _dev.ptr(out),
                                                      _dev.stream_ptr()))
        res = _dev.to_host(out).astype(np.result_type(in_dtype, np.int64))
        # the data part is returned as given (the reference appends to the caller's values, it does not mask them)
        src = np.asarray(bits)
        if flat:
            res = res[0]
            res[:n] = src
        else:
            res[:, :n] = src
        return res
