"""``LdpcEncoder`` / ``LdpcDecoder`` -- drop-in for neoradium/ldpc.py:670-1619 on top of libnrldpc (CUDA, sm_100a).

Same class names, constructor arguments, methods, attributes, exceptions, shapes and dtypes as the reference, so
``harq.HarqEntity`` and the Playground notebooks run unmodified on top of these classes.  Host code here only derives
the integer parameters (as the reference does) and moves caller-owned NumPy arrays to/from device buffers; every
operation on payload bits or LLRs is a CUDA kernel behind the C-ABI of include/nrldpc.h.  There is no CPU fallback.

Extensions (all opt-in, defaults reproduce the reference):
  * ``precision``  'fp64' (default: the reference's float64 arithmetic, bit for bit) or 'fp32' (the same operation order
    evaluated in float32 -- the throughput mode the benchmarks use)
  * ``LdpcDecoder.decodeLLRs(llrs, txBlockSize, numIter, harq=None)``: the fused chain of harq.py:165-173 in one
    kernel, for one transport block or a batch ([numTb, G]) of equally configured ones
  * ``earlyStop``  stop a code block once all parity checks hold (the reference always runs numIter iterations);
    ``earlyStopFrom=k`` tests the syndrome from iteration k on (a block runs at least k iterations);
    ``earlyStopFrom='auto'`` lets every launch start the tests one iteration before the first block of the previous
    launch on the same handle converged (a test costs ~15 % of an iteration and is wasted before that point)
Deliberate deviations ("do not copy"): the base graph is re-derived whenever (Zc, iLS) change
(the reference caches a stale one, ldpc.py:777) and ``isValidCodedBlock`` checks all rows (ldpc.py:841-843 returns
after the first).
"""
import functools
import warnings

import numpy as np
import torch

from . import _dev, _native, params
from .chancodebase import ChanCodeBase


def deprecated(replacement=None):
    # same behaviour as neoradium/utils.py:145-165
    def decorator(func):
        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            message = f"Call to deprecated function {func.__name__}."
            if replacement:
                message += f" Use {replacement} instead."
            warnings.warn(message, category=DeprecationWarning, stacklevel=2)
            return func(*args, **kwargs)
        return wrapper
    return decorator


_TORCH_F = {'fp32': torch.float32, 'fp64': torch.float64}
_NATIVE_F = {'fp32': _native.F32, 'fp64': _native.F64}


# **********************************************************************************************************************
class LdpcBase(ChanCodeBase):
    """Base class of the LDPC encoder and decoder (neoradium/ldpc.py:670-892)."""

    def __init__(self, baseGraphNo=1, modulation='QPSK', txLayers=1, nRef=0):
        super().__init__()
        self.baseGraphNo = baseGraphNo
        if self.baseGraphNo not in [1, 2]:
            raise ValueError("'baseGraphNo' must be 1 or 2!")
        self.modulation = modulation
        if self.modulation not in params.MOD_ORDER:
            raise ValueError("Invalid 'modulation' value!")
        self.qm = params.MOD_ORDER[self.modulation]
        self.maxCodeBlockSize = 8448 if baseGraphNo == 1 else 3840     # Kcb
        self.txBlockSize = 0        # B, including the 24-bit transport-block CRC
        self.numCodeBlocks = 0      # C
        self.codeBlockSize = 0      # K
        self.liftingSize = 0        # Zc
        self.setIndex = -1          # iLS
        self._baseGraph = None
        self._baseGraphKey = None
        self.numFillerBits = 0      # F
        self.txLayers = txLayers
        self.nRef = nRef

    # ------------------------------------------------------------------------------------------------------------------
    def __repr__(self):
        return self.print(getStr=True)

    def print(self, indent, title, getStr):
        repStr = "\n" if indent == 0 else ""
        repStr += indent * ' ' + title + "\n"
        repStr += indent * ' ' + "  Base Graph:         %d\n" % (self.baseGraphNo)
        repStr += indent * ' ' + "  Modulation:         %s\n" % (self.modulation)
        repStr += indent * ' ' + "  Number of layers:   %d\n" % (self.txLayers)
        if getStr:
            return repStr
        print(repStr)

    # ------------------------------------------------------------------------------------------------------------------
    @property
    def baseGraph(self):
        """int16 [P, n] lifted base graph, -1 = no edge (ldpc.py:775-789), from the library's TS 38.212 tables."""
        assert (self.setIndex >= 0 and self.liftingSize > 0), \
            "'Base Graph' not available. Encoder not initialized yet!"
        key = (self.baseGraphNo, self.setIndex, self.liftingSize)
        if self._baseGraph is None or self._baseGraphKey != key:
            P, n, _ = params.bg_dims(self.baseGraphNo)
            bg = np.empty((P, n), dtype=np.int16)
            _native.check(_native.lib().nrldpc_base_graph(self.baseGraphNo, self.setIndex, self.liftingSize,
                                                          bg.ctypes.data))
            self._baseGraph, self._baseGraphKey = bg, key
        return self._baseGraph

    # ------------------------------------------------------------------------------------------------------------------
    @deprecated(replacement="isValidCodedBlock")
    def isValidCodeword(self, codeWord):
        return self.isValidCodedBlock(codeWord)

    def isValidCodedBlock(self, codedBlock, firstRowOnly=False):
        """True if ``codedBlock`` (length n*Zc, un-punctured) satisfies every parity check (ldpc.py:825-843).
        ``firstRowOnly=True`` (compatibility switch) tests what the reference's loop actually tests: it returns from inside
        the first base-graph row (ldpc.py:841-843), so only those Zc checks are looked at."""
        cb = np.asarray(codedBlock).reshape(-1)
        z = self.liftingSize
        P, n, _ = params.bg_dims(self.baseGraphNo)
        assert cb.shape[0] == n * z
        d = _dev.to_dev(cb.astype(np.int64) if cb.dtype.kind == 'f' else cb, torch.int8).reshape(1, -1)
        ok = torch.empty((1,), dtype=torch.uint8, device=d.device)
        _native.check(_native.lib().nrldpc_parity_check_rows(_dev.handle(), self.baseGraphNo, z, _dev.ptr(d), 1,
                                                             1 if firstRowOnly else 0, _dev.ptr(ok), _dev.stream_ptr()))
        return bool(_dev.to_host(ok)[0])

    # ------------------------------------------------------------------------------------------------------------------
    def getRateMatchedCbLens(self, g, c):
        """E_r for each code block (ldpc.py:846-856)."""
        return np.int32(params.rate_matched_cb_lens(g, c, self.txLayers, self.qm))

    # ------------------------------------------------------------------------------------------------------------------
    def initialize(self, txBlockSize):
        """Derive C, Zc, iLS, K from B (ldpc.py:859-892)."""
        if self.txBlockSize == txBlockSize:
            return
        self.txBlockSize = txBlockSize
        self.numCodeBlocks, self.liftingSize, self.setIndex, self.codeBlockSize = \
            params.segmentation_params(self.baseGraphNo, txBlockSize)

    # ------------------------------------------------------------------------------------------------------------------
    def _tb_config(self, rv=0, g=0):
        z = self.liftingSize
        nz = (66 * z) if self.baseGraphNo == 1 else (50 * z)
        ncb = nz if self.nRef == 0 else min(nz, self.nRef)
        return _native.TbConfig(bg=self.baseGraphNo, zc=z, K=self.codeBlockSize, F=self.numFillerBits,
                                C=self.numCodeBlocks, qm=self.qm, nl=self.txLayers, ncb=ncb, rv=rv, reserved=0,
                                G=int(g)), nz, ncb


# **********************************************************************************************************************
class LdpcEncoder(LdpcBase):
    """LDPC encoder: segmentation, encoding, rate matching (neoradium/ldpc.py:896-1217)."""

    def __init__(self, baseGraphNo=1, modulation='QPSK', txLayers=1, nRef=0, targetRate=449 / 1024):
        super().__init__(baseGraphNo, modulation, txLayers, nRef)
        self.targetRate = targetRate

    def __repr__(self):
        return self.print(getStr=True)

    def print(self, indent=0, title=None, getStr=False):
        if title is None:
            title = "LDPC Encoder Properties:"
        repStr = super().print(indent, title, True)
        repStr += indent * ' ' + "  Target Rate:        %s\n" % (str(self.targetRate))
        if getStr:
            return repStr
        print(repStr)

    # ------------------------------------------------------------------------------------------------------------------
    def doSegmentation(self, txBlock, fillerBit=0):
        """Transport block (with its CRC) -> C x K code blocks (ldpc.py:981-1030).  ``fillerBit`` is ignored."""
        txBlock = np.asarray(txBlock)
        self.initialize(len(txBlock))
        c = self.numCodeBlocks
        bitsPerCodeBlock = int(np.ceil(self.txBlockSize / c))
        self.numFillerBits = self.codeBlockSize - bitsPerCodeBlock - (24 if c > 1 else 0)
        cfg, _, _ = self._tb_config()
        d = _dev.to_dev(txBlock, torch.int8).reshape(1, -1)
        # dtype: int8 when C > 1 (ldpc.py:1020), otherwise the input's dtype promoted with int8 (ldpc.py:1016)
        rdt = np.dtype(np.int8) if c > 1 else np.result_type(txBlock.dtype, np.int8)
        managed = rdt == np.int8 and _dev.managed_ok(c * self.codeBlockSize)      # stays on the device for encode() (see _managed.py)
        out = (_dev.managed_out((c, self.codeBlockSize), np.int8) if managed
               else torch.empty((c, self.codeBlockSize), dtype=torch.int8, device=d.device))
        _native.check(_native.lib().nrldpc_segment(_dev.handle(), cfg, _dev.ptr(d), 1, len(txBlock), len(txBlock),
                                                   _dev.ptr(out), _dev.stream_ptr()))
        if managed:
            _dev.sync()
            return out
        return _dev.to_host(out).astype(rdt)

    # ------------------------------------------------------------------------------------------------------------------
    def encode(self, codeBlocks, puncture=True):
        """C x K code blocks -> C x N encoded blocks (ldpc.py:1033-1090)."""
        codeBlocks = np.asanyarray(codeBlocks)      # (keeps a ManagedArray: its pages are read on the device in place)
        z = self.liftingSize
        P, n2, k = params.bg_dims(self.baseGraphNo)
        c, kk = codeBlocks.shape
        assert kk == k * z
        d = _dev.dev_in(codeBlocks, torch.int8, np.int8)
        cols = n2 - 2 if puncture else n2
        rdt = np.result_type(codeBlocks.dtype, np.int8)
        # the encoded blocks are HarqCW.encBuffer (harq.py:160): a ManagedArray keeps them on the device for the
        # rateMatch calls of every (re)transmission
        managed = rdt == np.int8 and _dev.managed_ok(c * cols * z)
        out = (_dev.managed_out((c, cols * z), np.int8) if managed
               else torch.empty((c, cols * z), dtype=torch.int8, device=d.device))
        _native.check(_native.lib().nrldpc_encode(_dev.handle(), self.baseGraphNo, z, _dev.ptr(d), c, _dev.ptr(out),
                                                  1 if puncture else 0, _dev.stream_ptr()))
        if managed:
            _dev.sync()
            return out
        return _dev.to_host(out).astype(rdt)

    # ------------------------------------------------------------------------------------------------------------------
    def rateMatch(self, codedBlocks, g=None, concatCBs=True, rv=0):
        """C x N encoded blocks -> rate-matched, interleaved bits (ldpc.py:1093-1159)."""
        codedBlocks = np.asanyarray(codedBlocks)      # (keeps a ManagedArray: its pages are read on the device in place)
        c, nz = codedBlocks.shape
        z = self.liftingSize
        assert nz in [66 * z, 50 * z]
        if rv not in [0, 1, 2, 3]:
            raise ValueError("Invalid 'rv' value! It must be one of 0, 1, 2, or 3.")
        if g is None:
            g = int(np.ceil((self.txBlockSize - 24) / self.targetRate))
        cfg, _, _ = self._tb_config(rv=rv, g=g)
        cfg.C = c
        lens = params.rate_matched_cb_lens(g, c, self.txLayers, self.qm)
        total = int(sum(lens))
        d = _dev.dev_in(codedBlocks, torch.int8, np.int8)
        out = torch.empty((total,), dtype=torch.int8, device=d.device)
        _native.check(_native.lib().nrldpc_rate_match(_dev.handle(), cfg, _dev.ptr(d), 1, _dev.ptr(out), total,
                                                      _dev.stream_ptr()))
        res = _dev.to_host(out).astype(codedBlocks.dtype)
        if concatCBs:
            return res
        offs = np.cumsum([0] + lens)
        return [res[offs[r]:offs[r + 1]] for r in range(c)]

    # ------------------------------------------------------------------------------------------------------------------
    @deprecated(replacement="getRateMatchedCodeBlocks")
    def getRateMatchedCodeWords(self, txBlock, g=None, concatCBs=True, addCrc=True):
        return self.getRateMatchedCodeBlocks(txBlock, g, concatCBs, addCrc)

    def getRateMatchedCodeBlocks(self, txBlock, g=None, concatCBs=True, addCrc=True):
        """CRC attach -> segmentation -> encoding -> rate matching in one call (ldpc.py:1167-1204).  The intermediate
        arrays stay on the device."""
        txBlock = np.asarray(txBlock)
        L = _native.lib()
        h, s = _dev.handle(), _dev.stream_ptr()
        d = _dev.to_dev(txBlock, torch.int8).reshape(1, -1)
        if addCrc:
            tb = torch.empty((1, d.shape[1] + 24), dtype=torch.int8, device=d.device)
            _native.check(L.nrldpc_crc_attach(h, _dev.ptr(d), 1, d.shape[1], d.shape[1], _native.CRC_IDS['24A'],
                                              _dev.ptr(tb), s))
        else:
            tb = d
        B = tb.shape[1]
        self.initialize(B)
        c, z = self.numCodeBlocks, self.liftingSize
        self.numFillerBits = self.codeBlockSize - int(np.ceil(B / c)) - (24 if c > 1 else 0)
        if g is None:
            g = int(np.ceil((self.txBlockSize - 24) / self.targetRate))
        cfg, nz, _ = self._tb_config(rv=0, g=g)
        cbs = torch.empty((c, self.codeBlockSize), dtype=torch.int8, device=d.device)
        _native.check(L.nrldpc_segment(h, cfg, _dev.ptr(tb), 1, B, B, _dev.ptr(cbs), s))
        coded = torch.empty((c, nz), dtype=torch.int8, device=d.device)
        _native.check(L.nrldpc_encode(h, self.baseGraphNo, z, _dev.ptr(cbs), c, _dev.ptr(coded), 1, s))
        lens = params.rate_matched_cb_lens(g, c, self.txLayers, self.qm)
        total = int(sum(lens))
        out = torch.empty((total,), dtype=torch.int8, device=d.device)
        _native.check(L.nrldpc_rate_match(h, cfg, _dev.ptr(coded), 1, _dev.ptr(out), total, s))
        # dtype as the reference chain produces it: int8 when C > 1, else the (CRC-appended) input dtype
        if c > 1:
            dt = np.int8
        else:
            dt = np.result_type(txBlock.dtype, np.int64) if addCrc else np.result_type(txBlock.dtype, np.int8)
        res = _dev.to_host(out).astype(dt)
        if concatCBs:
            return res
        offs = np.cumsum([0] + lens)
        return [res[offs[r]:offs[r + 1]] for r in range(c)]

    # ------------------------------------------------------------------------------------------------------------------
    def getDecoder(self, **kwargs):
        """An ``LdpcDecoder`` configured like this encoder (ldpc.py:1207-1217)."""
        return LdpcDecoder(self.baseGraphNo, self.modulation, self.txLayers, self.nRef, **kwargs)


# **********************************************************************************************************************
class _PendingLLRs:
    """Result of ``LdpcDecoder.decodeLLRsAsync``; ``result()`` blocks until the outputs are on the host."""

    def __init__(self, pend, unpack):
        self._pend, self._unpack = pend, unpack

    def done(self):
        return self._pend.done()

    def result(self):
        return self._unpack(self._pend.result())


class LdpcDecoder(LdpcBase):
    """LDPC decoder: rate recovery, layered min-sum decoding, CRC check and merge (neoradium/ldpc.py:1220-1619)."""

    def __init__(self, baseGraphNo=1, modulation='QPSK', txLayers=1, nRef=0, precision='fp64', earlyStop=False,
                 earlyStopFrom=1):
        super().__init__(baseGraphNo, modulation, txLayers, nRef)
        if precision not in _TORCH_F:
            raise ValueError("'precision' must be 'fp64' or 'fp32'!")
        self.precision = precision
        self.earlyStop = earlyStop
        self.earlyStopFrom = earlyStopFrom   # with earlyStop: first iteration after which the syndrome is tested
        self.lastIterations = None      # per-code-block iteration counts of the last decode (extension)
        self.rowStarts = None           # reference attributes of the undocumented decode2 (ldpc.py:1287-1290)
        self.rowNZcounts = None
        self.colIndexes = None

    def __repr__(self):
        return self.print(getStr=True)

    def print(self, indent=0, title=None, getStr=False):
        if title is None:
            title = "LDPC Decoder Properties:"
        repStr = super().print(indent, title, True)
        if getStr:
            return repStr
        print(repStr)

    # ------------------------------------------------------------------------------------------------------------------
    def _rx_setup(self, txBlockSize):
        self.initialize(txBlockSize + 24)                                           # ldpc.py:1365
        bitsPerCodeBlock = int(np.ceil(self.txBlockSize / self.numCodeBlocks))
        if self.numCodeBlocks > 1:
            bitsPerCodeBlock += 24
        self.numFillerBits = self.codeBlockSize - bitsPerCodeBlock                  # ldpc.py:1368

    def recoverRate(self, rxBlock, txBlockSize, harq=None):
        """LLRs [G] -> float64 [C, N] ready for ``decode`` (ldpc.py:1330-1418).  With ``harq`` the LLRs are combined
        into ``harq.decBuffer`` (float64 [C, Ncb-F], updated in place) using ``harq.rv``."""
        self._rx_setup(txBlockSize)
        c = self.numCodeBlocks
        rxBlock = np.asarray(rxBlock)
        rv = 0 if harq is None else harq.rv
        cfg, nz, ncb = self._tb_config(rv=rv, g=len(rxBlock))
        cirBufSize = ncb - self.numFillerBits
        hostBuf = None if harq is None else harq.decBuffer
        if hostBuf is not None:
            assert hostBuf.shape == (c, cirBufSize), \
                f"HARQ buffer shape mismatch! It must be a {c}x{cirBufSize} NumPy array!"
        x = _dev.to_dev(rxBlock, torch.float64).reshape(-1)
        need_buf = harq is not None
        managed = _dev.managed_ok(c * ncb * 8)
        # HARQ soft buffer (HarqCW.decBuffer, harq.py:121): a ManagedArray created here stays on the device between the
        # transmissions of a transport block; a caller-supplied plain array is combined and updated IN PLACE like the
        # reference does (ldpc.py:1410), which costs one upload and one download
        mbuf = None
        if need_buf:
            if hostBuf is None and managed:
                mbuf = buf = _dev.managed_out((c, cirBufSize), np.float64, zero=True)
            elif hostBuf is not None and managed and _dev.dev_in(hostBuf, torch.float64, np.float64) is hostBuf:
                mbuf = buf = hostBuf
            else:
                buf = (torch.zeros((c, cirBufSize), dtype=torch.float64, device=x.device) if hostBuf is None
                       else _dev.to_dev(hostBuf, torch.float64))
        else:
            buf = None
        out = (_dev.managed_out((c, ncb), np.float64) if managed
               else torch.empty((c, ncb), dtype=torch.float64, device=x.device))
        _native.check(_native.lib().nrldpc_rate_recover(_dev.handle(), cfg, _native.F64, _dev.ptr(x), 1, x.numel(),
                                                        x.numel(), _dev.ptr(buf), _dev.ptr(out), _dev.stream_ptr()))
        if need_buf:
            if mbuf is not None:
                harq.decBuffer = mbuf
            else:
                newBuf = _dev.to_host(buf)
                if hostBuf is not None and isinstance(hostBuf, np.ndarray) and hostBuf.dtype == np.float64:
                    hostBuf[...] = newBuf           # the reference mutates the caller's array in place (ldpc.py:1410)
                    harq.decBuffer = hostBuf
                else:
                    harq.decBuffer = newBuf
        if managed:
            _dev.sync()
            return out
        return _dev.to_host(out)

    # ------------------------------------------------------------------------------------------------------------------
    def decode(self, rxCodeBlock, numIter=5, onlyInfoBits=True, outputBelief=False):
        """Layered normalised min-sum decoding of C x N LLRs (ldpc.py:1495-1581)."""
        rxCodeBlock = np.asanyarray(rxCodeBlock)      # (keeps a ManagedArray: its pages are read on the device in place)
        c, nIn = rxCodeBlock.shape
        z = self.liftingSize
        P, n, k = params.bg_dims(self.baseGraphNo)
        assert nIn % z == 0 and nIn // z + 2 == n
        in64 = rxCodeBlock.dtype != np.float32
        if rxCodeBlock.dtype in (np.float32, np.float64):     # (a ManagedArray from recoverRate is read in place)
            x = _dev.dev_in(rxCodeBlock, torch.float64 if in64 else torch.float32, rxCodeBlock.dtype)
        else:
            x = _dev.to_dev(rxCodeBlock.astype(np.float64))
        outCols = k if onlyInfoBits else n
        tdt = _TORCH_F[self.precision]
        managed = _dev.managed_ok(c * outCols * z * (8 if outputBelief else 1))
        bits = beliefs = None
        if outputBelief:
            beliefs = (_dev.managed_out((c, outCols * z), np.float64) if managed and self.precision == 'fp64'
                       else torch.empty((c, outCols * z), dtype=tdt, device=x.device))
        else:
            bits = (_dev.managed_out((c, outCols * z), np.int8) if managed
                    else torch.empty((c, outCols * z), dtype=torch.int8, device=x.device))
        iters = torch.empty((c,), dtype=torch.int32, device=x.device)
        flags = _native.dec_flags(self.earlyStop, self.earlyStopFrom)
        _native.check(_native.lib().nrldpc_decode(
            _dev.handle(), self.baseGraphNo, z, _native.F64 if in64 else _native.F32, _NATIVE_F[self.precision],
            _dev.ptr(x), c, nIn, n - 2, int(numIter), flags, outCols, _dev.ptr(bits), _dev.ptr(beliefs),
            _dev.ptr(iters), _dev.stream_ptr()))
        self.lastIterations = _dev.to_host(iters)      # (synchronises the stream: managed results are complete)
        res = beliefs if outputBelief else bits
        if isinstance(res, np.ndarray):
            return res
        return _dev.to_host(res).astype(np.float64) if outputBelief else _dev.to_host(res)

    # ------------------------------------------------------------------------------------------------------------------
    def decode2(self, rxCodeBlock, maxIter=6, onlyInfoBits=True, outputBelief=False, alpha=0.75, stopOnGoodParity=True,
                firstRowOnly=False, beta=0.0):
        """The reference's undocumented verification decoder (ldpc.py:1421-1492): the same layered schedule walked one
        lifted row at a time, with the true second minimum (no "+100000" term) and a caller-chosen ``alpha``.
        ``stopOnGoodParity`` stops a block after the first iteration whose hard decisions satisfy EVERY parity check; the
        reference's own test looks at the first base-graph row only (``isValidCodedBlock``, ldpc.py:841-843), so with
        ``stopOnGoodParity=True`` it may stop earlier than this one; ``firstRowOnly=True`` (compatibility switch) applies
        the reference's first-row-only stop test instead.  ``beta`` > 0 (extension) turns the rule into offset min-sum,
        |message| = max(alpha * min - beta, 0).  ``lastIterations`` holds the per-block counts."""
        rxCodeBlock = np.asanyarray(rxCodeBlock)      # (keeps a ManagedArray: its pages are read on the device in place)
        c, nIn = rxCodeBlock.shape
        z = self.liftingSize
        P, n, k = params.bg_dims(self.baseGraphNo)
        assert nIn % z == 0 and nIn // z + 2 == n
        x = _dev.to_dev(rxCodeBlock if rxCodeBlock.dtype in (np.float32, np.float64) else rxCodeBlock.astype(np.float64))
        outCols = k if onlyInfoBits else n
        tdt = _TORCH_F[self.precision]
        bits = beliefs = None
        if outputBelief:
            beliefs = torch.empty((c, outCols * z), dtype=tdt, device=x.device)
        else:
            bits = torch.empty((c, outCols * z), dtype=torch.int8, device=x.device)
        iters = torch.empty((c,), dtype=torch.int32, device=x.device)
        _native.check(_native.lib().nrldpc_decode2_offset(
            _dev.handle(), self.baseGraphNo, z, _native.F64 if x.dtype == torch.float64 else _native.F32,
            _NATIVE_F[self.precision], _dev.ptr(x), c, nIn, nIn // z, int(maxIter), float(alpha), float(beta),
            (2 if firstRowOnly else 1) if stopOnGoodParity else 0, outCols, _dev.ptr(bits), _dev.ptr(beliefs), _dev.ptr(iters),
            _dev.stream_ptr()))
        self.lastIterations = _dev.to_host(iters)
        out = _dev.to_host(beliefs if outputBelief else bits)
        if onlyInfoBits:
            out = out[:, :self.codeBlockSize]
        return out.astype(np.float64) if outputBelief else out

    # ------------------------------------------------------------------------------------------------------------------
    def checkCrcAndMerge(self, rxCodedBlocks):
        """CRC check of every decoded code block and re-assembly of the transport block (ldpc.py:1584-1619)."""
        rxCodedBlocks = np.asanyarray(rxCodedBlocks)      # (keeps a ManagedArray: its pages are read on the device in place)
        c = self.numCodeBlocks
        cfg, _, _ = self._tb_config()
        d = _dev.dev_in(rxCodedBlocks, torch.int8, np.int8)
        assert tuple(d.shape) == (c, self.codeBlockSize)
        per = self.codeBlockSize - self.numFillerBits - (24 if c > 1 else 0)
        tb = torch.empty((c * per,), dtype=torch.int8, device=d.device)
        ok = torch.empty((c,), dtype=torch.uint8, device=d.device)
        _native.check(_native.lib().nrldpc_check_crc_and_merge(_dev.handle(), cfg, _dev.ptr(d), 1, _dev.ptr(tb),
                                                               c * per, _dev.ptr(ok), _dev.stream_ptr()))
        tbBits = _dev.to_host(tb).astype(rxCodedBlocks.dtype)
        okHost = _dev.to_host(ok).astype(np.bool_)
        if c == 1:
            return tbBits, [okHost[0]]          # the reference returns a list here (ldpc.py:1616)
        return tbBits, okHost

    # ------------------------------------------------------------------------------------------------------------------
    def decodeLLRsAsync(self, llrs, txBlockSize, numIter=5, out=None, slot=0):
        """Non-blocking ``decodeLLRs`` for a [numTb, G] HOST batch (extension): queues the H2D copies, the fused decode
        and the D2H copies and returns a pending result (``.result()`` -> (txBlocks, cbCrc, tbCrc)).  Keep at most two
        calls in flight, alternating ``slot`` 0/1 and the (pinned) host input/output buffers: the next batch's H2D
        then overlaps the current batch's decode and D2H."""
        isT = isinstance(llrs, torch.Tensor)
        if (llrs.dim() if isT else np.ndim(llrs)) != 2 or (isT and llrs.device.type != 'cpu'):
            raise ValueError("decodeLLRsAsync takes a [numTb, G] host batch")
        return self.decodeLLRs(llrs, txBlockSize, numIter, out=out, _async=int(slot))

    def decodeSymbolsAsync(self, symbols, noiseVar, txBlockSize, numIter=5, out=None, slot=0):
        """``decodeLLRsAsync`` one step further upstream (extension): ``symbols`` is a complex64 [numTb, G/qm] HOST batch of
        equalised symbols (what PDSCH.getLLRsFromGrid hands to Modem.getLLRsFromSymbols, pdsch.py:935-1000), ``noiseVar``
        the demapper's noise variance.  The symbols cross PCIe (2 bytes per coded bit at 16QAM, 1 at 256QAM, instead of 4
        for fp32 LLRs); max-log demapping (modulation.py:159-204) and the fused rate-recovery / decode / CRC chain run on
        the device.  LLRs are computed in float64 from the float32 symbol coordinates and rounded to float32 (a few ulp from
        the reference's float64 demapper, tests/test_gpu_link.py); decoding is the fp32 chain.  Same result tuple and
        in-flight rules as ``decodeLLRsAsync``."""
        isT = isinstance(symbols, torch.Tensor)
        x = symbols if isT else torch.from_numpy(np.ascontiguousarray(symbols))
        if x.dim() != 2 or x.device.type != 'cpu' or x.dtype != torch.complex64:
            raise ValueError("decodeSymbolsAsync takes a complex64 [numTb, G/qm] host batch")
        return self.decodeLLRs(x, txBlockSize, numIter, out=out, _async=int(slot), _noiseVar=float(noiseVar))

    def decodeSymbols(self, symbols, noiseVar, txBlockSize, numIter=5, out=None):
        """Blocking form of ``decodeSymbolsAsync``."""
        return self.decodeSymbolsAsync(symbols, noiseVar, txBlockSize, numIter, out=out).result()

    def decodeLLRs(self, llrs, txBlockSize, numIter=5, harq=None, precision=None, returnDevice=False, out=None,
                   _async=None, _noiseVar=None):
        """Fused RX chain (extension; the operation sequence of HarqCW.decodeLLRs, harq.py:165-173):
        recoverRate -> decode -> checkCrcAndMerge -> checkCrc('24A') in one kernel pass.

        ``llrs``: [G] for one transport block or [numTb, G] for a batch of equally configured blocks (float32, float64
        or float16 -- half values are widened exactly on the device and halve the PCIe traffic; NumPy array or a CUDA
        torch tensor).  Returns (txBlocks, cbCrc, tbCrc): decoded transport block(s)
        WITHOUT the 24 CRC bits ([A] or [numTb, A]), per-code-block CRC results ([C] / [numTb, C]) and the
        transport-block CRC24A result(s).  ``harq`` (single block only) supplies ``rv`` and the soft buffer exactly
        as in ``recoverRate``.  A [numTb, G] HOST batch takes the pipelined path (``TbBatchCodec.decode_host``: chunked
        H2D / decode / D2H overlap); ``out`` may then carry preallocated (ideally pinned) host arrays
        {tb [numTb, C*per], cbOk [numTb, C], tbOk [numTb], iters [numTb, C]} that receive the results."""
        if _async is not None:
            assert harq is None and not returnDevice
        self._rx_setup(txBlockSize)
        c, z = self.numCodeBlocks, self.liftingSize
        precision = precision or self.precision
        isT = isinstance(llrs, torch.Tensor)
        single = (llrs.dim() if isT else np.ndim(llrs)) == 1
        onHost = (not isT) or llrs.device.type == 'cpu'
        hdt = llrs.dtype if isT else np.asarray(llrs).dtype
        if (not single) and onHost and harq is None and not returnDevice and \
                hdt in (torch.float32, torch.float64, torch.float16, torch.complex64, np.float32, np.float64, np.float16):
            from .batch import TbBatchCodec
            x = llrs if isT else np.ascontiguousarray(llrs)
            gBits = x.shape[1] * (self.qm if _noiseVar is not None else 1)
            key = (txBlockSize, gBits, precision, self.earlyStop, self.earlyStopFrom)
            codec = getattr(self, '_hostCodec', None)
            if codec is None or self._hostCodecKey != key:
                # (a private library handle: the pipeline issues work on its own streams, include/nrldpc.h allows one handle
                # per (device, stream); the shared handle stays with the calls on the current stream)
                codec = TbBatchCodec(self.baseGraphNo, self.modulation, txBlockSize, gBits, self.txLayers, self.nRef,
                                     0, precision, self.earlyStop, ownHandle=True, earlyStopFrom=self.earlyStopFrom)
                self._hostCodec, self._hostCodecKey = codec, key
            def unpack(res):
                self.lastIterations = res['iters'].numpy().reshape(-1)
                return (res['tb'].numpy()[:, :txBlockSize], res['cbOk'].numpy().view(np.bool_),
                        res['tbOk'].numpy().view(np.bool_))
            if _async is not None:
                pend = codec.decode_host(x, numIter, out=out, wait=False, slot=_async, noiseVar=_noiseVar)
                return _PendingLLRs(pend, unpack)
            return unpack(codec.decode_host(x, numIter, out=out, noiseVar=_noiseVar))
        x = _dev.to_dev(llrs)
        if x.dtype not in (torch.float32, torch.float64, torch.float16):
            x = x.to(torch.float64)
        x = x.reshape(1, -1) if single else x
        numTb, G = x.shape
        rv = 0 if harq is None else harq.rv
        cfg, nz, ncb = self._tb_config(rv=rv, g=G)
        tdt = _TORCH_F[precision]
        buf = None
        if harq is not None:
            assert single, "HARQ combining is per transport block"
            cirBufSize = ncb - self.numFillerBits
            hostBuf = harq.decBuffer
            if hostBuf is not None:
                assert hostBuf.shape == (c, cirBufSize), \
                    f"HARQ buffer shape mismatch! It must be a {c}x{cirBufSize} NumPy array!"
                buf = _dev.to_dev(hostBuf, tdt)
            else:
                buf = torch.zeros((c, cirBufSize), dtype=tdt, device=x.device)
        per = self.codeBlockSize - self.numFillerBits - (24 if c > 1 else 0)
        tb = torch.empty((numTb, c * per), dtype=torch.int8, device=x.device)
        cbOk = torch.empty((numTb, c), dtype=torch.uint8, device=x.device)
        tbOk = torch.empty((numTb,), dtype=torch.uint8, device=x.device)
        iters = torch.empty((numTb, c), dtype=torch.int32, device=x.device)
        flags = _native.dec_flags(self.earlyStop, self.earlyStopFrom)
        _native.check(_native.lib().nrldpc_decode_tb(
            _dev.handle(), cfg, {torch.float64: _native.F64, torch.float16: _native.F16}.get(x.dtype, _native.F32),
            _NATIVE_F[precision], _dev.ptr(x), numTb, G, x.stride(0), _dev.ptr(buf), int(numIter), flags, _dev.ptr(tb), c * per,
            _dev.ptr(cbOk), _dev.ptr(tbOk), _dev.ptr(iters), _dev.stream_ptr()))
        if harq is not None:
            newBuf = _dev.to_host(buf).astype(np.float64)
            if isinstance(harq.decBuffer, np.ndarray) and harq.decBuffer.dtype == np.float64:
                harq.decBuffer[...] = newBuf
            else:
                harq.decBuffer = newBuf
        if returnDevice:
            self.lastIterations = iters
            return tb[:, :txBlockSize], cbOk, tbOk
        self.lastIterations = _dev.to_host(iters).reshape(-1)
        tbH = _dev.to_host(tb)[:, :txBlockSize]
        cbH = _dev.to_host(cbOk).astype(np.bool_)
        tbOkH = _dev.to_host(tbOk).astype(np.bool_)
        if single:
            return tbH[0], cbH[0], tbOkH[0]
        return tbH, cbH, tbOkH
