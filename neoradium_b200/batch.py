"""Device-resident batched front end of libnrldpc for throughput work (BLER sweeps, benchmarks, multi-GPU sharding).

``TbBatchCodec`` handles a batch of EQUALLY configured transport blocks (same base graph, A, G, modulation, layers) --
configs 2, 4 and 5 of BASELINE.json; mixed batches (config 3) are a Python-level loop over one codec per group.
Inputs and outputs are CUDA torch tensors (PyTorch only carries the buffers); every operation is one C-ABI call of
include/nrldpc.h.  The operation sequence is the reference's: TX = appendCrc('24A') -> doSegmentation -> encode ->
rateMatch (ldpc.py:1200-1204); RX = recoverRate -> decode -> checkCrcAndMerge -> checkCrc('24A') (ldpc.py:1234-1251).
"""
import math
import os

import torch

from . import _dev, _native, params


_IN_DTYPE = {torch.float32: _native.F32, torch.float64: _native.F64, torch.float16: _native.F16}


class PendingDecode:
    """Result of ``TbBatchCodec.decode_host(..., wait=False)``: ``result()`` returns the host output dict once the last
    D2H copy has landed."""

    def __init__(self, hout, event):
        self._hout, self._event = hout, event

    def done(self):
        return self._event.query()

    def result(self):
        self._event.synchronize()
        return self._hout


class TbBatchCodec:
    def __init__(self, baseGraphNo, modulation, txBlockSize, g, txLayers=1, nRef=0, rv=0, precision='fp32',
                 earlyStop=False, device=None, ownHandle=False, earlyStopFrom=1):
        """``ownHandle=True`` gives the codec a private library handle (scratch, temporaries): required when several
        codecs issue work concurrently on DIFFERENT streams (include/nrldpc.h: one handle per (device, stream))."""
        if baseGraphNo not in (1, 2):
            raise ValueError("'baseGraphNo' must be 1 or 2!")
        if modulation not in params.MOD_ORDER:
            raise ValueError("Invalid 'modulation' value!")
        if rv not in (0, 1, 2, 3):
            raise ValueError("Invalid 'rv' value! It must be one of 0, 1, 2, or 3.")
        self.bg, self.qm, self.nl, self.nRef, self.rv = baseGraphNo, params.MOD_ORDER[modulation], txLayers, nRef, rv
        self.A, self.G = int(txBlockSize), int(g)
        self.B = self.A + 24
        self.C, self.Zc, self.iLS, self.K = params.segmentation_params(self.bg, self.B)
        per = int(math.ceil(self.B / self.C))
        self.F = self.K - per - (24 if self.C > 1 else 0)
        self.per = self.K - self.F - (24 if self.C > 1 else 0)         # payload bits per code block in the merged TB
        self.N = (66 if self.bg == 1 else 50) * self.Zc
        self.ncb = self.N if nRef == 0 else min(self.N, nRef)
        self.lens = params.rate_matched_cb_lens(self.G, self.C, self.nl, self.qm)
        self.sumE = int(sum(self.lens))
        self.precision = precision
        self.earlyStop = earlyStop
        self.earlyStopFrom = earlyStopFrom   # first iteration after which the syndrome is tested (extension, default: every one; 'auto': follows the previous launch of this handle -- use ownHandle=True)
        self.device = device if device is not None else _dev.device()
        self.cfg = _native.TbConfig(bg=self.bg, zc=self.Zc, K=self.K, F=self.F, C=self.C, qm=self.qm, nl=self.nl,
                                    ncb=self.ncb, rv=self.rv, reserved=0, G=self.G)
        devIdx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._ownHandle = bool(ownHandle)
        self._h = _native.new_handle(devIdx) if ownHandle else _native.handle(devIdx)

    def __del__(self):
        # a private handle dies with its codec (after the device has drained: its scratch may still be in use)
        try:
            if getattr(self, '_ownHandle', False) and self._h is not None:
                torch.cuda.synchronize(self.device)
                _native.lib().nrldpc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------------------------
    def encode(self, payload):
        """payload int8 [numTb, A] (device) -> rate-matched bits int8 [numTb, sumE]."""
        L = _native.lib()
        s = _dev.stream_ptr()
        numTb = payload.shape[0]
        assert payload.shape[1] == self.A and payload.dtype == torch.int8 and payload.is_contiguous()
        dev = payload.device
        tb = torch.empty((numTb, self.B), dtype=torch.int8, device=dev)
        _native.check(L.nrldpc_crc_attach(self._h, _dev.ptr(payload), numTb, self.A, self.A, _native.CRC_IDS['24A'],
                                          _dev.ptr(tb), s))
        cbs = torch.empty((numTb * self.C, self.K), dtype=torch.int8, device=dev)
        _native.check(L.nrldpc_segment(self._h, self.cfg, _dev.ptr(tb), numTb, self.B, self.B, _dev.ptr(cbs), s))
        coded = torch.empty((numTb * self.C, self.N), dtype=torch.int8, device=dev)
        _native.check(L.nrldpc_encode(self._h, self.bg, self.Zc, _dev.ptr(cbs), numTb * self.C, _dev.ptr(coded), 1, s))
        out = torch.empty((numTb, self.sumE), dtype=torch.int8, device=dev)
        _native.check(L.nrldpc_rate_match(self._h, self.cfg, _dev.ptr(coded), numTb, _dev.ptr(out), self.sumE, s))
        return out

    # ------------------------------------------------------------------------------------------------------------------
    def random_payload(self, numTb, seed, firstTb=0, out=None):
        """int8 [numTb, A] payload bits generated on the device (Philox4x32-10): transport block t of the call is global
        block firstTb + t of the stream under `seed`, independent of batch size and of the number of GPUs."""
        if out is None:
            out = torch.empty((numTb, self.A), dtype=torch.int8, device=self.device)
        assert out.is_contiguous() and out.shape == (numTb, self.A)
        _native.check(_native.lib().nrldpc_random_bits(self._h, int(seed) & (2 ** 64 - 1), int(firstTb) * self.A, _dev.ptr(out),
                                                       numTb * self.A, _dev.stream_ptr()))
        return out

    # ------------------------------------------------------------------------------------------------------------------
    def alloc_outputs(self, numTb):
        dev = self.device
        return dict(tb=torch.empty((numTb, self.C * self.per), dtype=torch.int8, device=dev),
                    cbOk=torch.empty((numTb, self.C), dtype=torch.uint8, device=dev),
                    tbOk=torch.empty((numTb,), dtype=torch.uint8, device=dev),
                    iters=torch.empty((numTb, self.C), dtype=torch.int32, device=dev))

    def decode(self, llr, numIter, out=None, softBuffer=None):
        """llr float32|float64|float16 [numTb, >=G'] (device, row pitch = stride(0)) -> dict(tb, cbOk, tbOk, iters).
        float16 LLRs are widened exactly to the compute type on load (half the HBM / PCIe bytes of float32).
        One fused kernel pass (+ a tiny per-TB CRC combine) on the current stream; nothing is synchronised."""
        numTb = llr.shape[0]
        if out is None:
            out = self.alloc_outputs(numTb)
        flags = _native.dec_flags(self.earlyStop, self.earlyStopFrom)
        _native.check(_native.lib().nrldpc_decode_tb(
            self._h, self.cfg, _IN_DTYPE[llr.dtype],
            _native.F64 if self.precision == 'fp64' else _native.F32, _dev.ptr(llr), numTb, llr.shape[1],
            llr.stride(0), _dev.ptr(softBuffer), int(numIter), flags, _dev.ptr(out['tb']), self.C * self.per,
            _dev.ptr(out['cbOk']), _dev.ptr(out['tbOk']), _dev.ptr(out['iters']), _dev.stream_ptr()))
        return out

    def decode_symbols(self, symbols, noiseVar, numIter, out=None):
        """symbols: complex64 [numTb, >= G'/qm] DEVICE tensor of equalised symbols (row pitch = stride(0)) -> the dict of
        ``decode``.  Max-log demapping (Modem.getLLRsFromSymbols, modulation.py:159-204) runs inside the decoder's load phase
        where the library has a fused form (one code block per CTA, no repetition) and as a separate pass into a scratch
        buffer otherwise; the LLRs -- hence every output -- equal ``decode(demap(symbols))`` in fp32 either way."""
        assert symbols.dtype == torch.complex64 and symbols.dim() == 2 and symbols.stride(1) == 1
        assert self.precision == 'fp32', "symbol input runs the fp32 chain"
        numTb = symbols.shape[0]
        if out is None:
            out = self.alloc_outputs(numTb)
        flags = _native.dec_flags(self.earlyStop, self.earlyStopFrom)
        _native.check(_native.lib().nrldpc_decode_tb_symbols(
            self._h, self.cfg, _dev.ptr(symbols), numTb, symbols.shape[1], symbols.stride(0), float(noiseVar), int(numIter),
            flags, _dev.ptr(out['tb']), self.C * self.per, _dev.ptr(out['cbOk']), _dev.ptr(out['tbOk']),
            _dev.ptr(out['iters']), _dev.stream_ptr()))
        return out

    # ------------------------------------------------------------------------------------------------------------------
    def decode_host(self, llr_host, numIter, out=None, chunks=None, wait=True, slot=0, noiseVar=None):
        """Host-buffer entry point: llr_host is a float32|float64|float16 [numTb, G'] HOST array (NumPy array or CPU torch
        tensor; pinned memory gives full PCIe speed).  The batch is cut into `chunks` groups of transport blocks and
        pipelined over three CUDA streams -- H2D copy of chunk i+1, fused decode of chunk i and D2H copy of the results of
        chunk i-1 overlap -- and the call returns after everything has landed on the host.
        `out` (optional) is a dict of preallocated host arrays/tensors {tb, cbOk, tbOk, iters} to write into (NumPy
        `out=` convention; pinned buffers avoid a staging copy); otherwise fresh pinned tensors are allocated.
        Returns the dict of host torch tensors (use .numpy() for zero-copy views).

        ``wait=False`` returns a ``PendingDecode`` right after the work is queued; its ``result()`` blocks until the
        outputs are on the host.  Calls queue behind each other on the same three streams, so with two calls in flight
        (``slot`` 0 / 1 select independent device staging buffers; the caller alternates its host buffers likewise) the
        H2D copy of the next batch overlaps the decode and D2H of the current one and PCIe never idles.

        ``noiseVar`` (with a complex64 [numTb, G'/qm] host array): the input is the EQUALISED SYMBOLS of the codewords, not
        their LLRs.  The symbols cross PCIe (8 bytes per symbol = 2 bytes per coded bit at 16QAM, 1 at 256QAM, instead of 4
        for fp32 LLRs) and the max-log demapper (Modem.getLLRsFromSymbols, modulation.py:159-204 -> nrldpc_demap_maxlog,
        fp32 LLRs) runs on the device in front of the fused decode, chunk by chunk in the same pipeline."""
        x = llr_host if isinstance(llr_host, torch.Tensor) else torch.from_numpy(llr_host)
        symbols = x.dtype == torch.complex64
        assert x.device.type == 'cpu' and x.dim() == 2 and (x.dtype in _IN_DTYPE or symbols)
        if symbols:
            assert noiseVar is not None and noiseVar > 0, "symbol input needs the noise variance of the demapper"
            x = torch.view_as_real(x).reshape(x.shape[0], -1)             # [numTb, 2 * symbols] float32, no copy
        numTb, Gp = x.shape
        if chunks is None:
            chunks = int(os.environ.get("NRLDPC_HOST_CHUNKS", "2"))   # measured on the bench batch: 2 chunks 14.8, 3: 13.4, 4: 14.1, 6: 9.9 Gbit/s
        chunks = max(1, min(int(chunks), numTb))
        bounds = [(numTb * i) // chunks for i in range(chunks + 1)]
        key = (numTb, Gp, x.dtype, chunks, symbols)
        if not hasattr(self, '_streams'):
            self._streams = tuple(torch.cuda.Stream(self.device) for _ in range(3))
            self._pipe = {}
        st = self._pipe.get(slot)
        if st is not None and st['key'] != key and st.get('done') is not None:
            st['done'].synchronize()          # the old staging buffers are about to be freed: let their last use finish
        if st is None or st['key'] != key:
            st = dict(key=key, done=None, h2d=self._streams[0], comp=self._streams[1], d2h=self._streams[2],
                      din=[torch.empty((bounds[i + 1] - bounds[i], Gp), dtype=x.dtype, device=self.device)
                           for i in range(chunks)],
                      dllr=[torch.empty((bounds[i + 1] - bounds[i], (Gp // 2) * self.qm), dtype=torch.float32, device=self.device)
                            for i in range(chunks)] if (symbols and self.precision != 'fp32') else None,
                      dout=[self.alloc_outputs(bounds[i + 1] - bounds[i]) for i in range(chunks)])
            self._pipe[slot] = st
        if out is None:
            out = dict(tb=torch.empty((numTb, self.C * self.per), dtype=torch.int8).pin_memory(),
                       cbOk=torch.empty((numTb, self.C), dtype=torch.uint8).pin_memory(),
                       tbOk=torch.empty((numTb,), dtype=torch.uint8).pin_memory(),
                       iters=torch.empty((numTb, self.C), dtype=torch.int32).pin_memory())
        hout = {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(v)) for k, v in out.items()}
        cur = torch.cuda.current_stream(self.device)
        for s_ in (st['h2d'], st['comp'], st['d2h']):
            s_.wait_stream(cur)
        if st['done'] is not None:
            st['h2d'].wait_event(st['done'])  # this slot's previous call still owns din / dout until its last D2H copy
        for i in range(chunks):
            lo, hi = bounds[i], bounds[i + 1]
            with torch.cuda.stream(st['h2d']):
                st['din'][i].copy_(x[lo:hi], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record()
            with torch.cuda.stream(st['comp']):
                st['comp'].wait_event(ev_in)
                if symbols and self.precision == 'fp32':   # demapper fused into the decoder's load phase (nrldpc_decode_tb_symbols)
                    self.decode_symbols(torch.view_as_complex(st['din'][i].view(hi - lo, Gp // 2, 2)), noiseVar, numIter,
                                        out=st['dout'][i])
                elif symbols:
                    _native.check(_native.lib().nrldpc_demap_maxlog(
                        self._h, self.qm, _native.F32, _dev.ptr(st['din'][i]), (hi - lo) * (Gp // 2), float(noiseVar),
                        _native.F32, _dev.ptr(st['dllr'][i]), _dev.stream_ptr()))
                    self.decode(st['dllr'][i], numIter, out=st['dout'][i])
                else:
                    self.decode(st['din'][i], numIter, out=st['dout'][i])
                ev_done = torch.cuda.Event()
                ev_done.record()
            with torch.cuda.stream(st['d2h']):
                st['d2h'].wait_event(ev_done)
                for k in ('tb', 'cbOk', 'tbOk', 'iters'):
                    hout[k][lo:hi].copy_(st['dout'][i][k], non_blocking=True)
        with torch.cuda.stream(st['d2h']):
            done = torch.cuda.Event()
            done.record()
        st['done'] = done
        if not wait:
            return PendingDecode(hout, done)
        st['d2h'].synchronize()
        return hout

    # ------------------------------------------------------------------------------------------------------------------
    def accumulate(self, out, counters, refPayload=None):
        """counters int64[8] += {CBs, CB CRC fails, TBs, TB CRC fails, bit errors, sum iterations, 0, 0}."""
        numTb = out['tb'].shape[0]
        ref = refPayload
        if ref is not None:
            assert ref.shape[1] == self.A and ref.stride(1) == 1 and ref.dtype == torch.int8
        _native.check(_native.lib().nrldpc_accumulate_counters_ref(
            self._h, numTb, self.C, _dev.ptr(out['cbOk']), _dev.ptr(out['tbOk']), _dev.ptr(out['iters']),
            _dev.ptr(out['tb']) if ref is not None else None, self.C * self.per, _dev.ptr(ref),
            ref.stride(0) if ref is not None else 0, self.A, _dev.ptr(counters), _dev.stream_ptr()))
        return counters


def decode_groups(codecs, llrs, numIter, outs=None, softBuffers=None):
    """Mixed-configuration batch in ONE library call (BASELINE configs[2]: the codewords of a PDSCH slot differ in base
    graph, lifting size, layers, redundancy version; harq.py:331-347 loops over them on the host).  ``codecs[i]`` describes
    group i (a ``TbBatchCodec``: its equally configured transport blocks), ``llrs[i]`` is its [numTb_i, G_i] device tensor,
    ``softBuffers[i]`` its device-resident HARQ buffer or None.  The descriptor array goes to ``nrldpc_decode_tb_groups``,
    which runs the groups concurrently on the library's internal streams (forked from / joined into the current stream).
    All codecs must share precision and early-termination settings.  Returns the list of output dicts."""
    n = len(codecs)
    assert n == len(llrs) and n > 0
    c0 = codecs[0]
    assert all(c.precision == c0.precision and c.earlyStop == c0.earlyStop and c.earlyStopFrom == c0.earlyStopFrom
               for c in codecs), "one call, one precision / early-termination setting"
    if outs is None:
        outs = [c.alloc_outputs(x.shape[0]) for c, x in zip(codecs, llrs)]
    if softBuffers is None:
        softBuffers = [None] * n
    arr = (_native.TbGroup * n)()
    for i, (c, x, o, sb) in enumerate(zip(codecs, llrs, outs, softBuffers)):
        g = arr[i]
        g.cfg = c.cfg
        g.in_dtype = _IN_DTYPE[x.dtype]
        g.llr = x.data_ptr()
        g.num_tb, g.llr_len, g.llr_stride = x.shape[0], x.shape[1], x.stride(0)
        g.soft_buffer = sb.data_ptr() if sb is not None else None
        g.tb_bits, g.tb_bits_stride = o['tb'].data_ptr(), c.C * c.per
        g.cb_crc_ok, g.tb_crc_ok, g.iters = o['cbOk'].data_ptr(), o['tbOk'].data_ptr(), o['iters'].data_ptr()
    _native.check(_native.lib().nrldpc_decode_tb_groups(
        c0._h, arr, n, _native.F64 if c0.precision == 'fp64' else _native.F32, int(numIter),
        _native.dec_flags(c0.earlyStop, c0.earlyStopFrom), _dev.stream_ptr()))
    return outs


# ----------------------------------------------------------------------------------------------------------------------
# Synthetic channel for workload generation (NOT the hot path; plain torch ops, outside every timed region):
# Gray-mapped QAM of TS 38.211 5.1 (neoradium/modulation.py:60-74), complex AWGN, max-log LLR (modulation.py:190-204),
# positive LLR => bit 0.
# ----------------------------------------------------------------------------------------------------------------------
def qam_awgn_llr(bits, qm, snr_db, generator=None, dtype=torch.float32):
    """bits int8 [..., n*qm] -> max-log LLRs [..., n*qm] after unit-energy QAM + AWGN at Es/N0 = snr_db."""
    assert qm in (2, 4, 6, 8, 10), "BPSK is not needed by the workloads"
    half = qm // 2
    shp = bits.shape
    b = bits.reshape(-1, qm).to(torch.float64)
    scale = 1.0 / math.sqrt({2: 2, 4: 10, 6: 42, 8: 170, 10: 682}[qm])
    n0 = 10.0 ** (-snr_db / 10.0)

    def pam(bb):   # bb [..., half]: amplitude recursion of 38.211 (real part uses bits 0,2,4,.. imag 1,3,5,..)
        a = torch.ones(bb.shape[0], dtype=torch.float64, device=bits.device)
        for q in range(half - 1, 0, -1):
            a = (1 << (half - q)) - (1 - 2 * bb[:, q]) * a
        return (1 - 2 * bb[:, 0]) * a

    re = pam(b[:, 0::2]) * scale
    im = pam(b[:, 1::2]) * scale
    noise = torch.randn((2, re.shape[0]), dtype=torch.float64, device=bits.device, generator=generator) * math.sqrt(n0 / 2)
    yr, yi = re + noise[0], im + noise[1]
    # per-dimension PAM levels and their bit labels
    lv = torch.arange(1 << half, device=bits.device)
    lb = ((lv[:, None] >> torch.arange(half - 1, -1, -1, device=bits.device)[None, :]) & 1).to(torch.float64)
    levels = pam(lb) * scale                                              # [2^half]
    out = torch.empty((re.shape[0], qm), dtype=torch.float64, device=bits.device)
    for comp, y in ((0, yr), (1, yi)):
        d2 = (y[:, None] - levels[None, :]) ** 2                          # [n, 2^half]
        for q in range(half):
            m0 = torch.where(lb[None, :, q] == 0, d2, torch.inf).min(1).values
            m1 = torch.where(lb[None, :, q] == 1, d2, torch.inf).min(1).values
            out[:, 2 * q + comp] = (m1 - m0) / n0
    return out.reshape(shp).to(dtype)
