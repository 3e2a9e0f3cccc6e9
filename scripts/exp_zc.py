"""Decoder throughput per lifting size (edge-updates/s), mode A (rate-recovered LLRs, all N columns sent), fp32, 8 iterations.
ROWS=r: only the first r base-graph rows carry non-zero extension LLRs (a rate-matched block of rate ~ k/(k+r-2)); the
decoder then schedules r rows (exact row skipping) and the rate is counted in EXECUTED edge-updates.  ZCS=a,b,.. picks
the lifting sizes, OUT= the result file.  NRLDPC_DEC_OCC=n (library knob) sets the resident CTAs per SM."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neoradium_b200 import _native, _dev
L, h = _native.lib(), _dev.handle()
res = {}
for bg, n, k, edges in ((1, 68, 22, 316), (2, 52, 10, 197)):
    for zc in [int(v) for v in os.environ.get('ZCS', '384,352,320,256,240,208,192,176,128,96,64,32,16,8').split(',')]:
        numCb = max(2048, min(65536, (1 << 24) // (n * zc)))
        # WAVES=w: at least w full waves of CTAs (3 per SM at most, 384 // zc blocks each): the steady-state rate; without it the
        # small lifting sizes run 1-3 waves and the figure is mostly the partly filled last one
        waves = int(os.environ.get('WAVES', '0'))
        if waves: numCb = max(numCb, waves * 148 * 3 * max(1, 384 // zc))
        f64 = os.environ.get('DT', 'f32') == 'f64'   # DT=f64: the float64 instantiation (the drop-in default precision)
        x = torch.randn((numCb, (n - 2) * zc), device='cuda', dtype=torch.float64 if f64 else torch.float32) * 2 + 1.5
        rows = int(os.environ.get('ROWS', '0'))
        flags = 2
        if rows:
            x[:, (k + rows - 2) * zc:] = 0
            flags = 0
            P = 46 if bg == 1 else 42
            hb = np.empty((P, n), np.int16)
            _native.check(L.nrldpc_base_graph(bg, -1, zc, hb.ctypes.data))
            edges = int((hb[:rows] >= 0).sum())   # executed edges per lifted check set
        bits = torch.empty((numCb, k * zc), dtype=torch.int8, device='cuda')
        s = _dev.stream_ptr()
        def run():
            _native.check(L.nrldpc_decode(h, bg, zc, _native.F64 if f64 else _native.F32, _native.F64 if f64 else _native.F32, _dev.ptr(x), numCb, (n - 2) * zc, n - 2, 8, flags, k,
                                          _dev.ptr(bits), None, None, s))
        run(); run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        geups = numCb * edges * zc * 8 / ms / 1e6
        print("BG%d Zc=%3d  %6d blocks  %8.3f ms  %7.1f G edge-updates/s  %6.2f Mcb/s" % (bg, zc, numCb, ms, geups, numCb / ms / 1e3), flush=True)
        res["bg%d_z%d" % (bg, zc)] = dict(blocks=numCb, ms=ms, g_edge_updates_per_s=geups)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", os.environ.get("OUT", "exp_zc.json")), "w"), indent=1)
