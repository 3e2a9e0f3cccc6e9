"""Experiment: last-wave tail of the decoder at 1024 code blocks.  (a) time per launch vs batch size (solo CTA per SM,
two CTAs per SM, multiples of the 296 CTA slots); (b) two batches in flight on two streams with their own handles."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neoradium_b200 import _native  # noqa: E402
from neoradium_b200.batch import TbBatchCodec  # noqa: E402
from neoradium_b200.modulation import awgn_llr  # noqa: E402

BG, MOD, QM, C = 1, "16QAM", 4, 16
A, G = 8424 * C - 24, 14040 * C
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)


def new_codec():
    c = TbBatchCodec(BG, MOD, A, G, precision="fp32", device=dev)
    p = ctypes.c_void_p()
    _native.check(_native.lib().nrldpc_create(0, ctypes.byref(p)))
    c._h = p
    return c


codec = new_codec()
gen = torch.Generator(device=dev)
gen.manual_seed(1)
res = {}


def make(tbs, nb):
    out = []
    for b in range(nb):
        pl = torch.randint(0, 2, (tbs, A), dtype=torch.int8, device=dev, generator=gen)
        out.append(awgn_llr(codec.encode(pl), QM, snr_db=9.0, seed=100 + b, offset=0))
    return out


def timeit(fn, n=40, w=5):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for tbs in (9, 18, 37, 55, 64, 74, 128, 148, 256):
    llrs = make(tbs, 4 if tbs <= 64 else 2)
    out = codec.alloc_outputs(tbs)
    ms = timeit(lambda i=0: codec.decode(llrs[i % len(llrs)], 8, out=out))
    res["serial_cb%d" % (tbs * C)] = ms
    print("cb=%5d  %.4f ms  %.2f Gbit/s" % (tbs * C, ms, tbs * A / ms / 1e6), flush=True)
    del llrs, out

# (b) K streams, each with its own handle and outputs, round-robin
tbs = 64
llrs = make(tbs, 4)
for ns in (1, 2, 3, 4):
    codecs = [new_codec() for _ in range(ns)]
    outs = [c.alloc_outputs(tbs) for c in codecs]
    streams = [torch.cuda.Stream(dev) for _ in range(ns)]

    def step(i=0):
        k = i % ns
        with torch.cuda.stream(streams[k]):
            codecs[k].decode(llrs[i % 4], 8, out=outs[k])

    def run(n):
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        for i in range(n):
            step(i)
        for s in streams:
            cur.wait_stream(s)

    run(8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 48
    e0.record()
    run(n)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    ok = all(int(o["tbOk"].sum().item()) == tbs for o in outs)
    res["streams%d" % ns] = ms
    print("streams=%d  %.4f ms/step  %.2f Gbit/s  ok=%s" % (ns, ms, tbs * A / ms / 1e6, ok), flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "exp_tail.json"), "w"))
