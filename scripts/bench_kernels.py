"""Per-kernel timing of the HBM-bound stages (K2 encode, K3a rate match, K3b rate recover, K4 CRC / segmentation / merge)
at BASELINE configs[1] sizes: algorithmic bytes / CUDA-event time vs the measured HBM peak.  Run under gpurun;
prints one JSON object.  (The decoder K1 is covered by bench.py.)"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
from neoradium_b200 import _dev, _native
from neoradium_b200.batch import TbBatchCodec

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
A, G, numTb = 8424 * 16 - 24, 14040 * 16, int(os.environ.get("NTB", "256"))   # 4096 code blocks: > L2 for most stages
codec = TbBatchCodec(1, '16QAM', A, G, precision='fp32')
L, h, s = _native.lib(), _dev.handle(), _dev.stream_ptr()
dev = codec.device
ncb = numTb * 16
gen = torch.Generator(device=dev); gen.manual_seed(1)
pl = torch.randint(0, 2, (numTb, A), dtype=torch.int8, device=dev, generator=gen)
tb = torch.empty((numTb, A + 24), dtype=torch.int8, device=dev)
cbs = torch.empty((ncb, 8448), dtype=torch.int8, device=dev)
coded = torch.empty((ncb, 66 * 384), dtype=torch.int8, device=dev)
rm = torch.empty((numTb, G), dtype=torch.int8, device=dev)
llr = torch.empty((numTb, G), dtype=torch.float32, device=dev)
rr = torch.empty((ncb, 66 * 384), dtype=torch.float32, device=dev)
merged = torch.empty((numTb, 16 * 8424), dtype=torch.int8, device=dev)
ok = torch.empty((ncb,), dtype=torch.uint8, device=dev)
okp = torch.empty((ncb,), dtype=torch.uint8, device=dev)
full = torch.empty((ncb, 68 * 384), dtype=torch.int8, device=dev)

def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3

P = _dev.ptr
stages = {
    "crc_attach_24A (TB)": (lambda: _native.check(L.nrldpc_crc_attach(h, P(pl), numTb, A, A, 3, P(tb), s)), numTb * (A + A + 24)),
    "segment (+CRC24B)": (lambda: _native.check(L.nrldpc_segment(h, codec.cfg, P(tb), numTb, A + 24, A + 24, P(cbs), s)), numTb * (A + 24) + ncb * 8448),
    "encode (K2)": (lambda: _native.check(L.nrldpc_encode(h, 1, 384, P(cbs), ncb, P(coded), 1, s)), ncb * (8448 + 25344)),
    "rate_match (K3a)": (lambda: _native.check(L.nrldpc_rate_match(h, codec.cfg, P(coded), numTb, P(rm), G, s)), ncb * (25344 + 14040)),
    "rate_recover f32 (K3b)": (lambda: _native.check(L.nrldpc_rate_recover(h, codec.cfg, 0, P(llr), numTb, G, G, None, P(rr), s)), ncb * (14040 * 4 + 25344 * 4)),
    "check_crc_and_merge (K4)": (lambda: _native.check(L.nrldpc_check_crc_and_merge(h, codec.cfg, P(cbs), numTb, P(merged), 16 * 8424, P(ok), s)), ncb * (8448 + 8424)),
    "crc_check_24B (CB)": (lambda: _native.check(L.nrldpc_crc_check(h, P(cbs), ncb, 8448, 8448, 4, P(ok), s)), ncb * 8448),
    "parity_check": (lambda: _native.check(L.nrldpc_parity_check(h, 1, 384, P(full), ncb, P(okp), s)), ncb * 68 * 384),
}
from neoradium_b200.modulation import awgn_llr
from neoradium_b200.scrambling import scramble_
stages["awgn_llr 16QAM (f1)"] = (lambda: awgn_llr(rm, 4, snr_db=9.0, seed=3, out=llr), numTb * G * (1 + 4))
stages["scramble_llrs f32 (f1)"] = (lambda: scramble_(12345, llr), numTb * G * 8)
# make inputs meaningful
stages["crc_attach_24A (TB)"][0](); stages["segment (+CRC24B)"][0](); stages["encode (K2)"][0](); stages["rate_match (K3a)"][0]()
llr.copy_((1 - 2 * rm.to(torch.float32)) * 3)
_native.check(L.nrldpc_encode(h, 1, 384, P(cbs), ncb, P(full), 0, s))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
out = {"code_blocks": ncb, "hbm_peak_gbs": peak, "stages": {}}
for name, (fn, nbytes) in stages.items():
    sec = t(fn)
    out["stages"][name] = {"ms": sec * 1e3, "algorithmic_MB": nbytes / 1e6, "GBps": nbytes / sec / 1e9, "frac_of_measured_hbm": nbytes / sec / 1e9 / peak,
                           "Mcb_per_s": ncb / sec / 1e6}
assert int(ok.sum()) == ncb and int(okp.sum()) == ncb
print(json.dumps(out))
