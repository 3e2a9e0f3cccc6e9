#!/bin/bash
# helper kernels after the check-free CRC loop and the warp-per-block encoder / parity check: parity first, then timing
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
NTB=1024 timeout 200 python scripts/bench_kernels.py > gpurun_out/s_helpers_16k.json 2> gpurun_out/s_helpers.err
NTB=256 timeout 200 python scripts/bench_kernels.py > gpurun_out/s_helpers_4k.json 2>> gpurun_out/s_helpers.err
NRLDPC_ENC_CTA_PER_CB=1 NTB=1024 timeout 200 python scripts/bench_kernels.py > gpurun_out/s_helpers_16k_cta.json 2>> gpurun_out/s_helpers.err
python - <<'PY'
import json
for f in ("s_helpers_16k", "s_helpers_4k", "s_helpers_16k_cta"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, {k: round(v["frac_of_measured_hbm"], 3) for k, v in d["stages"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/s_helpers.err
