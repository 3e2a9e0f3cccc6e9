#!/bin/bash
# helper kernels: timing only (16 384 and 4 096 code blocks)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NTB=1024 timeout 200 python scripts/bench_kernels.py > gpurun_out/u_helpers_16k.json 2> gpurun_out/u_helpers.err
NTB=256 timeout 200 python scripts/bench_kernels.py > gpurun_out/u_helpers_4k.json 2>> gpurun_out/u_helpers.err
python - <<'PY'
import json
for f in ("u_helpers_16k", "u_helpers_4k"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, {k: round(v["frac_of_measured_hbm"], 3) for k, v in d["stages"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/u_helpers.err
