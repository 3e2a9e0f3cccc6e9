#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
tail -5 gpurun_out/m_pytest.log
NTB=1024 timeout 300 python scripts/bench_kernels.py > gpurun_out/r02_helper_kernels_16k.json 2>gpurun_out/m_err.log
NTB=256 timeout 300 python scripts/bench_kernels.py > gpurun_out/r02_helper_kernels_4k.json 2>>gpurun_out/m_err.log
python - <<'PY'
import json
for f in ("gpurun_out/r02_helper_kernels_16k.json","gpurun_out/r02_helper_kernels_4k.json"):
    d=json.load(open(f)); print(d["code_blocks"], {k: round(v["frac_of_measured_hbm"],3) for k,v in d["stages"].items()})
PY
tail -3 gpurun_out/m_err.log
