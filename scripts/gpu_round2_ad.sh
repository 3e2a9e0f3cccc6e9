#!/bin/bash
# compute-sanitizer: memcheck over more of the GPU suite, racecheck + synccheck over smoke()
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_link.py tests/test_gpu_sweep.py tests/test_gpu_real_callers.py -m gpu -x -q > gpurun_out/ad_memcheck_tests2.log 2>&1; echo "memcheck tests2 rc=$?"
tail -3 gpurun_out/ad_memcheck_tests2.log
timeout 300 compute-sanitizer --tool synccheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ad_synccheck_smoke.log 2>&1; echo "synccheck rc=$?"
tail -3 gpurun_out/ad_synccheck_smoke.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ad_racecheck_smoke.log 2>&1; echo "racecheck rc=$?"
tail -5 gpurun_out/ad_racecheck_smoke.log
