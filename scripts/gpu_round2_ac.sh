#!/bin/bash
# compute-sanitizer memcheck over smoke() and the parity tests of the kernels touched this round
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ac_memcheck_smoke.log 2>&1; echo "smoke rc=$?"
tail -4 gpurun_out/ac_memcheck_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "crc or segment or row_skipping or early_stop or multi_block or merge" > gpurun_out/ac_memcheck_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/ac_memcheck_tests.log
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/ac_memcheck_smoke.log gpurun_out/ac_memcheck_tests.log
