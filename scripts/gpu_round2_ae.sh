#!/bin/bash
# racecheck of smoke() with the layer barrier as bar.sync (libnrldpc_bar0.so = -DNR_DEC_BAR_MODE=0): are the reports of the
# default build artefacts of the mbarrier?
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export NRLDPC_LIB=$PWD/neoradium_b200/libnrldpc_bar0.so
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ae_racecheck_bar0.log 2>&1; echo "racecheck bar0 rc=$?"
grep -c "Race reported" gpurun_out/ae_racecheck_bar0.log
tail -4 gpurun_out/ae_racecheck_bar0.log
grep "Race reported" -A2 gpurun_out/ae_racecheck_bar0.log | cut -c1-260 | head -12
