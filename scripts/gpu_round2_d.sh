#!/bin/bash
# round-2 call D: real callers on the drop-in (unmodified HarqEntity / SnrScheduler), HARQ notebook statistic, reference arm
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ls baseline/_ref | head -3
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
tail -15 gpurun_out/d_pytest.log
timeout 900 python scripts/run_harq_notebook.py --transmissions 1000 --ref-transmissions 40 --out gpurun_out/r02_harq_notebook.json > gpurun_out/d_harq.log 2>&1
tail -60 gpurun_out/d_harq.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/d_ref_arm.json 2> gpurun_out/d_ref_arm.err
cat gpurun_out/d_ref_arm.json
