#!/bin/bash
# BG2 Zc=160 (600 G edge-updates/s) against Zc=176 (753): ncu sections of the decoder launch
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for z in 160 176; do
WAVES=8 ROWS=17 ZCS=$z OUT=z_tmp.json timeout 300 ncu --section SpeedOfLight --section WarpStateStats --section Occupancy --section LaunchStats --section SchedulerStats --section MemoryWorkloadAnalysis --section InstructionStats \
  --clock-control none -k regex:nr_decode_kernel -s 9 -c 1 --csv --page raw --log-file gpurun_out/z_ncu_bg2_$z.csv python scripts/exp_zc.py > gpurun_out/z_$z.log 2>&1
done
ls -la gpurun_out/z_ncu*
