#!/bin/bash
# ncu sections for the CRC-check and merge launches of scripts/bench_kernels.py
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cap() {  # name skip
  NTB=1024 timeout 300 ncu --section SpeedOfLight --section WarpStateStats --section Occupancy --section LaunchStats --section SchedulerStats --section MemoryWorkloadAnalysis \
      --clock-control none -k regex:nr_bitstream -s $2 -c 1 --csv --page raw --log-file gpurun_out/t_ncu_$1.csv python scripts/bench_kernels.py > /dev/null 2> gpurun_out/t_ncu_$1.err
}
cap crc_check 44
cap merge 31
ls -la gpurun_out/t_ncu_*
