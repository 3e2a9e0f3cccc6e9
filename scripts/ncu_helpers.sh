#!/bin/bash
# ncu --set full capture of one launch of every HBM-bound helper kernel (scripts/bench_kernels.py, 16 384 code blocks).
# usage (under gpurun): bash scripts/ncu_helpers.sh TAG   -> gpurun_out/helper_TAG_<name>.ncu-rep
TAG=${1:-r1}
cap() {  # name regex skip
  NTB=1024 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/helper_${TAG}_$1 -f \
      python scripts/bench_kernels.py > /dev/null 2> gpurun_out/ncu_helper_$1.log
}
cap crc_attach nr_bitstream 5
cap segment nr_bitstream 18
cap merge nr_bitstream 31
cap crc_check nr_bitstream 44
cap encode nr_encode 3
cap rate_match nr_rate_match 3
cap rate_recover nr_rate_recover 3
cap parity nr_parity 3
cap awgn nr_awgn 3
cap scramble 'nr_scramble|nr_gold' 3
ls -la gpurun_out/helper_${TAG}_*.ncu-rep
