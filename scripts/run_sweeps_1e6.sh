#!/bin/bash
# BASELINE configs[4] at full size on one B200 (profiles/r01_bler_sweeps_1e6.jsonl): 4 sweeps x 10 SNR points x 1e6 code blocks
rm -f gpurun_out/sweeps_r1.jsonl
python scripts/bler_sweep.py --tbs 250000 --bg 1 --mod 16QAM --A 33672 --rate 0.6 --iters 8 --snrs=7.0,7.4,7.8,8.0,8.2,8.4,8.6,8.8,9.0,9.4 --batch-tbs 512 --out gpurun_out/sweeps_r1.jsonl
python scripts/bler_sweep.py --tbs 1000000 --bg 2 --mod QPSK --A 3000 --rate 0.3 --iters 8 --snrs=-3.3,-3.0,-2.7,-2.4,-2.1,-1.8,-1.5,-1.2,-0.9,-0.6 --batch-tbs 2048 --out gpurun_out/sweeps_r1.jsonl
python scripts/bler_sweep.py --tbs 250000 --bg 1 --mod 256QAM --A 33672 --rate 0.75 --iters 8 --snrs=19.5,20.0,20.5,21.0,21.5,22.0,22.5,23.0,23.5,24.0 --batch-tbs 512 --out gpurun_out/sweeps_r1.jsonl
python scripts/bler_sweep.py --tbs 250000 --bg 1 --mod 64QAM --A 33672 --rate 0.5 --iters 8 --snrs=9.5,10.0,10.5,11.0,11.5,12.0,12.5,13.0,13.5,14.0 --batch-tbs 512 --out gpurun_out/sweeps_r1.jsonl
