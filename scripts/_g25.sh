TBS=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:nr_decode -s 9 -c 1 -o gpurun_out/decode_es -f python scripts/exp_cfg4.py > gpurun_out/ncu_es.log 2>&1
tail -3 gpurun_out/ncu_es.log
