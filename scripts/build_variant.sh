#!/bin/bash
# Build a variant of the library for A/B runs: scripts/build_variant.sh NAME "-DNR_DEC_R2P=0 ..."  -> neoradium_b200/libnrldpc_NAME.so
# (only the decoder translation units are recompiled; the other objects come from neoradium_b200/build, so build the default first)
set -e
NAME=$1; shift
ROOT=$(cd $(dirname $0)/.. && pwd)
B=/tmp/nrldpc_variant_$NAME; mkdir -p $B
cd $ROOT/neoradium_b200/csrc
pids=()
for f in decode*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c $f -o $B/${f%.cu}.o 2>$B/${f%.cu}.log &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
objs=$(ls $ROOT/neoradium_b200/build/*.o | grep -v '/decode' )
nvcc -shared -o $ROOT/neoradium_b200/libnrldpc_$NAME.so $objs $B/decode*.o -gencode arch=compute_100a,code=sm_100a
ls -la $ROOT/neoradium_b200/libnrldpc_$NAME.so
