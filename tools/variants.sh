#!/bin/bash
# Development: device-resident GB/s of several builds of the library (make variant ...), one line each.
mkdir -p gpurun_out
: > gpurun_out/variants.log
for v in "$@"; do
  lib=phfpfac_b200/_build/libpfac_b200_$v.so
  [ "$v" = base ] && lib=phfpfac_b200/_build/libpfac_b200.so
  echo "== $v" >> gpurun_out/variants.log
  PFAC_B200_LIB=$PWD/$lib timeout 300 python tools/microbench.py --workload ${WL:-config3} --sizes ${SIZES:-1024} --iters 5 2>&1 | grep -E "GB/s|derived|rror" | sed -E "s/info=.*//" >> gpurun_out/variants.log
done
cat gpurun_out/variants.log
