#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/microbench.py --workload config4 --sizes 512 --iters 3 > gpurun_out/c4_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:pfac_scan -s 3 -c 1 -f -o gpurun_out/c4_prof python tools/microbench.py --workload config4 --sizes 512 --iters 2 > gpurun_out/c4_ncu.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/c4_plain.log | head -1 | sed -E "s/info=.*//"
