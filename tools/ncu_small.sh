#!/bin/bash
# Development: where a 1 MiB scan spends its time -- variants + one full ncu capture of the detector at 1 MiB
mkdir -p gpurun_out
SIZES=1,1024 bash tools/variants.sh base noemit nos2 slot2 > /dev/null; cat gpurun_out/variants.log | grep -v derived
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pfac_scan -s 3 -c 1 -f -o gpurun_out/small_prof python tools/microbench.py --workload config3 --sizes 1 --iters 2 > gpurun_out/small_ncu.log 2>&1; echo "rc=$?"
