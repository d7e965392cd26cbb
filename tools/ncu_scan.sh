#!/bin/bash
# One full ncu capture of the detector kernel (development tool).  TAG names the outputs.
mkdir -p gpurun_out
TAG=${TAG:-r2}
WL=${WL:-config3}
MIB=${MIB:-1024}
KREGEX=${KREGEX:-pfac_scan}
timeout 300 python tools/microbench.py --workload $WL --sizes $MIB --iters 3 > gpurun_out/${TAG}_mbplain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s ${SKIP:-3} -c 1 -f -o gpurun_out/${TAG}_prof python tools/microbench.py --workload $WL --sizes $MIB --iters 2 > gpurun_out/${TAG}_ncu.log 2>&1; echo "full capture rc=$?"
cat gpurun_out/${TAG}_mbplain.log | tail -3
tail -3 gpurun_out/${TAG}_ncu.log
