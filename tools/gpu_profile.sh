#!/bin/bash
# bench line + ncu launch list + one full ncu capture of the scan kernel (development tool)
mkdir -p gpurun_out
TAG=${TAG:-r1}
BYTES=${BYTES:-1073741824}
timeout 600 python bench.py --steps ${STEPS:-20} --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json
timeout 300 python bench.py --steps 2 --warmup 1 --bytes $BYTES --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --bytes $BYTES --no-cpu-baseline --e2e-steps 1 > gpurun_out/${TAG}_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 300 python tools/microbench.py --workload config3 --sizes ${NCU_MIB:-1024} --iters 2 > gpurun_out/${TAG}_mbplain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pfac_scan -s 3 -c 1 -f -o gpurun_out/${TAG}_prof python tools/microbench.py --workload config3 --sizes ${NCU_MIB:-1024} --iters 2 > gpurun_out/${TAG}_ncu2.log 2>&1; echo "full capture rc=$?"
tail -3 gpurun_out/${TAG}_ncu2.log
