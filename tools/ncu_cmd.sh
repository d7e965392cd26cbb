#!/bin/bash
# Development: one full ncu capture of a kernel of any command.  usage: TAG=x KREGEX=pfac_dense SKIP=2 ncu_cmd.sh cmd...
mkdir -p gpurun_out
TAG=${TAG:-cap}
"$@" > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-pfac} -s ${SKIP:-2} -c 1 -f -o gpurun_out/${TAG}_prof "$@" > gpurun_out/${TAG}_ncu.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/${TAG}_ncu.log
