#!/bin/bash
# Bounded GPU check used during development: smoke, the GPU test-suite (per-test timeout, streamed
# log) and the device-resident micro-benchmark.  Every step has its own timeout.
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc" | tee -a gpurun_out/smoke.log
if [ $rc -ne 0 ]; then tail -5 gpurun_out/smoke.log; exit 1; fi
timeout ${PYTEST_TIMEOUT:-900} python -m pytest tests -m gpu -x -q --timeout ${TEST_TIMEOUT:-120} -p no:cacheprovider -v > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
if [ -n "$MB" ]; then
  for d in $MB; do PFAC_DEBUG=$d timeout 300 python tools/microbench.py --workload ${MB_WORKLOAD:-config3} --sizes ${MB_SIZES:-256,1024} >> gpurun_out/mb.log 2>&1; done
  tail -20 gpurun_out/mb.log
fi
