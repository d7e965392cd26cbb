#!/bin/bash
# Development: GPU parity suite + device-resident micro-benchmarks of the three synthetic configs (one gpurun call).
mkdir -p gpurun_out
TAG=${TAG:-it}
if [ -z "$NOTEST" ]; then
timeout ${PYTEST_TIMEOUT:-900} python -m pytest tests -m gpu -x -q --timeout ${TEST_TIMEOUT:-120} -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
fi
: > gpurun_out/${TAG}_mb.log
for w in ${WL:-config3 config2 config4}; do
  s=1024; [ $w = config4 ] && s=512
  timeout 300 python tools/microbench.py --workload $w --sizes $s --iters 5 2>&1 | grep -E "^detector|Error|error" | sed -E "s/info=.*//" >> gpurun_out/${TAG}_mb.log
done
cat gpurun_out/${TAG}_mb.log
