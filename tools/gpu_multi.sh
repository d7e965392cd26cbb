#!/bin/bash
# Multi-GPU bench lines of one box (run under gpurun --gpus N): weak and strong scaling of config 3, config 4,
# and the job-API sweep of large inputs.  usage: gpu_multi.sh N [steps]
N=$1; K=${2:-10}
mkdir -p gpurun_out
run() { # name, extra args...
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $K --warmup 3 "$@" > gpurun_out/r2_bench_${N}gpu_${name}.json 2> gpurun_out/r2_bench_${N}gpu_${name}.err
  echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_${N}gpu_${name}.json").read().strip().splitlines()[-1])
    print("  value", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "concurrent h2d", d["e2e"].get("h2d_concurrent_gbs_total"), "job", (d.get("e2e_job") or {}).get("value"), "parity", d["parity"]["equal"])
except Exception as e:
    print("  (no line)", e)
PY
}
run weak --no-cpu-baseline
run strong --scaling strong --no-cpu-baseline
[ "$N" = 8 ] && run config4 --workload config4 --no-cpu-baseline
if [ -n "$SWEEP" ]; then
  timeout 900 python tools/sweep.py --mib "" --job-gib ${SWEEP} --job-streams 1,4,8 > gpurun_out/r2_sweep_job_${N}gpu.md 2> gpurun_out/r2_sweep_job_${N}gpu.err; echo "sweep rc=$?"; cat gpurun_out/r2_sweep_job_${N}gpu.md
fi
nvidia-smi topo -m > gpurun_out/r2_topo_${N}gpu.txt 2>&1
