#!/bin/bash
# Development: a few ncu counters of the detector kernel for several builds of the library.
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__thread_inst_executed_per_inst_executed.ratio
first=1
for v in "$@"; do
  lib=phfpfac_b200/_build/libpfac_b200_$v.so
  [ "$v" = base ] && lib=phfpfac_b200/_build/libpfac_b200.so
  if [ $first = 1 ]; then PFAC_B200_LIB=$PWD/$lib timeout 300 python tools/microbench.py --workload ${WL:-config3} --sizes ${MIB:-1024} --iters 2 > gpurun_out/ncuv_plain.log 2>&1 || { tail -5 gpurun_out/ncuv_plain.log; exit 1; }; first=0; fi
  PFAC_B200_LIB=$PWD/$lib timeout 600 ncu --metrics $M --clock-control none -k regex:${KREGEX:-pfac_scan} -s 3 -c 1 --csv --log-file gpurun_out/ncuv_$v.csv python tools/microbench.py --workload ${WL:-config3} --sizes ${MIB:-1024} --iters 2 > gpurun_out/ncuv_$v.log 2>&1
  echo "== $v rc=$?"
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncuv_$v.csv")) if len(r)>10]
hdr=rows[0]
for r in rows[1:]:
    print(f"{r[hdr.index('Metric Name')]:85s} {r[hdr.index('Metric Value')]}")
PY
done
