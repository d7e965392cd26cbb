#!/bin/bash
# Round evidence on ONE GPU (run under gpurun): smoke, the GPU parity suite, the bench lines of every workload, the
# reference arm, ncu launch lists and full captures of the three kernels, small inputs, the 1-GPU sweep, the CLI
# end to end.  Everything lands under gpurun_out/ with the prefix $TAG; copy what is to be kept to profiles/.
TAG=${TAG:-r2}
O=gpurun_out
mkdir -p $O
step() { echo "== $1"; }
step smoke; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "rc=$?"; tail -2 $O/${TAG}_smoke.log
step pytest; timeout 900 python -m pytest tests -m gpu -q --timeout 180 -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -2 $O/${TAG}_pytest.log
step bench
timeout 600 python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench_config3.json 2> $O/${TAG}_bench_config3.err; echo "config3 rc=$?"
for w in config2 config4 config1 dictionary; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err; echo "$w rc=$?"
done
for w in config3 config1 dictionary; do
  timeout 600 python bench.py --impl reference --workload $w --steps 3 --warmup 1 > $O/${TAG}_reference_arm_$w.json 2> $O/${TAG}_reference_arm_$w.err; echo "reference $w rc=$?"
done
step "launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/${TAG}_launches_config3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > $O/${TAG}_ll1.log 2>&1; echo "rc=$?"
for w in config1 dictionary; do   # warm L2 (--cache-control none): the product's and the reference's kernel under the same measurement
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 60 --csv --log-file $O/${TAG}_launches_warm_$w.csv python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/${TAG}_ll2.log 2>&1; echo "$w rc=$?"
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 60 --csv --log-file $O/${TAG}_launches_warm_reference_$w.csv python bench.py --impl reference --workload $w --steps 1 --warmup 0 > $O/${TAG}_ll3.log 2>&1; echo "reference $w rc=$?"
done
step "ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pfac_scan -s 3 -c 1 -f -o $O/${TAG}_prof_detector_config3 python tools/microbench.py --workload config3 --sizes 1024 --iters 2 > $O/${TAG}_ncu1.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pfac_finalize -s 3 -c 1 -f -o $O/${TAG}_prof_finalize_config3 python tools/microbench.py --workload config3 --sizes 1024 --iters 2 > $O/${TAG}_ncu2.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:pfac_scan -s 3 -c 1 -f -o $O/${TAG}_prof_detector_config4_warm python tools/microbench.py --workload config4 --sizes 512 --iters 2 > $O/${TAG}_ncu3.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pfac_dense -s 2 -c 1 -f -o $O/${TAG}_prof_dense_dictionary python tools/dense_bench.py --patterns dictionary --mib 64 --iters 2 > $O/${TAG}_ncu4.log 2>&1; echo "rc=$?"
step "small inputs"; timeout 300 python tools/small_inputs.py > $O/${TAG}_small_inputs.txt 2>&1; tail -5 $O/${TAG}_small_inputs.txt | cut -c1-150
step "dense fixtures tiled"; for pset in dictionary experimentpattern; do timeout 300 python tools/dense_bench.py --patterns $pset --mib 64 2>&1 | grep -v derived | tail -1 | cut -c1-200; done > $O/${TAG}_dense_64MiB.txt; cat $O/${TAG}_dense_64MiB.txt
step sweep; timeout 600 python tools/sweep.py > $O/${TAG}_sweep_1gpu.md 2> $O/${TAG}_sweep_1gpu.err; cat $O/${TAG}_sweep_1gpu.md
step "cli e2e"; timeout 900 bash tools/cli_e2e.sh ${CLI_GIB:-8} > $O/${TAG}_cli_e2e.txt 2>&1; grep -E "reader|Wall|match progress" $O/${TAG}_cli_e2e.txt
