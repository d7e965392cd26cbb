#!/bin/bash
# Development: ncu launch list (device time of every kernel launch) of a command.  usage: launch_list.sh TAG cmd...
mkdir -p gpurun_out
TAG=$1; shift
"$@" > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c ${COUNT:-60} --csv --log-file gpurun_out/${TAG}_launches.csv "$@" > gpurun_out/${TAG}_ncu.log 2>&1
echo "rc=$?"
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_launches.csv")) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    print(f"{r[h.index('Kernel Name')][:60]:60s} grid={r[h.index('Grid Size')]:>12s} {float(r[h.index('Metric Value')])/1000:9.1f} us")
PY
