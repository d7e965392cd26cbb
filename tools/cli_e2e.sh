#!/bin/bash
# Development / profiles: the CLI end to end on a large generated file: file -> GPU_match_result.txt wall time,
# for the three input readers.  usage: cli_e2e.sh <GiB> [gpus]
GIB=${1:-8}
GPUS=${2:-1}
W=/tmp/gphf_e2e
mkdir -p $W gpurun_out
python - <<PY
import sys, os
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import pfac_synth as synth
from bench import WORKLOADS
pk, cnt, pseed, lo, hi, tk, tseed, nbytes, desc = WORKLOADS["config3"]
pats = synth.synth_patterns(pk, cnt, pseed, lo, hi)
open("$W/patterns", "wb").write(pats)
blk = synth.synth_text(tk, tseed, 1 << 30, patterns=pats)
with open("$W/input", "wb") as f:
    for i in range($GIB):
        f.write(blk.tobytes())
print("wrote", os.path.getsize("$W/input"), "bytes")
PY
cd $W
G=$OLDPWD/phfpfac_b200/_build/gphf
for r in stream mmap fread; do
  echo "== reader $r (file just written: page cache warm as far as it fits)"
  GPHF_GPUS=$GPUS GPHF_READER=$r $G patterns 4 256 input 2>&1 | grep -E "input reader|Time for|Wall time"
  md5sum GPU_match_result.txt
done
sync; (echo 3 > /proc/sys/vm/drop_caches) 2>/dev/null && echo "(page cache dropped)"
echo "== reader stream, O_DIRECT"
PFAC_READER_ODIRECT=1 GPHF_GPUS=$GPUS GPHF_READER=stream $G patterns 4 256 input 2>&1 | grep -E "Time for|Wall time"
md5sum GPU_match_result.txt
