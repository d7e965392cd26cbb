#!/usr/bin/env python
"""Survivors per stage of the scan kernel's filter cascade on a synthetic workload (host model of
the derived tables; development diagnostics, no GPU needed)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import phfpfac_b200 as pf
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import pfac_synth as synth
from bench import WORKLOADS
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config3")
ap.add_argument("--mib", type=int, default=8)
ap.add_argument("--t2", type=int, default=32768)
ap.add_argument("--t3", type=int, default=32768)
ap.add_argument("--tm2", type=int, default=32768)
a = ap.parse_args()
pk, cnt, pseed, lo, hi, tk, tseed, nbytes, desc = WORKLOADS[a.workload]
pats = synth.synth_patterns(pk, cnt, pseed, lo, hi)
t = pf.Tables.from_bytes(pats)
text = synth.synth_text(tk, tseed, a.mib << 20, patterns=pats)
print(t.derive_check(0, a.t2, a.t3, a.tm2))
c = t.filter_profile(text, 0, a.t2, a.t3, a.tm2)
n = c["positions"]
for k, v in c.items():
    print(f"{k:15s} {v:12d}  {v / n:.5f} per byte")
