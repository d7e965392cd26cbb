#!/usr/bin/env python
"""Input-size / stream-count sweep on one GPU (BASELINE.json configs[4]): device-resident and
end-to-end GB/s of the config-3 pattern set over 1 MiB .. 4 GiB.  Prints a markdown table."""
import argparse, ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import phfpfac_b200 as pf
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import pfac_synth as synth
from bench import WORKLOADS

ap = argparse.ArgumentParser()
ap.add_argument("--mib", default="1,4,16,64,256,1024,3072")
ap.add_argument("--streams", default="1,2,4,8")
ap.add_argument("--workload", default="config3")
a = ap.parse_args()
pk, cnt, pseed, lo, hi, tk, tseed, nbytes, desc = WORKLOADS[a.workload]
pats = synth.synth_patterns(pk, cnt, pseed, lo, hi)
tables = pf.Tables.from_bytes(pats)
sizes = [int(x) << 20 for x in a.mib.split(",")]
nmax = max(sizes)
h_text = torch.empty(nmax, dtype=torch.uint8, pin_memory=True)
synth.synth_text(tk, tseed, nmax, patterns=pats, out=h_text.numpy())
d_text = h_text.cuda()
cap = max(nmax // 8, 1 << 16)
d_out = torch.empty((cap, 2), dtype=torch.int32, device="cuda")
d_cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
h_out = torch.empty((cap, 2), dtype=torch.int32, pin_memory=True)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
print(f"| input | device-resident GB/s | " + " | ".join(f"e2e GB/s, {s} streams" for s in a.streams.split(",")) + " |")
print("|---|---|" + "---|" * len(a.streams.split(",")))
for n in sizes:
    m = pf.Matcher(tables, device=0, n_streams=1)
    iters = max(3, min(50, (2 << 30) // n))
    for _ in range(3):
        m.scan_device_raw(d_text.data_ptr(), n, n, 0, d_out.data_ptr(), cap, d_cnt.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        m.scan_device_raw(d_text.data_ptr(), n, n, 0, d_out.data_ptr(), cap, d_cnt.data_ptr(), st.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    dev = n * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9
    m.close()
    cols = []
    for s in [int(x) for x in a.streams.split(",")]:
        chunk = max(1 << 20, min(32 << 20, n // max(s, 1)))
        m = pf.Matcher(tables, device=0, n_streams=s, chunk_bytes=chunk)
        c = C.c_uint64(0)
        for _ in range(2):
            pf.check(pf.lib.pfac_scan_host(m._h, h_text.data_ptr(), n, n, 0, h_out.data_ptr(), cap, C.byref(c)))
        it2 = max(2, min(20, (1 << 30) // n))
        t0 = time.perf_counter()
        for _ in range(it2):
            pf.check(pf.lib.pfac_scan_host(m._h, h_text.data_ptr(), n, n, 0, h_out.data_ptr(), cap, C.byref(c)))
        cols.append(n * it2 / (time.perf_counter() - t0) / 1e9)
        m.close()
    print(f"| {n >> 20} MiB | {dev:.1f} | " + " | ".join(f"{x:.1f}" for x in cols) + " |", flush=True)
