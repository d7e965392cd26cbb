#!/usr/bin/env python
"""Device-resident timing of pfac_scan_device for one workload at several sizes (CUDA events on
the launching stream).  Development tool; bench.py is the contract."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import phfpfac_b200 as pf
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import pfac_synth as synth
from bench import WORKLOADS

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config3")
ap.add_argument("--sizes", default="32,256,1024")
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
pk, cnt, pseed, lo, hi, tk, tseed, nbytes, desc = WORKLOADS[a.workload]
pats = synth.synth_patterns(pk, cnt, pseed, lo, hi)
tables = pf.Tables.from_bytes(pats)
sizes = [int(s) << 20 for s in a.sizes.split(",")]
text = synth.synth_text(tk, tseed, max(sizes), patterns=pats)
d = torch.from_numpy(text).cuda()
m = pf.Matcher(tables, device=0)
print("derived:", m.derived_info(), flush=True)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
cap = max(sizes) // 8
out = torch.empty((cap, 2), dtype=torch.int32, device="cuda"); cntd = torch.zeros(1, dtype=torch.int64, device="cuda")
for n in sizes:
    for _ in range(2):
        m.scan_device_raw(d.data_ptr(), n, n, 0, out.data_ptr(), cap, cntd.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        m.scan_device_raw(d.data_ptr(), n, n, 0, out.data_ptr(), cap, cntd.data_ptr(), st.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    m.set_timing(True)
    for _ in range(a.iters):
        m.scan_device_raw(d.data_ptr(), n, n, 0, out.data_ptr(), cap, cntd.data_ptr(), st.cuda_stream)
    kms, kn = m.kernel_time()
    m.set_timing(False)
    print(f"detector {kms / max(kn, 1):.3f} ms = {n / (kms / max(kn, 1)) / 1e6:.1f} GB/s;", end=" ")
    print(f"{a.workload} debug={os.environ.get('PFAC_DEBUG','0')} n={n>>20}MiB {ms:.3f} ms {n/ms/1e6:.1f} GB/s matches={int(cntd.item())} info={m.last_info()}", flush=True)
import ctypes as C
cnt = C.c_uint64(0)
pf.lib.pfac_scan_device_sync(m._h, d.data_ptr(), sizes[0], sizes[0], 0, out.data_ptr(), cap, C.byref(cnt), None)
print("sync scan:", cnt.value, m.last_info(), flush=True)
m.close(); tables.close()
