#!/usr/bin/env python
"""Fixed cost of a device-resident scan at small input sizes: whole step vs the detector kernel alone
(CUDA events).  Development tool."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import phfpfac_b200 as pf
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import pfac_synth as synth
from bench import WORKLOADS
pk, cnt, pseed, lo, hi, tk, tseed, nbytes, desc = WORKLOADS["config3"]
pats = synth.synth_patterns(pk, cnt, pseed, lo, hi)
tables = pf.Tables.from_bytes(pats)
text = synth.synth_text(tk, tseed, 128 << 20, patterns=pats)
d = torch.from_numpy(text).cuda()
m = pf.Matcher(tables, device=0)
m.set_timing(True)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
cap = 1 << 20
out = torch.empty((cap, 2), dtype=torch.int32, device="cuda"); cntd = torch.zeros(1, dtype=torch.int64, device="cuda")
for mib in (1, 4, 16, 64, 128):
    n = mib << 20
    for _ in range(3):
        m.scan_device_raw(d.data_ptr(), n, n, 0, out.data_ptr(), cap, cntd.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize(); m.kernel_time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        m.scan_device_raw(d.data_ptr(), n, n, 0, out.data_ptr(), cap, cntd.data_ptr(), st.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    kms, kn = m.kernel_time()
    print(f"{mib} MiB: step {ms*1e3:.1f} us, detector {kms/kn*1e3:.1f} us, emit+finalize+gaps {(ms-kms/kn)*1e3:.1f} us, {n/ms/1e6:.1f} GB/s")
