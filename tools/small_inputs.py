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
    # the same without the event pair around the detector (which breaks the chain of dependent launches),
    # and the host's own time to enqueue a scan
    m.set_timing(False)
    import time
    for _ in range(3):
        m.scan_device_raw(d.data_ptr(), n, n, 0, out.data_ptr(), cap, cntd.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    e0.record()
    t0 = time.perf_counter()
    for _ in range(200):
        m.scan_device_raw(d.data_ptr(), n, n, 0, out.data_ptr(), cap, cntd.data_ptr(), st.cuda_stream)
    t_enq = (time.perf_counter() - t0) / 200
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / 200
    # one scan at a time (host waits for each): latency of a single call
    t0 = time.perf_counter()
    for _ in range(50):
        m.scan_device_raw(d.data_ptr(), n, n, 0, out.data_ptr(), cap, cntd.data_ptr(), st.cuda_stream)
        st.synchronize()
    t_one = (time.perf_counter() - t0) / 50
    m.set_timing(True)
    print(f"{mib} MiB: step {ms2*1e3:.1f} us back to back ({n/ms2/1e6:.1f} GB/s; host enqueue {t_enq*1e6:.1f} us per scan; one call + sync {t_one*1e6:.1f} us); "
          f"with event timing: step {ms*1e3:.1f} us, detector {kms/kn*1e3:.1f} us")
