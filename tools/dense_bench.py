#!/usr/bin/env python
"""Device-resident timing of the reference's own (dense-match) fixtures tiled to a larger input:
dictionary.txt or experimentpattern over the 1M text.  Development tool."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import phfpfac_b200 as pf
from conftest import load_fixtures

ap = argparse.ArgumentParser()
ap.add_argument("--patterns", default="dictionary")
ap.add_argument("--mib", type=int, default=64)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
fx = load_fixtures()
tables = pf.Tables.from_bytes(fx[a.patterns])
text = np.frombuffer(fx["1M"] * a.mib, dtype=np.uint8)
d = torch.from_numpy(text.copy()).cuda()
m = pf.Matcher(tables, device=0)
print("derived:", m.derived_info(), flush=True)
n = len(text)
cap = n            # up to one record per byte
out = torch.empty((cap, 2), dtype=torch.int32, device="cuda"); cntd = torch.zeros(1, dtype=torch.int64, device="cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for _ in range(2):
    m.scan_device_raw(d.data_ptr(), n, n, 0, out.data_ptr(), cap, cntd.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
m.set_timing(True)
e0.record()
for _ in range(a.iters):
    m.scan_device_raw(d.data_ptr(), n, n, 0, out.data_ptr(), cap, cntd.data_ptr(), st.cuda_stream)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
print(f"{a.patterns} x 1M tiled n={a.mib}MiB {ms:.3f} ms {n/ms/1e6:.1f} GB/s matches={int(cntd.item())} ({int(cntd.item())/n:.3f}/byte) detector={m.kernel_time()} info={m.last_info()}", flush=True)
