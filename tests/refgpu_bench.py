#!/usr/bin/env python
"""The reference's own GPU path (GPU_Malloc_Memory / GPU_TraceTable / GPU_Free_memory of
master_kernel.cu, built for sm_100a by `make -C oracle refgpu`) on this box: a baseline number for
the same pattern set, and its dense result as one more parity check of the product's records.
Test infrastructure (lives under tests/): the product never loads oracle/_ref.  The reference's dense result is
4 * max_pat_len bytes per input byte and indexed in 32 bits, so the input is limited to
2^32 / (4 * max_pat_len) bytes (config 3, max_pat_len 64: 16 MiB)."""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pfac_synth as synth
from bench import WORKLOADS
from _oracle import Oracle
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="config3", help="a synthetic workload of bench.py, or a reference fixture: "
                "config1 (experimentpattern over 1M) / dictionary (xaa..xad over 1M)")
ap.add_argument("--mib", type=int, default=0, help="input MiB (default: the largest the reference can index)")
ap.add_argument("--no-compare", action="store_true", help="time the reference only (no product scan, no sift)")
a = ap.parse_args()
so = os.path.join(ROOT, "oracle", "_ref", "libphfpfac_refgpu.so")
if not os.path.exists(so):
    print(json.dumps({"impl": "reference-gpu", "unavailable": "oracle/_ref/libphfpfac_refgpu.so not built (make -C oracle refgpu)"}))
    sys.exit(0)
ref = C.CDLL(so)
ref.refgpu_scan.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]

if a.workload in ("config1", "dictionary"):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_fixtures
    fx = load_fixtures()
    pats = fx["experimentpattern" if a.workload == "config1" else "dictionary"]
    desc = ("config1: the reference's experimentpattern over its 1M text" if a.workload == "config1"
            else "dictionary: the reference's xaa..xad (7,989 words) over its 1M text")
    tables = Oracle(pats, 1, 256)       # the oracle's restatement of CreateTable + FFDM (bit-identical to the reference's)
    p = tables.part(0)
    mpl = tables.max_pat_len
    n_max = ((1 << 32) // (4 * mpl)) - 4096
    reps = max(a.mib, 1)
    text = np.frombuffer(fx["1M"] * reps, dtype=np.uint8)[:min(reps << 20, n_max)].copy()
    if reps == 1:
        text = text[:len(text) - 1]        # the CLI drops the last byte of the file (main.cc:138)
    n = len(text)
else:
    pk, cnt, pseed, lo, hi, tk, tseed, nbytes, desc = WORKLOADS[a.workload]
    pats = synth.synth_patterns(pk, cnt, pseed, lo, hi)
    tables = Oracle(pats, 1, 256)
    p = tables.part(0)
    mpl = tables.max_pat_len
    n_max = ((1 << 32) // (4 * mpl)) - 4096
    n = min(a.mib << 20, n_max) if a.mib else n_max
    text = synth.synth_text(tk, tseed, n, patterns=pats)

_keep = []
def pinned(nbytes):      # pinned host memory as main.cc:147,161 (cudaHostAlloc), through torch
    t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    _keep.append(t)
    return C.c_void_p(t.data_ptr())

h_in = pinned(n + 4096)
C.memmove(h_in, text.ctypes.data, n)
h_res = pinned(n * mpl * 4)
s0 = np.ascontiguousarray(p.s0, dtype=np.int32); r = np.ascontiguousarray(p.r, dtype=np.int32)
HT = np.ascontiguousarray(p.HT, dtype=np.int32); val = np.ascontiguousarray(p.val, dtype=np.int32)
best = None
for rep in range(2):
    ms = (C.c_double * 3)()
    rc = ref.refgpu_scan(h_in, n, p.state_num, p.n_final, p.ht_size, 256, s0.ctypes.data, mpl, r.ctypes.data, HT.ctypes.data,
                         val.ctypes.data, h_res, ms)
    assert rc == 0
    if best is None or sum(ms) < sum(best):
        best = list(ms)
same, n_rec, sift_s = None, None, None
if not a.no_compare:
    res = np.ctypeslib.as_array(C.cast(h_res, C.POINTER(C.c_uint32)), shape=(n, mpl))
    t0 = time.perf_counter()
    rows, cols = np.nonzero(res != 0xFFFFFFFF)       # the host-side sift of main.cc:304-350
    states = res[rows, cols]
    ids = np.asarray(p.idmap)[states]
    sift_s = time.perf_counter() - t0
    import phfpfac_b200 as pf            # the product, only for the cross-check
    m = pf.Matcher(pf.Tables.from_bytes(pats, 1, 256))
    ours = m.scan_host(text)
    same = bool(len(ours) == len(rows) and np.array_equal(ours["pos"].astype(np.int64), rows)
                and np.array_equal(ours["id"].astype(np.int64), ids.astype(np.int64)))
    n_rec = int(len(rows))
total_ms = sum(best)
print(json.dumps({
    "impl": "reference-gpu", "workload": desc, "bytes": int(n), "max_pat_len": int(mpl), "matches": n_rec,
    "records_equal_to_product": same,
    "ms": {"malloc_memset": best[0], "trace_h2d_kernel_d2h": best[1], "free": best[2], "numpy_sift_of_dense_result": sift_s * 1e3 if sift_s is not None else None},
    "gbs_end_to_end": n / total_ms / 1e6,
    "note": "GPU_TraceTable prints its own H2D / kernel / D2H split above; dense result = %d bytes per input byte" % (4 * mpl)}))
