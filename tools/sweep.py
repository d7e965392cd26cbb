#!/usr/bin/env python
"""Input-size / stream-count / GPU-count sweep (BASELINE.json configs[4]).  Part 1, one GPU: device-resident
and end-to-end (pfac_scan_host) GB/s of a pattern set over 1 MiB .. 3 GiB at 1-8 streams.  Part 2
(--job-gib): inputs of several GiB through the product's own scheduler, pfac_job_run, from ONE process over
1, 2, 4, 8 of the box's GPUs (pinned host input -> H2D -> scan -> D2H, wall clock), and the same input
device-resident on one GPU in 1 GiB calls.  Prints markdown tables."""
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
ap.add_argument("--job-gib", default="", help="part 2: input sizes in GiB, e.g. 1,4,8,16,32")
ap.add_argument("--job-streams", default="1,4,8")
a = ap.parse_args()
pk, cnt, pseed, lo, hi, tk, tseed, nbytes, desc = WORKLOADS[a.workload]
pats = synth.synth_patterns(pk, cnt, pseed, lo, hi)
tables = pf.Tables.from_bytes(pats)
sizes = [int(x) << 20 for x in a.mib.split(",") if x]
job_sizes = [int(x) << 30 for x in a.job_gib.split(",") if x]
nmax = max(sizes + [1 << 20])
print(f"workload {a.workload}: {desc}; {torch.cuda.device_count()} GPU(s)\n")
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

if job_sizes:
    del d_text, h_text
    torch.cuda.empty_cache()
    ngpu = torch.cuda.device_count()
    gl = [g for g in (1, 2, 4, 8) if g <= ngpu]
    streams = [int(x) for x in a.job_streams.split(",")]
    print()
    print("| input | device-resident GB/s (1 GPU, 1 GiB calls) | " + " | ".join(f"job e2e GB/s, {g} GPU x {s} streams" for g in gl for s in streams) + " |")
    print("|---|---|" + "---|" * (len(gl) * len(streams)))
    base = None
    for n in job_sizes:
        try:
            h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        except Exception as e:   # not enough host memory to pin
            print(f"| {n >> 30} GiB | (cannot pin {n >> 30} GiB of host memory: {type(e).__name__}) |")
            break
        hn = h.numpy()
        if base is None:
            base = np.empty(1 << 30, dtype=np.uint8)
            synth.synth_text(tk, tseed, 1 << 30, patterns=pats, out=base)
        for o in range(0, n, 1 << 30):
            hn[o:o + (1 << 30)] = base[:min(1 << 30, n - o)]
        # device-resident: the same bytes in 1 GiB calls on GPU 0
        m = pf.Matcher(tables, device=0, n_streams=1)
        d = torch.from_numpy(base).cuda()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        calls = n >> 30
        for _ in range(2):
            m.scan_device_raw(d.data_ptr(), 1 << 30, 1 << 30, 0, d_out.data_ptr(), cap, d_cnt.data_ptr(), st.cuda_stream)
        e0.record()
        for _ in range(calls):
            m.scan_device_raw(d.data_ptr(), 1 << 30, 1 << 30, 0, d_out.data_ptr(), cap, d_cnt.data_ptr(), st.cuda_stream)
        e1.record(); torch.cuda.synchronize()
        dev = n / (e0.elapsed_time(e1) * 1e-3) / 1e9
        m.close(); del d
        cols = []
        for g in gl:
            for s_ in streams:
                job = pf.Job(tables, devices=list(range(g)), streams_per_gpu=s_)
                nm = C.c_uint64(0)
                pf.check(pf.lib.pfac_job_run(job._h, h.data_ptr(), n, C.byref(nm)))   # warm-up (allocations)
                best = 1e9
                for _ in range(2):
                    t0 = time.perf_counter()
                    pf.check(pf.lib.pfac_job_run(job._h, h.data_ptr(), n, C.byref(nm)))
                    best = min(best, time.perf_counter() - t0)
                cols.append(n / best / 1e9)
                job.close()
        print(f"| {n >> 30} GiB | {dev:.1f} | " + " | ".join(f"{x:.1f}" for x in cols) + f" | ({nm.value} matches)", flush=True)
        del h, hn
