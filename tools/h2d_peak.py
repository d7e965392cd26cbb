#!/usr/bin/env python
"""Bare host-to-device copy rate of this box (pinned memory, CUDA events): the ceiling of bench.py's
e2e leg.  --gpus N copies to N GPUs at the same time (one thread each); --wc uses write-combined
pinned memory."""
import argparse, threading, time
from cuda import cudart

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--mib", type=int, default=1024)
ap.add_argument("--wc", action="store_true")
ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
n = a.mib << 20
flags = cudart.cudaHostAllocPortable | (cudart.cudaHostAllocWriteCombined if a.wc else 0)


def chk(r):
    if r[0] != cudart.cudaError_t.cudaSuccess:
        raise RuntimeError(str(r[0]))
    return r[1] if len(r) == 2 else r[1:]


bufs = []
for g in range(a.gpus):
    chk(cudart.cudaSetDevice(g))
    h = chk(cudart.cudaHostAlloc(n, flags))
    d = chk(cudart.cudaMalloc(n))
    st = chk(cudart.cudaStreamCreate())
    chk(cudart.cudaMemcpyAsync(d, h, n, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, st))
    chk(cudart.cudaStreamSynchronize(st))
    bufs.append((h, d, st))
bar = threading.Barrier(a.gpus + 1)
times = [0.0] * a.gpus


def work(g):
    chk(cudart.cudaSetDevice(g))
    h, d, st = bufs[g]
    bar.wait()
    t0 = time.perf_counter()
    for _ in range(a.reps):
        chk(cudart.cudaMemcpyAsync(d, h, n, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, st))
    chk(cudart.cudaStreamSynchronize(st))
    times[g] = time.perf_counter() - t0


th = [threading.Thread(target=work, args=(g,)) for g in range(a.gpus)]
for t in th:
    t.start()
bar.wait()
for t in th:
    t.join()
tot = a.gpus * a.reps * n / max(times) / 1e9
print(f"H2D {a.mib} MiB x {a.reps} to {a.gpus} GPU(s) at once, {'write-combined' if a.wc else 'plain'} pinned: "
      f"{tot:.1f} GB/s total, {tot / a.gpus:.1f} per GPU")
