"""Seeded synthetic workloads of BASELINE.json's configs (tools/synth/pfac_synth.cc): pattern sets and
texts for the tests and bench.py.  A tools library of its own -- nothing of it is in the product ABI."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "tools", "_build", "libpfac_synth.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            subprocess.run(["make", "-s", "-C", ROOT, "synth"], check=True)
        _lib = C.CDLL(SO)
        _lib.pfac_synth_patterns.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        _lib.pfac_synth_patterns.restype = C.c_longlong
        _lib.pfac_synth_text.argtypes = [C.c_int, C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    return _lib


def synth_patterns(kind, count, seed, min_len, max_len):
    lib = _load()
    n = lib.pfac_synth_patterns(kind, count, seed, min_len, max_len, None, 0)
    if n < 0:
        raise ValueError(f"pfac_synth_patterns failed: {n}")
    buf = C.create_string_buffer(int(n))
    lib.pfac_synth_patterns(kind, count, seed, min_len, max_len, buf, n)
    return buf.raw[:n]


def synth_text(kind, seed, n, patterns=None, n_threads=0, out=None):
    lib = _load()
    a = np.empty(n, dtype=np.uint8) if out is None else out
    pb = bytes(patterns) if patterns is not None else None
    rc = lib.pfac_synth_text(kind, seed, a.ctypes.data if n else None, n, pb, len(pb) if pb else 0, n_threads)
    if rc:
        raise ValueError(f"pfac_synth_text failed: {rc}")
    return a
