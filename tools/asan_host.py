#!/usr/bin/env python
"""Host code (table builder, derived filter tables, synthetic generators, writer) under
AddressSanitizer + UBSan.  No GPU, no CUDA: the four host translation units are built on their own.

    g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fPIC -shared -Iinclude -Iphfpfac_b200/csrc -pthread \
        -o /tmp/libpfac_host_asan.so phfpfac_b200/csrc/pfac_{tables,writer,derive}.cc tools/synth/pfac_synth.cc
    LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0 \
        python tools/asan_host.py /tmp/libpfac_host_asan.so
Development tool; last run clean (round 1)."""
import ctypes as C, sys, os, numpy as np, gzip, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(sys.argv[1])
vp = C.c_void_p
lib.pfac_tables_build_mem_ext.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.c_uint, C.POINTER(vp)]
lib.pfac_tables_destroy.argtypes = [vp]
lib.pfac_tables_derive_check.argtypes = [vp, C.c_int, C.c_uint, C.c_uint, C.c_uint, C.POINTER(C.c_uint64)]
lib.pfac_last_error.restype = C.c_char_p
lib.pfac_synth_patterns.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_char_p, C.c_size_t]
lib.pfac_synth_patterns.restype = C.c_longlong
def synth(kind, n, seed, lo, hi):
    need = lib.pfac_synth_patterns(kind, n, seed, lo, hi, None, 0)
    buf = C.create_string_buffer(need)
    assert lib.pfac_synth_patterns(kind, n, seed, lo, hi, buf, need) == need
    return buf.raw[:need]
rng = np.random.default_rng(1)
cases = [(b"aaaa\naa\na\naaa\n", 0), (gzip.decompress(open(os.path.join(ROOT, 'tests', 'golden', 'dictionary.txt.gz'), 'rb').read()), 0),
         (synth(1, 3000, 3, 4, 64), 0), (synth(0, 20000, 5, 8, 32), 0), (b"a\\x41\\n\\\nb\\0c\n", 1), (b"abc\n\nabd\n", 0), (b"abc", 0), (b"", 0),
         (b"x" * 1023 + b"\n", 0), (b"\\", 1), (b"ab\\", 1)]
for t in range(30):
    n = int(rng.integers(1, 300)); lines = set()
    while len(lines) < n:
        L = int(rng.integers(1, 20)); lines.add(bytes((rng.integers(0, 255, L) + 0).astype(np.uint8)).replace(b"\n", b"\x0b"))
    cases.append((b"".join(l + b"\n" for l in lines), 0))
for blob, flags in cases:
    for width in (1, 8, 256, 4096):
        for parts in (1, 3):
            h = vp()
            rc = lib.pfac_tables_build_mem_ext(blob, len(blob), parts, width, flags, C.byref(h))
            if rc == 0:
                st = (C.c_uint64 * 10)()
                for sizes in ((32768, 32768, 32768), (1024, 0, 0), (131072, 4096, 256)):
                    rc2 = lib.pfac_tables_derive_check(h, 0, *sizes, st)
                    assert rc2 == 0, (rc2, lib.pfac_last_error())
                lib.pfac_tables_destroy(h)
print("asan drive ok", len(cases))
# writer + synth text
lib.pfac_synth_text.argtypes = [C.c_int, C.c_uint64, C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int]
pats = synth(1, 500, 3, 4, 64)
for n in (0, 1, 65535, 65536, 300001):
    buf = np.zeros(max(n, 1), dtype=np.uint8)
    assert lib.pfac_synth_text(1, 7, buf.ctypes.data, n, pats, len(pats), 3) == 0
    assert lib.pfac_synth_text(0, 7, buf.ctypes.data, n, None, 0, 0) == 0
lib.pfac_format_records.restype = C.c_size_t
lib.pfac_format_records.argtypes = [C.c_uint64, C.c_void_p, C.c_uint64, C.c_char_p, C.c_size_t]
rec = np.zeros((1 << 20) + 77, dtype=[("pos", "<u4"), ("id", "<u4")]); rec["pos"] = np.arange(len(rec)) * 5; rec["id"] = np.arange(len(rec)) % 1000 + 1
for base in (0, 2 ** 40 + 3):
    need = lib.pfac_format_records(base, rec.ctypes.data, len(rec), None, 0)
    out = C.create_string_buffer(need)
    assert lib.pfac_format_records(base, rec.ctypes.data, len(rec), out, need) == need
lib.pfac_write_begin.argtypes = [C.c_char_p, C.POINTER(vp)]; lib.pfac_write_records.argtypes = [vp, C.c_uint64, C.c_void_p, C.c_uint64]; lib.pfac_write_end.argtypes = [vp]
w = vp(); assert lib.pfac_write_begin(b"/tmp/asan_out.txt", C.byref(w)) == 0
assert lib.pfac_write_records(w, 9, rec.ctypes.data, len(rec)) == 0 and lib.pfac_write_records(w, 9, rec.ctypes.data, 10) == 0
assert lib.pfac_write_end(w) == 0
print("asan writer/synth ok")
