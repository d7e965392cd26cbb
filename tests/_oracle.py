"""ctypes bindings for the TEST-ONLY checkers under oracle/.

* ``Oracle``  -> oracle/_build/libpfac_oracle.so (plain-C restatement, oracle/pfac_oracle.c)
* ``RefBuild`` -> oracle/_ref/libphfpfac_ref.so  (the reference's own table builder compiled
  from /root/reference by oracle/Makefile; present only where it was built)

Nothing under phfpfac_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libpfac_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libphfpfac_ref.so")
REFERENCE_TREE = "/root/reference/regex_GPU_PHF"

_i32p = C.POINTER(C.c_int32)


def build_oracle():
    """Compile oracle/ (and oracle/_ref when the reference tree is mounted)."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)
    if os.path.isdir(REFERENCE_TREE):
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref"], check=True)


def _load_oracle():
    if not os.path.exists(ORACLE_SO):
        build_oracle()
    lib = C.CDLL(ORACLE_SO)
    lib.oracle_build_mem.restype = C.c_void_p
    lib.oracle_build_mem.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.oracle_build_mem_ext.restype = C.c_void_p
    lib.oracle_build_mem_ext.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.oracle_build_file.restype = C.c_void_p
    lib.oracle_build_file.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.oracle_free.argtypes = [C.c_void_p]
    for name in ("oracle_n_parts", "oracle_n_patterns", "oracle_max_pat_len"):
        getattr(lib, name).argtypes = [C.c_void_p]
        getattr(lib, name).restype = C.c_int
    lib.oracle_part_info.argtypes = [C.c_void_p, C.c_int, _i32p]
    for name in ("oracle_part_r", "oracle_part_HT", "oracle_part_val", "oracle_part_idmap", "oracle_part_s0"):
        getattr(lib, name).argtypes = [C.c_void_p, C.c_int]
        getattr(lib, name).restype = _i32p
    lib.oracle_part_pfac_row.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.oracle_part_pfac_row.restype = _i32p
    lib.oracle_part_r_entries.argtypes = [C.c_void_p, C.c_int]
    lib.oracle_part_r_entries.restype = C.c_int
    for name in ("oracle_scan_dense", "oracle_scan_compact"):
        f = getattr(lib, name)
        f.restype = C.c_longlong
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong]
    lib.oracle_scan_tables_omp.restype = C.c_longlong
    lib.oracle_scan_tables_omp.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
        C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong]
    lib.oracle_scan_tables_omp_1pass.restype = C.c_longlong
    lib.oracle_scan_tables_omp_1pass.argtypes = lib.oracle_scan_tables_omp.argtypes
    lib.oracle_max_threads.restype = C.c_int
    lib.oracle_write_result.restype = C.c_int
    lib.oracle_write_result.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_longlong]
    lib.oracle_run_cli.restype = C.c_longlong
    lib.oracle_run_cli.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int]
    return lib


_ORACLE = None


def oracle_lib():
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = _load_oracle()
    return _ORACLE


def _arr(ptr, n):
    if n <= 0:
        return np.zeros(0, dtype=np.int32)
    return np.ctypeslib.as_array(ptr, shape=(n,)).copy()


class PartTables:
    """Canonical arrays of one partition (the thread_data fields, main.cc:19-32)."""

    def __init__(self, state_num, n_final, max_len, ht_size, width, s0, r, HT, val, idmap, extra=None):
        self.state_num, self.n_final, self.max_len = state_num, n_final, max_len
        self.ht_size, self.width = ht_size, width
        self.s0, self.r, self.HT, self.val, self.idmap = s0, r, HT, val, idmap
        self.extra = extra or {}

    @property
    def r_entries(self):
        return (self.state_num * 256) // self.width + 1

    def same_as(self, other):
        return (self.state_num == other.state_num and self.n_final == other.n_final
                and self.max_len == other.max_len and self.ht_size == other.ht_size
                and np.array_equal(self.s0, other.s0) and np.array_equal(self.r, other.r)
                and np.array_equal(self.HT, other.HT) and np.array_equal(self.val, other.val)
                and np.array_equal(self.idmap, other.idmap))


class Oracle:
    def __init__(self, pattern_bytes, n_parts=4, width=256, escapes=False):
        lib = oracle_lib()
        err = C.c_int(0)
        self._lib = lib
        build = lib.oracle_build_mem_ext if escapes else lib.oracle_build_mem
        self._h = build(pattern_bytes, len(pattern_bytes), n_parts, width, C.byref(err))
        if not self._h:
            raise ValueError(f"oracle build failed: {err.value}")
        self.err = err.value
        self.width = width
        self.n_parts = lib.oracle_n_parts(self._h)
        self.n_patterns = lib.oracle_n_patterns(self._h)
        self.max_pat_len = lib.oracle_max_pat_len(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.oracle_free(self._h)
            self._h = None

    def part(self, g):
        lib, h = self._lib, self._h
        info = (C.c_int32 * 9)()
        lib.oracle_part_info(h, g, info)
        state_num, n_final, max_len, ht_size, max_row, n_keys, max_key, max_off, width = list(info)
        n_r = lib.oracle_part_r_entries(h, g)
        return PartTables(
            state_num, n_final, max_len, ht_size, width,
            _arr(lib.oracle_part_s0(h, g), 256), _arr(lib.oracle_part_r(h, g), n_r),
            _arr(lib.oracle_part_HT(h, g), ht_size), _arr(lib.oracle_part_val(h, g), ht_size),
            _arr(lib.oracle_part_idmap(h, g), n_final),
            extra=dict(max_row=max_row, n_keys=n_keys, max_key=max_key, max_offset=max_off))

    def pfac_row(self, g, state):
        return _arr(self._lib.oracle_part_pfac_row(self._h, g, state), 256)

    def scan(self, data, n=None, dense=False):
        """-> (pos int64[], id int32[]) in the reference's emit order (main.cc:341-349)."""
        buf = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        n = len(buf) if n is None else n
        fn = self._lib.oracle_scan_dense if dense else self._lib.oracle_scan_compact
        p = buf.ctypes.data if len(buf) else None
        cnt = fn(self._h, p, n, None, None, 0)
        if cnt < 0:
            raise ValueError(f"oracle scan failed: {cnt}")
        pos = np.zeros(cnt, dtype=np.int64)
        ids = np.zeros(cnt, dtype=np.int32)
        if cnt:
            fn(self._h, p, n, pos.ctypes.data, ids.ctypes.data, cnt)
        return pos, ids


def scan_tables_cpu(t, idmap, max_pat_len, data, nthreads=1, ref_tile_bound=True, count_only=False):
    """oracle_scan_tables_omp over canonical arrays (any builder's)."""
    lib = oracle_lib()
    buf = data if isinstance(data, np.ndarray) else np.frombuffer(bytes(data), dtype=np.uint8)
    s0 = np.ascontiguousarray(t.s0, dtype=np.int32)
    r = np.ascontiguousarray(t.r, dtype=np.int32)
    HT = np.ascontiguousarray(t.HT, dtype=np.int32)
    val = np.ascontiguousarray(t.val, dtype=np.int32)
    im = np.ascontiguousarray(idmap, dtype=np.int32)
    args = [s0.ctypes.data, r.ctypes.data, HT.ctypes.data, val.ctypes.data, t.ht_size, t.width,
            t.n_final, im.ctypes.data, max_pat_len, int(ref_tile_bound), buf.ctypes.data, len(buf), nthreads]
    cnt = lib.oracle_scan_tables_omp(*args, None, None, 0)
    if count_only:
        return cnt
    pos = np.zeros(cnt, dtype=np.int64)
    ids = np.zeros(cnt, dtype=np.int32)
    if cnt:
        lib.oracle_scan_tables_omp(*args, pos.ctypes.data, ids.ctypes.data, cnt)
    return pos, ids


def scan_tables_cpu_1pass(t, idmap, max_pat_len, data, nthreads=1, ref_tile_bound=True, cap=None):
    """oracle_scan_tables_omp_1pass: the input is walked ONCE and the records written (the timed CPU
    legs of bench.py).  cap = record capacity (default: one per 8 input bytes, grown if exceeded)."""
    lib = oracle_lib()
    buf = data if isinstance(data, np.ndarray) else np.frombuffer(bytes(data), dtype=np.uint8)
    s0 = np.ascontiguousarray(t.s0, dtype=np.int32)
    r = np.ascontiguousarray(t.r, dtype=np.int32)
    HT = np.ascontiguousarray(t.HT, dtype=np.int32)
    val = np.ascontiguousarray(t.val, dtype=np.int32)
    im = np.ascontiguousarray(idmap, dtype=np.int32)
    args = [s0.ctypes.data, r.ctypes.data, HT.ctypes.data, val.ctypes.data, t.ht_size, t.width,
            t.n_final, im.ctypes.data, max_pat_len, int(ref_tile_bound), buf.ctypes.data, len(buf), nthreads]
    cap = cap or max(len(buf) // 8, 65536)
    while True:
        pos = np.empty(cap, dtype=np.int64)
        ids = np.empty(cap, dtype=np.int32)
        cnt = lib.oracle_scan_tables_omp_1pass(*args, pos.ctypes.data, ids.ctypes.data, cap)
        if cnt <= cap:
            return pos[:cnt], ids[:cnt]
        cap = cnt


def render_result(pos, ids):
    """main.cc:344 line format, as bytes."""
    return "".join("At position %4d, match pattern %d\n" % (p, i) for p, i in zip(pos.tolist(), ids.tolist())).encode()


# ---------------------------------------------------------------- reference build

def ref_available():
    return os.path.exists(REF_SO)


_REF = None


def ref_lib():
    global _REF
    if _REF is None:
        lib = C.CDLL(REF_SO)
        lib.ref_build.restype = C.c_void_p
        lib.ref_build.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
        lib.ref_build_single.restype = C.c_void_p
        lib.ref_build_single.argtypes = [C.c_char_p, C.c_int, C.c_int]
        lib.ref_build_single_ext.restype = C.c_void_p
        lib.ref_build_single_ext.argtypes = [C.c_char_p, C.c_int, C.c_int]
        lib.ref_free.argtypes = [C.c_void_p]
        lib.ref_n_parts.argtypes = [C.c_void_p]
        lib.ref_max_pat_len.argtypes = [C.c_void_p]
        lib.ref_part_info.argtypes = [C.c_void_p, C.c_int, _i32p]
        for name in ("ref_part_r", "ref_part_HT", "ref_part_val", "ref_part_idmap", "ref_part_s0"):
            getattr(lib, name).argtypes = [C.c_void_p, C.c_int]
            getattr(lib, name).restype = _i32p
        lib.ref_part_pfac_row.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.ref_part_pfac_row.restype = _i32p
        _REF = lib
    return _REF


class RefBuild:
    """Tables built by the reference's own create_PFAC_table_reorder / patternsToPFAC / FFDM."""

    def __init__(self, pattern_file, streamnum=1, width=256, single=False, pfac_rows=None, escapes=False):
        lib = ref_lib()
        if pfac_rows is None:
            pfac_rows = os.path.getsize(pattern_file) + 16
        self._lib = lib
        self.width = width
        if escapes:
            self._h = lib.ref_build_single_ext(pattern_file.encode(), width, pfac_rows)
        elif single:
            self._h = lib.ref_build_single(pattern_file.encode(), width, pfac_rows)
        else:
            self._h = lib.ref_build(pattern_file.encode(), streamnum, width, pfac_rows)
        self.n_parts = lib.ref_n_parts(self._h)
        self.max_pat_len = lib.ref_max_pat_len(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.ref_free(self._h)
            self._h = None

    def part(self, g):
        lib, h = self._lib, self._h
        info = (C.c_int32 * 4)()
        lib.ref_part_info(h, g, info)
        state_num, n_final, max_len, ht_size = list(info)
        n_r = (state_num * 256) // self.width + 1
        return PartTables(
            state_num, n_final, max_len, ht_size, self.width,
            _arr(lib.ref_part_s0(h, g), 256), _arr(lib.ref_part_r(h, g), n_r),
            _arr(lib.ref_part_HT(h, g), ht_size), _arr(lib.ref_part_val(h, g), ht_size),
            _arr(lib.ref_part_idmap(h, g), n_final))

    def pfac_row(self, g, state):
        return _arr(self._lib.ref_part_pfac_row(self._h, g, state), 256)
