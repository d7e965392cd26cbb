"""phfpfac_b200 -- B200-native PFAC matcher: thin Python binding over the C ABI.

Host-side mirror of the reference's operator interface for the scan path
(regex_GPU_PHF/main.cc:35-37 and the thread_data block, main.cc:19-32):

    reference                                   here
    ------------------------------------------  ---------------------------------------------
    create_PFAC_table_reorder + FFDM            Tables.from_file / Tables.from_bytes
    thread_data{s0Table,r,HT,val,HTSize,...}    Tables.part(g)  (canonical arrays, bit-compatible)
    GPU_Malloc_Memory(...)                      Matcher(tables, device=...)
    GPU_TraceTable(...)                         Matcher.scan_host(...) / Matcher.scan_device(...)
    GPU_Free_memory(...)                        Matcher.close()
    merge + fprintf loop (main.cc:304-350)      format_records / write_result

All compute happens in phfpfac_b200/_build/libpfac_b200.so (hand-written sm_100a CUDA); this
module only marshals buffers.  PyTorch, when used, only supplies device memory and streams.
"""
import ctypes as C

import numpy as np

from ._lib import PFAC_ERR_OUTPUT_FULL, PfacError, check, lib

MATCH_DTYPE = np.dtype([("pos", "<u4"), ("id", "<u4")])   # struct pfac_match

__all__ = ["Tables", "Matcher", "Job", "PfacError", "MATCH_DTYPE", "format_records", "write_result",
           "write_sidecar", "read_sidecar", "pattern_file_hash", "pinned", "device_count", "plan_shard"]


def _as_u8(data):
    """The caller's bytes as a flat uint8 array WITHOUT a copy where they already are one (pinned buffers keep
    their address); anything that is not made of single bytes is refused rather than reinterpreted."""
    if isinstance(data, np.ndarray):
        if data.dtype.itemsize != 1:
            raise TypeError(f"input must be an array of bytes (uint8), not {data.dtype}")
        return np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    return np.frombuffer(bytes(data), dtype=np.uint8)


def _arr(ptr, n):
    if n <= 0 or not ptr:
        return np.zeros(0, dtype=np.int32)
    return np.ctypeslib.as_array(ptr, shape=(n,)).copy()


class PartView:
    """Canonical arrays of one partition -- the table half of thread_data (main.cc:19-32)."""

    def __init__(self, t, g):
        info = (C.c_int32 * 9)()
        check(lib.pfac_tables_part_info(t._h, g, info))
        (self.state_num, self.n_final, self.max_len, self.ht_size, self.n_r, self.n_keys, self.max_key,
         self.max_offset, self.min_len) = list(info)
        self.width = t.width
        self.s0 = _arr(lib.pfac_tables_s0(t._h, g), 256)
        self.r = _arr(lib.pfac_tables_r(t._h, g), self.n_r)
        self.HT = _arr(lib.pfac_tables_HT(t._h, g), self.ht_size)
        self.val = _arr(lib.pfac_tables_val(t._h, g), self.ht_size)
        self.idmap = _arr(lib.pfac_tables_idmap(t._h, g), self.n_final)


class Tables:
    """PFAC trie + PHF arrays (create_PFAC_table_reorder.c:6, phf.c:151)."""

    def __init__(self, handle):
        self._h = handle
        self.n_parts = lib.pfac_tables_n_parts(handle)
        self.n_patterns = lib.pfac_tables_n_patterns(handle)
        self.max_pat_len = lib.pfac_tables_max_pat_len(handle)
        self.width = lib.pfac_tables_width(handle)

    @classmethod
    def from_file(cls, path, n_parts=1, width=256, escapes=False):
        h = C.c_void_p()
        check(lib.pfac_tables_build_file_ext(str(path).encode(), n_parts, width, 1 if escapes else 0, C.byref(h)))
        return cls(h)

    @classmethod
    def from_bytes(cls, data, n_parts=1, width=256, escapes=False):
        """escapes=True: read_pattern_ext / fgetc_ext front-end (create_table_reorder.c:131, ctdef.h:37)."""
        data = bytes(data)
        h = C.c_void_p()
        check(lib.pfac_tables_build_mem_ext(data, len(data), n_parts, width, 1 if escapes else 0, C.byref(h)))
        return cls(h)

    @classmethod
    def from_arrays(cls, s0, r, HT, val, width, state_num, n_final, idmap, max_pat_len):
        a = [np.ascontiguousarray(x, dtype=np.int32) for x in (s0, r, HT, val, idmap)]
        h = C.c_void_p()
        check(lib.pfac_tables_from_arrays(a[0].ctypes.data, a[1].ctypes.data, len(a[1]), a[2].ctypes.data,
                                          a[3].ctypes.data, len(a[2]), width, state_num, n_final,
                                          a[4].ctypes.data, max_pat_len, C.byref(h)))
        return cls(h)

    @classmethod
    def load(cls, path):
        """A table set saved with save() (checksummed cache of the canonical arrays)."""
        h = C.c_void_p()
        check(lib.pfac_tables_load(str(path).encode(), C.byref(h)))
        return cls(h)

    def save(self, path):
        check(lib.pfac_tables_save(self._h, str(path).encode()))

    def source_hash(self):
        """Hash of the pattern file image + reader flags the set was built from (0: wrapped from arrays)."""
        return int(lib.pfac_tables_source_hash(self._h))

    def part(self, g=0):
        return PartView(self, g)

    def derive_check(self, g=0, t2_bytes=32768, t3_bytes=32768, tm2_bytes=32768):
        """Host-side build + verification of the detector's shared-memory filters (no GPU needed)."""
        st = (C.c_uint64 * 10)()
        check(lib.pfac_tables_derive_check(self._h, g, t2_bytes, t3_bytes, tm2_bytes, st))
        keys = ("image_bytes", "t1_pairs", "t2_set", "prefixes4", "has_short", "has_t3", "tm_keys", "tm2_keys",
                "t3_set", "tm2_bits")
        return dict(zip(keys, list(st)))

    def filter_profile(self, text, g=0, t2_bytes=32768, t3_bytes=32768, tm2_bytes=32768):
        """Diagnostics: survivors per stage of the detector's filter cascade over `text` (host model, counts only)."""
        buf = np.ascontiguousarray(text, dtype=np.uint8)
        c = (C.c_uint64 * 12)()
        check(lib.pfac_tables_filter_profile(self._h, g, t2_bytes, t3_bytes, tm2_bytes, buf.ctypes.data, len(buf), c))
        keys = ("positions", "t1_pass", "prefix_found", "window1_pass", "window2_pass", "bypass", "to_emit",
                "slices_flagged", "slices")
        return dict(zip(keys, list(c)))

    def lookup(self, state, byte, g=0):
        return lib.pfac_tables_lookup(self._h, g, state, byte)

    def close(self):
        if self._h and lib is not None:
            lib.pfac_tables_destroy(self._h)
        self._h = None

    def __del__(self):
        self.close()


def pattern_file_hash(path, escapes=False):
    h = C.c_uint64(0)
    check(lib.pfac_pattern_file_hash(str(path).encode(), 1 if escapes else 0, C.byref(h)))
    return h.value


def device_count():
    n = C.c_int(0)
    check(lib.pfac_device_count(C.byref(n)))
    return n.value


class Matcher:
    """One device: tables uploaded once, then scans (GPU_Malloc_Memory/GPU_TraceTable/GPU_Free_memory)."""

    def __init__(self, tables, device=0, part=0, n_streams=4, chunk_bytes=0):
        self._h = None
        h = C.c_void_p()
        check(lib.pfac_ctx_create(device, tables._h, part, n_streams, chunk_bytes, C.byref(h)))
        self._h = h
        self.tables = tables
        self.device = device

    def close(self):
        if self._h and lib is not None:
            lib.pfac_ctx_destroy(self._h)
        self._h = None

    def __del__(self):
        self.close()

    def last_info(self):
        info = (C.c_uint64 * 8)()
        check(lib.pfac_ctx_last_scan_info(self._h, info))
        keys = ("launches", "tiles", "ctas", "smem_bytes", "h2d_bytes", "d2h_bytes", "chunks", "flagged_tiles")
        return dict(zip(keys, list(info)))

    def set_timing(self, enable=True):
        check(lib.pfac_ctx_set_timing(self._h, 1 if enable else 0))

    def kernel_time(self):
        """(total ms, launches) of the detector kernel since the last call (needs set_timing)."""
        ms, n = C.c_double(0), C.c_int(0)
        check(lib.pfac_ctx_kernel_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def derived_info(self):
        info = (C.c_uint64 * 16)()
        check(lib.pfac_ctx_derived_info(self._h, info))
        keys = ("image_bytes", "t1_pairs", "t2_bits", "t2_set", "prefixes4", "has_short", "tm_keys", "tm2_keys",
                "t3_bits", "t3_set", "smem_bytes", "table_bytes", "ring_stages", "tm2_bits", "mode", "tm_bits")
        return dict(zip(keys, list(info)))

    # -- device-resident input (raw pointers; torch tensors are accepted for convenience)
    def scan_device_raw(self, d_in, n_starts, n_valid, base_pos, d_out, cap, d_count, stream=0):
        check(lib.pfac_scan_device(self._h, d_in, n_starts, n_valid, base_pos, d_out, cap, d_count, stream))

    def scan_device(self, t_in, n_starts=None, n_valid=None, base_pos=0, cap=None, offset=0):
        """t_in: torch uint8 CUDA tensor.  Returns a MATCH_DTYPE array (synchronous)."""
        import torch
        total = t_in.numel() - offset
        n_valid = total if n_valid is None else n_valid
        n_starts = n_valid if n_starts is None else n_starts
        cap = max(1024, n_starts // 8) if cap is None else cap
        while True:
            out = torch.empty((max(cap, 1), 2), dtype=torch.int32, device=t_in.device)
            cnt = C.c_uint64(0)
            # on torch's current stream: ordered after whatever produced t_in (and allocated `out`) there
            stream = torch.cuda.current_stream(t_in.device).cuda_stream
            if stream == 0:     # the legacy default stream: 0 means "the library's own stream" to the C ABI
                torch.cuda.current_stream(t_in.device).synchronize()
            rc = lib.pfac_scan_device_sync(self._h, t_in.data_ptr() + offset, n_starts, n_valid, base_pos,
                                           out.data_ptr(), cap, C.byref(cnt), stream or None)
            if rc == PFAC_ERR_OUTPUT_FULL:
                cap = cnt.value
                continue
            check(rc)
            rec = out[:cnt.value].cpu().numpy().astype(np.uint32).reshape(-1, 2)
            res = np.zeros(cnt.value, dtype=MATCH_DTYPE)
            res["pos"], res["id"] = rec[:, 0], rec[:, 1]
            return res

    # -- host input: H2D + kernel + D2H pipeline
    def scan_host(self, data, n_starts=None, base_pos=0, cap=None):
        buf = _as_u8(data)
        n_valid = len(buf)
        n_starts = n_valid if n_starts is None else n_starts
        cap = max(1024, n_starts // 8) if cap is None else cap
        while True:
            out = np.zeros(max(cap, 1), dtype=MATCH_DTYPE)
            cnt = C.c_uint64(0)
            rc = lib.pfac_scan_host(self._h, buf.ctypes.data if len(buf) else None, n_starts, n_valid, base_pos,
                                    out.ctypes.data, cap, C.byref(cnt))
            if rc == PFAC_ERR_OUTPUT_FULL:
                cap = cnt.value
                continue
            check(rc)
            return out[:cnt.value]


class Job:
    """All GPUs of the box: input sharded in contiguous chunks (+halo) -- the loop of main.cc:171-272."""

    def __init__(self, tables, devices=None, streams_per_gpu=4, chunk_bytes=0):
        self._h = None
        if devices is None:
            devices = list(range(device_count()))
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        check(lib.pfac_job_create(tables._h, arr, len(devices), streams_per_gpu, chunk_bytes, C.byref(h)))
        self._h = h
        self.tables = tables

    def run(self, data):
        """-> (total matches, [(base position, uint32 array [count, 2] of (pos - base, id)), ...] in position order)."""
        buf = _as_u8(data)
        n = C.c_uint64(0)
        check(lib.pfac_job_run(self._h, buf.ctypes.data if len(buf) else None, len(buf), C.byref(n)))
        return n.value, self._segments()

    def _segments(self):
        segs = []
        for i in range(lib.pfac_job_n_segments(self._h)):
            base, cnt, ptr = C.c_uint64(0), C.c_uint64(0), C.c_void_p()
            check(lib.pfac_job_segment(self._h, i, C.byref(base), C.byref(ptr), C.byref(cnt)))
            if cnt.value:
                a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint32)), shape=(cnt.value, 2)).copy()
            else:
                a = np.zeros((0, 2), dtype=np.uint32)
            segs.append((base.value, a))
        return segs

    def run_file(self, path, n):
        """pfac_job_run_file: scan the first n bytes of a file, read chunk by chunk while scanning."""
        nm = C.c_uint64(0)
        check(lib.pfac_job_run_file(self._h, str(path).encode(), n, C.byref(nm)))
        return nm.value, self._segments()

    def close(self):
        if self._h and lib is not None:
            lib.pfac_job_destroy(self._h)
        self._h = None

    def __del__(self):
        self.close()


class pinned:
    """Context manager: pin a numpy array in place (pfac_host_register) for link-speed H2D copies."""

    def __init__(self, arr, read_only=False):
        self.arr, self.read_only, self.ok = arr, read_only, False

    def __enter__(self):
        check(lib.pfac_host_register(self.arr.ctypes.data, self.arr.nbytes, int(self.read_only)))
        self.ok = True
        return self.arr

    def __exit__(self, *exc):
        if self.ok:
            lib.pfac_host_unregister(self.arr.ctypes.data)
        return False


def plan_shard(n, n_shards, max_pat_len, i):
    """pfac_job_plan: (start, n_starts, n_valid) of shard i -- contiguous chunk + halo of max_pat_len-1."""
    a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    check(lib.pfac_job_plan(n, n_shards, max_pat_len, i, C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def format_records(records, base_pos=0):
    """main.cc:344 line format for a MATCH_DTYPE array -> bytes."""
    rec = np.ascontiguousarray(records, dtype=MATCH_DTYPE)
    need = lib.pfac_format_records(base_pos, rec.ctypes.data if len(rec) else None, len(rec), None, 0)
    buf = C.create_string_buffer(need + 1)
    lib.pfac_format_records(base_pos, rec.ctypes.data if len(rec) else None, len(rec), buf, need)
    return buf.raw[:need]


def write_result(path, segments):
    """segments: iterable of (base_pos, MATCH_DTYPE array).  Writes GPU_match_result.txt (main.cc:335-350)."""
    w = C.c_void_p()
    check(lib.pfac_write_begin(str(path).encode(), C.byref(w)))
    try:
        for base, rec in segments:
            rec = np.ascontiguousarray(rec, dtype=MATCH_DTYPE)
            check(lib.pfac_write_records(w, base, rec.ctypes.data if len(rec) else None, len(rec)))
    finally:
        check(lib.pfac_write_end(w))


def write_sidecar(path, segments):
    """segments: iterable of (base_pos, MATCH_DTYPE array).  Writes the binary sidecar of the compact records
    (include/pfac_b200.h, pfac_sidecar_*): 8 bytes per match; the text file is a pure function of it."""
    w = C.c_void_p()
    check(lib.pfac_sidecar_begin(str(path).encode(), C.byref(w)))
    try:
        for base, rec in segments:
            rec = np.ascontiguousarray(rec, dtype=MATCH_DTYPE)
            check(lib.pfac_sidecar_records(w, base, rec.ctypes.data if len(rec) else None, len(rec)))
    finally:
        check(lib.pfac_sidecar_end(w))


def read_sidecar(path):
    """-> (pos: uint64 absolute start positions, id: uint32 pattern ids) in file (= position) order."""
    n = C.c_uint64(0)
    check(lib.pfac_sidecar_read(str(path).encode(), None, None, 0, C.byref(n)))
    pos = np.empty(n.value, dtype=np.uint64)
    ids = np.empty(n.value, dtype=np.uint32)
    check(lib.pfac_sidecar_read(str(path).encode(), pos.ctypes.data, ids.ctypes.data, n.value, C.byref(n)))
    return pos, ids
