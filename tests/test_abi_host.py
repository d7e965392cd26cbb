"""C-ABI surface and host-side logic that needs no GPU: exported symbols == include/*.h,
loud failure without a device (no CPU fallback), the result writer, the synthetic generators and
the shard plan."""
import ctypes as C
import hashlib
import os
import re
import subprocess

import numpy as np
import pytest

import phfpfac_b200 as pf
import pfac_synth as synth
from phfpfac_b200._lib import ABI_SYMBOLS, LIB_PATH, lib
from _oracle import Oracle, render_result

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("pfac_b200.h",):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(pfac_[A-Za-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    decl = declared_symbols()
    assert decl == set(ABI_SYMBOLS), decl ^ set(ABI_SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (pfac_[A-Za-z0-9_]+)", out))
    assert decl <= exported, decl - exported
    for name in decl:
        getattr(lib, name)
    assert lib.pfac_abi_version() == 1


def test_header_has_no_cuda_or_torch_types():
    src = open(os.path.join(ROOT, "include", "pfac_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert "cudaStream_t" not in code and "torch" not in code and "#include <cuda" not in code


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under phfpfac_b200/ or include/ may include, import,
    load or call it (comments may cite it)."""
    for base in ("phfpfac_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if "_build" in dp or "__pycache__" in dp:
                continue
            for f in files:
                text = open(os.path.join(dp, f), errors="replace").read()
                bad = re.findall(r"#include[^\n]*oracle|import[^\n]*oracle|libpfac_oracle|\boracle_[a-z_]+\s*\(|dlopen|CDLL\([^)]*oracle",
                                 text)
                assert not bad, (os.path.join(dp, f), bad)
    out = subprocess.run(["ldd", LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_no_gpu_fails_loudly():
    """No CPU fallback: without a device the scan entry points return PFAC_ERR_NO_DEVICE."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    t = pf.Tables.from_bytes(b"abc\n", 1, 256)
    with pytest.raises(pf.PfacError) as e:
        pf.Matcher(t, device=0)
    assert e.value.code == -9
    with pytest.raises(pf.PfacError) as e:
        pf.Job(t, devices=[0])
    assert e.value.code == -9


def test_writer_matches_reference_format(tmp_path, golden):
    g = golden["results"]["experimentpattern_x_experimentinput"]
    rec = np.zeros(g["lines"], dtype=pf.MATCH_DTYPE)
    for i, line in enumerate(g["text"].splitlines()):
        m = re.fullmatch(r"At position +(\d+), match pattern (\d+)", line)
        rec[i] = (int(m.group(1)), int(m.group(2)))
    assert pf.format_records(rec).decode() == g["text"]
    f = tmp_path / "GPU_match_result.txt"
    pf.write_result(f, [(0, rec[:10]), (0, rec[10:])])
    assert f.read_text() == g["text"]
    # 64-bit positions keep the %4d rule (main.cc:344) at every magnitude
    rng = np.random.default_rng(1)
    big = np.zeros(1000, dtype=pf.MATCH_DTYPE)
    big["pos"] = np.sort(rng.integers(0, 2 ** 32, 1000, dtype=np.uint64)).astype(np.uint32)
    big["id"] = rng.integers(1, 2 ** 31, 1000)
    base = 5 * 2 ** 32 + 7
    want = render_result(big["pos"].astype(np.int64) + base, big["id"].astype(np.int64))
    assert pf.format_records(big, base_pos=base) == want
    pf.write_result(f, [(base, big)])
    assert f.read_bytes() == want
    assert pf.format_records(np.zeros(0, dtype=pf.MATCH_DTYPE)) == b""
    # the multi-threaded path (>= 2^20 records) writes the same bytes as the sequential formatter
    many = np.zeros((1 << 20) + 12345, dtype=pf.MATCH_DTYPE)
    many["pos"] = np.arange(len(many), dtype=np.uint32) * 3
    many["id"] = (np.arange(len(many)) % 9973) + 1
    pf.write_result(f, [(7, many[:1000]), (7, many[1000:])])
    assert f.read_bytes() == pf.format_records(many, base_pos=7)


def test_binary_sidecar_round_trip(tmp_path, golden):
    """pfac_sidecar_*: the compact records as a binary file.  The text result is a pure function of
    it: reading the sidecar back and formatting it gives the golden GPU_match_result.txt of the reference flow."""
    g = golden["results"]["experimentpattern_x_experimentinput"]
    rec = np.zeros(g["lines"], dtype=pf.MATCH_DTYPE)
    for i, line in enumerate(g["text"].splitlines()):
        m = re.fullmatch(r"At position +(\d+), match pattern (\d+)", line)
        rec[i] = (int(m.group(1)), int(m.group(2)))
    f = tmp_path / "records.pfacrec"
    pf.write_sidecar(f, [(0, rec[:10]), (0, rec[:0]), (0, rec[10:])])   # the empty segment leaves no block
    raw = f.read_bytes()
    assert raw[:8] == b"PFACREC1" and np.frombuffer(raw[8:16], dtype="<u4").tolist() == [1, 8]
    assert np.frombuffer(raw[16:32], dtype="<u8").tolist() == [2, g["lines"]]
    assert len(raw) == 32 + 2 * 16 + 8 * g["lines"]
    pos, ids = pf.read_sidecar(f)
    assert pos.dtype == np.uint64 and np.array_equal(pos, rec["pos"]) and np.array_equal(ids, rec["id"])
    assert render_result(pos.astype(np.int64), ids.astype(np.int64)).decode() == g["text"]
    # 64-bit base positions, several blocks
    rng = np.random.default_rng(2)
    segs, want_pos, want_id = [], [], []
    for k in range(5):
        r = np.zeros(int(rng.integers(1, 3000)), dtype=pf.MATCH_DTYPE)
        r["pos"] = np.sort(rng.integers(0, 2 ** 32, len(r), dtype=np.uint64)).astype(np.uint32)
        r["id"] = rng.integers(1, 2 ** 31, len(r))
        base = k * 2 ** 33 + 5
        segs.append((base, r))
        want_pos.append(r["pos"].astype(np.uint64) + np.uint64(base))
        want_id.append(r["id"])
    pf.write_sidecar(f, segs)
    pos, ids = pf.read_sidecar(f)
    assert np.array_equal(pos, np.concatenate(want_pos)) and np.array_equal(ids, np.concatenate(want_id))
    # no records at all
    pf.write_sidecar(f, [])
    pos, ids = pf.read_sidecar(f)
    assert len(pos) == 0 and len(ids) == 0 and f.stat().st_size == 32
    # damaged files are refused: wrong magic, a file cut short, a capacity that is too small
    pf.write_sidecar(f, segs)
    raw = f.read_bytes()
    bad = tmp_path / "bad.pfacrec"
    bad.write_bytes(b"XFACREC1" + raw[8:])
    with pytest.raises(pf.PfacError) as e:
        pf.read_sidecar(bad)
    assert e.value.code == -1
    bad.write_bytes(raw[:-4])
    with pytest.raises(pf.PfacError) as e:
        pf.read_sidecar(bad)
    assert e.value.code == -1
    bad.write_bytes(raw[:32 + 8] + np.array([2 ** 40], dtype="<u8").tobytes() + raw[48:])   # a block count beyond the total
    with pytest.raises(pf.PfacError) as e:
        pf.read_sidecar(bad)
    assert e.value.code == -1
    n = C.c_uint64(0)
    buf = np.zeros(4, dtype=np.uint64)
    assert lib.pfac_sidecar_read(str(f).encode(), buf.ctypes.data, None, 4, C.byref(n)) == -8
    assert n.value == sum(len(r) for _, r in segs) > 4   # the required capacity is reported
    with pytest.raises(pf.PfacError):
        pf.read_sidecar(tmp_path / "nope")


def test_input_buffers_are_passed_without_a_copy():
    """Matcher.scan_host / Job.run hand the caller's bytes to the library where they lie (a pinned buffer must
    keep its address); arrays that are not made of bytes are refused."""
    a = np.arange(4096, dtype=np.uint8)
    assert pf._as_u8(a).ctypes.data == a.ctypes.data and pf._as_u8(a[16:]).ctypes.data == a.ctypes.data + 16
    assert pf._as_u8(a.reshape(64, 64)).ctypes.data == a.ctypes.data and pf._as_u8(a.reshape(64, 64)).shape == (4096,)
    assert pf._as_u8(a.view(np.int8)).dtype == np.uint8
    assert pf._as_u8(a[::2]).tobytes() == a[::2].tobytes()            # strided: copied, same bytes
    assert pf._as_u8(b"abc").tobytes() == b"abc" and pf._as_u8(bytearray(b"xy")).tobytes() == b"xy"
    assert len(pf._as_u8(b"")) == 0
    with pytest.raises(TypeError):
        pf._as_u8(np.zeros(4, dtype=np.int32))


def test_synth_is_deterministic_and_shaped():
    p1 = synth.synth_patterns(1, 2000, 3, 4, 64)
    assert p1 == synth.synth_patterns(1, 2000, 3, 4, 64) and p1 != synth.synth_patterns(1, 2000, 4, 4, 64)
    lines = p1.split(b"\n")[:-1]
    assert len(lines) == 2000 == len(set(lines)) and all(4 <= len(x) <= 64 for x in lines)
    p0 = synth.synth_patterns(0, 1000, 1, 8, 32).split(b"\n")[:-1]
    assert len(set(p0)) == 1000 and all(8 <= len(x) <= 32 and all(0x21 <= c <= 0x7E for c in x) for x in p0)
    a = synth.synth_text(1, 4, 300000, patterns=p1, n_threads=1)
    b = synth.synth_text(1, 4, 300000, patterns=p1, n_threads=5)
    assert np.array_equal(a, b)
    assert np.array_equal(a[:131072], synth.synth_text(1, 4, 131072, patterns=p1))   # prefix-stable in 64 KiB blocks
    c = synth.synth_text(0, 2, 200000)
    assert c.min() >= 0x0A and c.max() <= 0x7E and hashlib.md5(c.tobytes()).hexdigest() == \
        hashlib.md5(synth.synth_text(0, 2, 200000).tobytes()).hexdigest()
    # one planted pattern per 64 KiB block really is there
    o = Oracle(p1, 1, 256)
    pos, ids = o.scan(a)
    assert len(np.unique(pos // 65536)) == 5


def test_shard_plan_partitions_the_input():
    for n, g, mpl in ((0, 1, 5), (1, 4, 1), (65536, 2, 9), (10 ** 6 + 3, 3, 64), (2 ** 33 + 5, 8, 33), (100, 8, 4)):
        nxt = 0
        for i in range(g):
            start, ns, nv = pf.plan_shard(n, g, mpl, i)
            assert start == nxt and ns <= nv <= ns + mpl - 1 and start + nv <= n
            if start + ns < n:
                assert nv == min(ns + mpl - 1, n - start) and ns % 65536 == 0
            nxt = start + ns
        assert nxt == n
    with pytest.raises(pf.PfacError):
        pf.plan_shard(10, 2, 3, 2)


def test_sharded_oracle_scan_equals_whole_scan(fixtures):
    """The halo rule: shard results concatenated == one scan, incl. matches that
    straddle shard boundaries and the one ending exactly at input_size."""
    pats = fixtures["xab"] + b"England were\ncricket than England in the past four years.\n"
    o = Oracle(pats, 1, 256)
    data = np.frombuffer(fixtures["1M"][:-1], dtype=np.uint8)[:400000]
    pos, ids = o.scan(data)
    for g in (2, 3, 7):
        ps, is_ = [], []
        for i in range(g):
            start, ns, nv = pf.plan_shard(len(data), g, o.max_pat_len, i)
            p, d = o.scan(data[start:start + nv])
            keep = p < ns
            ps.append(p[keep] + start)
            is_.append(d[keep])
        assert np.array_equal(np.concatenate(ps), pos) and np.array_equal(np.concatenate(is_), ids)


def _build_c_example(tmp_path):
    import subprocess
    exe = tmp_path / "host_c_abi"
    build = os.path.join(ROOT, "phfpfac_b200", "_build")
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "host_c_abi.c"), "-L" + build, "-lpfac_b200",
                        "-Wl,-rpath," + build, "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_plain_c_and_example_links(tmp_path, fixtures):
    """include/pfac_b200.h compiles as C11 with no CUDA or C++ in sight, and the INTEGRATION.md host
    program links against the library; its table half runs without a GPU."""
    import subprocess
    exe = _build_c_example(tmp_path)
    pat = tmp_path / "pat"
    pat.write_bytes(fixtures["experimentpattern"])
    r = subprocess.run([str(exe), str(pat), "256"], capture_output=True, text=True)
    assert r.returncode == 0 and "4 patterns, max length 4" in r.stdout, (r.stdout, r.stderr)
    assert subprocess.run([str(exe), str(tmp_path / "nope"), "256"], capture_output=True).returncode == 1
    assert subprocess.run([str(exe)], capture_output=True).returncode == 255
