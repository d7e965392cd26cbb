#!/usr/bin/env python
"""Host model of the detector's stage 1 (numpy, no GPU): what DESIGN.md 3.3 / 8.1 says about probe strides,
window counts, index folding and bank-split table copies, reproducible.

    python tools/host_model.py [--workload config3] [--mib 16]

For a stride S the windows (2 input bytes) at offsets = 0 mod S of the stream are looked up in a table of bit
planes; a start at phase f = start mod S sees its bytes o_j, o_j + 1 with o_j = (S - f) mod S + j S.  It passes
when some pattern agrees with every window it covers:

    pass = W_0(w_0) and (E_0(w_0) or (W_1(w_1) and (E_1(w_1) or W_2(w_2) ...)))

  W_j : the pair is bytes o_j, o_j + 1 of some pattern (path of the trie)
  E_j : the pair is bytes o_j, o_j + 1 of a pattern that ends before it covers window j + 1

S = 2 with three windows is the built design (its ShX plane is shared by both phases; `built` below restates it
bit for bit and is checked against the C++ model behind pfac_tables_filter_profile).  The bank model counts, per
shared-memory load of a warp, the largest number of distinct 32-bit words that fall into one of the 32 banks.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import pfac_synth as synth  # noqa: E402
from bench import WORKLOADS  # noqa: E402


def pair_planes(pats, offsets):
    """W[j], E[j] (bool[65536], index = b0 | b1 << 8) for the windows at byte offsets `offsets` of a start."""
    W = [np.zeros(65536, dtype=bool) for _ in offsets]
    E = [np.zeros(65536, dtype=bool) for _ in offsets]
    for p in pats:
        n = len(p)
        for j, o in enumerate(offsets):
            if n >= o + 2:
                idx = p[o] | (p[o + 1] << 8)
                W[j][idx] = True
                nxt = offsets[j + 1] if j + 1 < len(offsets) else None
                if nxt is not None and n < nxt + 2:
                    E[j][idx] = True
    return W, E


def survivors(text, pats, stride, windows, fold=None):
    """Starts that pass stage 1 (count) with separate planes per (phase, window).  fold: maps a pair index to the
    table index (planes are OR-ed over the pairs that share an index)."""
    n = len(text) - 16
    pair = text[:-1].astype(np.uint32) | (text[1:].astype(np.uint32) << 8)
    total = 0
    min_len = min(len(p) for p in pats)
    for phase in range(stride):
        offs = [((stride - phase) % stride) + j * stride for j in range(windows)]
        assert offs[0] + 2 <= min_len, "the first window must lie inside every pattern"
        W, E = pair_planes(pats, offs)
        if fold is not None:
            f = fold(np.arange(65536, dtype=np.uint32))
            for planes in (W, E):
                for k, pl in enumerate(planes):
                    t = np.zeros(int(f.max()) + 1, dtype=bool)
                    np.logical_or.at(t, f, pl)
                    planes[k] = t[f]          # back to pair space: a pair passes if its folded entry does
        starts = np.arange(phase, n, stride)
        ok = None
        for j in reversed(range(windows)):
            w = pair[starts + offs[j]]
            here = W[j][w]
            if ok is None:
                ok = here
            else:
                ok = here & (E[j][w] | ok)
        total += int(ok.sum())
    return total, n


def built(text, pats, w3):
    """The built mode-0 rule (pfac_derive.cc stage1_pass) from the pattern list: even / odd starts, ShX shared."""
    n = len(text) - 16
    pair = text[:-1].astype(np.uint32) | (text[1:].astype(np.uint32) << 8)
    P = {k: np.zeros(65536, dtype=bool) for k in ("01", "12", "23", "34", "45", "56", "shortc", "shx")}
    for p in pats:
        L = len(p)
        for k, o in (("01", 0), ("12", 1), ("23", 2), ("34", 3), ("45", 4), ("56", 5)):
            if L >= o + 2:
                P[k][p[o] | (p[o + 1] << 8)] = True
        if L == 4:
            P["shortc"][p[1] | (p[2] << 8)] = True
        if L in (4, 5):
            P["shx"][p[2] | (p[3] << 8)] = True
        if L in (5, 6):
            P["shx"][p[3] | (p[4] << 8)] = True
    ev = np.arange(0, n, 2)
    od = np.arange(1, n, 2)
    if w3:
        e = P["01"][pair[ev]] & P["23"][pair[ev + 2]] & (P["shx"][pair[ev + 2]] | P["45"][pair[ev + 4]])
        o = P["shortc"][pair[od + 1]] | (P["12"][pair[od + 1]] & P["34"][pair[od + 3]] &
                                         (P["shx"][pair[od + 3]] | P["56"][pair[od + 5]]))
    else:
        e = P["01"][pair[ev]] & P["23"][pair[ev + 2]]
        o = P["shortc"][pair[od + 1]] | (P["12"][pair[od + 1]] & P["34"][pair[od + 3]])
    return int(e.sum()) + int(o.sum()), n


def rot2(c):
    return ((c << 2) | (c >> 6)) & 0xFF


def wavefronts(text, stride, lane_bytes, entry_bytes, index, copies=1, sample=200000):
    """Mean over warps of max-over-banks #distinct 32-bit words per load instruction.  Lane l of a warp probes the
    window at base + l * lane_bytes + k * stride (k-th load of the warp); `copies` bank-split replicas: lane l uses
    replica l % copies, which lives in its own 32 / copies banks."""
    n = len(text) - 2
    per_warp = 32 * lane_bytes
    n_warps = min(sample, n // per_warp - 1)
    rng = np.random.default_rng(1)
    bases = rng.integers(0, n // per_warp - 1, n_warps) * per_warp
    loads = lane_bytes // stride
    tot = 0.0
    lanes = np.arange(32)
    banks_per_copy = 32 // copies
    for k in range(loads):
        pos = bases[:, None] + lanes[None, :] * lane_bytes + k * stride
        x, y = text[pos].astype(np.uint32), text[pos + 1].astype(np.uint32)
        addr = index(x, y) * entry_bytes
        word = addr >> 2
        bank = (word % banks_per_copy) + (lanes[None, :] % copies) * banks_per_copy
        # distinct (bank, word) pairs per warp, then the fullest bank
        key = bank.astype(np.int64) * (1 << 32) + word + (lanes[None, :] % copies).astype(np.int64) * (1 << 24) * 0
        key.sort(axis=1)
        first = np.ones_like(key, dtype=bool)
        first[:, 1:] = key[:, 1:] != key[:, :-1]
        b = (key >> 32).astype(np.int64)
        cnt = np.zeros((n_warps, 32), dtype=np.int32)
        rows = np.repeat(np.arange(n_warps), 32).reshape(n_warps, 32)
        np.add.at(cnt, (rows[first], b[first]), 1)
        tot += cnt.max(axis=1).mean()
    return tot / loads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config3")
    ap.add_argument("--mib", type=int, default=16)
    a = ap.parse_args()
    pk, cnt, pseed, lo, hi, tk, tseed, _nbytes, desc = WORKLOADS[a.workload]
    blob = synth.synth_patterns(pk, cnt, pseed, lo, hi)
    pats = [np.frombuffer(x, dtype=np.uint8).astype(np.uint32) for x in blob.split(b"\n")[:-1]]
    text = synth.synth_text(tk, tseed, a.mib << 20, patterns=blob)
    print(f"{desc}: {len(pats)} patterns, lengths {min(map(len, pats))}..{max(map(len, pats))}, {a.mib} MiB of text")

    # ---- the built design, checked against the C++ host model of the shipped tables
    import phfpfac_b200 as pf
    t = pf.Tables.from_bytes(blob)
    n16 = len(text) - 16
    cpp = t.filter_profile(text[:n16 + 16])
    s3, n = built(text, pats, True)
    s2, _ = built(text, pats, False)
    print(f"built design (stride 2): two windows {s2 / n:.4%} of the starts survive stage 1, three windows {s3 / n:.4%}"
          f"   [C++ model over the same text: {cpp['t1_pass'] / cpp['positions']:.4%}]")

    # ---- strides and window counts with ideal (separate) planes
    min_len = min(map(len, pats))
    print("\nstride  windows  planes  probes per 128 B  stage-1 survivors")
    for stride, windows in ((2, 2), (2, 3), (3, 2), (3, 3), (4, 2), (6, 2), (6, 3)):
        if stride > min_len - 1:
            continue            # a stride of S needs S + 1 <= the shortest pattern (the first window must be inside)
        s, n = survivors(text, pats, stride, windows)
        print(f"{stride:6d}  {windows:7d}  {stride * (2 * windows - 1):6d}  {128 / stride:16.1f}  {s / n:.4%}")

    def fold15(i):
        return (i ^ (i >> 15)) & 0x7FFF
    if min_len >= 4:
        s, n = survivors(text, pats, 3, 3, fold=fold15)
        print(f"     3        3  (15-bit folded index, 16-bit entries: 64 KiB)          {s / n:.4%}")

    # ---- bank model of the T1 gather
    idx8 = lambda x, y: rot2(x) | (rot2(y) << 8)                       # noqa: E731  (the built index)
    idx15 = lambda x, y: fold15(rot2(x) | (rot2(y) << 8))              # noqa: E731
    print("\nshared-memory wavefronts per T1 load of a warp (1.0 = conflict-free)")
    w_built = wavefronts(text, 2, 16, 1, idx8)
    print(f"  built: u8 entries, 64 KiB, stride 2, 16 bytes per lane           {w_built:.2f}   -> {64 / 32 * w_built:.1f} per 128 input bytes")
    w_plain = wavefronts(text, 2, 16, 1, lambda x, y: x | (y << 8))
    print(f"  the same without the rotate-by-2 of the index                    {w_plain:.2f}")
    w_split = wavefronts(text, 2, 16, 1, idx8, copies=2)
    print(f"  two bank-split copies (128 KiB)                                  {w_split:.2f}   -> {64 / 32 * w_split:.1f} per 128 input bytes")
    if min_len >= 4:
        w3 = wavefronts(text, 3, 24, 2, idx15)
        print(f"  stride 3, 16-bit entries, folded index (64 KiB)                  {w3:.2f}   -> {128 / 3 / 32 * w3:.1f} per 128 input bytes")
        w3s = wavefronts(text, 3, 24, 2, idx15, copies=2)
        print(f"  the same in two bank-split copies (128 KiB)                      {w3s:.2f}   -> {128 / 3 / 32 * w3s:.1f} per 128 input bytes")
    rng = np.random.default_rng(2)
    rnd = rng.integers(0, 256, len(text), dtype=np.uint8)
    print(f"  (32 uniformly random addresses: {wavefronts(rnd, 2, 16, 1, lambda x, y: x | (y << 8)):.2f})")


if __name__ == "__main__":
    main()
