#!/usr/bin/env python
"""Generates tests/golden/* from the reference tree (run in the build container only).

    python tests/golden/make_golden.py

What is written, and where each number comes from:

* fixtures.json / dictionary.txt.gz -- the reference's shipped DATA fixtures needed on the GPU
  box, where /root/reference does not exist: experimentpattern (14 B), experimentinput (26 B),
  the 402-byte period of `1M` (+ the sha256 of the full file) and the English dictionary
  xaa+xab+xac+xad used as a pattern set.  No reference SOURCE
  is copied.
* golden.json
    kat_counts   : state/key counts printed by the reference's own old runs
                   (experiment/xaarecord:2-6 ..., tmp.dat:2-8) -- parsed from those files.
    ref_tables   : sha256 of s0Table / r / HT / val / patternIdMap produced by the reference's
                   OWN create_PFAC_table_reorder + FFDM, compiled from /root/reference by
                   oracle/Makefile (oracle/_ref/libphfpfac_ref.so).
    results      : GPU_match_result.txt images (md5, line count; the full text for the small
                   case) obtained by walking the REFERENCE-BUILT tables with the oracle's
                   restatement of SUBSEG_MATCH and merging partitions like main.cc:304-324.
    format_lines : first lines of experiment/GPU_match_resultall.txt (pins main.cc:344's format).
"""
import gzip
import hashlib
import json
import os
import re
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _oracle import RefBuild, build_oracle, render_result, scan_tables_cpu  # noqa: E402

REF = "/root/reference/regex_GPU_PHF"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.int32).tobytes()).hexdigest()


def table_digest(p):
    return {"state_num": int(p.state_num), "n_final": int(p.n_final), "max_len": int(p.max_len),
            "ht_size": int(p.ht_size), "n_r": int(len(p.r)), "s0": sha(p.s0), "r": sha(p.r), "HT": sha(p.HT),
            "val": sha(p.val), "idmap": sha(p.idmap)}


def ref_scan(rb, data):
    """All partitions over `data`, merged in main.cc:304-324 order (position, partition, depth)."""
    buf = np.frombuffer(data, dtype=np.uint8)
    pos_all, id_all = [], []
    for g in range(rb.n_parts):
        p = rb.part(g)
        pos, ids = scan_tables_cpu(p, p.idmap, p.max_len, buf, nthreads=4, ref_tile_bound=True)
        pos_all.append(pos)
        id_all.append(ids)
    pos = np.concatenate(pos_all)
    ids = np.concatenate(id_all)
    order = np.argsort(pos, kind="stable")
    return pos[order], ids[order]


def parse_record(path):
    out = {}
    keys = {"state num": "state_num", "final state num": "n_final", "max pattern length": "max_len",
            "Number of keys": "n_keys", "Max Key": "max_key", "width value": "width", "r table size": "r_size"}
    for line in open(path, errors="replace"):
        m = re.match(r"\s*([A-Za-z ]+?)\s*:\s*(\d+)", line)
        if m and m.group(1) in keys and keys[m.group(1)] not in out:
            out[keys[m.group(1)]] = int(m.group(2))
    return out


def main():
    build_oracle()
    rd = lambda name: open(os.path.join(REF, name), "rb").read()
    one_m = rd("1M")
    period = one_m[:402]
    assert (period * (len(one_m) // 402 + 1))[:len(one_m)] == one_m
    dictionary = b"".join(rd(x) for x in ("xaa", "xab", "xac", "xad"))
    fixtures = {
        "experimentpattern_hex": rd("experimentpattern").hex(),
        "experimentinput_hex": rd("experimentinput").hex(),
        "one_m_period_hex": period.hex(),
        "one_m_size": len(one_m),
        "one_m_sha256": hashlib.sha256(one_m).hexdigest(),
        "dictionary_sha256": hashlib.sha256(dictionary).hexdigest(),
        "dictionary_parts": {x: len(rd(x)) for x in ("xaa", "xab", "xac", "xad")},
    }
    json.dump(fixtures, open(os.path.join(HERE, "fixtures.json"), "w"), indent=1)
    with open(os.path.join(HERE, "dictionary.txt.gz"), "wb") as f:
        f.write(gzip.compress(dictionary, 9, mtime=0))

    golden = {"kat_counts": {}, "ref_tables": {}, "results": {}}
    for name, rec in (("xaa", "experiment/xaarecord"), ("xab", "experiment/xabrecord"),
                      ("xac", "experiment/xacrecord"), ("xad", "experiment/xadrecord"),
                      ("dictionary", "experiment/englishdicall"), ("experimentpattern", "tmp.dat")):
        golden["kat_counts"][name] = dict(parse_record(os.path.join(REF, rec)), source=rec)

    with tempfile.TemporaryDirectory() as td:
        dpath = os.path.join(td, "dict")
        open(dpath, "wb").write(dictionary)
        cases = [
            ("experimentpattern", os.path.join(REF, "experimentpattern"), 1, 256, False),
            ("experimentpattern", os.path.join(REF, "experimentpattern"), 1, 1024, True),
            ("experimentpattern", os.path.join(REF, "experimentpattern"), 2, 64, False),
            ("xaa", os.path.join(REF, "xaa"), 1, 4096, True),
            ("xad", os.path.join(REF, "xad"), 1, 256, False),
            ("dictionary", dpath, 1, 256, False),
            ("dictionary", dpath, 1, 4096, True),
            ("dictionary", dpath, 2, 64, False),
            ("dictionary", dpath, 1, 256, True),
        ]
        builds = {}
        for name, path, streamnum, width, single in cases:
            rb = RefBuild(path, streamnum=streamnum, width=width, single=single)
            key = f"{name}|parts={'1' if single else 4 * streamnum}|width={width}"
            golden["ref_tables"][key] = {"max_pat_len": int(rb.max_pat_len),
                                         "parts": [table_digest(rb.part(g)) for g in range(rb.n_parts)]}
            builds[key] = rb
        runs = [
            ("experimentpattern_x_experimentinput", "experimentpattern|parts=4|width=256", rd("experimentinput")[:-1], True),
            ("experimentpattern_x_1M", "experimentpattern|parts=4|width=256", one_m[:-1], False),
            ("dictionary_x_1M", "dictionary|parts=4|width=256", one_m[:-1], False),
            ("dictionary_x_1M_single_w4096", "dictionary|parts=1|width=4096", one_m[:-1], False),
            ("dictionary_x_1M_8parts_w64", "dictionary|parts=8|width=64", one_m[:-1], False),
            ("xaa_x_1M_first64k", "xaa|parts=1|width=4096", one_m[:65536], False),
        ]
        for rname, key, data, keep_text in runs:
            pos, ids = ref_scan(builds[key], data)
            text = render_result(pos, ids)
            golden["results"][rname] = {"tables": key, "input_size": len(data), "lines": int(len(pos)),
                                        "md5": hashlib.md5(text).hexdigest()}
            if keep_text:
                golden["results"][rname]["text"] = text.decode()
    golden["format_lines"] = open(os.path.join(REF, "experiment/GPU_match_resultall.txt")).read().splitlines()[:5]
    json.dump(golden, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(golden["results"], indent=1))


if __name__ == "__main__":
    main()
