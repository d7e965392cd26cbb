#!/usr/bin/env python
"""SASS evidence of the built library: per kernel, how many of the instructions that carry the design are in
the code (TMA bulk copies, mbarrier operations, shared-memory loads by width, shuffles / votes, global
loads / stores / atomics, programmatic-dependent-launch control).  Needs cuobjdump, no GPU.
usage: sass_summary.py [LIB] > profiles/r2_sass_summary.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "phfpfac_b200/_build/libpfac_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
pats = [("UBLKCP (cp.async.bulk: TMA bulk copy)", r"\bUBLKCP"), ("SYNCS (mbarrier arrive / try_wait / expect_tx)", r"\bSYNCS"),
        ("LDS.U8", r"\bLDS\.U8"), ("LDS.U16", r"\bLDS\.U16"), ("LDS (32 bit)", r"\bLDS (?!\.)|\bLDS R"), ("LDS.64", r"\bLDS\.64"),
        ("LDS.128", r"\bLDS\.128"), ("STS*", r"\bSTS"), ("ATOMS*", r"\bATOMS"), ("SHFL*", r"\bSHFL"), ("VOTE*/MATCH*", r"\bVOTE|\bMATCH"),
        ("REDUX*", r"\bREDUX"), ("LDG*", r"\bLDG"), ("STG*", r"\bSTG"), ("ATOMG*/RED*", r"\bATOMG|\bRED\b|\bRED\."),
        ("SHF* (funnel shifts)", r"\bSHF"), ("LOP3*", r"\bLOP3"), ("IMAD*", r"\bIMAD"), ("PRMT", r"\bPRMT"),
        ("ACQBULK / griddepcontrol (PDL)", r"ACQBULK|DEPBAR\.DEP|GRIDDEP|PREEXIT"), ("NANOSLEEP", r"\bNANOSLEEP"), ("BAR*", r"\bBAR\.")]
arch = None
fn = None
counts = collections.OrderedDict()
total = collections.Counter()
for line in out.splitlines():
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    if fn and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        ins = line.split("*/", 1)[1]
        total[fn] += 1
        for name, rx in pats:
            if re.search(rx, ins):
                counts[fn][name] += 1
print(f"{lib}: arch {arch}; instruction counts in the SASS of each kernel (static, not executed counts)\n")
for fn, c in counts.items():
    dem = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip() or fn
    print(f"{dem}\n  {total[fn]} instructions")
    for name, _ in pats:
        if c[name]:
            print(f"  {c[name]:6d}  {name}")
    print()
