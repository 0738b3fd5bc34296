#!/usr/bin/env python
"""Static evidence about the built library (no GPU needed): per kernel, the registers / shared memory / stack / spills cuobjdump reports
and counts of the SASS mnemonics that matter for this path -- 16-byte global loads and stores, bulk-async (TMA) copies and mbarrier
waits, warp shuffles / votes / match, local-memory traffic, tensor-core instructions (there must be none).
usage: python tools/sass_report.py [path/to/liberlamsa_b200.so] > profiles/sass_r2.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "erlamsa_b200", "liberlamsa_b200.so")
CUOBJDUMP = os.environ.get("CUOBJDUMP", "/usr/local/cuda/bin/cuobjdump")
GROUPS = [
    ("LDG.E.128", r"\bLDG\.E(\.[A-Z0-9_]+)*\.128"), ("STG.E.128", r"\bSTG\.E(\.[A-Z0-9_]+)*\.128"), ("LDG (any)", r"\bLDG\."), ("STG (any)", r"\bSTG\."),
    ("UBLKCP (cp.async.bulk)", r"\bUBLKCP"), ("UTMALDG/UTMASTG (tensor-map TMA)", r"\bUTMA(LDG|STG)"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("LDGSTS (cp.async)", r"\bLDGSTS"),
    ("SHFL", r"\bSHFL\."), ("VOTE", r"\bVOTE"), ("MATCH", r"\bMATCH"), ("REDUX", r"\bREDUX"), ("ATOMS/ATOMG/RED", r"\b(ATOMS|ATOMG|RED)\b|\bATOM[SG]\."),
    ("LDL (local load)", r"\bLDL"), ("STL (local store)", r"\bSTL"), ("BAR (CTA barrier)", r"\bBAR\."), ("WARPSYNC", r"\bWARPSYNC"),
    ("DFMA/DMUL/DADD (FP64: AS183 divisions)", r"\bD(FMA|MUL|ADD)\b"), ("HMMA/IMMA/UTCMMA/tcgen05 (tensor cores)", r"\b(HMMA|IMMA|DMMA|UTC[A-Z]*MMA|QGMMA|HGMMA)"),
]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


def main():
    res = subprocess.run([CUOBJDUMP, "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for ln in res.splitlines():
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in ln:
            usage[cur] = ln.strip()
            cur = None
    sass = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True).stdout
    counts, total = collections.OrderedDict(), {}
    cur = None
    pats = [(n, re.compile(p)) for n, p in GROUPS]
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter(); total[cur] = 0
            continue
        if cur is None or "/*" not in ln:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if not m:
            continue
        ins = m.group(1)
        total[cur] += 1
        for n, p in pats:
            if p.search(ins):
                counts[cur][n] += 1
    print("# static report of %s (cuobjdump -res-usage / -sass), sm_100a" % os.path.relpath(LIB, ROOT))
    for fn in counts:
        short = demangle(fn)
        short = re.sub(r"\(.*", "", short)
        print("\n== %s\n   %s\n   SASS instructions: %d" % (short, usage.get(fn, "(no resource line)"), total[fn]))
        for n, _ in GROUPS:
            if counts[fn][n]:
                print("   %-46s %d" % (n, counts[fn][n]))
    tc = sum(c["HMMA/IMMA/UTCMMA/tcgen05 (tensor cores)"] for c in counts.values())
    print("\ntensor-core instructions in the library: %d (byte shuffling and integer / FP64 scalar arithmetic only)" % tc)
    print("bulk-async copy instructions (UBLKCP) in the library: %d  -- the EB200_TMA_WORKERS variant of the copy workers (A/B in variants_r2.txt section 6)"
          % sum(c["UBLKCP (cp.async.bulk)"] for c in counts.values()))


if __name__ == "__main__":
    main()
