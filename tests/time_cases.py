"""Where does the time of a batch go? (not a test; run on the GPU box)
usage: EB200_CASE_TIMES=1 python tests/time_cases.py c2|c4|c5|smoke [n_cases]
Runs one batch through the host path with per-case timing of the general program (eb200_debug_case_times) and prints the
total by first used mutator / pattern and the slowest cases."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("EB200_CASE_TIMES", "1")
import erlamsa_b200  # noqa: E402
from erlamsa_b200 import _native as N  # noqa: E402
import corpus  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
codes = erlamsa_b200.mutator_codes() if hasattr(erlamsa_b200, "mutator_codes") else None
from erlamsa_b200.options import mutator_codes, pattern_codes, default_mutations, default_patterns  # noqa: E402
MC, PC = mutator_codes(), pattern_codes()
if which == "c2":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    blobs = corpus.uniform_corpus(0xE21A0002, n, 4096, "bin")
    opts = {"seed": (1, 2, 3), "max_case_out": 128 << 10}
elif which == "c4":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    blobs = corpus.uniform_corpus(0xE21A0004, min(n, 64), 262144, "markup")
    opts = {"seed": (1, 2, 3), "mutations": {c: 1 for c in ("ab", "ad", "tr2", "td", "ts1", "ts2", "tr", "sgm", "js")}, "patterns": {"od": 1}}
elif which == "c5":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    r5 = corpus.rng(0xE21A0005)
    blobs = [corpus.structured_text(r5, 4096).ljust(4096, b" ")[:4096] for _ in range(256)]
    opts = {"seed": (1, 2, 3), "mutations": {"ft": 1, "fn": 1, "fo": 1}, "patterns": {"od": 1}}
else:
    n = 256
    blobs = corpus.mixed_corpus(0x5A0CE, 256, max_len=3000)
    opts = {"seed": (1, 2, 3), "patterns": {"od": 1, "nd": 2, "bu": 1}, "max_case_out": 1 << 20}
eng = erlamsa_b200.Engine(0)
for rep in range(2):
    t0 = time.time()
    outs, meta = eng.fuzz_batch(blobs, opts, n_cases=n)
    wall = time.time() - t0
st = eng.last_stats
us = (C.c_uint32 * n)()
N.lib().eb200_debug_case_times.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_uint64]
N.lib().eb200_debug_case_times.restype = C.c_uint64
got = N.lib().eb200_debug_case_times(eng._ctx, us, n)
print("%s: %d cases, wall %.1f ms, kernel %.1f ms, launches %d, flagged u/d/o %d/%d/%d, timed entries %d" % (
    which, n, wall * 1e3, st.ms_decide, st.kernels_launched, st.n_unsupported, st.n_died, st.n_overflow, got))
tot = sum(us)
print("sum of per-case general-program time: %.1f ms (kernel wall %.1f ms => ~%.1f warps busy)" % (tot / 1e3, st.ms_decide, tot / 1e3 / max(st.ms_decide, 1e-9)))
by = {}
for k in range(n):
    m = meta[k]
    used = [MC[u] for u in m.used if u >= 0]
    key = (PC[m.pattern] if 0 <= m.pattern < len(PC) else "?") + ":" + (",".join(used[:3]) if used else "-")
    a = by.setdefault(key, [0, 0, 0])
    a[0] += us[k]; a[1] += 1; a[2] = max(a[2], us[k])
print("top (pattern:first mutators) by total time:")
for key, a in sorted(by.items(), key=lambda kv: -kv[1][0])[:25]:
    print("  %-40s total %9.1f ms  cases %5d  max %8.1f ms" % (key, a[0] / 1e3, a[1], a[2] / 1e3))
print("slowest cases:")
for k in sorted(range(n), key=lambda k: -us[k])[:25]:
    m = meta[k]
    print("  case %6d  %9.2f ms  pat %-3s used %-40s fails %3d in %7d out %8d status %d/%d draws %d" % (
        k, us[k] / 1e3, PC[m.pattern] if 0 <= m.pattern < len(PC) else "?", ",".join(MC[u] for u in m.used if u >= 0), m.n_failed,
        len(blobs[k % len(blobs)]), len(outs[k]), m.status, m.pad, m.draws))
mt = (C.c_uint64 * 98)()
N.lib().eb200_debug_mutator_times.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
if N.lib().eb200_debug_mutator_times(eng._ctx, mt):
    print("time inside each mutator (all attempts, successful or not):")
    for i in sorted(range(41), key=lambda i: -mt[2 * i]):
        if mt[2 * i + 1]:
            print("  %-6s total %10.1f ms  calls %7d  mean %9.1f us" % (MC[i], mt[2 * i] / 1e6, mt[2 * i + 1], mt[2 * i] / 1e3 / mt[2 * i + 1]))
    print("fuse search by phase / document mutators by phase:")
    for j, name in enumerate(("class tables (> 128 per side)", "registers (33..128)", "packed (<= 32 per side)", "one suffix per side", "flat levels", "sgm: tokenize", "sgm: pair tags", "js: tokenize")):
        if mt[82 + 2 * j + 1]:
            print("  %-26s total %10.1f ms  steps %8d  mean %9.2f us" % (name, mt[82 + 2 * j] / 1e6, mt[82 + 2 * j + 1], mt[82 + 2 * j] / 1e3 / mt[82 + 2 * j + 1]))
# time by mutator TRIED is not recorded; failures dominate when `used` is short and n_failed large
