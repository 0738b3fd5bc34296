"""Development aid (run on the GPU box): the same batches through both builds of the general kernel (64 and 128 registers per
thread, EB200_WIDE); any case whose bytes or meta differ is printed with its pattern and mutators. The two builds run the same
sources, so a difference means the program depends on something the sources do not pin down (an ordering between lanes)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import erlamsa_b200  # noqa: E402
import corpus  # noqa: E402
from erlamsa_b200.options import mutator_codes, pattern_codes  # noqa: E402

MC, PC = mutator_codes(), pattern_codes()
os.environ["EB200_WIDE"] = "0"
narrow = erlamsa_b200.Engine(0)
os.environ["EB200_WIDE"] = "1"
wide = erlamsa_b200.Engine(0)
# the 64-register build with the wide build's launch geometry (512 threads, 12 deciding warps): tells apart "depends on the
# geometry" (temp-slot mapping, worker count) from "depends on the code generation"
os.environ["EB200_WIDE"] = "0"; os.environ["EB200_THREADS"] = "512"; os.environ["EB200_DECIDERS"] = "12"
narrow512 = erlamsa_b200.Engine(0)
del os.environ["EB200_THREADS"], os.environ["EB200_DECIDERS"]
n_each = int(sys.argv[1]) if len(sys.argv) > 1 else 300
table = {c: p for c, p in erlamsa_b200.default_mutations() if c in erlamsa_b200.supported_mutations()}
configs = [
    ("uri od", corpus.web_corpus(0xE21A0900, 300), {"mutations": {"uri": 1}, "patterns": {"od": 1}, "seed": (5, 5, 5)}),
    ("uri+bd nd/bu", corpus.web_corpus(0xE21A0900, 300), {"mutations": {"uri": 1, "bd": 1}, "patterns": {"nd": 1, "bu": 1}, "seed": (5, 6, 5)}),
    ("table nd", corpus.mixed_corpus(0xE21A0100, 240), {"mutations": table, "patterns": {"nd": 1}, "seed": (2, 7, 1)}),
]
for m in ("b64", "ab", "sr", "num", "ft", "sgm", "js", "snand", "lis", "tr"):
    configs.append((m + " nd", corpus.web_corpus(0xE21A0900, 120) + corpus.mixed_corpus(0xE21A0100, 120), {"mutations": {m: 1}, "patterns": {"nd": 1}, "seed": (2, 7, 1)}))
for name, blobs, opts in configs:
    n = min(n_each, len(blobs))
    a, ma = narrow.fuzz_batch(blobs, dict(opts), n_cases=n)
    b, mb = wide.fuzz_batch(blobs, dict(opts), n_cases=n)
    c3, mc3 = narrow512.fuzz_batch(blobs, dict(opts), n_cases=n)
    a2, ma2 = narrow.fuzz_batch(blobs, dict(opts), n_cases=n)
    print("%-14s narrow vs narrow@512/12: %3d differ | narrow vs itself again: %3d differ | wide vs narrow@512/12: %3d differ" % (
        name, sum(1 for k in range(n) if a[k] != c3[k] or ma[k].draws != mc3[k].draws), sum(1 for k in range(n) if a[k] != a2[k] or ma[k].draws != ma2[k].draws),
        sum(1 for k in range(n) if b[k] != c3[k] or mb[k].draws != mc3[k].draws)))
    bad = [k for k in range(n) if a[k] != b[k] or ma[k].draws != mb[k].draws or ma[k].status != mb[k].status]
    print("%-14s %4d cases, %3d differ" % (name, n, len(bad)))
    for k in bad[:2]:
        x, y = ma[k], mb[k]
        print("   case %4d in %5d narrow: len %6d draws %6d st %d/%d pat %s used %s | wide: len %6d draws %6d st %d/%d used %s" % (
            k, len(blobs[k % len(blobs)]), len(a[k]), x.draws, x.status, x.pad, PC[x.pattern] if 0 <= x.pattern < len(PC) else "?",
            ",".join(MC[u] for u in x.used if u >= 0), len(b[k]), y.draws, y.status, y.pad, ",".join(MC[u] for u in y.used if u >= 0)))
        if a[k] != b[k]:
            i = next((i for i in range(min(len(a[k]), len(b[k]))) if a[k][i] != b[k][i]), min(len(a[k]), len(b[k])))
            print("        first difference at byte %d: %r | %r" % (i, a[k][max(0, i - 8):i + 24], b[k][max(0, i - 8):i + 24]))
