"""Share of cases the engine flags instead of mutating, per corpus, with the reference's default mutators and patterns
(not a test; run on the GPU box): python tests/flag_rates.py"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import corpus  # noqa: E402
import erlamsa_b200  # noqa: E402

eng = erlamsa_b200.Engine(0)
muts = dict(erlamsa_b200.default_mutations())
pats = dict(erlamsa_b200.default_patterns())
sets = {
    "random bytes 4 KiB (C2)": corpus.uniform_corpus(1, 2000, 4096, "bin"),
    "numeric text 4 KiB": corpus.uniform_corpus(2, 2000, 4096, "num"),
    "mixed (bin/num/lines/tiny)": corpus.mixed_corpus(3, 2000, 3000),
    "structured text": corpus.text_corpus(4, 2000, 1500),
    "web (urls/base64/json/markup)": corpus.web_corpus(5, 2000),
}
for name, blobs in sets.items():
    outs, meta = eng.fuzz_batch(blobs, {"mutations": muts, "patterns": pats, "seed": (1, 2, 3), "max_case_out": 1 << 20, "scratch_bytes": 6 << 30}, n_cases=len(blobs))
    c = collections.Counter(m.status for m in meta)
    print("%-32s ok %5.1f%%  flagged-unsupported %5.1f%%  died %4.1f%%  over-cap %4.1f%%" % (
        name, 100.0 * c[0] / len(blobs), 100.0 * c[1] / len(blobs), 100.0 * c[2] / len(blobs), 100.0 * c[3] / len(blobs)), flush=True)
