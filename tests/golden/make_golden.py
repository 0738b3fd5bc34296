#!/usr/bin/env python
"""Regenerates tests/golden/oracle_vectors.json.

PROVENANCE: the reference (Erlang) cannot run in the build container (no OTP), and its tests contain no byte-level
golden vectors, so these vectors are outputs of the ORACLE (oracle/, the CPU restatement), pinned here
so that (a) the oracle cannot drift silently and (b) the CUDA path is checked against committed bytes on the GPU box.
They are NOT reference outputs; the oracle itself is pinned only as described in oracle/src/driver.cpp."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import corpus  # noqa: E402
import oracle_lib as O  # noqa: E402

GPU_MUTS = ["uw", "ui", "num", "bd", "bei", "bed", "bf", "bi", "ber", "br", "sp", "sr", "sd", "snand", "srnd",
            "ld", "lds", "lr2", "lri", "lr", "ls", "lp", "lis", "lrs", "nil"]
DEFAULT_PRI = dict(zip(O.MUT_CODES, [10, 3, 1, 2, 1, 1, 1, 1, 3, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 2, 2, 7, 1, 1, 0]))

ALL_PATS = {"od": 1, "nd": 2, "bu": 1, "sk": 2, "sz": 2, "cs": 1, "ar": 1, "cp": 1, "co": 0, "nu": 0}
MARKUP = [b"<a>x</a>", b"<p>1<b>2</p>3</b>", b'<r a="1" b=\'2\' c>t<br/><!-- c --></r>', b"<?xml v?><d><e k=v>text one</e><e>text two</e></d>", b"pre<a",
          b"<html><body><p class='x'>Hello <b>bold</b> world</p><img src=\"http://h/i.png\"/></body></html>"]
JSON = [b'{"a":[1,true,null],"b":"str"}', b"[1,2,3]", b"hello", b'"abc', b" 17", b'{"k":{"n":-5,"s":"http://x/y","l":[]}}', b"true", b"null"]

CONFIGS = [
    # name, blobs, mutations, patterns, seed, n_cases, first_case
    ("c1_hello_erlamsa", [b"Hello erlamsa!\n"], {c: DEFAULT_PRI[c] for c in GPU_MUTS}, {"od": 1, "nd": 2, "bu": 1}, (1, 2, 3), 24, 1),
    ("byte_level", corpus.mixed_corpus(101, 12, 300), {c: 1 for c in ("bd", "bei", "bed", "bf", "bi", "ber", "br")}, {"od": 1}, (1, 2, 3), 24, 1),
    ("c3_mix_small", corpus.mixed_corpus(102, 12, 600), {c: 1 for c in ("bd", "bei", "bed", "bf", "bi", "ber", "br", "num")}, {"od": 1}, (1, 2, 3), 24, 5),
    ("lines", [corpus.text_lines(corpus.rng(103), 400) for _ in range(6)], {c: 1 for c in ("ld", "lds", "lr2", "lri", "lr", "ls", "lp", "lis", "lrs")},
     {"od": 1, "nd": 1}, (4, 5, 6), 18, 1),
    ("seq_and_utf8", corpus.mixed_corpus(104, 10, 200), {c: 1 for c in ("sp", "sr", "sd", "snand", "srnd", "uw", "ui")}, {"od": 1, "bu": 1}, (7, 8, 9), 20, 100),
    ("default_supported_mix", corpus.mixed_corpus(105, 16, 500), {c: DEFAULT_PRI[c] for c in GPU_MUTS}, {"od": 1, "nd": 2, "bu": 1}, (9, 9, 9), 32, 1),
    # the whole table at the reference's default priorities and all ten patterns (C1 of SURVEY.md 8d is the first one)
    ("c1_full_default", [b"Hello erlamsa!\n"], dict(DEFAULT_PRI), ALL_PATS, (1, 2, 3), 32, 1),
    ("full_default_mixed", corpus.mixed_corpus(106, 16, 400) + corpus.web_corpus(107, 16), dict(DEFAULT_PRI), ALL_PATS, (2, 2, 2), 48, 1),
    ("sgm_only", MARKUP, {"sgm": 1}, {"od": 1}, (3, 3, 3), 36, 1),
    ("js_only", JSON, {"js": 1}, {"od": 1}, (4, 4, 4), 40, 1),
    ("uri_b64", corpus.web_corpus(108, 20), {"uri": 1, "b64": 7}, {"od": 1, "nd": 1}, (5, 5, 5), 40, 1),
]
# configs whose every case must come out of the engine unflagged; for the others the GPU test compares the cases the
# engine did not flag (documents that need an sgm / js AST mutation are flagged)
DEVICE_COMPLETE = {"c1_hello_erlamsa", "byte_level", "c3_mix_small", "lines", "seq_and_utf8", "default_supported_mix"}


def main():
    out = []
    for name, blobs, muts, pats, seed, n, first in CONFIGS:
        outs, meta = O.fuzzer(blobs, mutations=muts, patterns=pats, seed=seed, n_cases=n, first_case=first, max_case_out=1 << 20)
        assert all(m.status == 0 for m in meta) or name not in DEVICE_COMPLETE, name
        out.append({"name": name, "device_complete": name in DEVICE_COMPLETE, "status": [m.status for m in meta], "blobs": [b.hex() for b in blobs], "mutations": muts, "patterns": pats, "seed": list(seed),
                    "n_cases": n, "first_case": first,
                    # small outputs verbatim, every output by length + sha256 (runaway repeats reach the 1 MiB cap)
                    "outputs": [o.hex() if len(o) <= 1024 else None for o in outs],
                    "digests": [[len(o), hashlib.sha256(o).hexdigest()] for o in outs],
                    "draws": [m.draws for m in meta], "pattern": [m.pattern for m in meta],
                    "used": [[u for u in m.used if u >= 0] for m in meta]})
    with open(os.path.join(HERE, "oracle_vectors.json"), "w") as f:
        json.dump({"provenance": "oracle (CPU restatement) outputs; see make_golden.py", "vectors": out}, f, indent=0)
    print("wrote", len(out), "configs,", sum(len(v["outputs"]) for v in out), "cases")


if __name__ == "__main__":
    main()
