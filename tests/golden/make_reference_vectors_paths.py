#!/usr/bin/env python
"""Regenerates tests/golden/reference_vectors_paths.json: the reference's FILE and STDIN generators
(fuzzer(#{paths => Files | ["-"], output => return, ...}), src/erlamsa_gen.erl:59-121), run from the reference's own source
by oracle/erlref with in-memory files. Includes BASELINE config C1: `echo 'Hello erlamsa!' | ./erlamsa --seed 1,2,3`.
Same provenance and purpose as make_reference_vectors.py."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import corpus  # noqa: E402

MUT_CODES = ["sgm", "js", "uw", "ui", "ab", "ad", "tr2", "td", "num", "ts1", "tr", "ts2", "bd", "bei", "bed", "bf", "bi", "ber", "br", "sp", "sr", "sd",
             "snand", "srnd", "ld", "lds", "lr2", "lri", "lr", "ls", "lp", "lis", "lrs", "ft", "fn", "fo", "len", "b64", "uri", "zip", "nil"]
DEF_PRI = [10, 3, 1, 2, 1, 1, 1, 1, 3, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 2, 2, 7, 1, 1, 0]
DEFAULT = dict(zip(MUT_CODES, DEF_PRI))
ALL_PATS = {"od": 1, "nd": 2, "bu": 1, "sk": 2, "sz": 2, "cs": 1, "ar": 1, "cp": 1, "co": 0, "nu": 0}


def main():
    sys.setrecursionlimit(3000000)
    from erlref.refrun import Reference
    ref = Reference(budget=200_000_000)
    files = [corpus.text_lines(corpus.rng(1200), 9000), corpus.numeric_text(corpus.rng(1201), 5000), corpus.random_bytes(corpus.rng(1202), 7000),
             corpus.structured_text(corpus.rng(1203), 3000), b"", b"ab", b"x" * 256, corpus.text_lines(corpus.rng(1204), 20000)]
    light = {"bd": 1, "bf": 1, "num": 2, "ld": 1, "lr2": 1, "sr": 1, "sd": 1, "fn": 1, "ui": 1}
    cfgs = [
        # name, files, stdin, mutations, patterns, generators, seed, cases
        ("c1_stdin_hello_erlamsa", None, b"Hello erlamsa!\n", None, None, None, (1, 2, 3), [1]),
        ("stdin_seeds", None, b"Hello erlamsa!\n", None, None, None, (4, 5, 6), [1]),
        ("stdin_text", None, files[0], light, {"od": 1, "nd": 1, "bu": 1}, None, (7, 7, 7), [1]),
        ("files_light_all_patterns", files, None, light, ALL_PATS, {"file": 1000, "random": 1}, (1, 2, 3), list(range(1, 41))),
        ("files_default", files, None, DEFAULT, ALL_PATS, {"file": 1000, "random": 1}, (3, 2, 1), list(range(1, 25))),
        ("files_blockscale", files, None, light, {"od": 1, "nd": 1}, {"file": 1000}, (9, 9, 9), list(range(1, 13)), {"blockscale": 0.25}),
    ]
    out = []
    for c in cfgs:
        name, fl, stdin, muts, pats, gens, seed, cases = c[:8]
        extra = c[8] if len(c) > 8 else {}
        rows = []
        for i in cases:
            rr = ref.case_paths(fl, i, seed, muts, pats, stdin=stdin, generators=gens, **extra)
            rows.append((rr.status, rr.output, rr.draws, rr.detail[:100]))
            if rr.status != "ok":
                print(name, i, rr.status, rr.detail[:100], flush=True)
        out.append({"name": name, "files": [f.hex() for f in fl] if fl is not None else None, "stdin": stdin.hex() if stdin is not None else None,
                    "mutations": muts, "patterns": pats, "generators": gens, "seed": list(seed), "cases": cases, "extra": extra,
                    "status": [r[0] for r in rows], "outputs": [r[1].hex() if len(r[1]) <= 2048 else None for r in rows],
                    "digests": [[len(r[1]), hashlib.sha256(r[1]).hexdigest()] for r in rows], "draws": [r[2] for r in rows]})
        print(name, "done", flush=True)
    with open(os.path.join(HERE, "reference_vectors_paths.json"), "w") as f:
        json.dump({"provenance": "erlamsa reference sources (commit 4a844bcd) executed by oracle/erlref; see make_reference_vectors_paths.py", "vectors": out}, f, indent=0)
    print("wrote", len(out), "configs,", sum(len(v["cases"]) for v in out), "cases")


if __name__ == "__main__":
    main()
