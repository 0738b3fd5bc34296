#!/usr/bin/env python
"""Regenerates tests/golden/reference_vectors_jump.json: the reference's JUMP generator (fuzzer(#{paths => [F1, F2, ...], generators =>
[{jump, 100}, ...], output => return}), src/erlamsa_gen.erl:123-150), run from the reference's own source by oracle/erlref with in-memory
files. Same provenance and purpose as make_reference_vectors_paths.py. The oracle restates this generator (oracle/src/driver.cpp
jump_block); the engine does not implement it, so these vectors pin the oracle only."""
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
    files = [corpus.text_lines(corpus.rng(1300), 9000), corpus.numeric_text(corpus.rng(1301), 5000), corpus.random_bytes(corpus.rng(1302), 7000),
             corpus.structured_text(corpus.rng(1303), 3000), b"x" * 256, corpus.text_lines(corpus.rng(1304), 20000)]
    edge = [b"", b"ab", corpus.text_lines(corpus.rng(1305), 600)]      # an empty file: finish(0) may leave an empty block list -> size([]) kills the worker
    light = {"bd": 1, "bf": 1, "num": 2, "ld": 1, "lr2": 1, "sr": 1, "sd": 1, "fn": 1, "ui": 1}
    cfgs = [
        # name, files, mutations, patterns, generators, seed, cases, extra
        ("jump_light_all_patterns", files, light, ALL_PATS, {"jump": 100}, (1, 2, 3), list(range(1, 41)), {}),
        ("jump_default_table", files, DEFAULT, ALL_PATS, {"jump": 100}, (3, 2, 1), list(range(1, 25)), {}),
        ("jump_among_default_generators", files, light, {"od": 1, "nd": 1, "bu": 1}, {"random": 1, "jump": 100, "file": 1000}, (5, 6, 7), list(range(1, 41)), {}),
        ("jump_edge_files", edge, light, {"od": 1, "nd": 1, "sk": 1, "nu": 1}, {"jump": 100}, (9, 9, 9), list(range(1, 41)), {}),
        ("jump_blockscale", files, light, {"od": 1, "nd": 1}, {"jump": 100}, (2, 7, 1), list(range(1, 17)), {"blockscale": 0.25}),
    ]
    out = []
    for name, fl, muts, pats, gens, seed, cases, extra in cfgs:
        rows = []
        for i in cases:
            rr = ref.case_paths(fl, i, seed, muts, pats, generators=gens, **extra)
            rows.append((rr.status, rr.output, rr.draws, rr.detail[:100]))
            if rr.status != "ok":
                print(name, i, rr.status, rr.detail[:100], flush=True)
        out.append({"name": name, "files": [f.hex() for f in fl], "mutations": muts, "patterns": pats, "generators": gens, "seed": list(seed), "cases": cases,
                    "extra": extra, "status": [r[0] for r in rows], "digests": [[len(r[1]), hashlib.sha256(r[1]).hexdigest()] for r in rows], "draws": [r[2] for r in rows]})
        print(name, "done", flush=True)
    with open(os.path.join(HERE, "reference_vectors_jump.json"), "w") as f:
        json.dump({"provenance": "erlamsa reference sources (commit 4a844bcd) executed by oracle/erlref; see make_reference_vectors_jump.py", "vectors": out}, f, indent=0)
    print("wrote", len(out), "configs,", sum(len(v["cases"]) for v in out), "cases")


if __name__ == "__main__":
    main()
