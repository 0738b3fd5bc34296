#!/usr/bin/env python
"""Regenerates tests/golden/reference_vectors.json -- outputs of the REFERENCE ITSELF.

PROVENANCE: every vector is what erlamsa's own source (/root/reference/src/*.erl, commit 4a844bcd) returns for
erlamsa_main:fuzzer(#{paths => [direct], output => return, input => Blob, seed => Seed, n => I, skip => I - 1, ...}),
executed in the build container by oracle/erlref (an Erlang evaluator written for this purpose: the reference's modules
are loaded and run as they are; only OTP's own library functions are restated, see oracle/erlref/bifs.py and
oracle/erlref/otp/otp_lists.erl). /root/reference does not travel to the GPU box, so the vectors are committed.
tests/test_reference_vectors.py checks the C++ oracle (CPU) and the CUDA engine (GPU) against them.

usage: python tests/golden/make_reference_vectors.py [--procs 7]"""
import argparse
import hashlib
import json
import multiprocessing as mp
import os
import signal
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
TIME_LIMIT = 240


def framed(r, n):
    """length-prefixed / checksummed blobs so that sz, cs and len find something"""
    import zlib
    body = corpus.random_bytes(r, n)
    k = int(r.integers(0, 4))
    if k == 0:
        return b"HD" + len(body).to_bytes(2, "big") + body + b"TAIL"
    if k == 1:
        return bytes([len(body) & 255]) + body[:255]
    if k == 2:
        x = 0
        for c in body:
            x ^= c
        return b"P" + body + bytes([x])
    return body + (zlib.crc32(body) & 0xFFFFFFFF).to_bytes(4, "big")


def configs():
    r = corpus.rng(20260923)
    out = []
    text = [corpus.structured_text(corpus.rng(500 + i), 300) for i in range(6)]
    lines = [corpus.text_lines(corpus.rng(520 + i), 300) for i in range(6)]
    nums = [corpus.numeric_text(corpus.rng(540 + i), 400) for i in range(6)]
    binr = [corpus.random_bytes(corpus.rng(560 + i), int(50 + 60 * i)) for i in range(6)]
    web = corpus.web_corpus(570, 10)
    sg = [corpus.sgml_doc(corpus.rng(580 + i), 500) for i in range(4)] + [b"<a>x</a>", b"<p>1<b>2</p>3</b>", b'<r a="1" b=\'2\' c>t<br/><!-- c --></r>']
    js = [corpus.json_doc(corpus.rng(590 + i), 500) for i in range(4)] + [b'{"a":[1,true,null],"b":"str"}', b"[1,2,3]", b" 17", b'{"k":{"n":-5,"s":"http://x/y","l":[]}}']
    fr = [framed(corpus.rng(600 + i), 40 + 30 * i) for i in range(8)]
    mixed = corpus.mixed_corpus(610, 12, 400)
    suited = {"sgm": sg, "js": js, "uw": binr + text, "ui": binr + text, "ab": text, "ad": text, "tr2": text, "td": text, "num": nums, "ts1": text, "tr": text, "ts2": text,
              "ld": lines, "lds": lines, "lr2": lines, "lri": lines, "lr": lines, "ls": lines, "lp": lines, "lis": lines, "lrs": lines,
              "ft": text + lines, "fn": text + lines, "fo": text + lines, "len": fr, "b64": web, "uri": web, "zip": binr, "nil": binr}
    # C1 of BASELINE.json (the direct-path equivalent) and the README's soft known answer
    out.append(("c1_hello_erlamsa_default", [b"Hello erlamsa!\n"], DEFAULT, ALL_PATS, (1, 2, 3), 24, 1))
    out.append(("readme_hello_default", [b"hello"], DEFAULT, ALL_PATS, (1, 2, 3), 12, 1))
    out.append(("reference_defaults_omitted", [b"Hello erlamsa!\n", b"hello 100\n"], None, None, (1, 2, 3), 8, 1))
    # every mutator on its own under `od`
    for m in MUT_CODES:
        blobs = suited.get(m, binr + nums[:2])
        out.append(("single_" + m, blobs, {m: 1}, {"od": 1}, (3, 1, 4), max(6, len(blobs)), 1))
    # every pattern on its own with a mixed table
    table = {"bd": 1, "bf": 1, "num": 3, "sr": 1, "sd": 1, "lr2": 1, "ld": 1, "ui": 2, "ft": 1, "ab": 1, "td": 1, "len": 2}
    for p in ALL_PATS:
        out.append(("pattern_" + p, mixed[:6] + fr[:4] + text[:2], table, {p: 1}, (2, 7, 1), 12, 3))
    # the reference's defaults on everything
    out.append(("default_mixed", mixed + web + sg[:3] + js[:3] + fr[:4], DEFAULT, ALL_PATS, (9, 8, 7), 64, 1))
    out.append(("default_text", text + lines + nums, DEFAULT, ALL_PATS, (5, 5, 5), 36, 1000))
    # BASELINE configs, scaled to sizes the evaluator finishes
    out.append(("c2_shape_4096_uniform", [corpus.random_bytes(corpus.rng(700 + i), 4096) for i in range(12)], DEFAULT, ALL_PATS, (1, 2, 3), 24, 1))
    c3 = {c: 1 for c in ("bd", "bei", "bed", "bf", "bi", "ber", "br", "num")}
    out.append(("c3_mutators_uniform", [corpus.random_bytes(corpus.rng(720 + i), 3000) for i in range(8)], c3, {"od": 1}, (1, 2, 3), 32, 1))
    out.append(("c3_mutators_numeric", [corpus.numeric_text(corpus.rng(730 + i), 3000) for i in range(8)], c3, {"od": 1}, (1, 2, 3), 32, 50))
    c4 = {c: 1 for c in ("ab", "ad", "tr2", "td", "ts1", "ts2", "tr", "sgm", "js")}
    out.append(("c4_mutators_markup", [corpus.sgml_doc(corpus.rng(740 + i), 1500) for i in range(4)] + [corpus.json_doc(corpus.rng(750 + i), 1500) for i in range(4)], c4, {"od": 1}, (1, 2, 3), 32, 1))
    out.append(("c5_fuse", text + lines, {"ft": 2, "fn": 1, "fo": 2}, {"od": 1, "nd": 1}, (1, 2, 3), 24, 1))
    out.append(("lines_stateful_rounds", lines, {"lis": 1, "lrs": 1, "lp": 1, "ls": 1}, {"nd": 1, "bu": 1}, (4, 5, 6), 18, 1))
    out.append(("seq_rounds", binr + nums[:3], {c: 1 for c in ("sp", "sr", "sd", "snand", "srnd", "uw", "ui")}, {"od": 1, "bu": 1}, (7, 8, 9), 24, 100))
    out.append(("random_generator_only", binr[:2], {"bd": 1, "bf": 1, "num": 1}, {"od": 1}, (6, 6, 6), 6, 1, {"generators": {"random": 1}}))
    out.append(("blockscale_half", lines[:3], {"ld": 1, "bd": 1}, {"od": 1}, (8, 1, 8), 6, 1, {"blockscale": 0.5}))
    return out


_ref = None


class Timeout(BaseException):
    pass


def _alarm(signum, frame):
    raise Timeout()


def run_case(job):
    global _ref
    name, k, blob, case_no, seed, muts, pats, extra = job
    from erlref.refrun import Reference
    if _ref is None:
        _ref = Reference(budget=400_000_000)
    signal.signal(signal.SIGALRM, _alarm)
    signal.alarm(TIME_LIMIT)
    try:
        rr = _ref.case(blob, case_no, seed, muts, pats, **(extra or {}))
        return name, k, rr.status, rr.detail[:120], rr.output, rr.draws
    except Timeout:
        _ref = None
        return name, k, "budget", "wall clock", b"", 0
    finally:
        signal.alarm(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=7)
    a = ap.parse_args()
    import resource
    sys.setrecursionlimit(3000000)
    resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))
    cfgs = configs()
    jobs = []
    for c in cfgs:
        name, blobs, muts, pats, seed, n, first = c[:7]
        extra = c[7] if len(c) > 7 else None
        for k in range(n):
            i = first + k
            jobs.append((name, k, blobs[(i - 1) % len(blobs)], i, seed, muts, pats, extra))
    res = {}
    with mp.Pool(a.procs, maxtasksperchild=40) as pool:
        for name, k, st, detail, out, draws in pool.imap_unordered(run_case, jobs, chunksize=1):
            res[(name, k)] = (st, detail, out, draws)
            if st != "ok":
                print(name, k, st, detail, flush=True)
    vectors = []
    for c in cfgs:
        name, blobs, muts, pats, seed, n, first = c[:7]
        extra = c[7] if len(c) > 7 else None
        rows = [res[(name, k)] for k in range(n)]
        vectors.append({"name": name, "blobs": [b.hex() for b in blobs], "mutations": muts, "patterns": pats, "seed": list(seed), "n_cases": n, "first_case": first,
                        "extra": extra or {},
                        "status": [r[0] for r in rows], "detail": [r[1] for r in rows],
                        "outputs": [r[2].hex() if len(r[2]) <= 2048 else None for r in rows],
                        "digests": [[len(r[2]), hashlib.sha256(r[2]).hexdigest()] for r in rows],
                        "draws": [r[3] for r in rows]})
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump({"provenance": "erlamsa reference sources (commit 4a844bcd) executed by oracle/erlref; see make_reference_vectors.py",
                   "vectors": vectors}, f, indent=0)
    nok = sum(1 for r in res.values() if r[0] == "ok")
    print("wrote %d configs, %d cases (%d ok)" % (len(vectors), len(res), nok))


if __name__ == "__main__":
    main()
