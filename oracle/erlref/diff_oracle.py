#!/usr/bin/env python
"""Differential run: the reference's own source (oracle/erlref evaluator) vs the C++ oracle (oracle/src), on random
(seed, input, options) triples.  Every divergence is a bug in one of the two to be argued from the .erl text.
usage: python -m erlref.diff_oracle [--n 2000] [--seed 1] [--procs 8] [--mode default|single|pattern|mixed|paths]
mode paths: fuzzer(#{paths => Files, ...}) -- the file, jump and random generators over random sets of in-memory files, random block scale"""
import argparse
import os
import sys
import multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

MUT_CODES = ["sgm", "js", "uw", "ui", "ab", "ad", "tr2", "td", "num", "ts1", "tr", "ts2", "bd", "bei", "bed", "bf", "bi", "ber", "br", "sp", "sr", "sd",
             "snand", "srnd", "ld", "lds", "lr2", "lri", "lr", "ls", "lp", "lis", "lrs", "ft", "fn", "fo", "len", "b64", "uri", "zip", "nil"]
DEF_PRI = [10, 3, 1, 2, 1, 1, 1, 1, 3, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 2, 2, 7, 1, 1, 0]
PAT_CODES = ["od", "nd", "bu", "sk", "sz", "cs", "ar", "cp", "co", "nu"]
PAT_PRI = [1, 2, 1, 2, 2, 1, 1, 1, 0, 0]


def make_case(idx, mode, base_seed):
    import numpy as np
    import corpus
    r = corpus.rng(base_seed * 1000003 + idx)
    kind = int(r.integers(0, 9))
    n = int(r.integers(0, 400)) if r.random() < 0.8 else int(r.integers(400, 2600))
    if kind == 0:
        blob = corpus.random_bytes(r, n)
    elif kind == 1:
        blob = corpus.numeric_text(r, n)
    elif kind == 2:
        blob = corpus.text_lines(r, n)
    elif kind == 3:
        blob = corpus.structured_text(r, max(n, 1))
    elif kind == 4:
        blob = corpus.web_corpus(int(r.integers(0, 1 << 30)), 5)[int(r.integers(0, 5))]
    elif kind == 5:
        blob = corpus.sgml_doc(r, max(n, 16))
    elif kind == 6:
        blob = corpus.json_doc(r, max(n, 16))
    elif kind == 7:
        blob = corpus.framed_blob(r, max(n, 8)) if hasattr(corpus, "framed_blob") else corpus.random_bytes(r, n)
    else:
        blob = corpus.text_lines(r, int(r.integers(0, 12)))
    seed = (int(r.integers(0, 100000)), int(r.integers(0, 100000)), int(r.integers(0, 100000)))
    case_no = int(r.integers(1, 40))
    if mode == "default":
        muts = dict(zip(MUT_CODES, DEF_PRI))
        pats = dict(zip(PAT_CODES, PAT_PRI))
    elif mode == "single":
        m = MUT_CODES[idx % 40]
        muts = {m: 1}
        pats = {"od": 1}
    elif mode == "pattern":
        k = int(r.integers(1, 8))
        pick = [MUT_CODES[int(i)] for i in r.choice(40, size=k, replace=False)]
        muts = {c: int(r.integers(1, 5)) for c in pick}
        p = PAT_CODES[idx % 10]
        pats = {p: 1}
    else:
        k = int(r.integers(1, 12))
        pick = [MUT_CODES[int(i)] for i in r.choice(41, size=k, replace=False)]
        muts = {c: int(r.integers(0, 8)) for c in pick}
        kp = int(r.integers(1, 5))
        pp = [PAT_CODES[int(i)] for i in r.choice(10, size=kp, replace=False)]
        pats = {c: int(r.integers(1, 4)) for c in pp}
    return blob, seed, case_no, muts, pats


def make_paths_case(idx, base_seed):
    """random file set + generator list for the file / jump / random generators (src/erlamsa_gen.erl:59-150)"""
    import corpus
    r = corpus.rng(base_seed * 1000003 + idx)
    nf = int(r.integers(1, 6))
    files = []
    for _ in range(nf):
        kind = int(r.integers(0, 6))
        n = int(r.integers(0, 600)) if r.random() < 0.5 else int(r.integers(600, 12000))
        if kind == 0:
            files.append(corpus.random_bytes(r, n))
        elif kind == 1:
            files.append(corpus.numeric_text(r, n))
        elif kind == 2:
            files.append(corpus.text_lines(r, n))
        elif kind == 3:
            files.append(corpus.structured_text(r, max(n, 1)))
        elif kind == 4:
            files.append(b"" if r.random() < 0.3 else b"ab"[: int(r.integers(1, 3))])
        else:
            files.append(corpus.sgml_doc(r, max(n, 16)))
    g = int(r.integers(0, 4))
    # lists in the order of erlamsa_gen:generators/0 (random, jump, direct, file, ...): make_generator walks the option list as given and
    # sort_by_priority keeps that order among equal priorities; the engine's options carry priorities only, i.e. they mean this order
    gens = [{"random": 1, "file": 1000}, {"random": 1, "jump": 100, "file": 1000}, {"jump": 100}, {"random": int(r.integers(0, 5)), "jump": int(r.integers(1, 50)), "file": int(r.integers(1, 50))}][g]
    if nf < 2 and g == 2:
        gens = {"random": 1, "jump": 100, "file": 1000}      # jump alone with one path leaves "No generators!"
    seed = (int(r.integers(0, 100000)), int(r.integers(0, 100000)), int(r.integers(0, 100000)))
    case_no = int(r.integers(1, 30))
    light = ["bd", "bei", "bf", "bi", "ber", "br", "num", "ld", "lds", "lr2", "lri", "lr", "ls", "lp", "sr", "sd", "sp", "fn", "fo", "ft", "ui", "uw", "ab", "ad", "len", "lis", "lrs"]
    k = int(r.integers(1, 9))
    muts = {light[int(i)]: int(r.integers(1, 5)) for i in r.choice(len(light), size=k, replace=False)}
    kp = int(r.integers(1, 5))
    pats = {PAT_CODES[int(i)]: int(r.integers(1, 4)) for i in r.choice(10, size=kp, replace=False)}
    extra = {"blockscale": [0.25, 0.5, 2.0][int(r.integers(0, 3))]} if r.random() < 0.25 else {}
    return files, gens, seed, case_no, muts, pats, extra


_ref = None
TIME_LIMIT = 45


def work(args):
    global _ref
    idx, mode, base_seed, cap = args
    from erlref.refrun import Reference
    from erlref.interp import run_with_big_stack
    import oracle_lib as O
    if _ref is None:
        _ref = Reference()
    if mode == "paths":
        files, gens, seed, case_no, muts, pats, extra = make_paths_case(idx, base_seed)
    else:
        blob, seed, case_no, muts, pats = make_case(idx, mode, base_seed)

    import signal

    class Timeout(BaseException):
        pass

    def on_alarm(signum, frame):
        raise Timeout()
    signal.signal(signal.SIGALRM, on_alarm)
    signal.alarm(TIME_LIMIT)
    try:
        if mode == "paths":
            rr = _ref.case_paths(files, case_no, seed, muts, pats, generators=gens, **extra)
        else:
            rr = _ref.case(blob, case_no, seed, muts, pats)     # main thread of the pool worker: the alarm can interrupt it
    except Timeout:
        _ref = None                                         # evaluator state may be inconsistent: start afresh
        return (idx, "skip", "ref=timeout", muts, pats)
    except Exception as e:   # evaluator gap: report, do not hide
        return (idx, "evaluator-error", repr(e)[:300], muts, pats)
    finally:
        signal.alarm(0)
    if mode == "paths":
        ogens = dict(gens)
        if len(files) < 2:
            ogens.pop("jump", None)                       # make_generator_fun: `jump when length(Args) > 1`, otherwise dropped
        if not ogens:
            return (idx, "skip", "no generator left", muts, pats)
        outs, meta = O.fuzzer(files, mutations=muts, patterns=pats, seed=seed, generators=ogens, n_cases=1, first_case=case_no, max_case_out=cap, **extra)
        if rr.status == "died" and meta[0].status == 2:
            return (idx, "ok", "", None, None)            # the worker dies on both sides (e.g. jump over an empty block list)
    else:
        outs, meta = O.fuzzer([blob], mutations=muts, patterns=pats, seed=seed, n_cases=1, first_case=case_no, max_case_out=cap)
    m = meta[0]
    if rr.status != "ok" or m.status != 0:
        # both sides must agree that the case is outside the comparable class
        return (idx, "skip", "ref=%s(%s) oracle_status=%d/%d" % (rr.status, rr.detail[:80], m.status, m.pad), muts, pats)
    if rr.output != outs[0] or rr.draws != m.draws:
        return (idx, "MISMATCH", "ref len %d draws %d | oracle len %d draws %d used %s pattern %d | ref %r | ora %r" % (
            len(rr.output), rr.draws, len(outs[0]), m.draws, [MUT_CODES[u] for u in m.used if u >= 0], m.pattern, rr.output[:80], outs[0][:80]), muts, pats)
    return (idx, "ok", "", None, None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=400)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--procs", type=int, default=8)
    ap.add_argument("--mode", default="default")
    ap.add_argument("--start", type=int, default=0)
    ap.add_argument("--cap", type=int, default=1 << 20)
    a = ap.parse_args()
    import resource
    sys.setrecursionlimit(3000000)
    try:
        resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))
    except Exception:
        pass
    jobs = [(i, a.mode, a.seed, a.cap) for i in range(a.start, a.start + a.n)]
    counts = {}
    with mp.Pool(a.procs, maxtasksperchild=50) as pool:
        for idx, st, detail, muts, pats in pool.imap_unordered(work, jobs, chunksize=4):
            counts[st] = counts.get(st, 0) + 1
            if st not in ("ok",):
                print(idx, st, detail, "muts=%s pats=%s" % (muts, pats) if st != "skip" else "", flush=True)
    print("summary:", counts)


if __name__ == "__main__":
    main()
