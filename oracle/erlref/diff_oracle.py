#!/usr/bin/env python
"""Differential run: the reference's own source (oracle/erlref evaluator) vs the C++ oracle (oracle/src), on random
(seed, input, options) triples.  Every divergence is a bug in one of the two to be argued from the .erl text.
usage: python -m erlref.diff_oracle [--n 2000] [--seed 1] [--procs 8] [--mode default|single|pattern]"""
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
    blob, seed, case_no, muts, pats = make_case(idx, mode, base_seed)

    import signal

    class Timeout(BaseException):
        pass

    def on_alarm(signum, frame):
        raise Timeout()
    signal.signal(signal.SIGALRM, on_alarm)
    signal.alarm(TIME_LIMIT)
    try:
        rr = _ref.case(blob, case_no, seed, muts, pats)     # main thread of the pool worker: the alarm can interrupt it
    except Timeout:
        _ref = None                                         # evaluator state may be inconsistent: start afresh
        return (idx, "skip", "ref=timeout", muts, pats)
    except Exception as e:   # evaluator gap: report, do not hide
        return (idx, "evaluator-error", repr(e)[:300], muts, pats)
    finally:
        signal.alarm(0)
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
