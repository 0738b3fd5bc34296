"""The parent-process part of erlamsa_main:fuzzer/1 as the ENGINE's host code restates it (compute_batch_params in eb_engine.cu, reached
without a GPU through eb200_debug_parent_draws) against the oracle, which is pinned to the reference's source: for random seeds, mutator /
pattern selections and generator lists the generator chosen by the parent's draw and the first case's thread seed must agree. Covers the
jump generator's place in that draw (src/erlamsa_gen.erl:194-199,220): kept only with two or more paths, refused when the draw lands on it."""
import ctypes as C

import numpy as np
import pytest


def _first_thread_seed(a1, a2, a3):
    out = []
    for _ in range(3):                      # erlamsa_rnd:gen_predictable_seed/0 :65 = 3 x erand(99999) over OTP random (AS183)
        a1, a2, a3 = a1 * 171 % 30269, a2 * 172 % 30307, a3 * 170 % 30323
        r = a1 / 30269 + a2 / 30307 + a3 / 30323
        out.append(int((r - int(r)) * 99999) + 1)
    return out


def _engine_parent(opts, n_blobs, n_cases=1):
    import erlamsa_b200
    from erlamsa_b200 import _native as N
    o = erlamsa_b200.make_opts(opts)
    out = (C.c_int64 * 8)()
    rc = N.lib().eb200_debug_parent_draws(C.byref(o), n_blobs, n_cases, out)
    return rc, list(out)


def test_parent_draws_match_the_oracle(oracle):
    import erlamsa_b200
    r = np.random.Generator(np.random.PCG64(20260923))
    codes = [c for c, _ in erlamsa_b200.default_mutations() if c in erlamsa_b200.supported_mutations()]
    pcodes = ["od", "nd", "bu", "sk", "sz", "cs", "co", "nu"]
    files = [b"alpha 1\nbeta 22\n", b"gamma 333\ndelta\n" * 40, b"x" * 300]
    n_jump = n_checked = 0
    for it in range(400):
        seed = tuple(int(x) for x in r.integers(0, 100000, size=3))
        muts = {c: int(r.integers(1, 6)) for c in r.choice(codes, size=int(r.integers(1, 12)), replace=False)}
        pats = {c: int(r.integers(1, 4)) for c in r.choice(pcodes, size=int(r.integers(1, 4)), replace=False)}
        kind = it % 4
        if kind == 0:
            gens, blobs = {"direct": 500, "random": 1}, files[:1]
        elif kind == 1:
            gens, blobs = {"file": 1000, "random": 1}, files
        elif kind == 2:
            gens, blobs = {"random": 1, "jump": 100, "file": 1000}, files
        else:      # jump asked for with ONE path: make_generator drops it
            gens, blobs = {"random": 1, "jump": int(r.integers(1, 3000)), "file": int(r.integers(1, 2000))}, files[:1]
        rc, out = _engine_parent({"seed": seed, "mutations": muts, "patterns": pats, "generators": gens}, len(blobs))
        ogens = dict(gens)
        if len(blobs) < 2:
            ogens.pop("jump", None)
        _, meta = oracle.fuzzer(blobs, mutations=muts, patterns=pats, seed=seed, generators=ogens, n_cases=1, max_case_out=1 << 20)
        assert out[0] == meta[0].generator, (it, seed, gens, out, meta[0].generator)
        if meta[0].generator == 4:
            n_jump += 1
            assert rc == -3            # EB200_ERR_UNSUPPORTED: the draw is the reference's, the generator is not on the device
            continue
        assert rc == 0
        assert _first_thread_seed(out[4], out[5], out[6]) == list(meta[0].thread_seed), (it, seed)
        assert out[2] == len(muts) and out[3] == len(pats)
        n_checked += 1
    assert n_checked > 300 and n_jump >= 3


def test_mirror_defaults_for_file_paths_include_jump_like_the_reference(oracle):
    """erlamsa_main.py hands the engine {random, jump, file} for two or more paths: the draw of tests/test_file_frontend.py's run
    (seed 1,2,3) lands on `file`, as it does in the oracle with the reference's default list"""
    muts = {"bd": 1, "bf": 1, "num": 2, "ld": 1, "sr": 1, "fn": 1}
    pats = {"od": 1, "nd": 1, "sk": 1}
    gens = {"random": 1, "jump": 100, "file": 1000}
    rc, out = _engine_parent({"seed": (1, 2, 3), "mutations": muts, "patterns": pats, "generators": gens}, 3, 24)
    _, meta = oracle.fuzzer([b"a\n", b"b\n", b"c\n"], mutations=muts, patterns=pats, seed=(1, 2, 3), generators=gens, n_cases=1)
    assert rc == 0 and out[0] == 2 == meta[0].generator
