"""REFERENCE vectors (tests/golden/reference_vectors.json): outputs of erlamsa's own source, executed in the build
container by oracle/erlref (provenance: tests/golden/make_reference_vectors.py). They pin

  CPU: the C++ oracle       == the reference   (bytes + RNG draw count per case)
  GPU: the CUDA engine      == the reference   (same, through the C ABI) -- directly, not via the oracle.

Cases the engine flags (status != 0: documented device gaps and capacity limits) are not compared;
WHICH cases those are is pinned exactly in tests/golden/expected_flags.json (case index -> [status, reason]), so a
regression that flags one more case fails. Set EB200_DUMP_FLAGS=1 to write what a run saw to gpurun_out/flags_seen.json."""
import hashlib
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(HERE, "golden", "reference_vectors.json")
VEC = json.load(open(PATH))["vectors"] if os.path.exists(PATH) else []
FLAGS_PATH = os.path.join(HERE, "golden", "expected_flags.json")
EXPECTED_FLAGS = json.load(open(FLAGS_PATH)) if os.path.exists(FLAGS_PATH) else {}
CAP = 1 << 22


def digest(o):
    return [len(o), hashlib.sha256(o).hexdigest()]


def test_vectors_present():
    assert len(VEC) >= 60 and sum(v["n_cases"] for v in VEC) >= 700
    ok = sum(1 for v in VEC for s in v["status"] if s == "ok")
    assert ok >= 0.9 * sum(v["n_cases"] for v in VEC)


@pytest.mark.parametrize("v", VEC, ids=[v["name"] for v in VEC])
def test_oracle_matches_reference(v, oracle):
    blobs = [bytes.fromhex(b) for b in v["blobs"]]
    kw = {}
    if "generators" in v["extra"]:
        kw["generators"] = v["extra"]["generators"]
    if "blockscale" in v["extra"]:
        kw["blockscale"] = v["extra"]["blockscale"]
    outs, meta = oracle.fuzzer(blobs, mutations=v["mutations"], patterns=v["patterns"], seed=tuple(v["seed"]), n_cases=v["n_cases"], first_case=v["first_case"],
                               max_case_out=CAP, **kw)
    bad = []
    for k in range(v["n_cases"]):
        st = v["status"][k]
        if st == "ok":
            if meta[k].status == 3 and v["digests"][k][0] > CAP // 4:
                continue      # the oracle's own output cap
            if meta[k].status != 0 or digest(outs[k]) != v["digests"][k] or meta[k].draws != v["draws"][k]:
                bad.append((k, meta[k].status, len(outs[k]), v["digests"][k][0], meta[k].draws, v["draws"][k]))
        elif st == "died":
            if meta[k].status != 2:
                bad.append((k, "reference died (%s), oracle status %d" % (v["detail"][k], meta[k].status)))
        # "budget" / "unsupported": the evaluator gave up (runaway growth, zip / zlib), nothing to compare
    assert not bad, "oracle differs from the reference at %r" % bad[:8]


@pytest.mark.gpu
@pytest.mark.parametrize("v", VEC, ids=[v["name"] for v in VEC])
def test_engine_matches_reference(v, engine):
    blobs = [bytes.fromhex(b) for b in v["blobs"]]
    opts = {"seed": tuple(v["seed"]), "first_case": v["first_case"], "max_case_out": CAP}
    if v["mutations"] is not None:
        opts["mutations"] = v["mutations"]
    if v["patterns"] is not None:
        opts["patterns"] = v["patterns"]
    opts.update(v["extra"])
    outs, meta = engine.fuzz_batch(blobs, opts, n_cases=v["n_cases"])
    flagged = {str(k): [meta[k].status, meta[k].pad] for k in range(v["n_cases"]) if meta[k].status not in (0, 2)}
    if os.environ.get("EB200_DUMP_FLAGS"):
        p = os.path.join(os.path.dirname(HERE), "gpurun_out", "flags_seen.json")
        seen = json.load(open(p)) if os.path.exists(p) else {}
        seen[v["name"]] = flagged
        os.makedirs(os.path.dirname(p), exist_ok=True)
        json.dump(seen, open(p, "w"), indent=0, sort_keys=True)
    bad = []
    for k in range(v["n_cases"]):
        st = v["status"][k]
        if str(k) in flagged:
            continue
        if st == "ok":
            if meta[k].status != 0 or digest(outs[k]) != v["digests"][k] or meta[k].draws != v["draws"][k]:
                bad.append((k, meta[k].status, len(outs[k]), v["digests"][k][0], meta[k].draws, v["draws"][k]))
        elif st == "died":
            if meta[k].status != 2:
                bad.append((k, "reference died (%s), engine status %d" % (v["detail"][k], meta[k].status)))
    assert not bad, "engine differs from the reference at %r" % bad[:8]
    assert flagged == EXPECTED_FLAGS.get(v["name"], {}), "flagged set changed: %r" % flagged


OTP_PATH = os.path.join(HERE, "golden", "otp_vectors.txt")


@pytest.mark.skipif(not os.path.exists(OTP_PATH), reason="no tests/golden/otp_vectors.txt (made by tools/dump_reference_vectors.erl on a box with Erlang/OTP <= 23)")
def test_otp_run_matches_committed_vectors():
    """closes the last gap: the same cases run by a REAL Erlang/OTP must give the bytes the evaluator gave"""
    want = {(v["name"], k): v for v in VEC for k in range(v["n_cases"])}
    bad, n = [], 0
    for ln in open(OTP_PATH):
        f = ln.split()
        if len(f) < 3:
            continue
        name, k, st = f[0], int(f[1]), f[2]
        v = want[(name, k)]
        if v["status"][k] != "ok":
            continue
        n += 1
        out = bytes.fromhex(f[3]) if (st == "ok" and len(f) > 3) else b""
        if st != "ok" or digest(out) != v["digests"][k]:
            bad.append((name, k, st))
    assert n > 0 and not bad, bad[:10]


# ---------------------------------------------------------------- file / stdin generators (reference_vectors_paths.json)
PATHS_PATH = os.path.join(HERE, "golden", "reference_vectors_paths.json")
PVEC = json.load(open(PATHS_PATH))["vectors"] if os.path.exists(PATHS_PATH) else []


def _paths_opts(v):
    gens = v["generators"]
    if v["stdin"] is not None:
        gens = gens or {"stdin": 100000, "random": 1}       # what make_generator keeps of the defaults for paths = ["-"]
        blobs = [bytes.fromhex(v["stdin"])]
    else:
        blobs = [bytes.fromhex(b) for b in v["files"]]
    return blobs, gens


@pytest.mark.parametrize("v", PVEC, ids=[v["name"] for v in PVEC])
def test_oracle_matches_reference_file_and_stdin_generators(v, oracle):
    blobs, gens = _paths_opts(v)
    kw = dict(v["extra"])
    bad = []
    for idx, i in enumerate(v["cases"]):
        outs, meta = oracle.fuzzer(blobs, mutations=v["mutations"], patterns=v["patterns"], seed=tuple(v["seed"]), generators=gens, n_cases=1, first_case=i,
                                   max_case_out=1 << 26, **kw)
        if v["status"][idx] == "ok" and (meta[0].status != 0 or digest(outs[0]) != v["digests"][idx] or meta[0].draws != v["draws"][idx]):
            bad.append((i, meta[0].status, len(outs[0]), v["digests"][idx][0], meta[0].draws, v["draws"][idx]))
    assert not bad, bad[:8]


@pytest.mark.gpu
@pytest.mark.parametrize("v", PVEC, ids=[v["name"] for v in PVEC])
def test_engine_matches_reference_file_and_stdin_generators(v, engine):
    blobs, gens = _paths_opts(v)
    bad, flagged = [], {}
    for idx, i in enumerate(v["cases"]):
        opts = {"seed": tuple(v["seed"]), "first_case": i, "max_case_out": 1 << 26, "generators": gens}
        if v["mutations"] is not None:
            opts["mutations"] = v["mutations"]
        if v["patterns"] is not None:
            opts["patterns"] = v["patterns"]
        opts.update(v["extra"])
        if v["stdin"] is not None:
            opts["first_case"] = 1
        outs, meta = engine.fuzz_batch(blobs, opts, n_cases=1)
        if meta[0].status not in (0, 2):
            flagged[str(i)] = [meta[0].status, meta[0].pad]
            continue
        if v["status"][idx] == "ok" and (meta[0].status != 0 or digest(outs[0]) != v["digests"][idx] or meta[0].draws != v["draws"][idx]):
            bad.append((i, meta[0].status, len(outs[0]), v["digests"][idx][0], meta[0].draws, v["draws"][idx]))
    if os.environ.get("EB200_DUMP_FLAGS"):
        p = os.path.join(os.path.dirname(HERE), "gpurun_out", "flags_seen.json")
        seen = json.load(open(p)) if os.path.exists(p) else {}
        seen[v["name"]] = flagged
        json.dump(seen, open(p, "w"), indent=0, sort_keys=True)
    assert not bad, "engine differs from the reference at %r" % bad[:8]
    assert flagged == EXPECTED_FLAGS.get(v["name"], {}), "flagged set changed: %r" % flagged


# ---------------------------------------------------------------- jump generator (reference_vectors_jump.json): oracle only
JUMP_PATH = os.path.join(HERE, "golden", "reference_vectors_jump.json")
JVEC = json.load(open(JUMP_PATH))["vectors"] if os.path.exists(JUMP_PATH) else []


@pytest.mark.parametrize("v", JVEC, ids=[v["name"] for v in JVEC])
def test_oracle_matches_reference_jump_generator(v, oracle):
    """src/erlamsa_gen.erl:123-150 -- the engine does not implement this generator (its host mirror refuses the option), the oracle
    does: every case the reference completes agrees in bytes and draw count, every case whose worker dies there dies here."""
    blobs = [bytes.fromhex(b) for b in v["files"]]
    bad = []
    for idx, i in enumerate(v["cases"]):
        outs, meta = oracle.fuzzer(blobs, mutations=v["mutations"], patterns=v["patterns"], seed=tuple(v["seed"]), generators=v["generators"], n_cases=1,
                                   first_case=i, max_case_out=1 << 26, **v["extra"])
        if v["status"][idx] == "ok":
            if meta[0].status != 0 or digest(outs[0]) != v["digests"][idx] or meta[0].draws != v["draws"][idx]:
                bad.append((i, meta[0].status, len(outs[0]), v["digests"][idx][0], meta[0].draws, v["draws"][idx]))
        elif v["status"][idx] == "died" and meta[0].status != 2:
            bad.append((i, "reference died", meta[0].status))
    assert not bad, bad[:8]


def test_engine_refuses_a_run_whose_generator_draw_lands_on_jump():
    """the option is accepted (it takes part in the parent's draw, tests/test_parent_draws.py); with jump as the only generator the
    draw can only land on it, and the engine's host code answers EB200_ERR_UNSUPPORTED before anything is launched"""
    import ctypes as C
    import erlamsa_b200
    from erlamsa_b200 import _native as N
    o = erlamsa_b200.make_opts({"generators": {"jump": 100}, "seed": (1, 2, 3)})
    out = (C.c_int64 * 8)()
    assert N.lib().eb200_debug_parent_draws(C.byref(o), 2, 1, out) == -3 and out[0] == 4
    with pytest.raises(ValueError):
        erlamsa_b200.make_opts({"generators": {"genfuz": 10000}})
