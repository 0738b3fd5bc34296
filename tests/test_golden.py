"""Committed golden vectors (tests/golden/oracle_vectors.json, provenance in make_golden.py):
CPU: the oracle still reproduces them;  GPU: the CUDA path reproduces the same bytes through the C ABI."""
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
VEC = json.load(open(os.path.join(HERE, "golden", "oracle_vectors.json")))["vectors"]


@pytest.mark.parametrize("v", VEC, ids=[v["name"] for v in VEC])
def test_oracle_reproduces_golden(v, oracle):
    blobs = [bytes.fromhex(b) for b in v["blobs"]]
    outs, meta = oracle.fuzzer(blobs, mutations=v["mutations"], patterns=v["patterns"], seed=tuple(v["seed"]),
                               n_cases=v["n_cases"], first_case=v["first_case"])
    assert [o.hex() for o in outs] == v["outputs"]
    assert [m.draws for m in meta] == v["draws"]


@pytest.mark.gpu
@pytest.mark.parametrize("v", VEC, ids=[v["name"] for v in VEC])
def test_engine_reproduces_golden(v, engine):
    blobs = [bytes.fromhex(b) for b in v["blobs"]]
    outs, meta = engine.fuzz_batch(blobs, {"mutations": v["mutations"], "patterns": v["patterns"], "seed": tuple(v["seed"]),
                                           "first_case": v["first_case"], "max_case_out": 1 << 28}, n_cases=v["n_cases"])
    assert [m.status for m in meta] == [0] * v["n_cases"]
    assert [o.hex() for o in outs] == v["outputs"]
    assert [m.draws for m in meta] == v["draws"]
    assert [m.pattern for m in meta] == v["pattern"]
    assert [[u for u in m.used if u >= 0] for m in meta] == [u[:16] for u in v["used"]]
