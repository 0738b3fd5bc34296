"""Committed golden vectors (tests/golden/oracle_vectors.json, provenance in make_golden.py):
CPU: the oracle still reproduces them;  GPU: the CUDA path reproduces the same bytes through the C ABI."""
import hashlib
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
VEC = json.load(open(os.path.join(HERE, "golden", "oracle_vectors.json")))["vectors"]


def digests(outs):
    return [[len(o), hashlib.sha256(o).hexdigest()] for o in outs]


@pytest.mark.parametrize("v", VEC, ids=[v["name"] for v in VEC])
def test_oracle_reproduces_golden(v, oracle):
    blobs = [bytes.fromhex(b) for b in v["blobs"]]
    outs, meta = oracle.fuzzer(blobs, mutations=v["mutations"], patterns=v["patterns"], seed=tuple(v["seed"]),
                               n_cases=v["n_cases"], first_case=v["first_case"], max_case_out=1 << 20)
    assert digests(outs) == v["digests"]
    assert [o.hex() for o, w in zip(outs, v["outputs"]) if w is not None] == [w for w in v["outputs"] if w is not None]
    assert [m.status for m in meta] == v["status"]
    assert [m.draws for m in meta] == v["draws"]


@pytest.mark.gpu
@pytest.mark.parametrize("v", VEC, ids=[v["name"] for v in VEC])
def test_engine_reproduces_golden(v, engine):
    blobs = [bytes.fromhex(b) for b in v["blobs"]]
    outs, meta = engine.fuzz_batch(blobs, {"mutations": v["mutations"], "patterns": v["patterns"], "seed": tuple(v["seed"]),
                                           "first_case": v["first_case"], "max_case_out": 1 << 20}, n_cases=v["n_cases"])
    if v["device_complete"]:
        assert [m.status for m in meta] == [0] * v["n_cases"]
    # cases the engine flags (documented gaps / caps) or the oracle could not finish are excluded, everything else is exact
    keep = [k for k in range(v["n_cases"]) if meta[k].status == 0 and v["status"][k] == 0]
    assert len(keep) * 2 >= v["n_cases"], "engine flagged too many cases: %r" % [m.status for m in meta]
    assert [digests([outs[k]])[0] for k in keep] == [v["digests"][k] for k in keep]
    assert [outs[k].hex() for k in keep if v["outputs"][k] is not None] == [v["outputs"][k] for k in keep if v["outputs"][k] is not None]
    assert [meta[k].draws for k in keep] == [v["draws"][k] for k in keep]
    assert [meta[k].pattern for k in keep] == [v["pattern"][k] for k in keep]
    assert [[u for u in meta[k].used if u >= 0] for k in keep] == [v["used"][k][:16] for k in keep]
