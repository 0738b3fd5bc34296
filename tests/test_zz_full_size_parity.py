"""GPU parity at the configurations' REAL sizes: the exact C2 shape x 2 000 cases and C4-size (256 KiB)
documents, CUDA engine against the oracle run on all host threads. Kept in a file of its own that sorts after the others: these are
the heaviest parity tests, and under `pytest -x` a capacity surprise here must not hide the reference-vector results.
Run on the B200 box: python -m pytest tests -m gpu"""
import pytest

import corpus
from test_parity_gpu import compare

pytestmark = pytest.mark.gpu


# every per-case capacity limit the engine documents: scratch / output arena, candidate segments, block runs, output cap,
# split pieces, edit-script segments, fuse tables, sizer / checksum wrapper depth, generator stream runs -- flagged, bounded by compare(), never wrong bytes
ALL_CAPACITY_REASONS = (1, 2, 4, 5, 6, 7, 8, 9, 10, 11)


class ThreadedOracle(object):
    """the oracle over windows of the case loop on all host threads (cases are independent and numbered globally, so the
    windows concatenate to the single-call result); for the full-size configurations, where one thread would need minutes"""

    def __init__(self, oracle):
        self.o = oracle

    def fuzzer(self, blobs, n_cases=None, first_case=1, **kw):
        import os
        from concurrent.futures import ThreadPoolExecutor
        n = len(blobs) if n_cases is None else n_cases
        threads = max(1, min(os.cpu_count() or 1, 128, n))
        per = (n + threads - 1) // threads
        wins = [(first_case + lo, min(per, n - lo)) for lo in range(0, n, per)]
        with ThreadPoolExecutor(len(wins)) as ex:                        # ctypes releases the GIL inside the C++ restatement
            parts = list(ex.map(lambda w: self.o.fuzzer(blobs, n_cases=w[1], first_case=w[0], **kw), wins))
        return [x for p in parts for x in p[0]], [m for p in parts for m in p[1]]


def test_c2_exact_shape_2000_cases(engine, oracle):
    """BASELINE config C2 as it is benched, 2 000 cases of it: 4 096-byte uniform blocks (corpus seed 0xE21A0002), the whole default
    mutator table (every code with a device implementation) at default priorities, all ten patterns at default priorities"""
    import erlamsa_b200
    muts = {c: p for c, p in erlamsa_b200.default_mutations() if c in erlamsa_b200.supported_mutations()}
    pats = dict(erlamsa_b200.default_patterns())
    blobs = corpus.uniform_corpus(0xE21A0002, 2000, 4096, "bin")
    n = compare(engine, ThreadedOracle(oracle), blobs, muts, pats, seed=(1, 2, 3), allow_unsupported=True, capacity_reasons=ALL_CAPACITY_REASONS)
    assert n >= 1700


def test_c4_size_documents(engine, oracle):
    """BASELINE config C4's document size: 262 144-byte SGML and JSON documents (32 distinct, two cases each) under C4's mutator set"""
    muts = {c: 1 for c in ("ab", "ad", "tr2", "td", "ts1", "ts2", "tr", "sgm", "js")}
    blobs = corpus.uniform_corpus(0xE21A0004, 32, 262144, "markup")
    assert all(len(b) == 262144 for b in blobs)
    n = compare(engine, ThreadedOracle(oracle), blobs, muts, {"od": 1}, seed=(1, 2, 3), n_cases=64, allow_unsupported=True, capacity_reasons=ALL_CAPACITY_REASONS)
    assert n >= 56
