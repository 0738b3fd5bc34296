"""Run in a subprocess with EB200_LIB = the mock library (tests/test_async_harness.py): the host mirror's multi-threaded mode end to end --
plan, one batch per worker range with its stream seed, file numbering, skip -- against a mock engine whose output encodes the case number
and the worker stream it was given."""
import os
import sys
import tempfile

from erlamsa_b200 import _native as N
from erlamsa_b200 import erlamsa_main
from erlamsa_b200.workers import worker_plan

assert b"MOCK" in N.lib().eb200_version()
files = [bytes([i]) * (10 + i) for i in range(1, 4)]
with tempfile.TemporaryDirectory() as tmp:
    paths = []
    for i, b in enumerate(files):
        p = os.path.join(tmp, "seed%d" % i)
        open(p, "wb").write(b)
        paths.append(p)
    for n, w, same, skip in ((11, 3, False, 0), (10, 3, True, 0), (12, 4, False, 5), (5, 8, False, 0)):
        out = os.path.join(tmp, "out_%d_%d_%d" % (n, w, skip))
        res = erlamsa_main.fuzzer({"paths": paths, "output": os.path.join(out, "%n"), "n": n, "seed": (1, 2, 3), "workers": w, "workers_same_seed": same, "skip": skip})
        assert res == []
        want = {}
        for wseed, first, cnt, stream_first in worker_plan((1, 2, 3), "x/%n", n, w, same):
            for k in range(cnt):
                i = first + k
                x = (i + wseed[0] + 7 * stream_first) & 255
                want[i] = bytes(c ^ x for c in files[(i - 1) % 3])
        assert sorted(want) == list(range(1, n + 1))
        got = {int(f): open(os.path.join(out, f), "rb").read() for f in os.listdir(out)}
        assert got == {i: b for i, b in want.items() if i > skip}, (n, w, same, skip)
    # single-threaded shapes keep the plain numbering (no stream override reaches the engine)
    out = os.path.join(tmp, "single")
    erlamsa_main.fuzzer({"paths": paths, "output": os.path.join(out, "%n"), "n": 4, "seed": (1, 2, 3), "workers": 1})
    assert {int(f): open(os.path.join(out, f), "rb").read() for f in os.listdir(out)} == {i: bytes(c ^ (i & 255) for c in files[(i - 1) % 3]) for i in range(1, 5)}
    assert erlamsa_main.fuzzer({"paths": ["direct"], "input": b"abc", "output": "return", "n": 6, "seed": (1, 2, 3), "workers": 3}) == [bytes(c ^ i for c in b"abc") for i in range(1, 7)]
print("OK")
