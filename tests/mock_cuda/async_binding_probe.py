"""Run in a subprocess with EB200_LIB = the mock library (tests/test_async_harness.py): drives Engine.submit_device / Engine.collect
exactly as tests/test_zzz_async_gpu.py does on the GPU, with numpy arrays standing in for device memory."""
import ctypes as C
import sys

import numpy as np

import erlamsa_b200
from erlamsa_b200 import _native as N

assert b"MOCK" in N.lib().eb200_version(), "this probe must never run against the real engine"
eng = erlamsa_b200.Engine(0)
blobs = [bytes([i % 251] * (1 + (i * 37) % 300)) for i in range(200)]
data = np.frombuffer(b"".join(blobs) + b"\0" * 64, dtype=np.uint8).copy()
offs = np.zeros(len(blobs) + 1, dtype=np.uint64)
offs[1:] = np.cumsum([len(b) for b in blobs])
nbytes, n, nb = int(offs[-1]), len(blobs), 5
cap = 2 * nbytes + 16 * n + 4096
msz = C.sizeof(N.Meta)


def aligned(nbytes_):
    raw = np.zeros(nbytes_ + 64, dtype=np.uint8)
    sh = (-raw.ctypes.data) % 16
    return raw[sh:sh + nbytes_]


def bufs():
    return aligned(cap), np.zeros(n + 1, dtype=np.uint64), np.zeros(n, dtype=np.uint64), np.zeros(n * msz, dtype=np.uint8)


def opts(b):
    return {"mutations": {"bd": 1, "num": 3}, "patterns": {"od": 1, "nd": 1}, "seed": (3 + b, 1, 4), "first_case": 1 + b * n, "max_case_out": 1 << 20}


def cases(o, f, l):
    return [bytes(o[int(f[k]):int(f[k]) + int(l[k])]) for k in range(n)]


din = aligned(len(data)); din[:] = data
want = []
for b in range(nb):
    o, f, l, m = bufs()
    st = eng.fuzz_batch_device(opts(b), din.ctypes.data, offs.ctypes.data, n, nbytes, n, o.ctypes.data, cap, f.ctypes.data, l.ctypes.data, m.ctypes.data, 0)
    want.append((cases(o, f, l), m.tobytes(), st.n_cases))
    assert want[-1][0][7] == bytes(x ^ ((3 + b) & 255) for x in blobs[(b * n + 7) % n])
held = [bufs() for _ in range(nb)]
tickets = [eng.submit_device(opts(b), din.ctypes.data, offs.ctypes.data, n, nbytes, n, held[b][0].ctypes.data, cap, held[b][1].ctypes.data,
                             held[b][2].ctypes.data, held[b][3].ctypes.data) for b in range(nb)]
assert N.lib().eb200_async_lanes(eng._ctx) == 2
for b in (3, 0, 4, 1, 2):
    st = eng.collect(tickets[b])
    assert st.n_cases == want[b][2] and st.kernels_launched == 5
    o, f, l, m = held[b]
    assert cases(o, f, l) == want[b][0] and m.tobytes() == want[b][1], b
try:
    eng.collect(tickets[0])
    sys.exit("a ticket was collected twice")
except erlamsa_b200.EngineError as e:
    assert e.code == -2
t = eng.submit_device(opts(0), din.ctypes.data, offs.ctypes.data, n, nbytes, n, held[0][0].ctypes.data, 10, held[0][1].ctypes.data, held[0][2].ctypes.data, 0)
try:
    eng.collect(t)
    sys.exit("an output arena of 10 bytes was accepted")
except erlamsa_b200.EngineError as e:
    assert e.code == -4
eng.close()
print("OK")
