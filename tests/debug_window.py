import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import corpus, erlamsa_b200
from erlamsa_b200 import _native as N
eng = erlamsa_b200.Engine(0)
blobs = corpus.mixed_corpus(0xE21A0300, 37)
muts = {c: 1 for c in ("bd", "bf", "num", "sr", "ld")}
o = erlamsa_b200.make_opts({"mutations": muts, "patterns": {"od": 1, "nd": 1}, "seed": (1,2,3), "first_case": 1000, "max_case_out": 1 << 28})
n_cases = 150
data = b"".join(blobs)
off = (C.c_uint64 * (len(blobs) + 1))(); acc = 0
for i, b in enumerate(blobs): off[i] = acc; acc += len(b)
off[len(blobs)] = acc
buf = C.create_string_buffer(data, len(data) + 1)
out_p = C.c_void_p(); out_off = (C.c_uint64 * (n_cases + 1))(); out_len = (C.c_uint64 * n_cases)(); meta = (N.Meta * n_cases)(); st = N.Stats()
rc = N.lib().eb200_fuzz_batch(eng._ctx, C.byref(o), C.cast(buf, C.c_void_p), off, len(blobs), n_cases, C.byref(out_p), out_off, out_len, meta, C.byref(st))
print("rc", rc, "total", out_off[n_cases])
for k in range(n_cases):
    if out_len[k] > (1 << 30) or out_off[k] > (1 << 40) or meta[k].status:
        print(k, out_off[k], out_len[k], meta[k].status, meta[k].pad, meta[k].draws, list(meta[k].used)[:6])
