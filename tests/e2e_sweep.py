"""e2e (host buffers) sweep of chunk size / upload look-ahead on the C3 shape (not a test; run on the GPU box).
usage: python tests/e2e_sweep.py [cases]"""
import ctypes as C
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
size = 65536
print(subprocess.run("nvidia-smi topo -m | head -12; lscpu | grep -i numa", shell=True, capture_output=True, text=True).stdout)
for chunk_mb, ahead in ((128, 1), (64, 2), (32, 2), (16, 3), (256, 2), (64, 4)):
    os.environ["EB200_CHUNK_MB"] = str(chunk_mb); os.environ["EB200_H2D_AHEAD"] = str(ahead)
    import importlib
    import erlamsa_b200
    from erlamsa_b200 import _native as N
    eng = erlamsa_b200.Engine(0)
    L = N.lib()
    in_b, out_b = n * size + 64, n * size + n * size // 12 + 512 * n + (128 << 20)
    p_in, p_out = L.eb200_host_alloc(eng._ctx, in_b), L.eb200_host_alloc(eng._ctx, out_b)
    hoff = (C.c_uint64 * (n + 1))(*[i * size for i in range(n + 1)])
    ho_off = (C.c_uint64 * (n + 1))(); ho_len = (C.c_uint64 * n)()
    st = N.Stats()
    o = erlamsa_b200.make_opts({"mutations": {c: 1 for c in ("bd", "bei", "bed", "bf", "bi", "ber", "br", "num")}, "patterns": {"od": 1}, "seed": (1, 2, 3), "scratch_bytes": 512 << 20})
    ts = []
    for i in range(3):
        o.first_case = 1 + i * n
        t0 = time.perf_counter()
        rc = L.eb200_fuzz_batch_into(eng._ctx, C.byref(o), p_in, C.cast(hoff, C.c_void_p), n, n, p_out, out_b, C.cast(ho_off, C.c_void_p), C.cast(ho_len, C.c_void_p), None, C.byref(st))
        ts.append(time.perf_counter() - t0)
        assert rc == 0, rc
    dt = min(ts[1:])
    print("chunk %4d MB ahead %d: %.1f ms  %.0f cases/s  %.1f GB/s each way (numa node %d)" % (chunk_mb, ahead, dt * 1e3, n / dt, n * size / dt / 1e9, L.eb200_numa_node(eng._ctx)), flush=True)
    L.eb200_host_free(eng._ctx, p_in); L.eb200_host_free(eng._ctx, p_out)
    eng.close()
