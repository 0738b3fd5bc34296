"""Per-pattern device time on the C2 corpus with the default mutators (not a test; run on the GPU box).
usage: python tests/time_patterns.py [n_cases]   -- every pattern runs in its own process under a timeout"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
if len(sys.argv) > 2:
    sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
    import time
    import corpus, erlamsa_b200
    pat, n = sys.argv[1], int(sys.argv[2])
    muts = dict(erlamsa_b200.default_mutations()) if len(sys.argv) < 4 else {c: 1 for c in sys.argv[3].split(",")}
    eng = erlamsa_b200.Engine(0)
    blobs = corpus.uniform_corpus(0xE21A0003, min(n, 2000), 4096, "bin")
    for rep in range(2):
        t = time.time()
        outs, meta = eng.fuzz_batch(blobs, {"mutations": muts, "patterns": {pat: 1}, "seed": (1, 2, 3), "max_case_out": 1 << 20,
                                            "scratch_bytes": 4 << 30, "first_case": 1 + rep * n}, n_cases=n)
        dt = time.time() - t
    import collections
    c = collections.Counter(m.status for m in meta)
    print("%-3s %6d cases %8.1f ms host wall  status %s" % (pat, n, dt * 1e3, dict(c)), flush=True)
else:
    n = sys.argv[1] if len(sys.argv) > 1 else "2000"
    specs = [("od", "len"), ("od", "ft"), ("od", "fo"), ("od", "fn"), ("sz", "bd"), ("cs", "bd"), ("nd", "bd,sr,num,sp,ab,td"), ("nd", None), ("sk", None)]
    for pat, muts in specs:
        try:
            r = subprocess.run([sys.executable, __file__, pat, n] + ([muts] if muts else []), capture_output=True, text=True, timeout=45)
            print((muts or "default").ljust(20), (r.stdout.strip() or r.stderr.strip()[-300:]), flush=True)
        except subprocess.TimeoutExpired:
            print("%-3s %s TIMEOUT (45 s) -- stopping" % (pat, muts), flush=True)
            break
