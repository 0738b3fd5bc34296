"""A/B timing of engine configurations on the C3 workload (not a test; run on the GPU box).
usage: python tests/perf_variants.py "deciders:threads[:fronts[:depth[:tma_workers]]],..." [cases]
Every configuration is (EB200_DECIDERS, EB200_THREADS) of eb_case_kernel: how many of the CTA's warps decide cases, the
rest being copy/scan workers (deciders == threads/32 is the round-1 arrangement: every warp does its own byte work)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import erlamsa_b200  # noqa: E402

configs = [tuple(int(x) for x in c.split(":")) for c in (sys.argv[1] if len(sys.argv) > 1 else "8:1024,12:1024,16:1024,32:1024").split(",")]
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
size = 65536
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
data = torch.randint(0, 256, (n_cases * size + 64,), dtype=torch.uint8, device=dev, generator=g)
off = torch.arange(0, (n_cases + 1) * size, size, dtype=torch.int64, device=dev)
out_cap = n_cases * size + n_cases * size // 12 + 512 * n_cases + (256 << 20)
d_out = torch.empty(out_cap, dtype=torch.uint8, device=dev)
d_off = torch.empty(n_cases + 1, dtype=torch.int64, device=dev)
d_len = torch.empty(n_cases, dtype=torch.int64, device=dev)
muts = {c: 1 for c in os.environ.get("PERF_MUTS", "bd,bei,bed,bf,bi,ber,br,num").split(",")}
for cfg in configs:
    dec, thr = cfg[0], cfg[1]
    fr = cfg[2] if len(cfg) > 2 else 1
    os.environ["EB200_FRONT_DEPTH"] = str(cfg[3] if len(cfg) > 3 else 32)
    os.environ["EB200_TMA_WORKERS"] = str(cfg[4] if len(cfg) > 4 else 0)
    os.environ["EB200_DECIDERS"] = str(dec)
    os.environ["EB200_THREADS"] = str(thr)
    os.environ["EB200_FRONTS"] = str(fr)
    eng = erlamsa_b200.Engine(0)
    res = []
    for i in range(int(os.environ.get("PERF_ITERS", "7"))):
        st = eng.fuzz_batch_device({"mutations": muts, "patterns": {"od": 1}, "seed": (1, 2, 3), "first_case": 1 + i * n_cases, "scratch_bytes": 256 << 20},
                                   data.data_ptr(), off.data_ptr(), n_cases, n_cases * size, n_cases, d_out.data_ptr(), out_cap, d_off.data_ptr(),
                                   d_len.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
        res.append((st.ms_decide, st.ms_scan, st.ms_apply, st.ms_total))
    print("per-step kernel ms:", " ".join("%.3f" % r[0] for r in res), flush=True)
    res = res[2:]
    avg = [sum(r[k] for r in res) / len(res) for k in range(4)]
    gbs = (2 * n_cases * size) / ((avg[2] or avg[0]) * 1e-3) / 1e9
    print("tma %s depth %s fronts %d deciders %d threads %d: kernel %.3f  scan %.3f  apply %.3f ms (%.0f GB/s = %.2f of 6572)  total %.3f  launches %d"
          % (os.environ["EB200_TMA_WORKERS"], os.environ["EB200_FRONT_DEPTH"], fr, dec, thr, avg[0], avg[1], avg[2], gbs, gbs / 6572.2, avg[3], st.kernels_launched), flush=True)
    eng.close()
