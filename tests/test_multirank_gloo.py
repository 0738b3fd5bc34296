"""N > 1 host logic on CPU (gloo, world_size 2): the case loop is cut into per-rank windows (erlamsa_b200.sharding),
each rank runs only its window, and the gathered result equals the single-process run -- the property that makes the
multi-GPU path collective-free. The per-rank compute here is the oracle (no GPU in this container); on the GPU box the
same windows drive the CUDA engine (bench.py --gpus N; tests/test_parity_gpu.py checks window independence on one GPU)."""
import hashlib
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_cases, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import corpus
    import oracle_lib as O
    from erlamsa_b200.sharding import shard_window
    blobs = corpus.mixed_corpus(0xD15, 40, 800)
    muts = {"bd": 1, "bf": 1, "num": 2, "sr": 1, "ld": 1, "lis": 1}
    first, cnt = shard_window(n_cases, rank, world, first_case=1)
    outs, _ = O.fuzzer(blobs, mutations=muts, patterns={"od": 1, "nd": 1}, seed=(1, 2, 3), n_cases=cnt, first_case=first)
    digest = [int.from_bytes(hashlib.sha256(o).digest()[:7], "big") for o in outs]
    t = torch.zeros(n_cases, dtype=torch.int64)
    t[first - 1:first - 1 + cnt] = torch.tensor(digest, dtype=torch.int64)
    dist.all_reduce(t)                      # test-only gather; the data path itself needs no collective
    if rank == 0:
        ret.put(t.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_windows_equal_single_process():
    import corpus
    import oracle_lib as O
    n_cases = 101
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_cases, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    blobs = corpus.mixed_corpus(0xD15, 40, 800)
    muts = {"bd": 1, "bf": 1, "num": 2, "sr": 1, "ld": 1, "lis": 1}
    outs, _ = O.fuzzer(blobs, mutations=muts, patterns={"od": 1, "nd": 1}, seed=(1, 2, 3), n_cases=n_cases, first_case=1)
    want = [int.from_bytes(hashlib.sha256(o).digest()[:7], "big") for o in outs]
    assert got == want
