"""Config C5: cross-seed donor pool for `fo`.
CPU: pool semantics in the oracle; a gloo world-size-2 run (each rank samples its shard, pools all-gathered, each rank runs
its window of cases) equals the single-process run with the concatenated pool.
GPU: eb200_sample_donors == the host sampling rule; the CUDA engine with an explicit pool == the oracle with the same pool."""
import hashlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

MUTS = {"ft": 2, "fn": 1, "fo": 2}
PATS = {"od": 1, "nd": 1}
STRIDE = 256
D = 16


def _corpus():
    import corpus
    return [corpus.structured_text(corpus.rng(900 + i), 500) for i in range(24)] + [corpus.text_lines(corpus.rng(950 + i), 400) for i in range(8)]


def test_pool_changes_fo_and_only_fo(oracle):
    from erlamsa_b200.donors import sample_windows
    blobs = _corpus()
    wins = sample_windows(blobs, D, STRIDE)
    assert len(wins) == D and all(0 < len(w) <= STRIDE for w in wins)
    base, mb = oracle.fuzzer(blobs, mutations={"fo": 1}, patterns={"od": 1}, seed=(1, 2, 3), n_cases=32)
    pooled, mp_ = oracle.fuzzer(blobs, opts=oracle.make_opts(mutations={"fo": 1}, patterns={"od": 1}, seed=(1, 2, 3), donors=(wins, STRIDE)), n_cases=32)
    assert sum(1 for a, b in zip(base, pooled) if a != b) >= 16          # the donor takes part in the splice
    # the pool must not touch anything but fo
    a, _ = oracle.fuzzer(blobs, mutations={"ft": 1, "bd": 1, "num": 1}, patterns=PATS, seed=(4, 5, 6), n_cases=32)
    b, _ = oracle.fuzzer(blobs, opts=oracle.make_opts(mutations={"ft": 1, "bd": 1, "num": 1}, patterns=PATS, seed=(4, 5, 6), donors=(wins, STRIDE)), n_cases=32)
    assert a == b


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_cases, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    from erlamsa_b200.sharding import shard_window
    from erlamsa_b200.donors import sample_windows, all_gather_pool
    blobs = _corpus()
    half = len(blobs) // world
    shard = blobs[rank * half:(rank + 1) * half]                       # this rank's corpus shard
    wins = sample_windows(shard, D, STRIDE)
    pool = torch.zeros((D, STRIDE), dtype=torch.uint8)
    for i, w in enumerate(wins):
        pool[i, :len(w)] = torch.frombuffer(bytearray(w), dtype=torch.uint8)
    lens = torch.tensor([len(w) for w in wins], dtype=torch.int32)
    gp, gl = all_gather_pool(pool, lens)                                # the exchange step (gloo here, NCCL on the GPU box)
    gwins = [bytes(gp[i, :int(gl[i])].tolist()) for i in range(gp.shape[0])]
    first, cnt = shard_window(n_cases, rank, world, first_case=1)
    outs, _ = O.fuzzer(blobs, opts=O.make_opts(mutations=MUTS, patterns=PATS, seed=(1, 2, 3), donors=(gwins, STRIDE)), n_cases=cnt, first_case=first)
    digest = [int.from_bytes(hashlib.sha256(o).digest()[:7], "big") for o in outs]
    t = torch.zeros(n_cases, dtype=torch.int64)
    t[first - 1:first - 1 + cnt] = torch.tensor(digest, dtype=torch.int64)
    dist.all_reduce(t)
    if rank == 0:
        ret.put(t.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_donor_exchange_equals_single_process(oracle):
    from erlamsa_b200.donors import sample_windows
    n_cases = 48
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_cases, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    blobs = _corpus()
    half = len(blobs) // 2
    gwins = sample_windows(blobs[:half], D, STRIDE) + sample_windows(blobs[half:2 * half], D, STRIDE)    # rank order
    outs, _ = oracle.fuzzer(blobs, opts=oracle.make_opts(mutations=MUTS, patterns=PATS, seed=(1, 2, 3), donors=(gwins, STRIDE)), n_cases=n_cases)
    assert got == [int.from_bytes(hashlib.sha256(o).digest()[:7], "big") for o in outs]


@pytest.mark.gpu
def test_engine_with_pool_matches_oracle(engine, oracle):
    from erlamsa_b200.donors import sample_windows
    blobs = _corpus()
    data = b"".join(blobs)
    dev = torch.device("cuda", 0)
    d_data = torch.frombuffer(bytearray(data + b"\0" * 64), dtype=torch.uint8).to(dev)
    offs = [0]
    for b in blobs:
        offs.append(offs[-1] + len(b))
    d_off = torch.tensor(offs, dtype=torch.int64, device=dev)
    d_pool = torch.zeros((D, STRIDE), dtype=torch.uint8, device=dev)
    d_len = torch.zeros((D,), dtype=torch.int32, device=dev)
    engine.sample_donors(d_data.data_ptr(), d_off.data_ptr(), len(blobs), D, STRIDE, d_pool.data_ptr(), d_len.data_ptr())
    torch.cuda.synchronize()
    wins = sample_windows(blobs, D, STRIDE)
    got = [bytes(d_pool[i, :int(d_len[i])].cpu().tolist()) for i in range(D)]
    assert got == wins                                                   # sampling kernel == host rule
    n = 64
    want, wmeta = oracle.fuzzer(blobs, opts=oracle.make_opts(mutations=MUTS, patterns=PATS, seed=(1, 2, 3), donors=(wins, STRIDE), max_case_out=1 << 22), n_cases=n)
    outs, meta = engine.fuzz_batch(blobs, {"mutations": MUTS, "patterns": PATS, "seed": (1, 2, 3), "max_case_out": 1 << 22,
                                           "donor_pool": (d_pool.data_ptr(), d_len.data_ptr(), D, STRIDE)}, n_cases=n)
    assert [m.status for m in meta] == [m.status for m in wmeta]
    bad = [k for k in range(n) if meta[k].status == 0 and (outs[k] != want[k] or meta[k].draws != wmeta[k].draws)]
    assert not bad, bad[:8]
    base, _ = engine.fuzz_batch(blobs, {"mutations": MUTS, "patterns": PATS, "seed": (1, 2, 3), "max_case_out": 1 << 22}, n_cases=n)
    assert sum(1 for a, b in zip(base, outs) if a != b) >= 8              # the pool really takes part
