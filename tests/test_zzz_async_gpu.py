"""GPU: the asynchronous pair returns, for every case, exactly what the synchronous device call returns (it runs the same entry
point on a lane's own context and stream), with several batches in flight and collected out of order. Sorted last on purpose."""
import ctypes as C

import pytest

import corpus

# the timeout's thread method ends the run instead of hanging it if a lane never comes back (a blocked C call cannot be interrupted by a signal)
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]


def _pack(torch, blobs, dev):
    data = b"".join(blobs)
    t = torch.zeros(len(data) + 64, dtype=torch.uint8, device=dev)
    t[:len(data)] = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
    offs, acc = [0], 0
    for b in blobs:
        acc += len(b)
        offs.append(acc)
    return t, torch.tensor(offs, dtype=torch.int64, device=dev), len(data)


def _cases(out, off, ln, n):
    o, f, l = out.cpu().numpy(), off.cpu().numpy(), ln.cpu().numpy()
    return [bytes(o[int(f[k]):int(f[k]) + int(l[k])]) for k in range(n)]


def test_async_pair_equals_synchronous_call(engine):
    import torch
    from erlamsa_b200 import _native as N
    dev = torch.device("cuda:0")
    blobs = corpus.mixed_corpus(0xA51C, 600, max_len=3000)
    n = len(blobs)
    data, off, nbytes = _pack(torch, blobs, dev)
    # no repeat mutators: every result stays near its input size, so no case depends on how full the shared overflow region is
    muts = {"bd": 1, "bf": 1, "bi": 1, "ber": 1, "num": 3, "ld": 1, "lr2": 1, "sd": 1, "ui": 1}
    nb = 5
    cap = 4 * nbytes + (64 << 20)
    msz = C.sizeof(N.Meta)

    def bufs():
        return (torch.zeros(cap, dtype=torch.uint8, device=dev), torch.zeros(n + 1, dtype=torch.int64, device=dev),
                torch.zeros(n, dtype=torch.int64, device=dev), torch.zeros(n * msz, dtype=torch.uint8, device=dev))

    def opts(b):
        return {"mutations": muts, "patterns": {"od": 1, "nd": 1}, "seed": (3, 1, 4), "first_case": 1 + b * n, "max_case_out": 1 << 20}

    want = []
    for b in range(nb):
        o, f, l, m = bufs()
        st = engine.fuzz_batch_device(opts(b), data.data_ptr(), off.data_ptr(), n, nbytes, n, o.data_ptr(), cap, f.data_ptr(), l.data_ptr(), m.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        want.append((_cases(o, f, l, n), m.cpu().numpy().tobytes(), st.n_cases))
    torch.cuda.synchronize()
    held = [bufs() for _ in range(nb)]
    torch.cuda.synchronize()          # the zero fills above ran on torch's stream; the lanes use their own
    tickets = [engine.submit_device(opts(b), data.data_ptr(), off.data_ptr(), n, nbytes, n, held[b][0].data_ptr(), cap, held[b][1].data_ptr(),
                                    held[b][2].data_ptr(), held[b][3].data_ptr()) for b in range(nb)]
    assert N.lib().eb200_async_lanes(engine._ctx) >= 1
    for b in (3, 0, 4, 1, 2):
        st = engine.collect(tickets[b])
        assert st.n_cases == want[b][2] and st.kernels_launched >= 5
        o, f, l, m = held[b]
        assert _cases(o, f, l, n) == want[b][0], "batch %d differs between the asynchronous and the synchronous call" % b
        assert m.cpu().numpy().tobytes() == want[b][1]
    with pytest.raises(Exception):
        engine.collect(tickets[0])       # a ticket is collected once
