"""The Erlang NIF shim (erlang/erlamsa_b200_nif.c) compiled against a mock erl_nif.h (erlang/mock/) and driven from C:
CPU: it compiles, links against the C ABI, and refuses to load without a GPU (the engine has no CPU fallback);
GPU: fuzz_batch_nif/10 returns, for every case, the bytes the Python binding of the same C ABI returns -- including the
     flagged-case tuples -- and rejects malformed arguments."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BLOBS = [b"hello 100 world\n", b"line one\nline two 42\nline three\n", b"<a href=\"http://x/y\">t</a>", b"{\"k\":[1,2,3],\"s\":\"v\"}", b"AAAABBBBCCCCDDDD 7 8 9",
         b"kittenslartibartfasterthaneelslartibartfastenyourseatbelts", b"(x (Y x))", b"A\n B\n C\n D\n"]


def build(tmp):
    exe = os.path.join(tmp, "nif_harness")
    src = [os.path.join(ROOT, "erlang", "erlamsa_b200_nif.c"), os.path.join(ROOT, "erlang", "mock", "erl_nif_mock.c"), os.path.join(ROOT, "erlang", "mock", "nif_harness.c")]
    libdir = os.path.join(ROOT, "erlamsa_b200")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "erlang", "mock"), "-I" + os.path.join(ROOT, "include")] + src +
                          ["-L" + libdir, "-lerlamsa_b200", "-Wl,-rpath," + libdir, "-lpthread", "-o", exe])
    return exe


def fnv(b):
    h = 1469598103934665603
    for c in b:
        h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def test_nif_compiles_and_refuses_to_load_without_gpu(tmp_path):
    import torch
    exe = build(str(tmp_path))
    r = subprocess.run([exe, "4"], capture_output=True, text=True)
    if not torch.cuda.is_available():
        assert r.returncode == 3 and "load_failed" in r.stdout


@pytest.mark.gpu
def test_nif_returns_what_the_c_abi_returns(tmp_path, engine):
    exe = build(str(tmp_path))
    n = 48
    r = subprocess.run([exe, str(n), "1", "2", "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [ln.split() for ln in r.stdout.strip().splitlines()]
    assert len(lines) == n
    outs, meta = engine.fuzz_batch(BLOBS, {"seed": (1, 2, 3)}, n_cases=n)
    for k, f in enumerate(lines):
        if meta[k].status in (1, 3):
            assert f[1] == "flagged" and int(f[2]) == meta[k].status
        else:
            assert int(f[1]) == len(outs[k]) and int(f[2], 16) == fnv(outs[k]), k
    bad = subprocess.run([exe, "4", "1", "2", "3", "bad"], capture_output=True, text=True)
    assert bad.returncode == 0 and "error badarg" in bad.stdout
