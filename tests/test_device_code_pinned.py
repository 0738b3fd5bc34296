"""The device code that ships is the device code the GPU suite ran: per-kernel SASS digests of the in-tree library against
profiles/sass_digest_r2.txt (tools/sass_digest.py). A change to a .cu / .cuh file that alters a shipped kernel fails here until the GPU
suite has been re-run on the new build and the digests refreshed with `python tools/sass_digest.py --write` -- host-side changes pass.
(nvcc of the image is deterministic: the same sources give the same SASS.)"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_shipped_kernels_match_the_gpu_tested_build():
    import sass_digest as S
    if not os.path.exists(S.CUOBJDUMP) or not os.path.exists(S.LIB):
        pytest.skip("no cuobjdump / library not built")
    try:
        got = S.digests()
    except subprocess.CalledProcessError:
        pytest.skip("cuobjdump cannot read the library")
    if S.ref_nvcc() != S.nvcc_release():
        pytest.skip("digests were taken with nvcc %s, this toolchain is %s" % (S.ref_nvcc(), S.nvcc_release()))
    ref = S.read_ref()
    assert set(got) == set(ref), sorted(set(got) ^ set(ref))
    bad = [k for k in sorted(ref) if got[k] != ref[k]]
    assert not bad, "device code changed since the last GPU-tested build: %r" % bad
