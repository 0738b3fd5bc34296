"""eb200_submit_device / eb200_collect (erlamsa_b200/csrc/eb_async.cpp) on the CPU: the file is host code over the engine's
synchronous entry point, so it is compiled against a stand-in cuda_runtime.h and a mock engine (tests/mock_cuda/) and its ticket
protocol, lane overlap, error propagation and teardown are exercised -- once plainly, once under ThreadSanitizer.
The GPU side (async results == synchronous results, bit for bit) is tests/test_zzz_async_gpu.py."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "tests", "mock_cuda")
SRC = [os.path.join(ROOT, "erlamsa_b200", "csrc", "eb_async.cpp"), os.path.join(MOCK, "async_harness.cpp")]


def build(tmp, extra):
    exe = os.path.join(tmp, "async_harness" + ("_tsan" if extra else ""))
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-Wall", "-Werror", "-I" + MOCK, "-I" + os.path.join(ROOT, "include")] + extra + SRC + ["-lpthread", "-o", exe])
    return exe


def test_async_lanes_protocol(tmp_path):
    r = subprocess.run([build(str(tmp_path), [])], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr


def test_async_lanes_under_thread_sanitizer(tmp_path):
    try:
        exe = build(str(tmp_path), ["-fsanitize=thread"])
    except subprocess.CalledProcessError:
        pytest.skip("no ThreadSanitizer runtime in this toolchain")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    if "FATAL: ThreadSanitizer" in r.stderr and "unexpected memory mapping" in r.stderr:
        pytest.skip("ThreadSanitizer cannot map its shadow memory on this kernel")
    assert r.returncode == 0 and "OK" in r.stdout and "WARNING: ThreadSanitizer" not in r.stderr, r.stdout + r.stderr


def test_python_binding_of_the_async_pair_against_a_mock_library(tmp_path):
    """Engine.submit_device / Engine.collect (ctypes signatures, ticket handling, error mapping) driven the way the GPU test drives
    them, against a stand-in library = the real eb_async.cpp over a mock engine whose device memory is host memory"""
    import sys
    lib = os.path.join(str(tmp_path), "libmock_erlamsa_b200.so")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-shared", "-fPIC", "-Wall", "-Werror", "-I" + MOCK, "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "erlamsa_b200", "csrc", "eb_async.cpp"), os.path.join(MOCK, "mock_lib.cpp"), "-lpthread", "-o", lib])
    env = dict(os.environ, EB200_LIB=lib, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, os.path.join(MOCK, "async_binding_probe.py")], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
    # the same stand-in library under the host mirror's multi-threaded mode (erlamsa_main.py + workers.py): plan, batches, file numbering, skip
    r = subprocess.run([sys.executable, os.path.join(MOCK, "workers_mirror_probe.py")], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
