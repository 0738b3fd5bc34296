#!/usr/bin/env python
"""Per-kernel digests of the SASS of the built library (cuobjdump -sass, addresses and encodings stripped): the way to say that a
host-side change left the device code alone, or which kernels a change touched.
  python tools/sass_digest.py                      print "<sha256-16> <instructions> <kernel>" for the in-tree library
  python tools/sass_digest.py --write              refresh profiles/sass_digest_r2.txt (do this only after the GPU suite passed on that build)
  python tools/sass_digest.py --check              compare with profiles/sass_digest_r2.txt, exit 1 on a difference"""
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "erlamsa_b200", "liberlamsa_b200.so")
REF = os.path.join(ROOT, "profiles", "sass_digest_r2.txt")
CUOBJDUMP = os.environ.get("CUOBJDUMP", "/usr/local/cuda/bin/cuobjdump")


def digests(lib=LIB):
    out = subprocess.run([CUOBJDUMP, "-sass", lib], capture_output=True, text=True, check=True).stdout
    d, cur = {}, None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            d[cur] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?;)", ln)
        if cur and m:
            d[cur].append(m.group(1))
    return {k: (hashlib.sha256("\n".join(v).encode()).hexdigest()[:16], len(v)) for k, v in d.items()}


def nvcc_release():
    try:
        out = subprocess.run([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "--version"], capture_output=True, text=True).stdout
        m = re.search(r"release [0-9.]+, V([0-9.]+)", out)
        return m.group(1) if m else "unknown"
    except OSError:
        return "unknown"


def ref_nvcc():
    for ln in open(REF):
        m = re.match(r"# nvcc (\S+)", ln)
        if m:
            return m.group(1)
    return None


def read_ref():
    ref = {}
    for ln in open(REF):
        if ln.startswith("#") or not ln.strip():
            continue
        h, n, k = ln.split()
        ref[k] = (h, int(n))
    return ref


def main():
    d = digests()
    if "--write" in sys.argv:
        with open(REF, "w") as f:
            f.write("# sha256[:16] of the SASS text, instruction count, kernel -- of the build whose GPU test run is recorded in the round's GPUTEST / profiles\n")
            f.write("# nvcc %s\n" % nvcc_release())
            for k in sorted(d):
                f.write("%s %d %s\n" % (d[k][0], d[k][1], k))
        return 0
    if "--check" in sys.argv:
        ref = read_ref()
        bad = [k for k in sorted(set(d) | set(ref)) if d.get(k) != ref.get(k)]
        for k in bad:
            print("DIFFERS", k, d.get(k), ref.get(k))
        return 1 if bad else 0
    for k in sorted(d):
        print(d[k][0], d[k][1], k)
    return 0


if __name__ == "__main__":
    sys.exit(main())
