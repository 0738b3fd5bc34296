"""Seeded synthetic corpora shared by the parity tests and bench.py (numpy PCG64)."""
import numpy as np


def rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def random_bytes(r, n):
    return r.integers(0, 256, size=n, dtype=np.uint8).tobytes()


def numeric_text(r, n):
    """printable ASCII with a 1-12 digit decimal (5% negative) every ~32 bytes (BASELINE C3 ii)"""
    out = bytearray()
    while len(out) < n:
        k = int(r.integers(8, 48))
        out += bytes(r.integers(0x20, 0x7f, size=k, dtype=np.uint8))
        if r.random() < 0.1:
            out += b"\n"
        d = int(r.integers(1, 13))
        num = int(r.integers(0, 10 ** d))
        out += (b" -" if r.random() < 0.05 else b" ") + str(num).encode() + b" "
    return bytes(out[:n])


def text_lines(r, n):
    out = bytearray()
    while len(out) < n:
        k = int(r.integers(0, 60))
        out += bytes(r.integers(0x20, 0x7f, size=k, dtype=np.uint8)) + b"\n"
    return bytes(out[:n])


def structured_text(r, n):
    """text with quotes, brackets, numbers, URIs and base64-looking runs (exercises strlex / tree / fuse / uri)"""
    words = [b"alpha", b"beta", b"kitten", b"slartibartfast", b"x", b"42", b"-7", b"http://example.com/a/b?q=1", b"file://etc/passwd",
             b"aGVsbG8gd29ybGQ=", b"QUJDREVGR0g=", b"lorem ipsum", b"", b"\\", b"a\\\"b", b"%s", b"\t"]
    opens = [(b"(", b")"), (b"[", b"]"), (b"<", b">"), (b"{", b"}"), (b'"', b'"'), (b"'", b"'")]
    out = bytearray()
    depth = []
    while len(out) < n:
        k = int(r.integers(0, 10))
        if k < 4:
            out += words[int(r.integers(0, len(words)))] + (b" " if r.random() < 0.7 else b"")
        elif k < 6:
            o, c = opens[int(r.integers(0, len(opens)))]
            out += o
            depth.append(c)
        elif k < 8 and depth:
            out += depth.pop()
        elif k == 8:
            out += b"\n"
        else:
            out += bytes(r.integers(0x20, 0x7f, size=int(r.integers(1, 9)), dtype=np.uint8))
        if r.random() < 0.01:
            out += bytes([int(r.integers(0, 256))])          # a stray binary byte
    while depth and r.random() < 0.7:
        out += depth.pop()
    return bytes(out[:n]) if r.random() < 0.5 else bytes(out)


def text_corpus(seed, count, max_len=1500):
    r = rng(seed)
    return [structured_text(r, int(r.integers(1, max_len))) for _ in range(count)]


def mixed_corpus(seed, count, max_len=5000, kinds=("bin", "num", "lines", "tiny")):
    r = rng(seed)
    blobs = []
    for i in range(count):
        kind = kinds[i % len(kinds)]
        if kind == "tiny":
            n = int(r.integers(0, 12))
            blobs.append(random_bytes(r, n) if i % 8 else text_lines(r, n))
            continue
        n = int(r.integers(1, max_len))
        blobs.append({"bin": random_bytes, "num": numeric_text, "lines": text_lines}[kind](r, n))
    return blobs


def uniform_corpus(seed, count, size, kind="bin"):
    r = rng(seed)
    f = {"bin": random_bytes, "num": numeric_text, "lines": text_lines}[kind]
    return [f(r, size) for _ in range(count)]
