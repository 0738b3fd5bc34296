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
    if kind == "markup":   # C4: SGML / JSON documents, blank-padded to the exact size (both tokenizers skip trailing blanks)
        return [(sgml_doc(r, size) if i % 2 == 0 else json_doc(r, size)).ljust(size, b" ") for i in range(count)]
    f = {"bin": random_bytes, "num": numeric_text, "lines": text_lines}[kind]
    return [f(r, size) for _ in range(count)]


def web_corpus(seed, count):
    """blobs for uri / b64 / js / sgm: URLs inside text, base64 runs fenced by binary bytes (so that a whole
    strlex text chunk decodes), JSON scalars, and small markup / JSON documents"""
    import base64
    r = rng(seed)
    words = [b"alpha", b"beta", b"gamma", b"delta", b"x", b"42", b"lorem", b"ipsum"]
    urls = [b"http://example.com/a/b?q=1", b"file://etc/passwd", b"https://host.tld//x///y/z", b"ftp://h", b"a://", b"s3://bucket/key/with/parts",
            b"file://////", b"://leading", b"gopher://a.b/c"]
    scalars = [b"12345", b"-17", b" 0", b"+5", b"true", b"false", b"null", b'"a string value"', b'"unterminated', b"hello", b"1e9", b"007", b'""',
               b"123456789012345678901234567890", b"  null  ", b"3.14"]
    docs = [b"<a>x</a>", b"<p>1<b>2</p>3</b>", b'<r a="1" b=\'2\' c>t<br/><!-- c --></r>', b"<?xml v?><a/>", b"<!-- x", b"<a", b"text < 3 and > 2",
            b'{"a":[1,true,null],"b":"str"}', b"[1,2,3]", b'{"k" 1}', b"[1,2", b"{}", b"[]", b'{"a":{"b":[{"c":"d"}]}}']
    out = []
    for i in range(count):
        k = i % 5
        if k == 0:      # text with URLs
            parts = []
            for _ in range(int(r.integers(1, 12))):
                parts.append(urls[int(r.integers(0, len(urls)))] if r.random() < 0.4 else words[int(r.integers(0, len(words)))])
                parts.append([b" ", b"\n", b"\x00", b"\xfe", b"\"", b"'"][int(r.integers(0, 6))])
            out.append(b"".join(parts))
        elif k == 1:    # base64 runs fenced by non-text bytes
            parts = []
            for _ in range(int(r.integers(1, 6))):
                raw = random_bytes(r, int(r.integers(0, 40))) if r.random() < 0.5 else text_lines(r, int(r.integers(1, 60)))
                enc = base64.b64encode(raw)
                if r.random() < 0.2:
                    enc = enc[:len(enc) // 2] + b"\n" + enc[len(enc) // 2:]      # white space inside is skipped by the decoder
                if r.random() < 0.15:
                    enc = enc[:-1]                                               # broken quantum
                parts.append(enc)
                parts.append(bytes([int(r.integers(0x80, 0x100))]) * int(r.integers(1, 3)))
            out.append(b"".join(parts))
        elif k == 2:
            out.append(scalars[int(r.integers(0, len(scalars)))])
        elif k == 3:
            out.append(docs[int(r.integers(0, len(docs)))])
        else:
            out.append(structured_text(r, int(r.integers(1, 400))))
    return out


def sgml_doc(r, n):
    """random tag tree (depth <= 8, 0-3 attributes, text leaves 8-64 chars), SURVEY.md 8d config C4"""
    names = [b"a", b"div", b"p", b"item", b"Node", b"x:y", b"li", b"span"]
    attrs = [b"id", b"class", b"href", b"xmlns", b"xmlns:xsi", b"data-k", b"checked"]
    out = bytearray(b"<?xml version=\"1.0\"?>")
    stack = []
    letters = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz ABCDEFGH 0123456789.,;-", dtype=np.uint8)
    while len(out) < n:
        k = int(r.integers(0, 10))
        if k < 4 and len(stack) < 8:
            t = names[int(r.integers(0, len(names)))]
            out += b"<" + t
            for _ in range(int(r.integers(0, 4))):
                a = attrs[int(r.integers(0, len(attrs)))]
                v = bytes(letters[r.integers(0, 36, size=int(r.integers(0, 12)))])
                q = int(r.integers(0, 3))
                out += b" " + a + (b"" if q == 2 and not v else b"=" + [b'"', b"'", b""][q] + (v.replace(b" ", b"_") if q == 2 else v) + [b'"', b"'", b""][q])
            if r.random() < 0.15:
                out += b"/>"
            else:
                out += b">"
                stack.append(t)
        elif k < 7 and stack:
            out += b"</" + stack.pop() + b">"
        elif k == 7:
            out += b"<!-- " + bytes(letters[r.integers(0, len(letters), size=int(r.integers(4, 24)))]) + b" -->"
        else:
            out += bytes(letters[r.integers(0, len(letters), size=int(r.integers(8, 65)))])
    while stack and r.random() < 0.9:
        out += b"</" + stack.pop() + b">"
    return bytes(out[:n])


def json_doc(r, n):
    """nested objects / arrays with strings, ints, null / bool until n bytes (C4); always a complete document"""
    letters = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz_ ABCDEF0123456789", dtype=np.uint8)

    def scalar():
        k = int(r.integers(0, 6))
        if k == 0:
            return str(int(r.integers(-10**9, 10**9))).encode()
        if k == 1:
            return [b"true", b"false", b"null"][int(r.integers(0, 3))]
        if k == 2:
            return b'"http://h.example/' + bytes(letters[r.integers(0, 26, size=int(r.integers(1, 10)))]) + b'"'
        return b'"' + bytes(letters[r.integers(0, len(letters), size=int(r.integers(0, 24)))]) + b'"'

    def value(depth, budget):
        k = int(r.integers(0, 10))
        if depth >= 6 or budget < 16 or k < 4:
            return scalar()
        parts = []
        used = 2
        is_obj = k < 7
        while used < budget and (not parts or r.random() < 0.85):
            v = value(depth + 1, (budget - used) // 2)
            item = (b'"' + bytes(letters[r.integers(0, 27, size=int(r.integers(1, 9)))]) + b'":' + (b" " if r.random() < 0.3 else b"") + v) if is_obj else v
            parts.append(item)
            used += len(item) + 1
        sep = b"," if r.random() < 0.7 else b", "
        return (b"{" + sep.join(parts) + b"}") if is_obj else (b"[" + sep.join(parts) + b"]")

    parts = []
    used = 2
    while used < n - 64:
        v = value(1, min(4096, n - used))
        parts.append(v)
        used += len(v) + 1
    doc = b"[" + b",".join(parts) + b"]"
    return doc[:n] if len(doc) <= n else (b"[" + b",".join(parts[:-1]) + b"]")


def markup_corpus(seed, count, size):
    """C4: half SGML, half JSON documents of about `size` bytes"""
    r = rng(seed)
    return [sgml_doc(r, size) if i % 2 == 0 else json_doc(r, size) for i in range(count)]
