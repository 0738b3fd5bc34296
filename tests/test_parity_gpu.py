"""GPU parity: the CUDA engine (through the C ABI) against the oracle, byte for byte, on the same seeded
inputs -- outputs, RNG draw counts, per-case thread seeds, chosen pattern and used-mutator lists.
Run on the B200 box: python -m pytest tests -m gpu"""
import pytest

import corpus

pytestmark = pytest.mark.gpu

SUPPORTED_PATS = {"od": 1, "nd": 2, "bu": 1, "sk": 2, "co": 0, "nu": 0}


CAP = 1 << 20   # per-case output cap given to both sides (runaway repeats are flagged, not compared)


def compare(engine, oracle, blobs, mutations, patterns, seed=(1, 2, 3), n_cases=None, first_case=1, allow_unsupported=False, capacity_reasons=(1, 4, 5, 6)):
    o_out, o_meta = oracle.fuzzer(blobs, mutations=mutations, patterns=patterns, seed=seed, n_cases=n_cases, first_case=first_case, max_case_out=CAP)   # oracle: any object with this method (ThreadedOracle below)
    g_out, g_meta = engine.fuzz_batch(blobs, {"mutations": mutations, "patterns": patterns, "seed": seed, "first_case": first_case, "max_case_out": CAP},
                                      n_cases=n_cases)
    assert len(o_out) == len(g_out)
    bad = []
    n_cmp = 0
    n_big = 0
    for k, (a, b, ma, mb) in enumerate(zip(o_out, g_out, o_meta, g_meta)):
        assert list(ma.thread_seed) == list(mb.thread_seed), "thread seed differs at case %d" % k
        if mb.status == 1 and allow_unsupported:   # engine walked into a path without a device implementation
            continue
        if ma.status != 0:
            continue
        if ma.status == 3 or (mb.status == 3 and mb.pad in capacity_reasons):
            # documented capacity limits: a case that blows up past the output cap (sr/lr repeats compounded by
            # nd/bu rounds) or the engine's run/piece tables is flagged on either side, never silently wrong
            n_big += 1
            continue
        n_cmp += 1
        if a != b or ma.draws != mb.draws or ma.pattern != mb.pattern or ma.n_used != mb.n_used or list(ma.used) != list(mb.used) \
                or mb.status != 0:
            bad.append((k, len(blobs[(first_case - 1 + k) % len(blobs)]), len(a), len(b), ma.draws, mb.draws, ma.pattern, mb.pattern,
                        list(ma.used)[:4], list(mb.used)[:4], mb.status, mb.pad))
    assert not bad, "mismatches (case, in_len, oracle_len, gpu_len, o_draws, g_draws, o_pat, g_pat, o_used, g_used, g_status): %r" % bad[:8]
    assert n_big * 8 <= len(o_out), "too many capacity overflows: %d" % n_big
    return n_cmp


@pytest.mark.parametrize("code", ["bd", "bei", "bed", "bf", "bi", "ber", "br", "uw", "ui", "sd", "sr", "sp", "snand", "srnd", "num",
                                  "ld", "lds", "lr2", "lri", "lr", "ls", "lp", "lis", "lrs", "nil"])
def test_single_mutator_od(engine, oracle, code):
    blobs = corpus.mixed_corpus(0xE21A0000 + len(code), 240)
    n = compare(engine, oracle, blobs, {code: 1}, {"od": 1})
    assert n == len(blobs)


@pytest.mark.parametrize("code", ["ab", "ad", "tr2", "td", "ts1", "ts2", "tr", "ft", "fn", "fo"])
def test_structure_mutators_od(engine, oracle, code):
    """strlex / parse-tree / fuse mutators on quoted, bracketed, line-structured text (plus binary and tiny blobs)"""
    blobs = corpus.text_corpus(0xE21A0500 + len(code), 200) + corpus.mixed_corpus(0xE21A0600, 40, 400)
    n = compare(engine, oracle, blobs, {code: 1}, {"od": 1}, seed=(3, 1, 4))
    assert n >= len(blobs) - 4


def test_structure_mutators_multi_round(engine, oracle):
    """closure state (fo's remembered block, lis/lrs slots) and re-chunked block lists across nd / bu rounds"""
    muts = {c: 1 for c in ("ab", "ad", "tr2", "td", "ts1", "ts2", "tr", "ft", "fn", "fo", "num", "lis", "bd")}
    blobs = corpus.text_corpus(0xE21A0700, 300, 900)
    compare(engine, oracle, blobs, muts, {"nd": 2, "bu": 1, "od": 1}, seed=(2, 7, 1), allow_unsupported=True)


def test_uri_mutator(engine, oracle):
    """every text chunk holding "://" is rewritten; the mutator then rebinds itself to b64 for the following rounds"""
    blobs = corpus.web_corpus(0xE21A0900, 300)
    n = compare(engine, oracle, blobs, {"uri": 1}, {"od": 1}, seed=(5, 5, 5))
    assert n >= len(blobs) - 10          # "scheme://" with nothing but slashes behind it kills the case on both sides (badmatch :745,:751)
    compare(engine, oracle, blobs, {"uri": 1, "bd": 1}, {"nd": 1, "bu": 1}, seed=(5, 6, 5), allow_unsupported=True)


def test_b64_mutator_nested_round(engine, oracle):
    """decodable chunks get one nested scheduler round over the whole default table; only a second level of nesting
    (or an sgm/js document inside the decoded bytes) may flag a case"""
    blobs = corpus.web_corpus(0xE21A0901, 300)
    n = compare(engine, oracle, blobs, {"b64": 1}, {"od": 1}, seed=(6, 5, 5), allow_unsupported=True)
    assert n >= len(blobs) * 0.85


def test_sgm_js_refusals_and_scalar_documents(engine, oracle):
    blobs = corpus.web_corpus(0xE21A0902, 300) + corpus.mixed_corpus(0xE21A0903, 60, 300)
    n = compare(engine, oracle, blobs, {"sgm": 10, "js": 3, "nil": 0}, {"od": 1}, seed=(7, 5, 5), allow_unsupported=True)
    assert n >= len(blobs) * 0.93
    n = compare(engine, oracle, blobs, {"js": 1}, {"od": 1}, seed=(8, 5, 5), allow_unsupported=True)
    assert n >= len(blobs) * 0.93


def test_sgm_js_documents(engine, oracle):
    """C4-style documents: every sgm mutation and every js mutation on arrays / objects, exact"""
    blobs = corpus.uniform_corpus(0xE21A0906, 60, 3000, "markup") + corpus.uniform_corpus(0xE21A0907, 20, 20000, "markup")
    for seed in ((1, 1, 1), (2, 3, 4)):
        n = compare(engine, oracle, blobs, {"sgm": 10, "js": 3}, {"od": 1}, seed=seed, n_cases=240, allow_unsupported=True)
        assert n >= 240 * 0.9           # inner-text rounds that open a second nested round (base64-looking text inside a document) are flagged
    n = compare(engine, oracle, blobs, {"sgm": 1, "js": 1, "ab": 1, "tr2": 1}, {"nd": 1, "bu": 1}, seed=(5, 1, 1), n_cases=160, allow_unsupported=True)
    assert n >= 160 * 0.85


def test_true_default_mutator_mix(engine, oracle):
    """all 41 mutators at the reference's default priorities (src/erlamsa_mutations.erl:1291-1331) and the default patterns"""
    import erlamsa_b200.options as eo
    muts = dict(eo.default_mutations())
    blobs = corpus.mixed_corpus(0xE21A0904, 200, 2000) + corpus.web_corpus(0xE21A0905, 100)
    n = compare(engine, oracle, blobs, muts, {"od": 1, "nd": 2, "bu": 1, "sk": 2, "sz": 2, "cs": 1, "ar": 1, "cp": 1}, seed=(9, 5, 5), allow_unsupported=True)
    assert n >= len(blobs) * 0.9


def framed_corpus(seed, count):
    """blobs that really carry length fields / xor8 / crc32 trailers, so the sizer and checksum searches FIND something"""
    import struct
    import zlib
    r = corpus.rng(seed)
    out = []
    for i in range(count):
        body = corpus.random_bytes(r, int(r.integers(4, 600))) if i % 2 else corpus.structured_text(r, int(r.integers(4, 600)))
        pre = corpus.random_bytes(r, int(r.integers(0, 12)))
        kind = i % 7
        if kind == 0:
            out.append(pre + struct.pack(">H", len(body)) + body)
        elif kind == 1:
            out.append(pre + struct.pack("<I", len(body)) + body + b"TAIL")
        elif kind == 2:
            out.append(pre + bytes([min(len(body), 255)]) + body[:255])
        elif kind == 3:
            x = 0
            for b in body:
                x ^= b
            out.append(pre + body + bytes([x]))
        elif kind == 4:
            out.append(pre + body + struct.pack(">I", zlib.crc32(body)))
        elif kind == 5:
            out.append(struct.pack(">Q", len(body) + 2) + body + b"xy")
        else:
            out.append(pre + body)
    return out


@pytest.mark.parametrize("pat", ["sz", "cs", "ar", "cp"])
def test_complex_patterns(engine, oracle, pat):
    muts = {c: 1 for c in ("bd", "bf", "bi", "num", "sd", "ld", "ab", "td", "ft")}
    blobs = framed_corpus(0xE21A0800, 210) + corpus.mixed_corpus(0xE21A0801, 60, 1200)
    n = compare(engine, oracle, blobs, muts, {pat: 1}, seed=(6, 6, 6), allow_unsupported=True)
    assert n > len(blobs) * 0.8


def test_len_mutator(engine, oracle):
    blobs = framed_corpus(0xE21A0900, 210) + corpus.mixed_corpus(0xE21A0901, 60, 3000)
    n = compare(engine, oracle, blobs, {"len": 1}, {"od": 1, "nd": 1}, seed=(8, 1, 8))
    assert n >= len(blobs) - 3


def test_default_pattern_mix_and_all_device_mutators(engine, oracle):
    """the reference's default patterns (all ten) over every mutator that has a device implementation"""
    import erlamsa_b200
    muts = {c: p for c, p in erlamsa_b200.default_mutations() if c in erlamsa_b200.supported_mutations()}
    pats = dict(erlamsa_b200.default_patterns())
    blobs = framed_corpus(0xE21A0A00, 140) + corpus.text_corpus(0xE21A0A01, 200, 900) + corpus.mixed_corpus(0xE21A0A02, 160, 1500)
    n = compare(engine, oracle, blobs, muts, pats, seed=(1, 2, 3), allow_unsupported=True)
    assert n > len(blobs) * 0.6          # the text third of this corpus is full of '<' and quoted runs: sgm / nested b64 documents are flagged


@pytest.mark.parametrize("pat", ["od", "nd", "bu", "sk", "co", "nu"])
def test_patterns_with_mix(engine, oracle, pat):
    import erlamsa_b200
    muts = {c: p for c, p in erlamsa_b200.default_mutations() if c in erlamsa_b200.supported_mutations()}
    blobs = corpus.mixed_corpus(0xE21A0100, 160 if pat in ("sk", "bu", "nd") else 400)
    n = compare(engine, oracle, blobs, muts, {pat: 1}, seed=(11, 22, 33), allow_unsupported=True)
    assert n > len(blobs) // 3


def test_default_supported_mix(engine, oracle):
    import erlamsa_b200
    muts = {c: p for c, p in erlamsa_b200.default_mutations() if c in erlamsa_b200.supported_mutations()}
    blobs = corpus.mixed_corpus(0xE21A0200, 600)
    compare(engine, oracle, blobs, muts, SUPPORTED_PATS, seed=(5, 6, 7), allow_unsupported=True)


def test_c3_mutators_64k(engine, oracle):
    muts = {c: 1 for c in ("bd", "bei", "bed", "bf", "bi", "ber", "br", "num")}
    blobs = corpus.uniform_corpus(0xE21A0003, 48, 65536, "bin") + corpus.uniform_corpus(0xE21A0004, 48, 65536, "num")
    n = compare(engine, oracle, blobs, muts, {"od": 1})
    assert n == len(blobs)


def test_light_and_full_kernel_flavours_agree(engine):
    """the host launches a kernel flavour without the heavy mutators in its call graph when the options allow;
    forcing the full one must not change a byte"""
    import os
    blobs = corpus.mixed_corpus(0xE21A0400, 300)
    opts = {"mutations": {c: 1 for c in ("bd", "bf", "num", "sr", "sp", "ld", "lis", "uw", "snand")}, "patterns": {"od": 1, "nd": 1, "bu": 1},
            "seed": (3, 3, 3), "max_case_out": CAP}
    a_out, a_meta = engine.fuzz_batch(blobs, opts, n_cases=600)
    os.environ["EB200_FORCE_FULL"] = "1"
    try:
        b_out, b_meta = engine.fuzz_batch(blobs, opts, n_cases=600)
    finally:
        del os.environ["EB200_FORCE_FULL"]
    assert a_out == b_out
    assert [(m.status, m.draws, m.pattern, list(m.used)) for m in a_meta] == [(m.status, m.draws, m.pattern, list(m.used)) for m in b_meta]


def test_case_window_and_corpus_wraparound(engine, oracle):
    """first_case / n_cases select a window of the reference's case loop; blobs are reused modulo the corpus size"""
    blobs = corpus.mixed_corpus(0xE21A0300, 37)
    muts = {c: 1 for c in ("bd", "bf", "num", "sr", "ld")}
    compare(engine, oracle, blobs, muts, {"od": 1, "nd": 1}, n_cases=150, first_case=1000)


def test_pipelined_host_path_equals_plain_host_path(engine):
    """eb200_fuzz_batch_into switches to the chunked copy/compute-overlapped pipeline for big batches; results must
    be identical to the single-shot host path (same case ids, same bytes)"""
    import ctypes as C
    import numpy as np
    from erlamsa_b200 import _native as N
    import erlamsa_b200
    n, size = 8192, 40960
    r = corpus.rng(0xE21A0B00)
    data = r.integers(0, 256, size=n * size, dtype=np.uint8)
    data[::97] = 0x31                                  # sprinkle digits so that num has work
    off = np.arange(0, (n + 1) * size, size, dtype=np.uint64)
    opts = erlamsa_b200.make_opts({"mutations": {c: 1 for c in ("bd", "bei", "bf", "bi", "num", "sd")}, "patterns": {"od": 1, "nd": 1},
                                   "seed": (4, 4, 4), "first_case": 1, "max_case_out": 1 << 22})
    out = np.zeros(n * size + n * size // 10 + (64 << 20), dtype=np.uint8)   # slots (input + slack) + overflow region
    o_off = np.zeros(n + 1, dtype=np.uint64); o_len = np.zeros(n, dtype=np.uint64)
    st = N.Stats()
    rc = N.lib().eb200_fuzz_batch_into(engine._ctx, C.byref(opts), data.ctypes.data, off.ctypes.data, n, n, out.ctypes.data, out.size,
                                       o_off.ctypes.data, o_len.ctypes.data, None, C.byref(st))
    assert rc == 0
    # plain path: same call through the malloc()ing entry point (never pipelined)
    out_p = C.c_void_p(); p_off = np.zeros(n + 1, dtype=np.uint64); p_len = np.zeros(n, dtype=np.uint64)
    rc = N.lib().eb200_fuzz_batch(engine._ctx, C.byref(opts), data.ctypes.data, C.cast(off.ctypes.data, C.POINTER(C.c_uint64)), n, n, C.byref(out_p),
                                  C.cast(p_off.ctypes.data, C.POINTER(C.c_uint64)), C.cast(p_len.ctypes.data, C.POINTER(C.c_uint64)), None, None)
    assert rc == 0
    try:
        assert (o_len == p_len).all()
        plain = np.ctypeslib.as_array(C.cast(out_p, C.POINTER(C.c_uint8)), shape=(int(p_off[n]),))
        for k in list(range(0, n, 257)) + [n - 1]:
            a = out[int(o_off[k]):int(o_off[k]) + int(o_len[k])]
            b = plain[int(p_off[k]):int(p_off[k]) + int(p_len[k])]
            assert a.tobytes() == b.tobytes(), k
    finally:
        N.lib().eb200_free(out_p)


def test_philox_mode_is_distribution_equivalent(engine, oracle):
    """Philox mode replaces the AS183 stream but keeps the decision logic: over many cases the histogram of the pattern
    chosen, of the mutator that got applied, and of output-length deltas must match the oracle's (AS183) histograms.
    (The reference's own tests pin distributions the same way, e.g. sed_tree_swap_one_test: 6 distinct outputs.)"""
    import collections
    muts = {c: 1 for c in ("bd", "bei", "bed", "bf", "bi", "ber", "br", "sd", "num", "ld", "lr2", "ls", "uw", "ui")}
    pats = {"od": 2, "nd": 1, "bu": 1}
    blobs = corpus.mixed_corpus(0xE21A0C00, 64, 800, kinds=("num", "lines"))
    n = 24000
    ref, rmeta = oracle.fuzzer(blobs, mutations=muts, patterns=pats, seed=(1, 2, 3), n_cases=n, max_case_out=CAP)
    got, gmeta = engine.fuzz_batch(blobs, {"mutations": muts, "patterns": pats, "seed": (1, 2, 3), "rng": "philox", "max_case_out": CAP}, n_cases=n)

    def hists(outs, meta):
        pat = collections.Counter(m.pattern for m in meta)
        first = collections.Counter(m.used[0] for m in meta)
        delta = collections.Counter(max(-3, min(3, len(o) - len(blobs[k % len(blobs)]))) for k, o in enumerate(outs))
        rounds = collections.Counter(min(m.n_used, 6) for m in meta)
        return pat, first, delta, rounds

    for name, a, b in zip(("pattern", "first mutator", "length delta", "rounds"), hists(ref, rmeta), hists(got, gmeta)):
        keys = sorted(set(a) | set(b))
        chi2 = sum((a[k] - b[k]) ** 2 / max(a[k] + b[k], 1) for k in keys)
        dof = max(len(keys) - 1, 1)
        assert chi2 < 3.0 * dof + 25, "%s histograms differ: chi2 %.1f on %d dof\n%r\n%r" % (name, chi2, dof, dict(a), dict(b))


def test_philox_mode_whole_table_distribution(engine, oracle):
    """the same check over the WHOLE default table (all 41 mutators, all patterns): which mutator gets applied first, which
    pattern runs and how many rounds a case takes must be distributed as under the reference stream"""
    import collections
    blobs = corpus.mixed_corpus(0xE21A0C41, 48, 300, kinds=("num", "lines", "bin")) + corpus.text_corpus(0xE21A0C42, 16, 300)
    n = 6000
    ref, rmeta = oracle.fuzzer(blobs, seed=(1, 2, 3), n_cases=n, max_case_out=1 << 18)
    got, gmeta = engine.fuzz_batch(blobs, {"seed": (1, 2, 3), "rng": "philox", "max_case_out": 1 << 18}, n_cases=n)
    keep_r = [m for m in rmeta if m.status == 0]
    keep_g = [m for m in gmeta if m.status == 0]
    assert len(keep_g) > 0.9 * n and len(keep_r) > 0.9 * n

    def hists(meta):
        return (collections.Counter(m.pattern for m in meta), collections.Counter(m.used[0] for m in meta), collections.Counter(min(m.n_used, 5) for m in meta))

    for name, a, b in zip(("pattern", "first mutator", "rounds"), hists(keep_r), hists(keep_g)):
        keys = sorted(set(a) | set(b))
        sa, sb = sum(a.values()), sum(b.values())
        chi2 = sum((a[k] / sa - b[k] / sb) ** 2 / max(a[k] / sa + b[k] / sb, 1e-9) for k in keys) * min(sa, sb)
        dof = max(len(keys) - 1, 1)
        assert chi2 < 3.0 * dof + 30, "%s histograms differ: chi2 %.1f on %d dof\n%r\n%r" % (name, chi2, dof, dict(a), dict(b))


def test_philox_mode_runs_and_differs(engine):
    blobs = corpus.mixed_corpus(0xE21A0400, 200)
    a, _ = engine.fuzz_batch(blobs, {"mutations": {"bd": 1, "bf": 1, "num": 1}, "patterns": {"od": 1}, "seed": (1, 2, 3), "rng": "philox"})
    b, _ = engine.fuzz_batch(blobs, {"mutations": {"bd": 1, "bf": 1, "num": 1}, "patterns": {"od": 1}, "seed": (1, 2, 3), "rng": "philox"})
    c, _ = engine.fuzz_batch(blobs, {"mutations": {"bd": 1, "bf": 1, "num": 1}, "patterns": {"od": 1}, "seed": (1, 2, 3)})
    assert a == b            # deterministic
    assert a != c            # a different stream than AS183


@pytest.mark.xfail(strict=False, reason="open issue (DESIGN.md section 9): tests/wide_diff.py saw 1 of 240 whole-table nd cases differ between two "
                                        "runs of the same build once; the 128-register build of the general kernel differs on b64's nested rounds")
def test_run_to_run_determinism_whole_table_nd(engine):
    """the same batch twice through the same engine: every byte and every draw count must repeat"""
    import erlamsa_b200
    muts = {c: p for c, p in erlamsa_b200.default_mutations() if c in erlamsa_b200.supported_mutations()}
    blobs = corpus.mixed_corpus(0xE21A0100, 240)
    opts = {"mutations": muts, "patterns": {"nd": 1}, "seed": (2, 7, 1), "max_case_out": CAP}
    a, ma = engine.fuzz_batch(blobs, dict(opts), n_cases=len(blobs))
    for _ in range(3):
        b, mb = engine.fuzz_batch(blobs, dict(opts), n_cases=len(blobs))
        bad = [k for k in range(len(blobs)) if a[k] != b[k] or ma[k].draws != mb[k].draws or ma[k].status != mb[k].status]
        assert not bad, "cases that did not repeat: %r" % bad[:8]
