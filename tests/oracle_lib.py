"""ctypes binding of the ORACLE (oracle/liberlamsa_oracle.so) -- test infrastructure only.

The oracle is the CPU restatement of erlamsa's mutation hot path; see oracle/src/driver.cpp.
Nothing under erlamsa_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liberlamsa_oracle.so")

MUT_CODES = ["sgm", "js", "uw", "ui", "ab", "ad", "tr2", "td", "num", "ts1", "tr", "ts2",
             "bd", "bei", "bed", "bf", "bi", "ber", "br", "sp", "sr", "sd", "snand", "srnd",
             "ld", "lds", "lr2", "lri", "lr", "ls", "lp", "lis", "lrs", "ft", "fn", "fo",
             "len", "b64", "uri", "zip", "nil"]
PAT_CODES = ["od", "nd", "bu", "sk", "sz", "cs", "ar", "cp", "co", "nu"]


class Opts(C.Structure):
    _fields_ = [("seed", C.c_int64 * 3), ("blockscale", C.c_double),
                ("muta_pri", C.c_int32 * 41), ("pat_pri", C.c_int32 * 10),
                ("gen_direct_pri", C.c_int32), ("gen_random_pri", C.c_int32),
                ("ssrf_host", C.c_char * 64), ("ssrf_port", C.c_int32), ("max_case_out", C.c_uint64),
                ("donor_pool", C.c_void_p), ("donor_len", C.c_void_p), ("n_donors", C.c_uint64), ("donor_stride", C.c_uint32), ("pad", C.c_uint32),
                ("gen_file_pri", C.c_int32), ("gen_stdin_pri", C.c_int32),
                ("gen_jump_pri", C.c_int32), ("pad2", C.c_int32),
                ("case_stream_seed", C.c_int64 * 3), ("case_stream_first", C.c_uint64)]


class Meta(C.Structure):
    _fields_ = [("pattern", C.c_int32), ("generator", C.c_int32), ("n_used", C.c_int32), ("n_failed", C.c_int32),
                ("used", C.c_int32 * 16), ("draws", C.c_uint64), ("status", C.c_int32), ("pad", C.c_int32),
                ("thread_seed", C.c_int64 * 3)]


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.eo_default_opts.argtypes = [C.POINTER(Opts)]
        L.eo_fuzzer.argtypes = [C.POINTER(Opts), C.c_void_p, C.POINTER(C.c_uint64), C.c_uint64, C.c_uint64, C.c_uint64,
                                C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(Meta)]
        L.eo_fuzzer.restype = C.c_int
        L.eo_run_mutator.argtypes = [C.POINTER(Opts), C.c_int, C.POINTER(C.c_int64), C.c_char_p, C.c_uint64,
                                     C.c_char_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
                                     C.POINTER(C.c_double)]
        L.eo_run_mutator.restype = C.c_int
        L.eo_rnd_seed.argtypes = [C.c_int64] * 3
        L.eo_rnd_uniform.restype = C.c_double
        L.eo_rnd_rand.argtypes = [C.c_uint64]
        L.eo_rnd_rand.restype = C.c_uint64
        L.eo_rnd_erand.argtypes = [C.c_uint64]
        L.eo_rnd_erand.restype = C.c_uint64
        L.eo_rnd_state.argtypes = [C.POINTER(C.c_int64)]
        L.eo_lists_sort.argtypes = [C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
        L.eo_lex_unlex.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]
        L.eo_mutator_code.restype = C.c_char_p
        L.eo_pattern_code.restype = C.c_char_p
        _lib = L
    return _lib


def make_opts(seed=(1, 2, 3), mutations=None, patterns=None, blockscale=1.0, generators=None,
              ssrf_host="localhost", ssrf_port=51234, max_case_out=0, donors=None, case_stream=None):
    """mutations / patterns: None = reference defaults, else dict or list of (code, pri) -- the
    reference's `[{Code, Pri}]` option lists (src/erlamsa_main.erl:129,156)."""
    o = Opts()
    lib().eo_default_opts(C.byref(o))
    o.seed[0], o.seed[1], o.seed[2] = seed
    o.blockscale = blockscale
    if mutations is not None:
        m = dict(mutations) if not isinstance(mutations, dict) else mutations
        for i, c in enumerate(MUT_CODES):
            o.muta_pri[i] = m.get(c, -1)
    if patterns is not None:
        p = dict(patterns) if not isinstance(patterns, dict) else patterns
        for i, c in enumerate(PAT_CODES):
            o.pat_pri[i] = p.get(c, -1)
    if generators is not None:
        g = dict(generators)
        o.gen_direct_pri = g.get("direct", -1)
        o.gen_random_pri = g.get("random", -1)
        o.gen_file_pri = g.get("file", -1)
        o.gen_stdin_pri = g.get("stdin", -1)
        o.gen_jump_pri = g.get("jump", -1)
    if case_stream is not None:
        # (worker seed {A,B,C}, number of the worker's first case): --workers, src/erlamsa_main.erl:254-280
        (o.case_stream_seed[0], o.case_stream_seed[1], o.case_stream_seed[2]), o.case_stream_first = case_stream
    o.ssrf_host = ssrf_host.encode()
    o.ssrf_port = ssrf_port
    o.max_case_out = max_case_out
    if donors is not None:
        # donors = (windows: list of bytes, stride): the cross-seed pool of BASELINE config C5
        wins, stride = donors
        pool = b"".join(w.ljust(stride, b"\0") for w in wins)
        o._pool = C.create_string_buffer(pool, len(pool))
        o._lens = (C.c_uint32 * len(wins))(*[len(w) for w in wins])
        o.donor_pool = C.cast(o._pool, C.c_void_p)
        o.donor_len = C.cast(o._lens, C.c_void_p)
        o.n_donors = len(wins)
        o.donor_stride = stride
    return o


def fuzzer(blobs, opts=None, n_cases=None, first_case=1, **kw):
    """Run cases first_case..first_case+n_cases-1 of one erlamsa_main:fuzzer/1 call; case I reads
    blobs[(I-1) % len(blobs)]. Returns (list of output bytes, list of Meta)."""
    if opts is None:
        opts = make_opts(**kw)
    if n_cases is None:
        n_cases = len(blobs)
    data = b"".join(blobs)
    off = (C.c_uint64 * (len(blobs) + 1))()
    acc = 0
    for i, b in enumerate(blobs):
        off[i] = acc
        acc += len(b)
    off[len(blobs)] = acc
    buf = C.create_string_buffer(data, len(data) + 1)
    out_p = C.c_void_p()
    out_off = (C.c_uint64 * (n_cases + 1))()
    meta = (Meta * n_cases)()
    rc = lib().eo_fuzzer(C.byref(opts), C.cast(buf, C.c_void_p), off, len(blobs), first_case, n_cases,
                         C.byref(out_p), out_off, meta)
    if rc != 0:
        raise RuntimeError("oracle eo_fuzzer rc=%d" % rc)
    total = out_off[n_cases]
    # (ctypes.string_at takes a C int size: a few GB of outputs -- bench.py's CPU legs -- would wrap negative)
    raw = bytes((C.c_char * total).from_address(out_p.value)) if total else b""
    outs = [raw[out_off[i]:out_off[i + 1]] for i in range(n_cases)]
    return outs, list(meta)


def fuzzer_total_bytes(blobs, opts, n_cases, first_case=1):
    """Same run as fuzzer() but the outputs stay in the library's buffer: returns only their total size. For timing
    the restatement itself (bench.py's CPU legs) without a multi-GB copy into Python objects on the clock."""
    data = b"".join(blobs)
    off = (C.c_uint64 * (len(blobs) + 1))()
    acc = 0
    for i, b in enumerate(blobs):
        off[i] = acc
        acc += len(b)
    off[len(blobs)] = acc
    buf = C.create_string_buffer(data, len(data) + 1)
    out_p = C.c_void_p()
    out_off = (C.c_uint64 * (n_cases + 1))()
    meta = (Meta * n_cases)()
    rc = lib().eo_fuzzer(C.byref(opts), C.cast(buf, C.c_void_p), off, len(blobs), first_case, n_cases,
                         C.byref(out_p), out_off, meta)
    if rc != 0:
        raise RuntimeError("oracle eo_fuzzer rc=%d" % rc)
    return int(out_off[n_cases])


def run_mutator(code, data, seed, next_block=None, rounds=1, opts=None):
    """Seed the RNG, apply one mutator `rounds` times to [data | next_block]; mirrors the reference's
    eunit helpers (src/erlamsa_mutations_test.erl:40-49). Returns (bytes, delta, rc)."""
    if opts is None:
        opts = make_opts()
    s = (C.c_int64 * 3)(*seed)
    out_p = C.c_void_p()
    out_len = C.c_uint64()
    delta = C.c_double()
    rc = lib().eo_run_mutator(C.byref(opts), MUT_CODES.index(code), s, data, len(data),
                              next_block, 0 if next_block is None else len(next_block), rounds,
                              C.byref(out_p), C.byref(out_len), C.byref(delta))
    if rc < 0:
        raise RuntimeError("oracle eo_run_mutator rc=%d" % rc)
    return (C.string_at(out_p, out_len.value) if out_len.value else b""), delta.value, rc
