"""Run the REFERENCE ITSELF (the .erl sources under /root/reference/src, executed by oracle/erlref's evaluator):
erlamsa_main:fuzzer/1 with paths=[direct], output=return -- the path the engine replaces.

TEST INFRASTRUCTURE (oracle/).  /root/reference exists only in the build container, so everything produced here is
committed as fixtures (tests/golden/reference_vectors.json, made by tests/golden/make_reference_vectors.py)."""
import os

from .interp import Runtime, BudgetExceeded, run_with_big_stack
from .terms import NIL, Cons, from_py, to_py, ErlError

REF_SRC = os.environ.get("ERLAMSA_REFERENCE_SRC", "/root/reference/src")
OTP_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "otp")


def available():
    return os.path.exists(os.path.join(REF_SRC, "erlamsa_main.erl"))


class RefResult(object):
    __slots__ = ("output", "draws", "status", "detail")

    def __init__(self, output, draws, status, detail=""):
        self.output = output      # bytes (b"" when the reference recorded nothing)
        self.draws = draws        # random:uniform calls inside the worker process of the case
        self.status = status      # "ok" | "died" | "budget" | "unsupported"
        self.detail = detail


class Reference(object):
    """one evaluator instance with the reference's modules loaded on demand"""

    def __init__(self, budget=30_000_000):
        self.rt = Runtime([REF_SRC, OTP_DIR])
        self.budget = budget

    def opts_map(self, blob, seed, mutations=None, patterns=None, n=1, skip=0, blockscale=None, generators=None, extra=None):
        d = {"paths": from_py(["direct"]), "output": "return", "input": blob, "seed": tuple(seed), "n": n}
        if skip:
            d["skip"] = skip
        if mutations is not None:
            d["mutations"] = from_py([(c, p) for c, p in (mutations.items() if isinstance(mutations, dict) else mutations)])
        if patterns is not None:
            d["patterns"] = from_py([(c, p) for c, p in (patterns.items() if isinstance(patterns, dict) else patterns)])
        if blockscale is not None:
            d["blockscale"] = blockscale
        if generators is not None:
            d["generators"] = from_py([(c, p) for c, p in (generators.items() if isinstance(generators, dict) else generators)])
        if extra:
            d.update(extra)
        return d

    def fuzzer(self, opts):
        """erlamsa_main:fuzzer(Opts) -> (list of bytes, per-worker draw counts, crash or None)"""
        rt = self.rt
        rt.steps = 0
        rt.budget = self.budget
        rt.child_draws = []
        rt.last_crash = None
        res = rt.call("erlamsa_main", "fuzzer", opts)
        return [x for x in to_py(res)], list(rt.child_draws), rt.last_crash

    def case(self, blob, case_no, seed, mutations=None, patterns=None, **kw):
        """test case number `case_no` (1-based) of fuzzer(#{input => blob, seed => seed, ...}): the case_no-th
        gen_predictable_seed() of the parent stream, as erlamsa_b200's batch semantics define case I."""
        opts = self.opts_map(blob, seed, mutations, patterns, n=case_no, skip=case_no - 1, **kw)
        try:
            outs, draws, crash = self.fuzzer(opts)
        except BudgetExceeded:
            return RefResult(b"", 0, "budget")
        except self.rt.Unsupported as e:
            return RefResult(b"", 0, "unsupported", str(e))
        except RecursionError:
            return RefResult(b"", 0, "budget", "recursion")
        if crash is not None:
            return RefResult(b"", draws[-1] if draws else 0, "died", str(crash))
        out = outs[0] if outs else b""
        return RefResult(out, draws[-1] if draws else 0, "ok")

    def run(self, blobs, seed, n_cases=None, first_case=1, mutations=None, patterns=None, **kw):
        """the engine's batch: case I (first_case..) reads blobs[(I-1) % len(blobs)]"""
        if n_cases is None:
            n_cases = len(blobs)
        out = []
        for k in range(n_cases):
            i = first_case + k
            out.append(self.case(blobs[(i - 1) % len(blobs)], i, seed, mutations, patterns, **kw))
        return out


    def case_paths(self, files, case_no, seed, mutations=None, patterns=None, stdin=None, generators=None, **kw):
        """test case `case_no` of fuzzer(#{paths => Files | ["-"], output => return, ...}): the file / stdin generators.
        files: list of byte strings (virtual files f0, f1, ...); stdin: bytes (paths = ["-"], only meaningful with case_no == 1).
        Non-direct paths do not record results (src/erlamsa_main.erl:143), so the written bytes are read back from the
        erlamsa_logger:log_data call the driver makes for every case (:199)."""
        rt = self.rt
        rt.vfs = {"f%d" % i: b for i, b in enumerate(files or [])}
        paths = ["-"] if stdin is not None else ["f%d" % i for i in range(len(files))]
        rt.stdin.data = stdin or b""
        rt.stdin.pos = 0
        opts = self.opts_map(b"", seed, mutations, patterns, n=case_no, skip=case_no - 1, generators=generators, **kw)
        del opts["input"]
        opts["paths"] = from_py([from_py([ord(c) for c in p]) for p in paths])
        opts["maxrunningtime"] = 600000
        rt.logged_data = []
        try:
            rt.steps = 0
            rt.budget = self.budget
            rt.child_draws = []
            rt.last_crash = None
            rt.call("erlamsa_main", "fuzzer", opts)
        except BudgetExceeded:
            return RefResult(b"", 0, "budget")
        except rt.Unsupported as e:
            return RefResult(b"", 0, "unsupported", str(e))
        if rt.last_crash is not None:
            return RefResult(b"", rt.child_draws[-1] if rt.child_draws else 0, "died", str(rt.last_crash))
        out = rt.logged_data[-1] if rt.logged_data else b""
        return RefResult(out, rt.child_draws[-1] if rt.child_draws else 0, "ok")
