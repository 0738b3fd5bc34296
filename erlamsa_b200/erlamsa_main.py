"""Python mirror of the reference's driver entry, erlamsa_main:fuzzer/1 (reference
src/erlamsa_main.erl:124-247), for paths == [direct] / output == return, backed by the CUDA engine.

    fuzzer(#{paths => [direct], input => Bin, seed => {A,B,C}, n => N, mutations => [...], patterns => [...]})
        -> [binary()]            (empty outputs are not recorded, record_result/2 :120-122)

`fuzz/1` is the spelling BASELINE.json's north star uses for the same entry. `input` may also be a
LIST of binaries (a corpus): case I then mutates element (I-1) rem length -- the batched
generalisation the engine exists for; n defaults to the corpus size.
"""
from .engine import Engine

_engine = None


def _get_engine():
    global _engine
    if _engine is None:
        _engine = Engine(0)
    return _engine


def fuzzer(opts):
    opts = dict(opts)
    paths = opts.get("paths", ["-"])
    if list(paths) != ["direct"]:
        raise NotImplementedError("only paths => [direct] is served by the batch engine; "
                                  "stdin/file/network front-ends stay in Erlang (SURVEY.md 8b)")
    if opts.get("output", "return") != "return":
        raise NotImplementedError("only output => return")
    inp = opts.get("input")
    if inp is None:
        raise ValueError("direct generator needs `input`")
    blobs = [bytes(inp)] if isinstance(inp, (bytes, bytearray, memoryview)) else [bytes(b) for b in inp]
    n = int(opts.get("n", len(blobs) if len(blobs) > 1 else 1))
    skip = int(opts.get("skip", 0))
    o = dict(opts)
    o["first_case"] = skip + 1
    outs, _meta = _get_engine().fuzz_batch(blobs, o, n_cases=max(n - skip, 0))
    return [x for x in outs if x != b""]


fuzz = fuzzer
