"""Python mirror of the reference's library API erlamsa_app:fuzz/1,2 (reference src/erlamsa_app.erl:255-263):
fuzz(Data) / fuzz(Data, Opts) -> mutated binary, a pure function of (data, opts incl. seed)."""
from . import erlamsa_main


def fuzz(data, opts=None):
    o = dict(opts or {})
    o.update({"paths": ["direct"], "output": "return", "input": bytes(data)})   # get_direct_fuzzing_opts/2
    res = erlamsa_main.fuzzer(o)
    # extract_function([X]) -> X ; [] stays [] in the reference (an empty result) -- b"" here
    return res[0] if res else b""
