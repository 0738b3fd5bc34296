"""erlamsa_b200 -- B200-native batched mutation engine for erlamsa's hot path
(erlamsa_mutations + erlamsa_rnd per test case over a corpus).

Product = the CUDA library behind include/erlamsa_b200.h (csrc/). This package is the thin host-side
mirror of the reference's own entry points (erlamsa_app:fuzz/1,2, erlamsa_main:fuzzer/1) used by the
parity tests and the benchmark; the Erlang-side binding is in erlang/ and INTEGRATION.md.
"""
from . import _native
from .engine import Engine, EngineError
from .options import (make_opts, default_mutations, default_patterns, supported_mutations, supported_patterns,
                      mutator_codes, pattern_codes, string_to_actions)
from . import erlamsa_app, erlamsa_main

__all__ = ["Engine", "EngineError", "make_opts", "default_mutations", "default_patterns", "supported_mutations",
           "supported_patterns", "mutator_codes", "pattern_codes", "string_to_actions", "erlamsa_app", "erlamsa_main"]
