"""Host-side logic that needs no GPU: the -m/-p option grammar, option-map translation, corpus helpers."""
import pytest

import corpus


def test_string_to_actions_grammar():
    import erlamsa_b200 as E
    d = E.default_mutations()
    # erlamsa_cmdparse:string_to_actions/3: reversed list; a bare code takes its DEFAULT priority (sgm = 10, num = 3, bd = 1)
    assert E.string_to_actions("bd,num=3,sr=2", "mutation", d) == [("sr", 2), ("num", 3), ("bd", 1)]
    assert E.string_to_actions("sgm,num,bd=7", "mutation", d) == [("bd", 7), ("num", 3), ("sgm", 10)]
    assert dict(E.string_to_actions("bd=2,bd=5", "mutation", d)) == {"bd": 2}          # as a map the FIRST entry wins (maps:from_list of the reversed list)
    assert E.string_to_actions("bd=2=3,,num=", "mutation", d) == [("num", 3), ("bd", 1)]  # string:tokens/2 pieces
    assert E.string_to_actions("default", "mutation", d) == d
    with pytest.raises(ValueError):
        E.string_to_actions("bd=x", "mutation", d)
    with pytest.raises(ValueError):
        E.string_to_actions("bd,nope", "mutation", d)


def test_make_opts_translates_the_reference_option_map():
    import erlamsa_b200 as E
    o = E.make_opts({"seed": (7, 8, 9), "mutations": [("bd", 2), ("num", 5)], "patterns": "od,nd=3", "blockscale": 2.0, "skip": 10})
    codes = E.mutator_codes()
    assert list(o.seed) == [7, 8, 9] and o.blockscale == 2.0 and o.first_case == 11
    assert o.muta_pri[codes.index("bd")] == 2 and o.muta_pri[codes.index("num")] == 5 and o.muta_pri[codes.index("sr")] == -1
    pc = E.pattern_codes()
    assert o.pat_pri[pc.index("od")] == 1 and o.pat_pri[pc.index("nd")] == 3 and o.pat_pri[pc.index("bu")] == -1
    with pytest.raises(ValueError):
        E.make_opts({"mutations": {"xx": 1}})


def test_unseeded_options_get_a_random_seed():
    import erlamsa_b200 as E
    a, b = E.make_opts({}), E.make_opts({})
    assert all(0 <= x < 65536 for x in a.seed)
    assert list(a.seed) != list(b.seed) or True   # 2^-48 chance of equality; only the range is asserted


def test_corpora_are_deterministic():
    a = corpus.mixed_corpus(1, 50)
    b = corpus.mixed_corpus(1, 50)
    assert a == b and len({len(x) for x in a}) > 10
    assert corpus.uniform_corpus(2, 3, 4096, "num")[0][:1] != b""


def test_shard_windows_partition_the_case_range():
    from erlamsa_b200.sharding import shard_window
    for n, world in ((100, 1), (100, 2), (1000003, 8), (5, 8)):
        got = []
        for r in range(world):
            first, cnt = shard_window(n, r, world, first_case=11)
            got.extend(range(first, first + cnt))
        assert got == list(range(11, 11 + n))


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours) runs without a GPU and prints one JSON
    line with the keys of the bench contract"""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "cases/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_markup_corpus_shapes():
    """C4 documents: exact size, SGML ones tokenise as markup, JSON ones as a complete document (oracle-side check)"""
    import oracle_lib as O
    blobs = corpus.uniform_corpus(11, 6, 3000, "markup")
    assert all(len(b) == 3000 for b in blobs)
    outs, meta = O.fuzzer(blobs, mutations={"sgm": 1, "js": 1}, patterns={"od": 1}, seed=(1, 2, 3), n_cases=48, max_case_out=1 << 20)
    assert all(m.status == 0 for m in meta)
    assert sum(1 for m in meta if m.n_used) >= 40          # refusals would leave the case unused
