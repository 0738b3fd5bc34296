"""Host-side logic that needs no GPU: the -m/-p option grammar, option-map translation, corpus helpers."""
import pytest

import corpus


def test_string_to_actions_grammar():
    import erlamsa_b200 as E
    d = E.default_mutations()
    assert E.string_to_actions("bd,num=3,sr=2", "mutation", d) == [("bd", 1), ("num", 3), ("sr", 2)]
    assert E.string_to_actions("default", "mutation", d) == d
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
