"""CPU tests of the ORACLE (the C++ restatement of the reference path): known answers for the RNG, the one
fixed-seed reference test, OTP lists:sort behaviour, and the reference's 31 eunit property tests
(reference src/erlamsa_mutations_test.erl) mirrored 1:1 -- same inputs, regexes, iteration budgets."""
import ctypes as C
import random
import re

import pytest

import oracle_lib as O


def seeds(n, base=1000):
    return [(base + 7 * i, 2 * base + 13 * i + 1, 3 * base + 29 * i + 2) for i in range(n)]


# ------------------------------------------------------------------ RNG (OTP `random`, AS183)
def test_as183_known_answers_default_seed():
    L = O.lib()
    L.eo_rnd_seed0()
    got = [L.eo_rnd_uniform() for _ in range(3)]
    assert got == [0.4435846174457203, 0.7230402056221108, 0.94581636451987]   # public OTP values


def test_seed_maps_like_otp():
    L = O.lib()
    L.eo_rnd_seed(1, 2, 3)
    s = (C.c_int64 * 3)()
    L.eo_rnd_state(s)
    assert list(s) == [2, 3, 4]
    L.eo_rnd_seed(-40000, 0, 30322)
    L.eo_rnd_state(s)
    assert list(s) == [40000 % 30268 + 1, 1, 1]


def test_rand_zero_bound_draws_nothing():
    L = O.lib()
    L.eo_rnd_seed(1, 2, 3)
    a = (C.c_int64 * 3)(); b = (C.c_int64 * 3)()
    L.eo_rnd_state(a)
    assert L.eo_rnd_rand(0) == 0 and L.eo_rnd_erand(0) == 0
    L.eo_rnd_state(b)
    assert list(a) == list(b)
    assert 0 <= L.eo_rnd_rand(10) < 10
    L.eo_rnd_state(b)
    assert list(a) != list(b)


# ------------------------------------------------------------------ lists:sort/2
def test_lists_sort_is_stable_for_a_valid_ordering_function():
    L = O.lib()
    rnd = random.Random(5)
    for _ in range(400):
        n = rnd.randint(0, 41)
        keys = [rnd.randint(0, 6) for _ in range(n)]
        arr = (C.c_int32 * max(n, 1))(*keys)
        out = (C.c_int32 * max(n, 1))()
        L.eo_lists_sort(arr, n, 1, out)          # fun(A,B) -> A >= B   (weighted_permutations, :1249)
        want = sorted(range(n), key=lambda i: -keys[i])   # python's sort is stable
        assert list(out)[:n] == want


def test_lists_sort_strict_comparator_orders_by_priority():
    L = O.lib()
    rnd = random.Random(6)
    for _ in range(200):
        n = rnd.randint(1, 12)
        keys = [rnd.randint(0, 3) for _ in range(n)]
        arr = (C.c_int32 * n)(*keys); out = (C.c_int32 * n)()
        L.eo_lists_sort(arr, n, 0, out)          # fun(A,B) -> A > B   (sort_by_priority, utils :114-117)
        got = list(out)
        assert sorted(got) == list(range(n))
        ks = [keys[i] for i in got]
        assert ks == sorted(ks, reverse=True)    # priorities descend whatever happens to ties


# ------------------------------------------------------------------ erlamsa_mutations_test.erl mirrors
def run_until(code, data, pred, tries, **kw):
    for s in seeds(tries):
        out, _d, _rc = O.run_mutator(code, data, s, **kw)
        if pred(out):
            return True
    return False


def test_sed_num():                                                   # :74-77
    assert run_until("num", b" 100 + 100 + 100 ", lambda o: b"101" in o, 1500)


def random_lex_string(rnd):                                           # :51-63
    out = bytearray()
    for _ in range(rnd.randint(0, 41)):
        t = rnd.randint(0, 7)
        out.append({0: 92, 1: 34, 2: 39, 3: 0}.get(t, rnd.randint(0, 255) if t == 4 else 97))
    return bytes(out)


def test_string_lexer_roundtrip():                                     # :84-93
    L = O.lib()
    rnd = random.Random(7)
    cases = [bytes([233, 39, 39, 97, 97, 97, 0])] + [random_lex_string(rnd) for _ in range(10000)]
    for data in cases:
        out_p = C.c_void_p(); out_len = C.c_uint64(); n = C.c_int32()
        L.eo_lex_unlex(data, len(data), C.byref(out_p), C.byref(out_len), C.byref(n))
        got = C.string_at(out_p, out_len.value) if out_len.value else b""
        assert got == data


DASHES = b"-" * 40 + b'""' + b"-" * 50


def test_ascii_bad():                                                  # :96-100
    rx = re.compile(rb'^-*".*[%|a].*"-*$', re.S)
    assert run_until("ab", DASHES, lambda o: rx.match(o) is not None, 50)


def test_ascii_delimeter():                                            # :102-109 (cm_port 12345, cm_host 127.0.0.1)
    rx = re.compile(rb'^-*"-*$', re.S)
    opts = O.make_opts(ssrf_host="127.0.0.1", ssrf_port=12345)
    assert run_until("ad", DASHES, lambda o: rx.match(o) is not None, 50, opts=opts)


def test_sed_fuse_this():                                              # :115-119
    src = b"kittenslartibartfasterthaneelslartibartfastenyourseatbelts"
    assert run_until("ft", src, lambda o: o == b"kittenslartibartfastenyourseatbelts", 500)


def test_sed_tree_stutter():                                           # :126-130
    assert run_until("tr", b"(x (Y x))", lambda o: b"(x (x (x (x (Y x)))))" in o, 500)


def distinct(code, data, n):
    return {O.run_mutator(code, data, s)[0] for s in seeds(n, 77)}


def test_sed_tree_dup():                                               # :145-146
    assert len(distinct("tr2", b"(a) (b)", 30)) == 2


def test_sed_tree_swap_one():                                          # :148-149
    assert len(distinct("ts1", b"A (a) (b) (c) B", 400)) == 6


def test_sed_tree_swap_two():                                          # :151-152
    assert len(distinct("ts2", b"(a) (b (c))", 30)) == 3


FIVE = b"1\n 2\n  3\n  4\n    5\n"


def tokens(b):
    return [t for t in b.split(b"\n") if t]


def test_line_del():                                                   # :167-169
    out, _, _ = O.run_mutator("ld", FIVE, (9, 8, 7))
    assert len(tokens(out)) + 1 == len(tokens(FIVE))


def test_line_del_seq_statistics():                                    # :171-181
    src = b"0\n1\n 2\n  3\n   4\n    5\n     6\n      7\n       8\n         9\n"
    tot = sum(len(tokens(O.run_mutator("lds", src, s)[0])) for s in seeds(1000, 31))
    assert tot / 1000.0 < 0.75 * len(tokens(src))


def test_line_dup():                                                   # :183-185
    assert O.run_mutator("lr2", b"1\n", (4, 5, 6))[0] == b"1\n1\n"


def test_line_clone():                                                 # :187-194
    for s in seeds(20):
        assert O.run_mutator("lri", b"1\n2\n", s)[0] in (b"1\n2\n", b"1\n1\n", b"2\n2\n")


def test_line_repeat():                                                # :196-200
    assert len(tokens(O.run_mutator("lr", FIVE, (1, 1, 1))[0])) > len(tokens(FIVE))


def test_line_swap_length_and_correct():                               # :202-213
    assert len(tokens(O.run_mutator("ls", FIVE, (3, 3, 3))[0])) == len(tokens(FIVE))
    assert O.run_mutator("ls", b"A\n B\n", (3, 3, 3))[0] == b" B\nA\n"


def test_line_perm_length():                                           # :215-217
    assert len(tokens(O.run_mutator("lp", FIVE, (8, 8, 8))[0])) == len(tokens(FIVE))


def test_st_line_ins_fixed_seed():                                     # :223-230 -- the reference's only fixed-seed test
    out, _, _ = O.run_mutator("lis", b"ABC DEF", (1, 2, 3))
    assert out == b"ABC DEFABC DEF"          # draws: 10 x uniform(1), Up=16, pick=7, P=1 (SURVEY.md 8c trace)
    h = len(out) // 2
    assert out[:h] == out[h:]


def rand_block(seed):
    r = random.Random(seed)
    return bytes(r.randint(0, 255) for _ in range(r.randint(1, 4096)))


@pytest.mark.parametrize("code,check", [
    ("bd", lambda x, y: len(x) - 1 == len(y)),                                        # :247-252
    ("bi", lambda x, y: len(x) + 1 == len(y)),                                        # :254-258
    ("bf", lambda x, y: len(x) == len(y)),                                            # :266-270
    ("bei", lambda x, y: sum(y) - sum(x) in (1, -255)),                               # :272-278
    ("bed", lambda x, y: sum(y) - sum(x) in (-1, 255)),                               # :280-286
    ("ber", lambda x, y: len(x) == len(y)),                                           # :288-292
    ("sp", lambda x, y: len(x) == len(y) and sorted(x) == sorted(y)),                 # :295-299
    ("sd", lambda x, y: len(x) > len(y)),                                             # :301-305
    ("sr", lambda x, y: len(x) < len(y)),                                             # :307-311
])
def test_byte_level_properties(code, check):
    blk = rand_block(hash(code) & 0xffff)
    for s in seeds(300, 5):
        out, _, _ = O.run_mutator(code, blk, s)
        assert check(blk, out), (code, s)


def test_sed_byte_repeat():                                            # :260-264
    for s in seeds(200):
        assert O.run_mutator("br", b"\x01", s)[0] == b"\x01\x01"


def test_funny_unicode_table_size():
    assert O.lib().eo_funny_unicode_count() == 179                     # 17 manual + 162 encoded code points


# ------------------------------------------------------------------ driver level
def test_fuzzer_is_a_pure_function_of_seed_and_input():
    muts = {"bd": 1, "bf": 1, "num": 3, "sr": 1, "ld": 1}
    a, ma = O.fuzzer([b"hello 100 world\nfoo\n"] * 3, mutations=muts, seed=(1, 2, 3), n_cases=50)
    b, mb = O.fuzzer([b"hello 100 world\nfoo\n"] * 3, mutations=muts, seed=(1, 2, 3), n_cases=50)
    c, _ = O.fuzzer([b"hello 100 world\nfoo\n"] * 3, mutations=muts, seed=(1, 2, 4), n_cases=50)
    assert a == b and a != c
    assert len(set(a)) > 10
    # a window of the case loop equals the same cases of the full run (skip / first_case semantics)
    w, _ = O.fuzzer([b"hello 100 world\nfoo\n"] * 3, mutations=muts, seed=(1, 2, 3), n_cases=10, first_case=21)
    assert w == a[20:30]


def test_empty_and_tiny_inputs():
    outs, meta = O.fuzzer([b"", b"a", b"\n", b"0"], mutations={"bd": 1, "num": 1, "ld": 1, "sr": 1}, seed=(3, 2, 1), n_cases=200)
    # status 1 only where the `cp` pattern (first, or as a continuation) meets a block too short for a zlib header: the reference's streaming
    # inflate accepts the incomplete header and re-deflates, which is not restated (oracle/src/driver.cpp, P_CP)
    assert all(m.status in (0, 1) for m in meta) and sum(1 for m in meta if m.status == 1) <= 40
    outs, meta = O.fuzzer([b"", b"a", b"\n", b"0"], mutations={"bd": 1, "num": 1, "ld": 1, "sr": 1}, patterns={"od": 1, "nd": 1, "bu": 1, "co": 1, "nu": 1},
                          seed=(3, 2, 1), n_cases=200)     # (sk / sz / cs / ar draw their continuation from all ten patterns, cp included)
    assert all(m.status == 0 for m in meta)


# ---- sgm / js (src/erlamsa_sgml.erl, src/erlamsa_json.erl): the reference has no tests for them; the expectations
# below are derived by hand from the cited clauses and pin the restatement against regressions.
def _outs(code, data, n=300):
    res = {}
    for s in range(n):
        out, d, rc = O.run_mutator(code, data, (s + 1, s * 7 + 3, s * 13 + 5))
        res.setdefault((out if rc == 0 else None, d if rc == 0 else None, rc), 0)
        res[(out if rc == 0 else None, d if rc == 0 else None, rc)] += 1
    return res


def test_sgml_refuses_non_markup():
    # tz(nil, <<>>) :83 -> throw(incorrect_sgml): unchanged, delta -1; binarish data never reaches the tokenizer (:185-186)
    for data in (b"hello", b"no tags here at all", b"\x00\x01<a>x</a>", b"<a", b"<a b='1"):
        assert set(_outs("sgm", data, 40)) == {(data, -1.0, 0)}


def test_sgml_unterminated_comment_as_first_tag_kills_the_case():
    # tz({'!--',_}, <<>>) has no clause (:99-100): function_clause is an error, not the throw sgml_mutate/2 catches (:754)
    assert set(_outs("sgm", b"<!-- x", 20)) == {(None, None, 2)}
    # ... but after the first tag it sits inside the catch-all try of :74-79 and degrades to text
    outs = _outs("sgm", b"<a><!-- x", 200)
    assert all(rc == 0 for (_, _, rc) in outs) and (b"<a><!-- x", -1.0, 0) in outs


def test_sgml_refold_normalises_and_drops_leading_text():
    # bytes before the first '<' are skipped by tz(nil,_) :82; attributes refold as " name=value" (:290-301), sc as " />" (:313)
    outs = _outs("sgm", b"pre<a b='1'  c >t<br/></a>post", 400)
    changed = [o for (o, d, rc) in outs if rc == 0 and o != b"pre<a b='1'  c >t<br/></a>post"]
    assert changed and all(not o.startswith(b"pre") for o in changed)
    assert any(o == b"<a b='1' c>t<br /></a>post" for o in changed)          # a no-op mutation still re-serialises


def test_sgml_breaktag_reverses_children():
    # sgml_breaktag :590-601: Internals ++ [{open,..} | Tree] is prepended to a list that is reversed afterwards
    outs = _outs("sgm", b"<p>1<b>2</p>3</b>", 400)
    assert any(o == b"<p>2<b>13</b>" for (o, d, rc) in outs)
    # closed-earlier recovery (:225-234, :210-213) keeps the document text identical when nothing is mutated
    assert any(o == b"<p>1<b>2</p>3</b>" and d == -1.0 for (o, d, rc) in outs)


def test_sgml_xmlns_injection_uses_ssrf_endpoint():
    outs = _outs("sgm", b"<a>x</a>", 300)
    want = b'<a xmlns="http://localhost:51234/" xmlns:xsi="http://localhost:51234/" xsi:schemaLocation="http://localhost:51234/">x</a>'
    assert any(o == want and d == 1.0 for (o, d, rc) in outs)


def test_json_tokenizer_tolerance():
    # a second top-level value is refused (ws/3 with an empty context, :103); EOF inside a structure drops it (:89)
    assert set(_outs("js", b"1 2", 40)) == {(b"1 2", -1.0, 0)}
    assert set(_outs("js", b'{"a" 1}', 40)) == {(b'{"a" 1}', -1.0, 0)}
    assert set(_outs("js", b'{"a":1', 40)) == {(b"", -1.0, 0)}
    # anything without separators is one "number" token (:188-195); dup of the lone token refolds as a bracketed list (:277-279)
    outs = _outs("js", b"hello", 400)
    assert (b"hello", -1.0, 0) in outs and any(o == b"[hello,hello]" and d == 1.0 for (o, d, rc) in outs)
    # unterminated string -> junkstring with the quote appended (:185-186), folded with both quotes again (:263-264)
    assert any(o == b'"abc""' for (o, d, rc) in _outs("js", b'"abc', 50))


def test_json_structure_and_value_mutations():
    doc = b'{"a":[1,true,null],"b":"str"}'
    outs = _outs("js", doc, 600)
    got = {o for (o, d, rc) in outs if rc == 0}
    assert b'{"a":[1,false,null],"b":"str"}' in got                         # basic_type_mutation(Boolean)
    assert b'{"a":[1,true,"%n%s"],"b":"str"}' in got                        # mutate_null/2 :640-643
    assert any(o.startswith(b'{"@class"') or o.startswith(b'{"$type"') or o.startswith(b'{"__type"') for o in got)   # :615-618
    assert all(rc in (0, 2) for (_, _, rc) in outs)                         # swap/insert may hit the badmatch of :576-577
    pumped = {o for (o, d, rc) in _outs("js", b"[1,2,3]", 400) if rc == 0 and d == -2.0 and o.startswith(b"[")}
    assert b"[1,2,[1,2,[1,2,[1,2,3]]]]" in pumped                           # json_pump, PumpCnt = 2 (:560): the 2nd round re-inserts the pumped tree


def test_linear_time_builders_match_the_clause_by_clause_forms():
    """the sgml stack-discipline AST builder and the two-pass bracket parser are re-derivations (needed for blocks with
    thousands of unclosed tags / brackets); they must agree with the literal restatements of the reference's clauses"""
    L = O.lib()
    L.eo_sgml_selfcheck.argtypes = [C.c_char_p, C.c_uint64]
    L.eo_tree_selfcheck.argtypes = [C.c_char_p, C.c_uint64]
    r = random.Random(5)
    tags = [b"a", b"B", b"p", b"div"]
    for _ in range(6000):
        out = b""
        for _ in range(r.randint(1, 40)):
            k = r.randint(0, 9)
            t = r.choice(tags)
            if k < 3:
                out += b"<" + t + (b" x='1'" if r.random() < 0.3 else b"") + b">"
            elif k < 6:
                out += b"</" + t + b">"
            elif k == 6:
                out += b"<" + t + b"/>"
            elif k == 7:
                out += b"<!-- c -->" if r.random() < 0.5 else b"<?q?>"
            else:
                out += r.choice([b"text", b" ", b"x<y", b"1 < 2 > 3", b"<"])
        assert L.eo_sgml_selfcheck(out, len(out)) == 0, out
    al = b"()[]<>{}\"'ab \n"
    for _ in range(8000):
        d = bytes(r.choice(al) for _ in range(r.randint(0, 40)))
        assert L.eo_tree_selfcheck(d, len(d)) == 0, d
