"""The C-ABI shared library loads on a machine without a GPU and exports every entry point include/erlamsa_b200.h
declares; the Python mirror of the reference's name surface agrees with the reference's tables. No compute calls."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "erlamsa_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(eb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from erlamsa_b200 import _native as N
    assert os.path.exists(N.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(N.LIB_PATH)
    declared = header_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), "missing export " + name
    assert sorted(N.EXPORTED_SYMBOLS) == declared


def test_struct_layouts_match_header_sizes():
    from erlamsa_b200 import _native as N
    assert C.sizeof(N.Meta) == 16 + 64 + 8 + 8 + 24
    assert C.sizeof(N.Opts) % 8 == 0
    o = N.Opts()
    N.lib().eb200_default_opts(o)
    assert list(o.seed) == [1, 2, 3] and o.blockscale == 1.0 and o.gen_direct_pri == 500 and o.gen_random_pri == 1
    assert o.ssrf_host == b"localhost" and o.ssrf_port == 51234 and o.first_case == 1


def test_name_surface_matches_reference_tables():
    import erlamsa_b200 as E
    # reference src/erlamsa_mutations.erl:1291-1331 and src/erlamsa_patterns.erl:395-404
    assert E.mutator_codes() == ["sgm", "js", "uw", "ui", "ab", "ad", "tr2", "td", "num", "ts1", "tr", "ts2", "bd", "bei", "bed", "bf",
                                 "bi", "ber", "br", "sp", "sr", "sd", "snand", "srnd", "ld", "lds", "lr2", "lri", "lr", "ls", "lp",
                                 "lis", "lrs", "ft", "fn", "fo", "len", "b64", "uri", "zip", "nil"]
    assert dict(E.default_mutations())["sgm"] == 10 and dict(E.default_mutations())["b64"] == 7 and dict(E.default_mutations())["nil"] == 0
    assert E.pattern_codes() == ["od", "nd", "bu", "sk", "sz", "cs", "ar", "cp", "co", "nu"]
    assert [p for _, p in E.default_patterns()] == [1, 2, 1, 2, 2, 1, 1, 1, 0, 0]
    assert set(E.supported_mutations()) <= set(E.mutator_codes())


def test_no_device_means_loud_failure_not_fallback():
    import erlamsa_b200 as E
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(E.EngineError) as ei:
        E.Engine(0)
    assert ei.value.code == -6
    with pytest.raises(E.EngineError):
        E.erlamsa_app.fuzz(b"hello", {"seed": (1, 2, 3)})


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "erlamsa_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "liberlamsa_oracle" not in txt and "oracle/" not in txt, f
