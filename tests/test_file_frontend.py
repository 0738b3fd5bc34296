"""`./erlamsa dir/* -n N -o out/%n` on the engine: file generator (src/erlamsa_gen.erl:104-121) +
file writer with %n templating (src/erlamsa_out.erl:103-123) through the Python mirror of erlamsa_main:fuzzer/1."""
import os

import pytest


def test_file_name_template():
    from erlamsa_b200.erlamsa_main import _file_name
    assert _file_name("out/fuzz-%n.bin", 7) == "out/fuzz-7.bin"
    assert _file_name("%n-%n", 12) == "12-12"
    assert _file_name("plain", 3) == "plain"


@pytest.mark.gpu
def test_files_in_files_out(tmp_path, engine, oracle):
    import corpus
    from erlamsa_b200 import erlamsa_main
    files = [corpus.text_lines(corpus.rng(77), 6000), corpus.numeric_text(corpus.rng(78), 3000), corpus.random_bytes(corpus.rng(79), 5000)]
    paths = []
    for i, b in enumerate(files):
        p = tmp_path / ("seed%d.bin" % i)
        p.write_bytes(b)
        paths.append(str(p))
    muts = {"bd": 1, "bf": 1, "num": 2, "ld": 1, "sr": 1, "fn": 1}
    pats = {"od": 1, "nd": 1, "sk": 1}
    n = 24
    res = erlamsa_main.fuzzer({"paths": paths, "output": str(tmp_path / "out" / "case-%n.fuzz"), "n": n, "seed": (1, 2, 3), "mutations": muts, "patterns": pats,
                               "max_case_out": 1 << 24})
    assert res == []
    want, wm = oracle.fuzzer(files, mutations=muts, patterns=pats, seed=(1, 2, 3), generators={"random": 1, "jump": 100, "file": 1000}, n_cases=n, max_case_out=1 << 24)   # the mirror's (= the reference's) defaults for several paths
    for k in range(n):
        got = (tmp_path / "out" / ("case-%d.fuzz" % (k + 1))).read_bytes()
        assert wm[k].status == 0 and got == want[k], k
    # skip: only the later cases are written, with their own numbers
    erlamsa_main.fuzzer({"paths": paths, "output": str(tmp_path / "o2" / "%n"), "n": 6, "skip": 4, "seed": (1, 2, 3), "mutations": muts, "patterns": pats})
    assert sorted(os.listdir(tmp_path / "o2")) == ["5", "6"]
    assert (tmp_path / "o2" / "5").read_bytes() == want[4]
