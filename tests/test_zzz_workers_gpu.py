"""GPU: the host mirror's multi-threaded mode (`workers` > 1 with a file output) writes, for every case number, the file the oracle gives
for the same worker plan -- which tests/test_workers.py pins to the reference's own run on the CPU. Sorted last on purpose."""
import os

import pytest

import corpus

# the timeout's thread method ends the run instead of hanging it if a lane never comes back (a blocked C call cannot be interrupted by a signal)
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900, method="thread")]


@pytest.mark.parametrize("n,w,same", [(11, 3, False), (10, 3, True)])
def test_workers_mode_files(tmp_path, engine, oracle, n, w, same):
    from erlamsa_b200 import erlamsa_main
    from erlamsa_b200.workers import worker_plan
    files = [corpus.text_lines(corpus.rng(1200), 900), corpus.numeric_text(corpus.rng(1201), 500), corpus.random_bytes(corpus.rng(3), 700)]
    paths = []
    for i, b in enumerate(files):
        p = tmp_path / ("seed%d.bin" % i)
        p.write_bytes(b)
        paths.append(str(p))
    muts = {"bd": 1, "bf": 1, "num": 2, "ld": 1, "sr": 1, "fn": 1}
    pats = {"od": 1, "nd": 1, "sk": 1}
    gens = {"random": 1, "file": 1000}
    seed = (1, 2, 3)
    erlamsa_main._engine = engine                     # the session's engine instead of a second context
    res = erlamsa_main.fuzzer({"paths": paths, "output": str(tmp_path / "out" / "%n"), "n": n, "seed": seed, "mutations": muts, "patterns": pats,
                               "generators": gens, "workers": w, "workers_same_seed": same, "max_case_out": 1 << 24})
    assert res == []
    assert sorted(os.listdir(tmp_path / "out"), key=int) == [str(i) for i in range(1, n + 1)]
    for wseed, first, cnt, stream_first in worker_plan(seed, "out/%n", n, w, same):
        want, wm = oracle.fuzzer(files, mutations=muts, patterns=pats, seed=seed, generators=gens, n_cases=cnt, first_case=first, max_case_out=1 << 24,
                                 case_stream=(wseed, stream_first))
        for k in range(cnt):
            assert wm[k].status == 0 and (tmp_path / "out" / str(first + k)).read_bytes() == want[k], first + k
