"""Development aid: a small batch for `compute-sanitizer --tool racecheck` (shared-memory hazards between lanes of a warp)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import erlamsa_b200  # noqa: E402
import corpus  # noqa: E402

eng = erlamsa_b200.Engine(0)
blobs = corpus.web_corpus(0xE21A0900, 300)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
outs, meta = eng.fuzz_batch(blobs, {"mutations": {"uri": 1, "bd": 1}, "patterns": {"nd": 1, "bu": 1}, "seed": (5, 6, 5)}, n_cases=n)
print("done", sum(len(o) for o in outs))
