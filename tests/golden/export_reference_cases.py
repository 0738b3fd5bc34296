#!/usr/bin/env python
"""tests/golden/reference_vectors.json -> tests/golden/reference_cases.term (file:consult/1 format) for
tools/dump_reference_vectors.erl, which re-runs the same cases on a real Erlang/OTP."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def plist(d):
    return "default" if d is None else "[" + ",".join("{%s,%d}" % (k, v) for k, v in d.items()) + "]"


def main():
    vec = json.load(open(os.path.join(HERE, "reference_vectors.json")))["vectors"]
    with open(os.path.join(HERE, "reference_cases.term"), "w") as f:
        for v in vec:
            blobs = v["blobs"]
            extra = []
            if "generators" in v["extra"]:
                extra.append("{generators,%s}" % plist(v["extra"]["generators"]))
            if "blockscale" in v["extra"]:
                extra.append("{blockscale,%r}" % float(v["extra"]["blockscale"]))
            for k in range(v["n_cases"]):
                i = v["first_case"] + k
                f.write('{"%s",%d,"%s",%d,{%d,%d,%d},%s,%s,[%s]}.\n' % (v["name"], k, blobs[(i - 1) % len(blobs)], i, v["seed"][0], v["seed"][1], v["seed"][2],
                                                                       plist(v["mutations"]), plist(v["patterns"]), ",".join(extra)))
    print("wrote reference_cases.term")


if __name__ == "__main__":
    main()
