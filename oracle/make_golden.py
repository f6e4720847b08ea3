"""oracle/make_golden.py -- TEST INFRASTRUCTURE.  Generates tests/golden/ from the UNMODIFIED reference
header (oracle/_ref/libsiftref.so, built from /root/reference by oracle/Makefile).  The reference ships
no tests or fixtures of its own, so these vectors -- outputs of the reference itself
run in this container -- are what pins the oracle and the CUDA path.

    python oracle/make_golden.py        (needs /root/reference; rerun only if the header changes)

Writes
  tests/golden/header_small.npz   full in-place ('dog'), Gaussian ('gauss') and K0 ('init') results for
                                  small sides, every octave, as flat float32 arrays
  tests/golden/header_hashes.json FNV-1a-64 of every (octave, slot) plane for larger sides + KAT values
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path = [ROOT] + [p for p in sys.path if os.path.abspath(p or '.') != os.path.join(ROOT, 'oracle')]
from oracle import oracle as O  # noqa: E402

spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "sift-parallel-optimization_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

SMALL = [(1, 3), (2, 3), (3, 3), (5, 2), (8, 3), (16, 2), (16, 3), (37, 3), (64, 0), (67, 3), (100, 5)]
LARGE = [(512, 3), (512, 2), (1080, 3), (2048, 3)]
KINDS = ("ones", "pattern", "noise")


def main() -> None:
    assert O.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    small = {}
    for n, S in SMALL:
        for kind in KINDS:
            img = synth.make(kind, n, n)
            for what in ("dog", "gauss", "init"):
                planes = O.header_run(img, S, what)
                small[f"{kind}_n{n}_S{S}_{what}"] = np.concatenate([p.reshape(-1) for p in planes])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "header_small.npz"), **small)

    hashes = {"_doc": "FNV-1a-64 of raw float32 bytes, row-major, per [octave][slot]; "
                      "dog = GenerateDoG() in-place layout, gauss = GaussFilter(o) on every octave"}
    for n, S in LARGE:
        for kind in ("pattern", "noise"):
            img = synth.make(kind, n, n)
            for what in ("dog", "gauss"):
                planes = O.header_run(img, S, what)
                hashes[f"{kind}_n{n}_S{S}_{what}"] = [[O.fnv1a64(p[s]) for s in range(S + 3)] for p in planes]
    # known-answer values: n=512, S=3, octave 0, centre pixel
    c = 256
    kat = {}
    for kind in ("ones", "pattern"):
        img = synth.make(kind, 512, 512)
        g, d = O.header_run(img, 3, "gauss"), O.header_run(img, 3, "dog")
        kat[kind] = {"gauss_center": [float(g[0][s][c][c]) for s in range(6)],
                     "inplace_center": [float(d[0][s][c][c]) for s in range(6)],
                     "corner": float(d[0][0][0][0])}
    hashes["kat_n512_S3"] = kat
    with open(os.path.join(ROOT, "tests", "golden", "header_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1)
    print("golden written:", len(small), "small arrays;", len(hashes) - 2, "hash sets")


if __name__ == "__main__":
    main()
