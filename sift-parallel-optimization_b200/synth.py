"""Deterministic synthetic grayscale frames (int32, values 0..255 -- exactly representable in fp32,
and the pixel type the reference constructor takes: ``GaussPyramid(int** img, int len, int S)``,
GuassDePyramid.h:36).

* ``ones``     what the reference driver feeds (main.cpp:31-35)
* ``pattern``  p[i][j] = (131 i + 71 j + (i j mod 13)) mod 256
* ``noise``    splitmix64(seed + linear index) >> 56, seed = 0x5EED0001 + frame  (counter based, so any
               row band of a huge frame can be generated on its own rank without the rest)
"""
from __future__ import annotations

import numpy as np

NOISE_SEED = 0x5EED0001


def ones(h: int, w: int) -> np.ndarray:
    return np.ones((h, w), dtype=np.int32)


def pattern(h: int, w: int, row0: int = 0) -> np.ndarray:
    i = (np.arange(h, dtype=np.int64) + row0)[:, None]
    j = np.arange(w, dtype=np.int64)[None, :]
    return ((131 * i + 71 * j + (i * j) % 13) % 256).astype(np.int32)


def noise(h: int, w: int, frame: int = 0, row0: int = 0, full_w: int | None = None) -> np.ndarray:
    full_w = full_w or w
    idx = (np.arange(h, dtype=np.uint64) + np.uint64(row0))[:, None] * np.uint64(full_w) \
        + np.arange(w, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        z = idx + np.uint64(NOISE_SEED + frame) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(56)).astype(np.int32)


def make(kind: str, h: int, w: int, frame: int = 0, row0: int = 0) -> np.ndarray:
    if kind == "ones":
        return ones(h, w)
    if kind == "pattern":
        return pattern(h, w, row0)
    if kind == "noise":
        return noise(h, w, frame, row0)
    raise ValueError(f"unknown synthetic input {kind!r}")
