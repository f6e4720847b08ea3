"""oracle/make_golden_conv.py -- TEST INFRASTRUCTURE.  Golden vectors for CONV mode from an INDEPENDENT statement of
its specification, written with scipy.ndimage only: no line of it is shared with
oracle/sspyr_oracle.c (orc_conv_build) or with the CUDA kernels.  The reference contains no convolution, so there is no
upstream vector to pin this mode to; this fixture is the second opinion both the C oracle (tests/test_oracle.py) and
the CUDA path (tests/test_gpu_conv.py) are compared with.

    python oracle/make_golden_conv.py            writes tests/golden/conv_scipy.npz

Specification restated: sigma_s = sigma0 * 2^(s/S); octave 0: G_0 = I (*) g(sqrt(max(sigma0^2 - sigma_in^2, 0.01)));
G_s = G_{s-1} (*) g(sqrt(sigma_s^2 - sigma_{s-1}^2)); taps exp(-k^2 / (2 sigma^2)) for |k| <= R = max(1, ceil(rs * sigma)),
normalised in double and rounded to float; clamp-to-edge border; row pass (stored as float) then column pass; octave
o+1: G_0 = G_S of octave o at even rows and columns (GuassDePyramid.h:80 phase); DoG_s = G_s - G_{s+1} (:143).
"""
from __future__ import annotations

import importlib.util
import os

import numpy as np
from scipy import ndimage as ndi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "sift-parallel-optimization_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

# name: (h, w, octaves, S, sigma0, sigma_in, radius_sigmas, pixel kind)
CASES = {
    "noise_96x128": (96, 128, 3, 3, 1.6, 0.5, 3.0, "i32"),
    "noise_135x241": (135, 241, 3, 3, 1.6, 0.5, 3.0, "i32"),
    "noise_80x72_S2": (80, 72, 3, 2, 1.2, 0.5, 4.0, "i32"),         # other sigma0 / S / radius
    "unit_64x96": (64, 96, 2, 3, 1.6, 0.5, 3.0, "f32"),          # [0,1]-normalised float pixels: the north_star's 1e-4 scale
    "pattern_140x420": (140, 420, 3, 3, 1.6, 0.5, 3.0, "i32"),   # a strip edge (420 > 3 x 128) and several 32-row steps
}


def pixels(name: str, h: int, w: int, kind: str) -> np.ndarray:
    if name.startswith("pattern"):
        return synth.pattern(h, w)
    img = synth.noise(h, w, frame=len(name))
    return (img / 255.0).astype(np.float32) if kind == "f32" else img


def taps(si: float, rs: float) -> np.ndarray:
    R = max(1, int(np.ceil(rs * si)))
    k = np.arange(-R, R + 1, dtype=np.float64)
    t = np.exp(-k * k / (2.0 * si * si))
    return (t / t.sum()).astype(np.float32).astype(np.float64)


def blur(a: np.ndarray, si: float, rs: float) -> np.ndarray:
    t = taps(si, rs)
    rows = ndi.correlate1d(a.astype(np.float64), t, axis=1, mode="nearest").astype(np.float32)
    return ndi.correlate1d(rows.astype(np.float64), t, axis=0, mode="nearest").astype(np.float32)


def pyramid(img: np.ndarray, octaves: int, S: int, sigma0: float, sigma_in: float, rs: float) -> list[np.ndarray]:
    sigma0, sigma_in, rs = float(np.float32(sigma0)), float(np.float32(sigma_in)), float(np.float32(rs))   # the C ABI takes floats
    sig = [sigma0 * 2.0 ** (s / S) for s in range(S + 3)]
    inc = [np.sqrt(max(sigma0 ** 2 - sigma_in ** 2, 0.01))] + [np.sqrt(sig[s] ** 2 - sig[s - 1] ** 2) for s in range(1, S + 3)]
    h, w = img.shape
    out, base = [], None
    for o in range(octaves):
        levels = [blur(img.astype(np.float32), inc[0], rs) if o == 0 else base]
        for s in range(1, S + 3):
            levels.append(blur(levels[-1], inc[s], rs))
        out.append(np.stack(levels))
        base = levels[S][::2, ::2][:h >> (o + 1), :w >> (o + 1)]
    return out


def main() -> None:
    blob = {}
    for name, (h, w, octs, S, s0, sin, rs, kind) in CASES.items():
        for o, g in enumerate(pyramid(pixels(name, h, w, kind), octs, S, s0, sin, rs)):
            blob[f"{name}_g{o}"] = g
    path = os.path.join(ROOT, "tests", "golden", "conv_scipy.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path) >> 10, "KiB,", len(blob), "arrays")


if __name__ == "__main__":
    main()
