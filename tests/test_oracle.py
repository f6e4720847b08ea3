"""CPU suite, part 1: pin the oracle.  The C restatement must reproduce, bit for bit, (a) the golden
vectors the unmodified reference header produced (tests/golden, oracle/make_golden.py) and (b) the
header itself when oracle/_ref is present (always in the build container)."""
from __future__ import annotations

import numpy as np
import pytest

from conftest import KINDS, SMALL_CASES, bits_equal, split_flat


# ---- (a) golden vectors: work without /root/reference -------------------------------------------------
@pytest.mark.parametrize("n,S", SMALL_CASES)
@pytest.mark.parametrize("kind", KINDS)
def test_port_matches_header_golden_small(O, synth, golden_small, n, S, kind):
    img = synth.make(kind, n, n)
    want_dog = split_flat(golden_small[f"{kind}_n{n}_S{S}_dog"], n, S)
    want_g = split_flat(golden_small[f"{kind}_n{n}_S{S}_gauss"], n, S)
    want_init = split_flat(golden_small[f"{kind}_n{n}_S{S}_init"], n, S)
    octs = len(want_dog)
    assert octs == O.octaves_all(n, n) == n.bit_length()          # GuassDePyramid.h:48-53
    mirror = O.ref_mirror(img, S=S)
    mirror_g = O.ref_mirror(img, S=S, do_dog=False)
    closed = O.ref_build(img, S=S)
    for o in range(octs):
        assert want_dog[o].shape == (S + 3, n >> o, n >> o)        # :64,:66
        assert bits_equal(mirror[o], want_dog[o]), f"mirror in-place, octave {o}"
        assert bits_equal(mirror_g[o], want_g[o]), f"mirror gauss, octave {o}"
        assert bits_equal(closed["inplace"][o], want_dog[o]), f"closed-form in-place, octave {o}"
        assert bits_equal(closed["gauss"][o], want_g[o]), f"closed-form gauss, octave {o}"
        assert bits_equal(closed["dog"][o], want_dog[o][:S + 2]), f"closed-form dog, octave {o}"
        # K0: every level is the decimated original (GuassDePyramid.h:76-86)
        dec = img[::1 << o, ::1 << o][:n >> o, :n >> o].astype(np.float32)
        for s in range(S + 3):
            assert bits_equal(want_init[o][s], dec)


@pytest.mark.parametrize("key", ["pattern_n512_S3", "noise_n512_S3", "pattern_n512_S2", "pattern_n1080_S3",
                                 "noise_n1080_S3"])
def test_port_matches_header_hashes(O, synth, golden_hashes, key):
    kind, n, S = key.split("_")
    n, S = int(n[1:]), int(S[1:])
    img = synth.make(kind, n, n)
    got = O.ref_build(img, S=S, want=("gauss", "inplace"))
    for what, name in (("inplace", "dog"), ("gauss", "gauss")):
        want = golden_hashes[f"{key}_{name}"]
        assert len(want) == len(got[what])
        for o, planes in enumerate(got[what]):
            assert [O.fnv1a64(planes[s]) for s in range(S + 3)] == want[o], f"{name} octave {o}"


def test_known_answers_from_the_survey(O, synth, golden_hashes):
    """SURVEY section 8c KATs (n=512, S=3, octave 0, centre 256) -- frozen from the unmodified header."""
    kat = golden_hashes["kat_n512_S3"]
    assert kat["ones"]["gauss_center"] == pytest.approx(
        [0.0373792462, 0.123953938, 0.204044923, 0.234206781, 0.20851095, 0.150978088], rel=1e-7)
    assert kat["ones"]["inplace_center"] == pytest.approx(
        [-0.0865746886, -0.0800909847, -0.0301618576, 0.0256958306, 0.0575328618, 0.150978088], rel=1e-7)
    assert kat["pattern"]["inplace_center"] == pytest.approx(
        [-0.259724081, -0.240272939, -0.0904855728, 0.0770874619, 0.1725986, 0.452934265], rel=1e-7)
    assert kat["ones"]["corner"] == 0.0                            # the window underflows at the corners
    assert golden_hashes["pattern_n512_S3_dog"][0][0] == "0a50f09285dc5d0a"
    assert golden_hashes["pattern_n512_S3_gauss"][0][0] == "30bd2dcf3d944cfd"
    for kind in ("ones", "pattern"):
        got = O.ref_build(synth.make(kind, 512, 512), S=3)
        assert [float(got["gauss"][0][s][256][256]) for s in range(6)] == kat[kind]["gauss_center"]
        assert [float(got["inplace"][0][s][256][256]) for s in range(6)] == kat[kind]["inplace_center"]


# ---- (b) the header itself ---------------------------------------------------------------------------
@pytest.mark.parametrize("n,S,kind", [(33, 3, "noise"), (128, 4, "pattern"), (250, 1, "noise"), (600, 3, "noise")])
def test_port_matches_live_header(O, synth, n, S, kind):
    if not O.have_ref():
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    img = synth.make(kind, n, n)
    hd, hg = O.header_run(img, S, "dog"), O.header_run(img, S, "gauss")
    got = O.ref_build(img, S=S)
    mir = O.ref_mirror(img, S=S)
    for o in range(len(hd)):
        assert bits_equal(got["inplace"][o], hd[o]) and bits_equal(got["gauss"][o], hg[o])
        assert bits_equal(mir[o], hd[o])


def test_reference_parallel_variants_agree_with_serial(O, synth):
    """Secondary oracles: pThread GenerateDoG_i, OpenMP GenerateDoG, AVX512xPTHREAD."""
    if not O.have_ref():
        pytest.skip("oracle/_ref not built here")
    img = synth.noise(256, 256)
    serial = O.header_run(img, 3, "dog")
    variants = ["pthread_i", "omp"] + (["a512xp"] if O.load_ref_avx512() is not None else [])
    for v in variants:
        got = O.header_run(img, 3, v, threads=4)
        for o in range(len(serial)):
            assert bits_equal(got[o], serial[o]), f"{v} octave {o}"


# ---- properties of the restatement itself --------------------------------------------------------------
def test_window_uses_the_serial_float_halving_rule(O):
    """GuassDePyramid.h:107-115: 1080 -> len 67.5 at octave 4 -> 67 samples centred on 33.25."""
    f = O.window(1080, 4, 0, 2.0)
    assert f.size == 67
    k = np.arange(67, dtype=np.float32)
    sig = np.float32(2.0)
    want = np.exp(-(k - np.float32(33.25)) ** 2 / (2 * sig * sig)).astype(np.float32)
    assert np.argmax(f) == 33 and f[33] > f[34] > f[32]            # 33.25 is closer to 33, then 34
    big = want > 1e-20                                             # skip the denormal tails
    np.testing.assert_allclose((f / f[33])[big], (want / want[33])[big], rtol=1e-5)
    assert O.window(512, 0, 0, 2.0)[255] == O.window(512, 0, 0, 2.0)[256]   # centre 255.5: symmetric


@pytest.mark.parametrize("h,w,octs", [(67, 120, 5), (135, 240, 3), (30, 17, 0), (1, 9, 1)])
def test_rectangular_closed_form_equals_mirror(O, synth, h, w, octs):
    img = synth.noise(h, w)
    octs = octs or O.octaves_all(h, w)
    mir = O.ref_mirror(img, octaves=octs, S=3)
    got = O.ref_build(img, octaves=octs, S=3)
    for o in range(octs):
        assert got["inplace"][o].shape == (6, h >> o, w >> o)
        assert bits_equal(got["inplace"][o], mir[o])


def test_row_bands_equal_slices_of_the_full_frame(O, synth):
    h, w, octs, S = 208, 96, 5, 3
    img = synth.noise(h, w)
    full = O.ref_build(img, octaves=octs, S=S)
    for row0, rows in ((0, 64), (64, 96), (160, 48)):
        band = O.ref_build(img[row0:row0 + rows], octaves=octs, S=S, row0=row0, full_h=h)
        for o in range(octs):
            lo, n = row0 >> o, rows >> o
            assert bits_equal(band["gauss"][o], full["gauss"][o][:, lo:lo + n])
            assert bits_equal(band["dog"][o], full["dog"][o][:, lo:lo + n])


def test_float_pixels_commute_with_the_int_path(O, synth):
    """[0,1]-normalised float input: the pipeline is linear in p, so oracle(int)/255 ~ oracle(float(p)/255)."""
    img = synth.noise(64, 64)
    a = O.ref_build(img, S=3)["inplace"]
    b = O.ref_build((img / 255.0).astype(np.float32), S=3)["inplace"]
    for o in range(len(a)):
        assert np.max(np.abs(a[o] / 255.0 - b[o])) <= 1e-4


# ---- CONV-mode specification (no upstream parity; these pin its internal consistency) ------------------
def test_conv_taps_are_normalised_and_symmetric(O):
    for s in range(6):
        t = O.conv_taps(s, 3, 1.6, 0.5, 3.0)
        assert t.size % 2 == 1 and abs(float(t.astype(np.float64).sum()) - 1.0) < 1e-6
        np.testing.assert_array_equal(t, t[::-1])
        assert np.argmax(t) == t.size // 2


def test_conv_oracle_basic_properties(O):
    h, w, octs, S = 48, 64, 3, 3
    const = np.full((h, w), 7, dtype=np.int32)
    out = O.conv_build(const, octs, S)
    for o in range(octs):
        assert out["gauss"][o].shape == (S + 3, h >> o, w >> o)
        np.testing.assert_allclose(out["gauss"][o], 7.0, rtol=1e-6)      # DC gain 1, clamp border
        np.testing.assert_allclose(out["dog"][o], 0.0, atol=1e-5)
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (h, w)).astype(np.int32)
    out = O.conv_build(img, octs, S)
    for o in range(octs):
        g = out["gauss"][o]
        np.testing.assert_array_equal(out["dog"][o], g[:-1] - g[1:])     # DoG_s = G_s - G_{s+1}
        if o:
            np.testing.assert_array_equal(g[0], out["gauss"][o - 1][S][::2, ::2][:h >> o, :w >> o])
        v = [float(g[s].var()) for s in range(S + 3)]
        assert all(v[s] > v[s + 1] for s in range(S + 2))                # blur only ever smooths


def test_extrema_oracle_finds_a_planted_peak(O):
    dog = np.zeros((5, 9, 9), dtype=np.float32)
    dog[2, 4, 4] = 1.0
    dog[1, 2, 6] = -2.0
    f = O.extrema_octave(dog, 0.5)
    assert f.shape == (3, 9, 9) and f.sum() == 2 and f[1, 4, 4] == 1 and f[0, 2, 6] == 1


@pytest.mark.parametrize("h,w,octs,S,sigma0,rs", [(96, 130, 3, 3, 1.6, 3.0), (61, 47, 2, 2, 1.2, 4.0), (128, 128, 4, 3, 2.0, 3.0)])
def test_conv_oracle_against_an_independent_scipy_restatement(O, h, w, octs, S, sigma0, rs):
    """CONV mode has no upstream parity, so its C oracle (orc_conv_build) is at least checked against a second,
    independent statement of the same specification (DESIGN.md 'CONV mode') written here with scipy.ndimage:
    sigma_s = sigma0 * 2^(s/S); G_0 = I * g(sqrt(max(sigma0^2 - sigma_in^2, 0.01))); G_s = G_{s-1} * g(sqrt(sigma_s^2 -
    sigma_{s-1}^2)); taps exp(-k^2 / 2 sigma^2) over |k| <= ceil(rs * sigma), normalised; clamp-to-edge border; row
    pass then column pass; next octave = G_S at even rows / columns; DoG_s = G_s - G_{s+1}."""
    ndi = pytest.importorskip("scipy.ndimage")
    sigma_in = 0.5
    sigma0, rs = float(np.float32(sigma0)), float(np.float32(rs))     # the API takes them as C floats
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (h, w)).astype(np.int32)

    def taps(si):
        R = max(1, int(np.ceil(rs * si)))
        k = np.arange(-R, R + 1, dtype=np.float64)
        t = np.exp(-k * k / (2.0 * si * si))
        return (t / t.sum()).astype(np.float32).astype(np.float64)

    def blur(a, si):
        t = taps(si)
        rows = ndi.correlate1d(a.astype(np.float64), t, axis=1, mode="nearest").astype(np.float32)   # row pass, stored as fp32
        return ndi.correlate1d(rows.astype(np.float64), t, axis=0, mode="nearest").astype(np.float32)

    sig = [sigma0 * 2.0 ** (s / S) for s in range(S + 3)]
    inc = [np.sqrt(max(sigma0 ** 2 - sigma_in ** 2, 0.01))] + [np.sqrt(sig[s] ** 2 - sig[s - 1] ** 2) for s in range(1, S + 3)]
    got = O.conv_build(img, octs, S, sigma0=sigma0, sigma_in=sigma_in, radius_sigmas=rs)
    base = None
    for o in range(octs):
        levels = [blur(img.astype(np.float32), inc[0]) if o == 0 else base]
        for s in range(1, S + 3):
            levels.append(blur(levels[-1], inc[s]))
            mine = O.conv_taps(s, S, sigma0, sigma_in, rs)                    # same radius, taps equal to an fp32 ulp
            assert mine.size == taps(inc[s]).size
            np.testing.assert_allclose(mine.astype(np.float64), taps(inc[s]), rtol=2e-7, atol=1e-10)
        want = np.stack(levels)
        assert got["gauss"][o].shape == want.shape == (S + 3, h >> o, w >> o)
        np.testing.assert_allclose(got["gauss"][o], want, rtol=0, atol=2e-5 * 255)
        np.testing.assert_allclose(got["dog"][o], want[:-1] - want[1:], rtol=0, atol=4e-5 * 255)
        base = got["gauss"][o][S][::2, ::2][:h >> (o + 1), :w >> (o + 1)]      # (the oracle's own G_S: errors do not compound)


# ---- CONV golden fixture (tests/golden/conv_scipy.npz, oracle/make_golden_conv.py): an independent scipy statement ----
def _conv_golden_cases():
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("make_golden_conv", os.path.join(ROOT, "oracle", "make_golden_conv.py"))
    try:
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)                       # needs scipy only to regenerate; CASES / pixels() do not
    except ImportError:
        return None
    return mod


@pytest.mark.parametrize("name", ["noise_96x128", "noise_135x241", "noise_80x72_S2", "unit_64x96", "pattern_140x420"])
def test_conv_oracle_against_the_committed_scipy_fixture(O, name):
    """orc_conv_build vs the committed golden vectors of the independent restatement: <= 2e-5 of full scale per
    level (double accumulation on both sides; the fixture rounds the row pass to float like the kernels do)."""
    import os
    from conftest import GOLDEN
    mod = _conv_golden_cases()
    if mod is None:
        pytest.skip("scipy not importable")
    gold = np.load(os.path.join(GOLDEN, "conv_scipy.npz"))
    h, w, octs, S, s0, sin, rs, kind = mod.CASES[name]
    img = mod.pixels(name, h, w, kind)
    got = O.conv_build(img, octs, S, sigma0=s0, sigma_in=sin, radius_sigmas=rs)
    scale = 1.0 if kind == "f32" else 255.0
    for o in range(octs):
        want = gold[f"{name}_g{o}"]
        assert got["gauss"][o].shape == want.shape
        assert np.max(np.abs(got["gauss"][o].astype(np.float64) - want)) <= 2e-5 * scale, f"octave {o}"
        assert np.max(np.abs(got["dog"][o].astype(np.float64) - (want[:-1].astype(np.float64) - want[1:]))) <= 4e-5 * scale
