"""GPU suite: parity of the CUDA path, called through the C ABI (libsspyr.so), against the golden vectors
of the reference header and the oracle.  REF mode is integer-geometry-exact and BIT-exact in fp32 (the
north_star's tolerance is 1e-4 of full scale; the kernels are built to meet 0 ulp)."""
from __future__ import annotations

import io

import numpy as np
import pytest

from conftest import KINDS, SMALL_CASES, bits_equal, split_flat

pytestmark = pytest.mark.gpu

TOL = 1e-4   # north_star: max abs error on [0,1]-normalised pixels, per level


# ---- the reference's own class surface ---------------------------------------------------------------
@pytest.mark.parametrize("n,S", SMALL_CASES)
@pytest.mark.parametrize("kind", KINDS)
def test_class_matches_header_golden(pkg, synth, golden_small, n, S, kind):
    img = synth.make(kind, n, n)
    want_init = split_flat(golden_small[f"{kind}_n{n}_S{S}_init"], n, S)
    want_g = split_flat(golden_small[f"{kind}_n{n}_S{S}_gauss"], n, S)
    want_d = split_flat(golden_small[f"{kind}_n{n}_S{S}_dog"], n, S)
    g = pkg.GaussPyramid(img, n, S)                               # GuassDePyramid.h:36
    try:
        assert g.layer == len(want_d) == n.bit_length()           # :48-53, exact
        assert g.initialized
        for o in range(g.layer):                                  # state after the ctor's GaussPyInit (:76-86)
            assert np.stack(g.GaussPy[o]).shape == (S + 3, n >> o, n >> o)
            assert bits_equal(np.stack(g.GaussPy[o]), want_init[o]), f"K0 octave {o}"
        for o in range(g.layer):                                  # public GaussFilter(o) (:106-134)
            g.GaussFilter(o)
            assert bits_equal(np.stack(g.GaussPy[o]), want_g[o]), f"GaussFilter octave {o}"
        g.GenerateDoG()                                           # :136-149
        for o in range(g.layer):
            assert bits_equal(np.stack(g.GaussPy[o]), want_d[o]), f"GenerateDoG octave {o}"
        g.GaussPyInit()                                           # re-callable reset (pThread.h:315-317)
        assert bits_equal(np.stack(g.GaussPy[0]), want_init[0])
    finally:
        g.close()


@pytest.mark.parametrize("key", ["pattern_n512_S3", "noise_n512_S3", "pattern_n512_S2", "pattern_n1080_S3",
                                 "noise_n1080_S3", "pattern_n2048_S3"])
def test_class_matches_header_hashes(pkg, O, synth, golden_hashes, key):
    kind, n, S = key.split("_")
    n, S = int(n[1:]), int(S[1:])
    g = pkg.GaussPyramid(synth.make(kind, n, n), n, S)
    try:
        g.GenerateDoG()
        want = golden_hashes[f"{key}_dog"]
        assert g.layer == len(want)
        for o in range(g.layer):
            assert [O.fnv1a64(g.GaussPy[o][s]) for s in range(S + 3)] == want[o], f"octave {o}"
    finally:
        g.close()


def test_known_answers(pkg, synth, golden_hashes):
    kat = golden_hashes["kat_n512_S3"]
    for kind in ("ones", "pattern"):
        g = pkg.GaussPyramid(synth.make(kind, 512, 512), 512, 3)
        g.GenerateDoG()
        assert [float(g.GaussPy[0][s][256][256]) for s in range(6)] == kat[kind]["inplace_center"]
        assert float(g.GaussPy[0][0][0][0]) == kat[kind]["corner"]
        g.close()


def test_output_prints_level0_of_every_octave(pkg, synth):
    g = pkg.GaussPyramid(synth.pattern(4, 4), 4, 1)
    buf = io.StringIO()
    g.output(buf)                                                  # GuassDePyramid.h:89-104
    lines = buf.getvalue().splitlines()
    assert lines[0].split() == [f"{v:g}" for v in synth.pattern(4, 4)[0]]
    assert lines[4] == "==" * 4 and lines[7] == "==" * 2 and lines[9] == "==" and len(lines) == 10
    g.close()


# ---- the superset C ABI: rectangles, chosen octaves, pixel types, outputs, slots, bands ------------------
CONFIGS = [  # (h, w, octaves) -- the BASELINE shapes at full or reduced size
    (512, 512, 4),      # C1
    (1080, 1920, 5),    # C2, full size
    (270, 480, 5),      # C3/C4 geometry / 8
    (135, 241, 3),      # odd width: scalar tails, 16-byte row alignment of odd rows
    (67, 120, 0),       # all octaves of a small rectangle
    (5, 1030, 3),
]


@pytest.mark.parametrize("h,w,octs", CONFIGS)
def test_rectangular_all_outputs_bit_exact(pkg, O, synth, h, w, octs):
    img = synth.noise(h, w)
    ref = O.ref_build(img, octaves=octs or None, S=3)
    with pkg.ScaleSpace(h, w, octs, 3) as ss:
        assert ss.octaves == len(ref["gauss"]) and ss.levels == 6 and ss.dogs == 5
        ss.upload(img)
        ss.build()
        gg, dd, ip = ss.download_gauss(), ss.download_dog(), ss.download_inplace()
        assert ss.last_launches() == 1
        for o in range(ss.octaves):
            r, c, pitch = ss.level_dims(o)
            assert (r, c) == (h >> o, w >> o) and pitch % 32 == 0 and pitch >= c
            assert bits_equal(gg[o], ref["gauss"][o]), f"gauss octave {o}"
            assert bits_equal(dd[o], ref["dog"][o]), f"dog octave {o}"
            assert bits_equal(ip[o], ref["inplace"][o]), f"in-place octave {o}"
        px = sum((h >> o) * (w >> o) for o in range(ss.octaves))
        assert ss.algorithmic_bytes() == 4 * h * w + 4 * 11 * px      # B_full, SURVEY section 8d


def test_window_tables_on_device_equal_the_oracle(pkg, O):
    h, w = 1080, 1920
    with pkg.ScaleSpace(h, w, 5, 3) as ss:
        for o in range(5):
            for s in range(6):
                assert bits_equal(ss.window_table(o, s, 0), O.window(h, o, s))
                assert bits_equal(ss.window_table(o, s, 1), O.window(w, o, s))


@pytest.mark.parametrize("S,sigma0", [(0, 2.0), (2, 2.0), (5, 2.0), (3, 1.6), (3, 7.5), (6, 2.0), (9, 2.0), (13, 2.0)])
def test_scales_and_sigma(pkg, O, synth, S, sigma0):
    img = synth.noise(96, 160)
    ref = O.ref_build(img, octaves=4, S=S, sigma0=sigma0)
    with pkg.ScaleSpace(96, 160, 4, S, sigma0=sigma0) as ss:
        ss.upload(img)
        ss.build()
        for o, a in enumerate(ss.download_inplace()):
            assert bits_equal(a, ref["inplace"][o])


def test_inplace_only_outputs_and_missing_planes(pkg, O, synth):
    img = synth.noise(128, 128)
    ref = O.ref_build(img, octaves=4, S=3)
    with pkg.ScaleSpace(128, 128, 4, 3, outputs=pkg.OUT_INPLACE) as ss:
        ss.upload(img)
        ss.build()
        for o, a in enumerate(ss.download_inplace()):
            assert bits_equal(a, ref["inplace"][o])
        assert ss.algorithmic_bytes() == 4 * 128 * 128 + 4 * 6 * sum((128 >> o) ** 2 for o in range(4))  # B_ref
        with pytest.raises(pkg.SspyrError) as e:
            ss.download(0, 1, pkg.KIND_GAUSS)
        assert e.value.code == pkg._lib.ERR_STATE
        with pytest.raises(pkg.SspyrError):
            ss.build(stage=pkg.STAGE_FILTER)


def test_u8_and_normalised_float_pixels(pkg, O, synth):
    h, w = 120, 200
    img = synth.noise(h, w)
    ref = O.ref_build(img, octaves=3, S=3)
    with pkg.ScaleSpace(h, w, 3, 3, pixel_type=pkg.PIXEL_U8) as ss:
        ss.upload(img.astype(np.uint8))
        ss.build()
        for o, a in enumerate(ss.download_inplace()):
            assert bits_equal(a, ref["inplace"][o])
    norm = (img / 255.0).astype(np.float32)
    fref = O.ref_build(norm, octaves=3, S=3)
    with pkg.ScaleSpace(h, w, 3, 3, pixel_type=pkg.PIXEL_F32) as ss:
        ss.upload(norm)
        ss.build()
        for o, a in enumerate(ss.download_inplace()):
            assert bits_equal(a, fref["inplace"][o])                       # same float pixels: still bit-exact
            err = np.max(np.abs(a - ref["inplace"][o] / 255.0), axis=(1, 2))    # vs the reference's int path
            assert np.all(err <= TOL), f"octave {o}: per-level max abs error {err}"


def test_frame_slots_batch_and_rebuild(pkg, O, synth):
    h, w, n = 72, 100, 5
    frames = [synth.noise(h, w, frame=f) for f in range(n)]
    with pkg.ScaleSpace(h, w, 3, 3, frames=n) as ss:
        with pytest.raises(pkg.SspyrError):
            ss.elapsed_ms()                                                  # timing is opt-in
        ss.set_tuning(timing=1)
        for f in range(n):
            ss.upload(frames[f], frame=f)
        ss.build_batch(0, n)
        assert ss.last_launches() == 1                                       # one launch for the batch (small frames: no L2 prefetch)
        for f in (0, 3, 4):
            ref = O.ref_build(frames[f], octaves=3, S=3)["inplace"]
            for o, a in enumerate(ss.download_inplace(frame=f)):
                assert bits_equal(a, ref[o]), f"frame {f} octave {o}"
        ss.upload(frames[0], frame=2)                                        # slot reuse
        with pytest.raises(pkg.SspyrError):
            ss.download(0, 0, pkg.KIND_DOG, frame=2)                         # stale until rebuilt
        ss.build_batch(4, 3)                                                 # wraps: slots 4, 0, 1
        ss.build(2)
        ref = O.ref_build(frames[0], octaves=3, S=3)["inplace"]
        assert bits_equal(ss.download_inplace(frame=2)[0], ref[0])
        assert ss.elapsed_ms() >= 0.0


def test_device_resident_input_via_torch(pkg, O, synth):
    import torch
    h, w = 96, 128
    img = synth.noise(h, w)
    t = torch.from_numpy(img).cuda()
    st = torch.cuda.Stream()
    with pkg.ScaleSpace(h, w, 4, 3) as ss:
        ss.set_stream(st.cuda_stream)
        st.wait_stream(torch.cuda.current_stream())
        ss.set_input_device(t.data_ptr(), t.stride(0) * 4)
        ss.build()
        ref = O.ref_build(img, octaves=4, S=3)["inplace"]
        for o, a in enumerate(ss.download_inplace()):
            assert bits_equal(a, ref[o])
        with pytest.raises(pkg.SspyrError):
            ss.set_input_device(t.data_ptr() + 4, 0)                         # misaligned


def test_producer_kernel_right_before_the_build(pkg, O, synth):
    """A build of a slot that reads a caller-produced device image is ordered after the producer's stores even though
    consecutive library launches normally overlap (programmatic dependent launch): several rounds of
    'producer kernel writes the image -> sspyr_build' on one stream with no sync in between."""
    import torch
    h, w = 1080, 1920
    st = torch.cuda.Stream()
    t = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    with pkg.ScaleSpace(h, w, 5, 3, outputs=pkg.OUT_INPLACE, frames=2) as ss:
        ss.set_stream(st.cuda_stream)
        ss.upload(synth.noise(h, w, frame=99), frame=1)
        ss.set_input_device(t.data_ptr(), t.stride(0) * 4, frame=0)
        src = [torch.from_numpy(synth.noise(h, w, frame=k)).cuda() for k in range(4)]
        torch.cuda.synchronize()
        with torch.cuda.stream(st):
            for k in range(4):
                ss.build(1)                               # a PDL-launched neighbour in the stream
                t.copy_(src[k])                           # producer kernel (device-to-device copy kernel + an elementwise op)
                t.add_(1).sub_(1)
                ss.build(0)
        ss.sync()
        ref = O.ref_build(synth.noise(h, w, frame=3), octaves=5, S=3, want=("inplace",))["inplace"]
        for o, a in enumerate(ss.download_inplace(0)):
            assert bits_equal(a, ref[o]), f"octave {o}"


@pytest.mark.parametrize("rpt,block,bx,pdl,occ", [(1, 128, 0, 1, 0), (2, 256, 32, 1, 0), (4, 64, 64, 0, 1), (4, 256, 128, 1, 3),
                                                  (1, 256, 96, 0, 2), (2, 96, 96, 1, 5), (1, 32, 32, 1, 1)])
def test_every_tuning_is_bit_exact(pkg, O, synth, rpt, block, bx, pdl, occ):
    h, w = 203, 330
    img = synth.noise(h, w)
    ref = O.ref_build(img, octaves=5, S=3)
    with pkg.ScaleSpace(h, w, 5, 3) as ss:
        ss.set_tuning(rows_per_thread=rpt, block=block, bx=bx, pdl=pdl, occ=occ)
        ss.upload(img)
        ss.build()
        for o, a in enumerate(ss.download_gauss()):
            assert bits_equal(a, ref["gauss"][o])
        for o, a in enumerate(ss.download_dog()):
            assert bits_equal(a, ref["dog"][o])


def test_row_bands_reproduce_the_full_frame(pkg, O, synth):
    """ROWBAND partition (C4/C5 geometry, reduced): each band is built by its own handle from its slice of the
    input and its slice of the row window; no halo in REF mode."""
    h, w, octs, world = 1080, 480, 5, 4
    img = synth.noise(h, w)
    full = O.ref_build(img, octaves=octs, S=3)
    for rank in range(world):
        row0, rows = pkg.band_rows(h, octs, world, rank)
        with pkg.ScaleSpace(rows, w, octs, 3, band_row0=row0, full_height=h) as ss:
            ss.upload(np.ascontiguousarray(img[row0:row0 + rows]))
            ss.build()
            gg, dd = ss.download_gauss(), ss.download_dog()
            for o in range(octs):
                lo, n = row0 >> o, rows >> o
                assert bits_equal(gg[o], full["gauss"][o][:, lo:lo + n]), f"rank {rank} octave {o}"
                assert bits_equal(dd[o], full["dog"][o][:, lo:lo + n])


def test_full_size_4k_frame_and_properties(pkg, O, synth):
    """C3 frame size (3840x2160, 5 octaves): full comparison against the threaded port plus size-independent
    properties: DoG_s == G_s - G_{s+1} recomputed from the downloaded planes, and exact linearity under x2."""
    h, w, octs = 2160, 3840, 5
    img = synth.noise(h, w) // 2                                     # 0..127 so that 2*img stays <= 255
    ref = O.ref_build(img, octaves=octs, S=3, want=("inplace",))
    with pkg.ScaleSpace(h, w, octs, 3, frames=2) as ss:
        ss.upload(img, frame=0)
        ss.upload(img * 2, frame=1)
        ss.build_batch(0, 2)
        a, b = ss.download_inplace(0), ss.download_inplace(1)
        gg = ss.download_gauss(0)
        for o in range(octs):
            assert bits_equal(a[o], ref["inplace"][o]), f"octave {o}"
            np.testing.assert_array_equal(a[o][:5], gg[o][:5] - gg[o][1:])
            normal = np.abs(a[o]) > 1e-30                                # x2 is exact away from denormals
            np.testing.assert_array_equal(b[o][normal], 2 * a[o][normal])


@pytest.mark.parametrize("mode_name", ["ref", "conv"])
@pytest.mark.parametrize("h,w,octs", [(135, 241, 4), (67, 515, 3), (200, 96, 5)])
def test_kernels_never_write_outside_the_level_rectangles(pkg, synth, mode_name, h, w, octs):
    """Own bounds check (compute-sanitizer is closed on this pool): fill the whole frame slot with a sentinel,
    build, and require every byte outside the H_o x W_o rectangles -- row padding up to the 128-byte pitch --
    to still hold it, while every in-range output was written."""
    import torch
    from sift_parallel_optimization_b200.exchange import device_view
    mode = pkg.MODE_REF if mode_name == "ref" else pkg.MODE_CONV
    S = 3
    with pkg.ScaleSpace(h, w, octs, S, mode=mode) as ss:
        dims = [ss.level_dims(o) for o in range(octs)]
        floats = sum((2 * (S + 3) - 1) * r * p for r, _, p in dims)
        base = ss.device_ptr(0, 0, pkg.KIND_GAUSS)
        slot = device_view(base, floats * 4, torch.device("cuda", torch.cuda.current_device())).view(torch.int32)
        slot.fill_(-1)                                            # 0xFFFFFFFF: a NaN no kernel produces
        ss.upload(synth.noise(h, w))
        ss.build()
        ss.sync()
        host = slot.cpu().numpy()
        off = 0
        for r, c, p in dims:
            planes = host[off:off + (2 * (S + 3) - 1) * r * p].reshape(2 * (S + 3) - 1, r, p)
            assert np.all(planes[:, :, c:] == -1), "a kernel wrote into the row padding"
            assert not np.any(planes[:, :, :c] == -1), "an output pixel was never written"
            off += planes.size


def test_reference_driver_on_the_cuda_class():
    """examples/drop_in_driver.cpp = the reference driver's call sequence on GaussPyramid_cuda, built by
    __graft_entry__.build() (linked against the serial header where /root/reference existed: then it exits 0
    only if max |cuda - serial| == 0)."""
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "build", "main_cuda")
    if not os.path.exists(exe):
        pytest.skip("build/main_cuda not built")
    for n in ("64", "512"):
        out = subprocess.run([exe, n], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stdout + out.stderr
        assert float(out.stdout.splitlines()[0]) > 0.0                     # mean ms per GenerateDoG(), as main.cpp:74 prints
        if "max |cuda - serial header|" in out.stdout:
            assert "max |cuda - serial header| = 0" in out.stdout


def test_pgm_ingest_example(pkg, O, synth, tmp_path):
    """examples/pgm_pyramid.cpp: 8-bit PGM in -> u8 ingest -> levels out as PGM, through the C ABI from C++."""
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "build", "pgm_pyramid")
    if not os.path.exists(exe):
        pytest.skip("build/pgm_pyramid not built")
    h, w = 120, 200
    img = synth.noise(h, w).astype(np.uint8)
    src = tmp_path / "in.pgm"
    src.write_bytes(b"P5\n# synthetic\n%d %d\n255\n" % (w, h) + img.tobytes())
    for mode in ("ref", "conv"):
        out = subprocess.run([exe, str(src), str(tmp_path / mode), mode, "3"], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stdout + out.stderr
        assert f"{w}x{h}, 3 octaves x 6 levels" in out.stdout
        for o in range(3):
            data = (tmp_path / f"{mode}_o{o}_g0.pgm").read_bytes()
            head = b"P5\n%d %d\n255\n" % (w >> o, h >> o)
            assert data.startswith(head) and len(data) == len(head) + (w >> o) * (h >> o)
    # CONV level 0 of octave 0 is a mild blur of the input: its PGM must match the specification to +-1 grey level
    got = np.frombuffer((tmp_path / "conv_o0_g0.pgm").read_bytes()[-h * w:], dtype=np.uint8).reshape(h, w).astype(np.float32)
    want = np.clip(O.conv_build(img.astype(np.int32), 3, 3)["gauss"][0][0] + 0.5, 0, 255).astype(np.uint8).astype(np.float32)
    assert np.max(np.abs(got - want)) <= 1


def _random_geometries(seed, n, max_h, max_w):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        h, w = int(rng.integers(1, max_h)), int(rng.integers(1, max_w))
        all_o = int(min(h, w)).bit_length()
        out.append((h, w, int(rng.integers(1, all_o + 1)), int(rng.integers(0, 6))))
    return out


@pytest.mark.parametrize("h,w,octs,S", _random_geometries(20261018, 24, 300, 700) + [(1, 1, 1, 3), (2, 3, 2, 0), (7, 4, 3, 2),
                                                                                      (3, 1029, 2, 3), (1025, 3, 2, 1)])
def test_random_geometries_ref_bit_exact(pkg, O, synth, h, w, octs, S):
    """Seeded sweep of odd shapes (widths below one quad, heights of one row, every octave/level count)."""
    img = synth.noise(h, w, frame=h * 1000 + w)
    ref = O.ref_build(img, octaves=octs, S=S)
    with pkg.ScaleSpace(h, w, octs, S) as ss:
        ss.upload(img)
        ss.build()
        for o, a in enumerate(ss.download_gauss()):
            assert bits_equal(a, ref["gauss"][o]), f"gauss octave {o}"
        for o, a in enumerate(ss.download_inplace()):
            assert bits_equal(a, ref["inplace"][o]), f"in-place octave {o}"
