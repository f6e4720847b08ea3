"""GPU suite, CONV mode: the true separable-blur scale space (north_star) against its CPU specification
(oracle/sspyr_oracle.c orc_conv_build).  There is no upstream parity for this mode -- the reference has no
convolution -- so the bar is the north_star's tolerance against OUR restated oracle: max abs error <= 1e-4 of
full scale per level (pixels 0..255 -> 0.0255; [0,1] floats -> 1e-4); level/DoG counts, dims, DoG order and
decimation phase follow the reference (GuassDePyramid.h:64,66,80,140,143) and are exact."""
from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-4


def check(got, want, scale, what):
    for o, (a, b) in enumerate(zip(got, want)):
        assert a.shape == b.shape, (what, o, a.shape, b.shape)
        err = np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)), axis=(1, 2))
        assert np.all(err <= TOL * scale), f"{what} octave {o}: per-level max abs error {err} > {TOL * scale}"


def test_taps_equal_the_specification(pkg, O):
    for S, sigma0, rs in ((3, 1.6, 3.0), (3, 1.6, 4.0), (2, 1.2, 3.0), (5, 2.0, 3.0)):
        with pkg.ScaleSpace(64, 64, 2, S, sigma0=sigma0, mode=pkg.MODE_CONV, radius_sigmas=rs) as ss:
            for s in range(S + 3):
                np.testing.assert_array_equal(ss.conv_taps(s), O.conv_taps(s, S, sigma0, 0.5, rs))


@pytest.mark.parametrize("h,w,octs,S,rs", [(270, 480, 5, 3, 3.0), (135, 241, 3, 3, 3.0), (1080, 1920, 5, 3, 3.0),
                                           (200, 333, 4, 3, 4.0), (96, 128, 3, 2, 3.0), (64, 70, 2, 5, 4.0),
                                           (40, 1000, 3, 2, 3.0)])
def test_full_pyramid_within_tolerance(pkg, O, synth, h, w, octs, S, rs):
    img = synth.noise(h, w)
    ref = O.conv_build(img, octs, S, radius_sigmas=rs)
    with pkg.ScaleSpace(h, w, octs, S, mode=pkg.MODE_CONV, radius_sigmas=rs) as ss:
        assert (ss.octaves, ss.levels, ss.dogs) == (octs, S + 3, S + 2)
        ss.upload(img)
        ss.build()
        gg, dd = ss.download_gauss(), ss.download_dog()
        for o in range(octs):
            assert ss.level_dims(o)[:2] == (h >> o, w >> o)
        check(gg, ref["gauss"], 255.0, "gauss")
        check(dd, ref["dog"], 255.0, "dog")
        for o in range(1, octs):            # exact: next octave base = even-phase decimation of G_S
            np.testing.assert_array_equal(gg[o][0], gg[o - 1][S][::2, ::2][:h >> o, :w >> o])
        for o in range(octs):               # exact up to the fused subtraction's own rounding
            np.testing.assert_allclose(dd[o], gg[o][:-1] - gg[o][1:], atol=1e-5 * 255)


def test_pixel_types_sigma_and_constant_image(pkg, O, synth):
    h, w, octs = 150, 260, 4
    img = synth.noise(h, w)
    ref = O.conv_build(img, octs, 3, sigma0=2.0, sigma_in=0.0)
    for pix, arr, scale in ((pkg.PIXEL_U8, img.astype(np.uint8), 255.0), (pkg.PIXEL_I32, img, 255.0)):
        with pkg.ScaleSpace(h, w, octs, 3, sigma0=2.0, sigma_in=0.0, mode=pkg.MODE_CONV, pixel_type=pix) as ss:
            ss.upload(arr)
            ss.build()
            check(ss.download_gauss(), ref["gauss"], scale, f"gauss pix{pix}")
    norm = (img / 255.0).astype(np.float32)
    fref = O.conv_build(norm, octs, 3)
    with pkg.ScaleSpace(h, w, octs, 3, mode=pkg.MODE_CONV, pixel_type=pkg.PIXEL_F32) as ss:
        ss.upload(norm)
        ss.build()
        check(ss.download_gauss(), fref["gauss"], 1.0, "gauss f32")       # the north_star's 1e-4 on [0,1] pixels
        check(ss.download_dog(), fref["dog"], 1.0, "dog f32")
    const = np.full((h, w), 9, dtype=np.int32)
    with pkg.ScaleSpace(h, w, octs, 3, mode=pkg.MODE_CONV) as ss:
        ss.upload(const)
        ss.build()
        for g in ss.download_gauss():
            np.testing.assert_allclose(g, 9.0, rtol=2e-6)                  # DC gain 1 with clamp-to-edge borders


def test_batched_frame_slots(pkg, O, synth):
    h, w, octs, n = 120, 200, 3, 4
    frames = [synth.noise(h, w, frame=f) for f in range(n)]
    with pkg.ScaleSpace(h, w, octs, 3, mode=pkg.MODE_CONV, frames=n) as ss:
        for f in range(n):
            ss.upload(frames[f], frame=f)
        ss.set_tuning(conv_cascade=2)
        ss.build_batch(0, n)
        assert ss.last_launches() == 1                       # cascade: one launch for the whole batch, all levels
        ss.set_tuning(conv_cascade=0)
        ss.build_batch(0, n)
        assert ss.last_launches() == 6 + 5 * (octs - 1)      # per-level path: one launch per level for the whole batch
        for f in (0, 3):
            check(ss.download_gauss(frame=f), O.conv_build(frames[f], octs, 3)["gauss"], 255.0, f"frame {f}")


@pytest.mark.parametrize("world", [2, 3])
def test_row_bands_with_halo_exchange(pkg, O, synth, world):
    """ROWBAND partition emulated on one GPU: one handle per band, neighbour halo rows copied before every
    level (LocalExchanger -- the same schedule DistExchanger runs over NCCL)."""
    h, w, octs, S = 416, 300, 4, 3
    img = synth.noise(h, w)
    ref = O.conv_build(img, octs, S)
    bands = [pkg.band_rows(h, octs, world, r) for r in range(world)]
    hs = []
    for row0, rows in bands:
        ss = pkg.ScaleSpace(rows, w, octs, S, mode=pkg.MODE_CONV, band_row0=row0, full_height=h)
        ss.upload(np.ascontiguousarray(img[row0:row0 + rows]))
        hs.append(ss)
    with pytest.raises(pkg.SspyrError):
        hs[0].build()                                         # banded CONV must go level by level
    pkg.LocalExchanger(hs).build()
    for (row0, rows), ss in zip(bands, hs):
        gg, dd = ss.download_gauss(), ss.download_dog()
        want_g = [ref["gauss"][o][:, row0 >> o:(row0 >> o) + (rows >> o)] for o in range(octs)]
        want_d = [ref["dog"][o][:, row0 >> o:(row0 >> o) + (rows >> o)] for o in range(octs)]
        check(gg, want_g, 255.0, f"band@{row0} gauss")
        check(dd, want_d, 255.0, f"band@{row0} dog")
        ss.close()


def test_bands_match_the_unbanded_gpu_result_exactly(pkg, synth):
    """Same kernels, same summation order: a banded build must equal the single-handle build bit for bit."""
    h, w, octs, S = 256, 200, 3, 3
    img = synth.noise(h, w)
    with pkg.ScaleSpace(h, w, octs, S, mode=pkg.MODE_CONV) as ss:
        ss.upload(img)
        ss.build()
        whole = ss.download_gauss()
    hs = []
    for r in range(2):
        row0, rows = pkg.band_rows(h, octs, 2, r)
        b = pkg.ScaleSpace(rows, w, octs, S, mode=pkg.MODE_CONV, band_row0=row0, full_height=h)
        b.upload(np.ascontiguousarray(img[row0:row0 + rows]))
        hs.append((row0, rows, b))
    pkg.LocalExchanger([b for _, _, b in hs]).build()
    for row0, rows, b in hs:
        for o, g in enumerate(b.download_gauss()):
            np.testing.assert_array_equal(g, whole[o][:, row0 >> o:(row0 >> o) + (rows >> o)])
        b.close()


@pytest.mark.parametrize("h,w,octs,S", [(96, 128, 3, 3), (135, 241, 3, 3), (270, 480, 4, 2), (64, 1030, 2, 4), (17, 70, 1, 3)])
def test_extrema_flags_and_keypoints(pkg, O, synth, h, w, octs, S):
    """The tiled DoG extremum scan (csrc/extrema.cu): flag planes equal the oracle's scan of the SAME DoG planes exactly,
    and the compacted keypoint list holds exactly the flagged pixels -- position, octave, level and DoG value -- in
    both modes; tiles that straddle the plane edge, planes smaller than a tile, several frame slots in one launch."""
    thresh = 0.5
    for mode in (pkg.MODE_CONV, pkg.MODE_REF):
        frames = 2
        with pkg.ScaleSpace(h, w, octs, S, mode=mode, outputs=pkg.OUT_ALL | pkg.OUT_EXTREMA | pkg.OUT_KEYPOINTS,
                            extrema_thresh=thresh, frames=frames) as ss:
            for f in range(frames):
                ss.upload(synth.noise(h, w, frame=f), frame=f)
            ss.build_batch(0, frames)
            for f in range(frames):
                dd = ss.download_dog(frame=f)
                rec, found = ss.download_keypoints(frame=f)
                want_list = []
                for o in range(octs):
                    want = O.extrema_octave(dd[o], thresh)               # scan of the GPU's own DoG planes: exact
                    got = np.stack([ss.download(o, s, pkg.KIND_EXTREMA, frame=f) for s in range(S)])
                    np.testing.assert_array_equal(got, want)
                    for s, y, x in np.argwhere(want):
                        want_list.append((int(x), int(y), (o << 16) | (int(s) + 1), int(dd[o][s + 1][y, x].view(np.int32))))
                assert found == len(want_list) == len(rec)
                assert sorted(map(tuple, rec.tolist())) == sorted(want_list)
                if mode == pkg.MODE_CONV and h * w > 5000:
                    assert found > 0


def test_keypoint_capacity_and_errors(pkg, synth):
    h, w = 200, 300
    with pkg.ScaleSpace(h, w, 3, 3, mode=pkg.MODE_CONV, outputs=pkg.OUT_DOG | pkg.OUT_KEYPOINTS, max_keypoints=5) as ss:
        ss.upload(synth.noise(h, w))
        ss.build()
        rec, found = ss.download_keypoints(capacity=1000)
        assert found > 5 and len(rec) == 5                           # counted beyond the capacity, stored up to it
        assert ss.device_ptr(0, 0, pkg.KIND_KEYPOINTS) != 0
        ss.build()                                                   # the cursor restarts with every build
        assert ss.download_keypoints(capacity=10)[1] == found
    with pytest.raises(pkg.SspyrError):                              # needs DoG
        pkg.ScaleSpace(h, w, 3, 3, outputs=pkg.OUT_GAUSS | pkg.OUT_KEYPOINTS)
    with pytest.raises(pkg.SspyrError) as e:                         # band seams would be scanned as image borders
        pkg.ScaleSpace(64, w, 2, 3, outputs=pkg.OUT_ALL | pkg.OUT_KEYPOINTS, band_row0=0, full_height=128)
    assert e.value.code == pkg._lib.ERR_UNSUPPORTED
    with pkg.ScaleSpace(h, w, 3, 3) as ss:
        ss.upload(synth.noise(h, w))
        ss.build()
        with pytest.raises(pkg.SspyrError):
            ss.download_keypoints()                                  # not configured


@pytest.mark.parametrize("h,w,octs,S,rs", [(270, 480, 4, 3, 3.0), (135, 1241, 3, 3, 3.0), (600, 700, 3, 3, 4.0),
                                           (97, 513, 2, 2, 3.0), (64, 2100, 2, 3, 3.0)])
def test_marching_kernel_equals_tile_kernel(pkg, O, synth, h, w, octs, S, rs):
    """The column-strip marching kernel (large levels) does the same FMA chains in the same order as the
    shared-memory tile kernel: bit-identical results, and both within tolerance of the specification."""
    img = synth.noise(h, w)
    out = {}
    for march in (0, 1):
        with pkg.ScaleSpace(h, w, octs, S, mode=pkg.MODE_CONV, radius_sigmas=rs, frames=2) as ss:
            ss.set_tuning(conv_cascade=0, conv_march=march)
            ss.upload(img, frame=0)
            ss.upload(synth.noise(h, w, frame=5), frame=1)
            ss.build_batch(0, 2)
            out[march] = (ss.download_gauss(0), ss.download_dog(0), ss.download_gauss(1))
    for a, b in zip(out[0], out[1]):
        for o in range(octs):
            np.testing.assert_array_equal(a[o], b[o])
    ref = O.conv_build(img, octs, S, radius_sigmas=rs)
    check(out[1][0], ref["gauss"], 255.0, "march gauss")
    check(out[1][1], ref["dog"], 255.0, "march dog")


def test_marching_kernel_on_row_bands(pkg, O, synth):
    h, w, octs, S = 512, 640, 3, 3
    img = synth.noise(h, w)
    ref = O.conv_build(img, octs, S)
    hs = []
    for r in range(2):
        row0, rows = pkg.band_rows(h, octs, 2, r)
        b = pkg.ScaleSpace(rows, w, octs, S, mode=pkg.MODE_CONV, band_row0=row0, full_height=h)
        b.set_tuning(conv_march=1)
        b.upload(np.ascontiguousarray(img[row0:row0 + rows]))
        hs.append((row0, rows, b))
    pkg.LocalExchanger([b for _, _, b in hs]).build()
    for row0, rows, b in hs:
        want = [ref["gauss"][o][:, row0 >> o:(row0 >> o) + (rows >> o)] for o in range(octs)]
        check(b.download_gauss(), want, 255.0, f"band@{row0}")
        b.close()


@pytest.mark.parametrize("world", [2, 3])
def test_row_bands_reading_neighbour_planes_in_place(pkg, O, synth, world):
    """Peer-memory halos: the blur kernels read the neighbour band's rows straight out of its planes (what runs
    over NVLink between GPUs); here all bands live on one GPU and share one stream.  Must equal the exchanged
    build bit for bit, and the specification within tolerance."""
    h, w, octs, S = 416, 300, 4, 3
    img = synth.noise(h, w)
    ref = O.conv_build(img, octs, S)
    bands = [pkg.band_rows(h, octs, world, r) for r in range(world)]

    def make():
        hs = []
        for row0, rows in bands:
            ss = pkg.ScaleSpace(rows, w, octs, S, mode=pkg.MODE_CONV, band_row0=row0, full_height=h)
            ss.upload(np.ascontiguousarray(img[row0:row0 + rows]))
            hs.append(ss)
        return hs

    a, b = make(), make()
    pkg.LocalExchanger(a).build()
    link = pkg.LocalPeerLink(b)
    link.build()
    link.build()                                              # counters keep counting across frames
    for (row0, rows), x, y in zip(bands, a, b):
        y.sync()
        gx, gy = x.download_gauss(), y.download_gauss()
        for o in range(octs):
            np.testing.assert_array_equal(gx[o], gy[o])
        check(y.download_dog(), [ref["dog"][o][:, row0 >> o:(row0 >> o) + (rows >> o)] for o in range(octs)], 255.0, "dog")
        x.close()
        y.close()


def _random_conv_geometries(seed, n):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        h, w = int(rng.integers(1, 260)), int(rng.integers(1, 600))
        all_o = int(min(h, w)).bit_length()
        out.append((h, w, int(rng.integers(1, min(all_o, 5) + 1)), int(rng.integers(2, 5)), float(rng.choice([3.0, 4.0])),
                    int(rng.integers(0, 2))))
    return out


@pytest.mark.parametrize("h,w,octs,S,rs,march", _random_conv_geometries(77, 20) + [(1, 1, 1, 3, 3.0, 1), (2, 300, 2, 3, 3.0, 1),
                                                                                     (300, 2, 2, 3, 3.0, 0), (33, 129, 3, 3, 4.0, 1)])
@pytest.mark.parametrize("cascade", [0, 1])
def test_random_geometries_conv_within_tolerance(pkg, O, synth, h, w, octs, S, rs, march, cascade):
    """Seeded sweep of odd shapes for all CONV kernels (planes smaller than a tile, than the blur radius, ...)."""
    if cascade and not march:
        pytest.skip("the cascade always marches")
    img = synth.noise(h, w, frame=h * 1000 + w)
    ref = O.conv_build(img, octs, S, radius_sigmas=rs)
    with pkg.ScaleSpace(h, w, octs, S, mode=pkg.MODE_CONV, radius_sigmas=rs) as ss:
        ss.set_tuning(conv_cascade=2 * cascade, conv_march=march)
        ss.upload(img)
        ss.build()
        check(ss.download_gauss(), ref["gauss"], 255.0, "gauss")
        check(ss.download_dog(), ref["dog"], 255.0, "dog")


def _all_planes(ss, frame=0):
    return ss.download_gauss(frame) + ss.download_dog(frame)


@pytest.mark.parametrize("h,w,octs,S,rs,frames", [(270, 480, 4, 3, 3.0, 1), (333, 1241, 3, 3, 3.0, 2), (600, 700, 3, 2, 4.0, 1),
                                                  (97, 513, 2, 3, 3.0, 3), (2160, 3840, 3, 3, 3.0, 1)])
def test_level_chaining_is_bit_identical(pkg, synth, h, w, octs, S, rs, frames):
    """Level chaining (a level's CTA waits for the segments of the previous level it reads, not for the whole
    previous grid; conv_march.cuh) is a schedule, not arithmetic: forced on (conv_chain=2, also for grids of less
    than a wave), automatic (1) and off (0) give the same bits, build after build, with new pixels every build
    (a CTA that ran ahead of its producer would blur the previous build's rows)."""
    imgs = [[synth.noise(h, w, frame=10 * b + f) for f in range(frames)] for b in range(3)]
    out = {}
    for chain in (0, 1, 2):
        with pkg.ScaleSpace(h, w, octs, S, mode=pkg.MODE_CONV, radius_sigmas=rs, frames=frames) as ss:
            ss.set_tuning(conv_cascade=0, conv_chain=chain)
            got = []
            for b in range(3):
                for f in range(frames):
                    ss.upload(imgs[b][f], frame=f)
                ss.build_batch(0, frames)
                ss.sync()
                got.append([_all_planes(ss, f) for f in range(frames)])
            out[chain] = got
    for chain in (1, 2):
        for b in range(3):
            for f in range(frames):
                for a, c in zip(out[0][b][f], out[chain][b][f]):
                    np.testing.assert_array_equal(a, c)


def test_level_chaining_survives_retuning_and_graph_replay(pkg, O, synth):
    """The per-segment build counters restart from zero when the segmentation may change (sspyr_set_tuning), and
    a captured launch sequence (CUDA graph, from the 2nd build of a slot on) replays them correctly."""
    h, w, octs = 500, 900, 3
    with pkg.ScaleSpace(h, w, octs, 3, mode=pkg.MODE_CONV, frames=2) as ss:
        ss.set_tuning(conv_cascade=0, conv_chain=2)
        for b in range(5):                                   # eager, capture, replays; alternating slots
            img = synth.noise(h, w, frame=b)
            ss.upload(img, frame=b % 2)
            ss.build(b % 2)
            if b in (0, 3, 4):
                check(ss.download_gauss(b % 2), O.conv_build(img, octs, 3)["gauss"], 255.0, f"build {b}")
        for waves, seg in ((2, 64), (5, 32), (0, 0)):        # new segment grids on the same handle
            ss.set_tuning(conv_waves=waves)
            ss.set_tuning(conv_seg_min=seg)
            for b in range(3):
                img = synth.noise(h, w, frame=100 + b)
                ss.upload(img, frame=0)
                ss.build(0)
            ref = O.conv_build(img, octs, 3)
            check(ss.download_gauss(0), ref["gauss"], 255.0, f"waves {waves}")
            check(ss.download_dog(0), ref["dog"], 255.0, f"waves {waves}")


def test_frame_lanes_keep_stream_order(pkg, synth):
    """CONV builds of different frame slots run on separate stream sets (lanes) and overlap; uploads, downloads,
    batch calls that span lanes and rebuilds of a slot must still behave as if everything ran in stream order:
    same bits as the one-build-at-a-time handle (conv_lanes=1) for a mixed sequence of calls."""
    h, w, octs, n = 300, 520, 4, 6

    def run(lanes):
        res = []
        with pkg.ScaleSpace(h, w, octs, 3, mode=pkg.MODE_CONV, frames=n, outputs=pkg.OUT_ALL | pkg.OUT_EXTREMA) as ss:
            ss.set_tuning(conv_cascade=0, conv_lanes=lanes)
            for rnd in range(3):
                for f in range(n):                       # upload -> build per slot, nothing waits in between
                    ss.upload(synth.noise(h, w, frame=100 * rnd + f), frame=f)
                    ss.build(f)
                ss.build_batch(1, 3)                     # slots 1..3 again, as one batch on one lane
                ss.upload(synth.noise(h, w, frame=100 * rnd + 50), frame=2)
                ss.build(2)                              # ... and slot 2 once more with new pixels, on its own lane
                ss.build(5)
                for f in (0, 2, 3, 5):
                    res.append(ss.download_gauss(f) + ss.download_dog(f))
                res.append([ss.download_inplace(1)])
        return res

    a, b = run(1), run(4)
    assert len(a) == len(b)
    for x, y in zip(a, b):
        for p, q in zip(x, y):
            if isinstance(p, list):
                for pp, qq in zip(p, q):
                    np.testing.assert_array_equal(pp, qq)
            else:
                np.testing.assert_array_equal(p, q)


@pytest.mark.parametrize("world,slots", [(2, 3), (3, 2)])
def test_peer_bands_with_several_builds_in_flight(pkg, synth, world, slots):
    """Row bands reading their neighbours' planes in place, whole-pyramid builds (one library call per band and
    slot, as bench.py runs them on N GPUs), several frame slots in flight on frame lanes: every slot has its own
    progress counters.  All bands live on one GPU here, each with its own stream; must equal the unbanded build
    bit for bit, slot by slot, build after build."""
    import torch
    h, w, octs, S = 384, 520, 3, 3
    bands = [pkg.band_rows(h, octs, world, r) for r in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    hs = []
    for (row0, rows), st in zip(bands, streams):
        b = pkg.ScaleSpace(rows, w, octs, S, mode=pkg.MODE_CONV, band_row0=row0, full_height=h, frames=slots)
        b.set_stream(st.cuda_stream)
        hs.append(b)
    for i, b in enumerate(hs):
        if i > 0:
            b.peer_attach_local(0, hs[i - 1])
        if i + 1 < world:
            b.peer_attach_local(1, hs[i + 1])
    with pkg.ScaleSpace(h, w, octs, S, mode=pkg.MODE_CONV) as whole:
        for rnd in range(3):
            imgs = [synth.noise(h, w, frame=10 * rnd + f) for f in range(slots)]
            for (row0, rows), b in zip(bands, hs):
                for f in range(slots):
                    b.upload(np.ascontiguousarray(imgs[f][row0:row0 + rows]), frame=f)
                b.sync()
            for f in range(slots):                       # nothing waits between these calls
                for b in hs:
                    b.build(f)
            for b in hs:
                b.sync()
            for f in range(slots):
                whole.upload(imgs[f])
                whole.build()
                want = whole.download_gauss() + whole.download_dog()
                for (row0, rows), b in zip(bands, hs):
                    got = b.download_gauss(f) + b.download_dog(f)
                    for o in range(octs):
                        np.testing.assert_array_equal(got[o], want[o][:, row0 >> o:(row0 >> o) + (rows >> o)])
                        np.testing.assert_array_equal(got[octs + o], want[octs + o][:, row0 >> o:(row0 >> o) + (rows >> o)])
    for b in hs:
        b.close()


# ---- the cascade: one launch per build, levels pipelined through L2 (conv_cascade.cuh) -----------------------------
@pytest.mark.parametrize("h,w,octs,S,rs,frames,seg", [(270, 480, 4, 3, 3.0, 1, 0), (333, 1241, 3, 3, 3.0, 2, 32), (600, 700, 3, 2, 4.0, 1, 64),
                                                      (97, 513, 2, 3, 3.0, 3, 0), (1080, 1920, 5, 3, 3.0, 2, 0), (2160, 3840, 5, 3, 3.0, 1, 0),
                                                      (1500, 300, 6, 4, 3.0, 1, 96), (40, 3000, 3, 3, 3.0, 2, 0)])
def test_cascade_equals_the_per_level_path_bit_for_bit(pkg, synth, h, w, octs, S, rs, frames, seg):
    """The cascade is a schedule, not arithmetic: every plane of every frame equals the one-launch-per-level build,
    build after build with new pixels every build (an item that ran ahead of its producer, or overwrote rows the
    previous build was still reading, would show up as stale or torn rows), for several segment heights."""
    imgs = [[synth.noise(h, w, frame=10 * b + f) for f in range(frames)] for b in range(3)]
    out = {}
    for cascade in (0, 1):
        with pkg.ScaleSpace(h, w, octs, S, mode=pkg.MODE_CONV, radius_sigmas=rs, frames=frames) as ss:
            ss.set_tuning(conv_cascade=2 * cascade, conv_casc_seg=seg)
            got = []
            for b in range(3):
                for f in range(frames):
                    ss.upload(imgs[b][f], frame=f)
                ss.build_batch(0, frames)
                if cascade and rs * 3.1 <= 12:                 # (radius_sigmas 4 gives radii up to 13: those handles stay on the
                    assert ss.last_launches() == 1             #  per-level path, which must of course agree with itself)
                ss.sync()
                got.append([_all_planes(ss, f) for f in range(frames)])
            out[cascade] = got
    for b in range(3):
        for f in range(frames):
            for a, c in zip(out[0][b][f], out[1][b][f]):
                np.testing.assert_array_equal(a, c)


def test_cascade_builds_overlap_safely(pkg, O, synth):
    """Back-to-back cascade launches with nothing in between overlap under programmatic dependent launch; the slot
    epoch keeps two builds of the SAME slot apart and the item counters order everything else.  Many builds over a
    ring of slots (and repeatedly over one slot) without a sync, then every slot must hold its last frame's pyramid."""
    h, w, octs, n = 540, 960, 4, 3
    frames = [synth.noise(h, w, frame=f) for f in range(n)]
    want = [O.conv_build(fr, octs, 3) for fr in frames]
    with pkg.ScaleSpace(h, w, octs, 3, mode=pkg.MODE_CONV, frames=n) as ss:
        ss.set_tuning(conv_cascade=2)
        for f in range(n):
            ss.upload(frames[f], frame=f)
        for rnd in range(6):
            for f in range(n):
                ss.build(f)
            ss.build(1)
            ss.build(1)                                  # the same slot twice in a row
        ss.sync()
        for f in range(n):
            check(ss.download_gauss(f), want[f]["gauss"], 255.0, f"slot {f} gauss")
            check(ss.download_dog(f), want[f]["dog"], 255.0, f"slot {f} dog")
        ss.set_tuning(conv_casc_seg=32)                  # retuning restarts the counters
        ss.build_batch(0, n)
        ss.build_batch(0, n)
        ss.sync()
        check(ss.download_gauss(2), want[2]["gauss"], 255.0, "after retuning")


def test_cascade_pixel_types_and_device_input(pkg, O, synth):
    import torch
    h, w, octs = 300, 700, 4
    img = synth.noise(h, w)
    ref = O.conv_build(img, octs, 3)
    with pkg.ScaleSpace(h, w, octs, 3, mode=pkg.MODE_CONV, pixel_type=pkg.PIXEL_U8) as ss:
        ss.set_tuning(conv_cascade=2)
        ss.upload(img.astype(np.uint8))
        ss.build()
        assert ss.last_launches() == 1
        check(ss.download_gauss(), ref["gauss"], 255.0, "u8")
    t = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()
    with pkg.ScaleSpace(h, w, octs, 3, mode=pkg.MODE_CONV, outputs=pkg.OUT_INPLACE) as ss:   # DoG + every Gaussian (CONV keeps them)
        ss.set_stream(st.cuda_stream)
        ss.set_tuning(conv_cascade=2)
        ss.set_input_device(t.data_ptr(), t.stride(0) * 4)
        with torch.cuda.stream(st):
            for rnd in range(3):                         # a producer kernel on the same stream right before each build
                t.copy_(torch.from_numpy(synth.noise(h, w, frame=rnd)).cuda(), non_blocking=True)
                t.add_(0)
                ss.build()
        ss.sync()
        check(ss.download_gauss(), O.conv_build(synth.noise(h, w, frame=2), octs, 3)["gauss"], 255.0, "device input")


# ---- independent fixture: the CUDA path against the committed scipy golden vectors (not only our own C oracle) --------
@pytest.mark.parametrize("cascade", [1, 0])
@pytest.mark.parametrize("name", ["noise_96x128", "noise_135x241", "noise_80x72_S2", "unit_64x96", "pattern_140x420"])
def test_cuda_against_the_committed_scipy_fixture(pkg, name, cascade):
    """tests/golden/conv_scipy.npz is generated by oracle/make_golden_conv.py from a statement of the CONV
    specification that shares no code with the C oracle or the kernels (scipy.ndimage.correlate1d, double
    accumulation).  Both CUDA schedules must meet the north_star's tolerance against it: 1e-4 of full scale per level."""
    import importlib.util
    import os
    from conftest import GOLDEN, ROOT
    gold = np.load(os.path.join(GOLDEN, "conv_scipy.npz"))
    spec = importlib.util.spec_from_file_location("make_golden_conv", os.path.join(ROOT, "oracle", "make_golden_conv.py"))
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except ImportError:
        pytest.skip("scipy not importable (only its CASES table and pixels() are used here)")
    h, w, octs, S, s0, sin, rs, kind = mod.CASES[name]
    img = mod.pixels(name, h, w, kind)
    pix = pkg.PIXEL_F32 if kind == "f32" else pkg.PIXEL_I32
    with pkg.ScaleSpace(h, w, octs, S, sigma0=s0, sigma_in=sin, radius_sigmas=rs, mode=pkg.MODE_CONV, pixel_type=pix) as ss:
        ss.set_tuning(conv_cascade=2 * cascade)
        ss.upload(img)
        ss.build()
        gg, dd = ss.download_gauss(), ss.download_dog()
    scale = 1.0 if kind == "f32" else 255.0
    want = [gold[f"{name}_g{o}"] for o in range(octs)]
    check(gg, want, scale, f"{name} gauss vs scipy")
    check(dd, [g[:-1].astype(np.float64) - g[1:] for g in want], scale, f"{name} dog vs scipy")
