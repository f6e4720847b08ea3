"""GPU suite, BASELINE.json's configurations at FULL size: C4 (7680x4320, 5 octaves), C5
(16384x16384, 8 octaves: plane offsets close to 2^32 floats, the last octaves 256/128 pixels wide) and C3 as specified
(256 frames of 3840x2160 through the frame-slot ring).  REF mode, bit-exact against the oracle's closed-form port
(pinned to the reference header by tests/test_oracle.py), through the C ABI."""
from __future__ import annotations

import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu


def _band_view(pkg, ss, octave, level, kind, r0, r1):
    """Rows [r0, r1) of one device plane -> host, without downloading the whole plane."""
    import torch
    from sift_parallel_optimization_b200.exchange import device_view
    rows, cols, pitch = ss.level_dims(octave)
    ptr = ss.device_ptr(octave, level, kind)
    t = device_view(ptr + r0 * pitch * 4, (r1 - r0) * pitch * 4, torch.device("cuda", torch.cuda.current_device()))
    return t.view(torch.float32).view(r1 - r0, pitch)[:, :cols].cpu().numpy()


def test_c4_full_size_every_plane_bit_exact(pkg, O, synth):
    """7680x4320, 5 octaves x 6 levels: all 11 planes of every octave against orc_ref_build."""
    h, w, octs = 4320, 7680, 5
    img = synth.noise(h, w, frame=4)
    ref = O.ref_build(img, octaves=octs, S=3, want=("gauss", "dog"))
    with pkg.ScaleSpace(h, w, octs, 3) as ss:
        ss.upload(img)
        ss.build()
        assert ss.last_launches() == 1
        gg = ss.download_gauss()
        for o in range(octs):
            assert ss.level_dims(o)[:2] == (h >> o, w >> o)
            assert bits_equal(gg[o], ref["gauss"][o]), f"gauss octave {o}"
        del gg
        for o in range(octs):
            for s in range(5):
                assert bits_equal(ss.download(o, s, pkg.KIND_DOG), ref["dog"][o][s]), f"DoG octave {o} level {s}"


def test_c5_full_size_band_by_band_bit_exact(pkg, O, synth):
    """16384x16384, 8 octaves x 6 levels in ONE handle (16.8 GB of planes; octave-0 planes are 2^28 floats and the
    last plane of octave 0 starts 2.7e9 floats into the slot).  Checked band by band -- 2048 full-image rows at a
    time, every plane of every octave -- against orc_ref_build(row0, full_height) on the same pixels."""
    n, octs, S, band = 16384, 8, 3, 2048
    with pkg.ScaleSpace(n, n, octs, S) as ss:
        img = synth.noise(n, n, frame=5)
        ss.upload(img)
        ss.build()
        ss.sync()
        assert ss.octaves == octs and ss.level_dims(7)[:2] == (128, 128)
        assert ss.algorithmic_bytes() == 4 * n * n + 4 * 11 * sum((n >> o) ** 2 for o in range(octs))
        for row0 in range(0, n, band):
            ref = O.ref_build(np.ascontiguousarray(img[row0:row0 + band]), octaves=octs, S=S, row0=row0, full_h=n,
                              want=("gauss", "dog"))
            for o in range(octs):
                r0, r1 = row0 >> o, (row0 + band) >> o
                for s in range(S + 3):
                    assert bits_equal(_band_view(pkg, ss, o, s, pkg.KIND_GAUSS, r0, r1), ref["gauss"][o][s]), \
                        f"rows {row0}.. gauss octave {o} level {s}"
                for s in range(S + 2):
                    assert bits_equal(_band_view(pkg, ss, o, s, pkg.KIND_DOG, r0, r1), ref["dog"][o][s]), \
                        f"rows {row0}.. DoG octave {o} level {s}"
            del ref


def test_c3_as_specified_256_frames_through_the_ring(pkg, O, synth):
    """C3: 256 frames of 3840x2160 x 5 octaves, streamed through a ring of frame slots in batched builds exactly as
    bench.py runs it.  Every frame's in-place result is reduced to a 64-bit checksum on the device and must equal the
    checksum of the same frame built alone on a second, single-slot handle (catches slot / batch / ring mix-ups and
    overlap races for all 256 frames); 12 frames -- first, last and the ring's wrap points -- are also compared bit
    for bit with the oracle port."""
    import torch
    from sift_parallel_optimization_b200.exchange import device_view
    h, w, octs, S, frames, slots = 2160, 3840, 5, 3, 256, 8
    dev = torch.device("cuda", torch.cuda.current_device())
    st = torch.cuda.current_stream()
    ring = pkg.ScaleSpace(h, w, octs, S, outputs=pkg.OUT_INPLACE, frames=slots)
    solo = pkg.ScaleSpace(h, w, octs, S, outputs=pkg.OUT_INPLACE, frames=1)
    ring.set_stream(st.cuda_stream)
    solo.set_stream(st.cuda_stream)
    try:
        dims = [ring.level_dims(o) for o in range(octs)]

        def checksum(ss, slot):                   # octave-weighted sum of the raw bits of the in-place tail of every octave
            acc = torch.zeros((), dtype=torch.int64, device=dev)
            for o, (r, c, p) in enumerate(dims):
                ptr = ss.device_ptr(o, 0, pkg.KIND_INPLACE, frame=slot)
                t = device_view(ptr, (S + 3) * r * p * 4, dev).view(torch.int32).view(S + 3, r, p)[:, :, :c]
                acc = acc + t.to(torch.int64).sum() * (o + 1)
            return acc

        picked = {0, 1, 7, 8, 9, 63, 64, 127, 128, 200, 254, 255}
        sums_ring, sums_solo = [], []
        noise = synth.noise(h, w, frame=3000)     # frame f = (noise + 7 f) mod 256: 256 different frames, cheap to make
        for base in range(0, frames, slots):
            imgs = [((noise + 7 * (base + k)) & 255).astype(np.int32) for k in range(slots)]
            for k in range(slots):
                ring.upload(imgs[k], frame=k)
            ring.build_batch(0, slots)            # one fused launch for the batch
            assert ring.last_launches() <= 2
            for k in range(slots):
                sums_ring.append(checksum(ring, k))
                solo.upload(imgs[k])
                solo.build()
                sums_solo.append(checksum(solo, 0))
                if base + k in picked:
                    ref = O.ref_build(imgs[k], octaves=octs, S=S, want=("inplace",))["inplace"]
                    for o, a in enumerate(ring.download_inplace(frame=k)):
                        assert bits_equal(a, ref[o]), f"frame {base + k} octave {o}"
            ring.sync()
        a = torch.stack(sums_ring).cpu().numpy()
        b = torch.stack(sums_solo).cpu().numpy()
        assert a.shape == (frames,) and np.array_equal(a, b), f"frames that differ: {np.nonzero(a != b)[0][:10]}"
        assert len(set(a.tolist())) == frames     # 256 different frames gave 256 different pyramids
    finally:
        ring.close()
        solo.close()
