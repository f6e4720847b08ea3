"""scripts/sanitize_case.py -- a small run of every kernel family (REF fused, CONV tile, CONV strip, extrema,
peer-pointer halos) for compute-sanitizer (one tool per gpurun call).  Checks results against the oracle too."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

pkg, O = entry.load_package(), entry.load_oracle()
h, w, octs, S = 150, 277, 4, 3
img = pkg.synth.noise(h, w)
ref = O.ref_build(img, octaves=octs, S=S)
with pkg.ScaleSpace(h, w, octs, S, frames=2, outputs=pkg.OUT_ALL | pkg.OUT_EXTREMA, extrema_thresh=0.5) as ss:
    ss.upload(img, 0)
    ss.upload(img, 1)
    ss.build_batch(0, 2)
    for o, a in enumerate(ss.download_inplace(1)):
        assert np.array_equal(a.view(np.uint32), ref["inplace"][o].view(np.uint32))
cref = O.conv_build(img, octs, S)
for march in (0, 1):
    with pkg.ScaleSpace(h, w, octs, S, mode=pkg.MODE_CONV, outputs=pkg.OUT_ALL | pkg.OUT_EXTREMA, extrema_thresh=0.5) as ss:
        ss.set_tuning(conv_march=march, conv_graph=0)
        ss.upload(img)
        ss.build()
        for o, a in enumerate(ss.download_gauss()):
            assert np.max(np.abs(a - cref["gauss"][o])) < 0.0255
H2 = 192
img2 = pkg.synth.noise(H2, w)
cref2 = O.conv_build(img2, 3, S)
hs = []
for r in range(2):
    row0, rows = pkg.band_rows(H2, 3, 2, r)
    b = pkg.ScaleSpace(rows, w, 3, S, mode=pkg.MODE_CONV, band_row0=row0, full_height=H2)
    b.upload(np.ascontiguousarray(img2[row0:row0 + rows]))
    hs.append((row0, rows, b))
pkg.LocalPeerLink([b for _, _, b in hs]).build()
for row0, rows, b in hs:
    b.sync()
    for o, a in enumerate(b.download_gauss()):
        assert np.max(np.abs(a - cref2["gauss"][o][:, row0 >> o:(row0 >> o) + (rows >> o)])) < 0.0255
    b.close()
print("sanitize case ok")
