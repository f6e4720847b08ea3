"""scripts/sweep_ref.py -- in-process tuning sweep of the fused REF kernel (one GPU).
Prints one line per (workload, outputs, rows_per_thread, block, bx, pdl): us/frame, GB/s, fraction of peak."""
from __future__ import annotations

import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
import torch  # noqa: E402

pkg = entry.load_package()
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

SHAPES = {"c1": (512, 512, 4), "c2": (1080, 1920, 5), "c3": (2160, 3840, 5), "c4": (4320, 7680, 5)}


def run(name, outputs, tunings, steps):
    h, w, octs = SHAPES[name]
    probe = pkg.ScaleSpace(h, w, octs, 3, outputs=outputs)
    fb = probe.algorithmic_bytes()
    probe.close()
    slots = max(2, min(8, -(-(4 * (126 << 20)) // fb)))
    ss = pkg.ScaleSpace(h, w, octs, 3, outputs=outputs, frames=slots)
    st = torch.cuda.current_stream()
    ss.set_stream(st.cuda_stream)
    for s in range(slots):
        ss.upload(pkg.synth.noise(h, w, frame=s), frame=s)
    ss.sync()
    rows = []
    for rpt, block, bx, pdl, occ in tunings:
        ss.set_tuning(rows_per_thread=rpt, block=block, bx=bx, pdl=pdl, occ=occ)
        for i in range(20):
            ss.build(i % slots)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(steps):
            ss.build(i % slots)
        e1.record(st)
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / steps * 1e3
        gbs = fb / us / 1e3
        rows.append((us, rpt, block, bx, pdl, occ, gbs))
        print(f"{name} out={outputs} rpt={rpt} block={block} bx={bx} pdl={pdl} occ={occ}: {us:8.2f} us/frame  {gbs:7.1f} GB/s  "
              f"frac={gbs / PEAK:.3f}  {h * w / us:9.1f} Mpix/s", flush=True)
    ss.close()
    best = min(rows)
    print(f"BEST {name} out={outputs}: rpt={best[1]} block={best[2]} bx={best[3]} pdl={best[4]} occ={best[5]} -> {best[0]:.2f} us, "
          f"{best[6]:.1f} GB/s, frac={best[6] / PEAK:.3f}", flush=True)


if __name__ == "__main__":
    names = sys.argv[1:] or ["c2", "c3"]
    tunings = list(itertools.product((1, 2), (96, 128), (0, 32), (1,), (0, 2, 3, 4, 5, 6, 8, 12)))
    for name in names:
        steps = {"c1": 2000, "c2": 1000, "c3": 300, "c4": 100}[name]
        run(name, pkg.OUT_ALL, tunings, steps)
    run("c2", pkg.OUT_INPLACE, [(1, 128, 0, 1, 0), (2, 128, 0, 1, 0), (1, 128, 0, 1, 4), (2, 128, 0, 1, 4)], 1000)
