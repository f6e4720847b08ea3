"""scripts/check_conv_bands_nccl.py -- run under torchrun on N GPUs: CONV-mode row bands with the per-level NCCL
halo exchange (exchange.DistExchanger) must reproduce the specification (oracle) on every rank's band, and the
REF-mode bands (no exchange) must be bit-exact.  Prints one PASS/FAIL line per rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
        scripts/check_conv_bands_nccl.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def main() -> int:
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pkg, O = entry.load_package(), entry.load_oracle()
    h, w, octs, S = 1088, 1500, 4, 3
    img = pkg.synth.noise(h, w, frame=11)
    row0, rows = pkg.band_rows(h, octs, world, rank)
    ok = True
    # CONV: halo exchange over NCCL P2P before every level
    ref = O.conv_build(img, octs, S)
    ss = pkg.ScaleSpace(rows, w, octs, S, mode=pkg.MODE_CONV, device=local, band_row0=row0, full_height=h)
    ss.set_stream(torch.cuda.current_stream().cuda_stream)
    ss.upload(np.ascontiguousarray(img[row0:row0 + rows]))
    ex = pkg.DistExchanger(ss, rank, world, dev)
    ex.build()
    ex.build()                                   # twice: the schedule must be re-entrant
    worst = 0.0
    for kind, got in (("gauss", ss.download_gauss()), ("dog", ss.download_dog())):
        for o in range(octs):
            want = ref[kind][o][:, row0 >> o:(row0 >> o) + (rows >> o)]
            err = float(np.max(np.abs(got[o].astype(np.float64) - want)))
            worst = max(worst, err)
            ok &= err <= 1e-4 * 255
    ss.close()
    # CONV again with peer-memory halos: neighbours' planes mapped through CUDA IPC, read inside the blur kernel
    ps = pkg.ScaleSpace(rows, w, octs, S, mode=pkg.MODE_CONV, device=local, band_row0=row0, full_height=h)
    ps.set_stream(torch.cuda.current_stream().cuda_stream)
    ps.upload(np.ascontiguousarray(img[row0:row0 + rows]))
    ps.sync()
    px = pkg.PeerExchanger(ps, rank, world)
    px.build()
    px.build_stepwise()
    px.build()
    ps.sync()
    worst_peer = 0.0
    for kind, got in (("gauss", ps.download_gauss()), ("dog", ps.download_dog())):
        for o in range(octs):
            want = ref[kind][o][:, row0 >> o:(row0 >> o) + (rows >> o)]
            err = float(np.max(np.abs(got[o].astype(np.float64) - want)))
            worst_peer = max(worst_peer, err)
            ok &= err <= 1e-4 * 255
    dist.barrier()
    ps.close()
    worst = max(worst, worst_peer)
    # REF: no exchange at all, bit-exact
    rref = O.ref_build(img, octaves=octs, S=S, want=("inplace",))["inplace"]
    with pkg.ScaleSpace(rows, w, octs, S, device=local, band_row0=row0, full_height=h) as rs:
        rs.upload(np.ascontiguousarray(img[row0:row0 + rows]))
        rs.build()
        for o, a in enumerate(rs.download_inplace()):
            want = rref[o][:, row0 >> o:(row0 >> o) + (rows >> o)]
            ok &= bool(np.array_equal(a.view(np.uint32), np.ascontiguousarray(want).view(np.uint32)))
    # C4 band geometry (7680x4320, 5 octaves: 540-row bands at N = 8), the way bench.py runs it: whole-pyramid builds over
    # peer memory, three frame slots in flight, several rounds back to back (from the third build of a slot on the launch
    # sequence replays as a CUDA graph).  Every band must equal the UNBANDED build of the same pixels on its own GPU bit for
    # bit, and no wait may have timed out (sspyr_sync reports that).
    # Both seam protocols: whole-level progress flags (default) and levels chained across the seam through the neighbours'
    # segment counters (conv_band_chain = 1).
    if "--c4" in sys.argv:
        H, W, O5, slots = 4320, 7680, 5, 3
        r0, nr = pkg.band_rows(H, O5, world, rank)
        whole = pkg.ScaleSpace(H, W, O5, S, mode=pkg.MODE_CONV, device=local)
        want = {}
        for chain in (0, 1):
            band = pkg.ScaleSpace(nr, W, O5, S, mode=pkg.MODE_CONV, device=local, band_row0=r0, full_height=H, frames=slots)
            band.set_stream(torch.cuda.current_stream().cuda_stream)
            band.set_tuning(conv_band_chain=chain)
            link = pkg.PeerExchanger(band, rank, world)
            try:
                for rnd in range(4):
                    imgs = [pkg.synth.noise(H, W, frame=100 * rnd + f) for f in range(slots)] if rnd in (0, 3) else imgs
                    for f in range(slots):
                        band.upload(np.ascontiguousarray(imgs[f][r0:r0 + nr]), frame=f)
                    band.sync()
                    dist.barrier()                           # every band's pixels are in place before anyone reads a halo
                    for f in range(slots):                   # nothing waits between these calls
                        link.build(f)
                    band.sync()                              # raises if a neighbour / level wait timed out
                    dist.barrier()
                    if rnd in (0, 3):
                        for f in (0, slots - 1):
                            key = (rnd, f)
                            if key not in want:
                                whole.upload(imgs[f])
                                whole.build()
                                want[key] = (whole.download_gauss(), whole.download_dog())
                            wg, wd = want[key]
                            bg, bd = band.download_gauss(f), band.download_dog(f)
                            for o in range(O5):
                                lo, n = r0 >> o, nr >> o
                                ok &= bool(np.array_equal(bg[o], wg[o][:, lo:lo + n])) and bool(np.array_equal(bd[o], wd[o][:, lo:lo + n]))
                    dist.barrier()
            finally:
                band.close()
        whole.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print(f"rank {rank}/{world} band rows [{row0},{row0 + rows}) CONV max|err|={worst:.3g} "
          f"{'PASS' if ok else 'FAIL'} (all ranks: {'PASS' if int(flag.item()) else 'FAIL'})", flush=True)
    dist.destroy_process_group()
    return 0 if int(flag.item()) else 1


if __name__ == "__main__":
    sys.exit(main())
