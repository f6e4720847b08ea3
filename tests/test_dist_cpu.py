"""CPU suite, part 3: the N>1 host logic on world_size-2 gloo (no GPU): bench.py's Dist plumbing (barrier, max/sum
over ranks) and the two partitions -- frame shards (BATCH) and row bands (ROWBAND) -- each rank computing only
its share (with the oracle standing in for the kernels) and the union reproducing the unpartitioned result bit
for bit.  REF mode needs no data-path collective, so the only traffic here is the final check itself."""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def _worker(rank: int, world: int, port: int, tmp: str) -> None:
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry
    import bench
    pkg, O = entry.load_package(), entry.load_oracle()

    d = bench.Dist(world, rank, backend="gloo")
    assert bench.dist_env() == (rank, world, rank)
    d.barrier()
    assert d.max(float(rank + 1)) == float(world)                 # timings are reduced as max over ranks
    assert d.sum(1.5) == 1.5 * world

    # ROWBAND: c4 geometry / 8 -> every rank builds its band from its slice of the frame only
    H, W, octs, S = 540, 960, 5, 3
    rows, row0, full_h, width, o2, frames = bench.rank_geometry(pkg, "c4", world, rank)
    assert (full_h, width, o2, frames) == (4320, 7680, 5, 1) and row0 % 16 == 0
    r0, nrows = pkg.band_rows(H, octs, world, rank)
    band_img = pkg.synth.noise(nrows, W, frame=3, row0=r0)          # generated on the rank, not scattered
    band = O.ref_build(band_img, octaves=octs, S=S, row0=r0, full_h=H, want=("inplace",))["inplace"]
    # BATCH: c3 geometry / 8, 5 frames sharded
    first, count = pkg.shard_frames(5, world, rank)
    mine = {f: O.ref_build(pkg.synth.noise(270, 480, frame=f), octaves=5, S=S, want=("inplace",))["inplace"]
            for f in range(first, first + count)}

    gathered = [None] * world
    dist.all_gather_object(gathered, {"band": (r0, nrows, band), "frames": mine})
    if rank == 0:
        full = O.ref_build(pkg.synth.noise(H, W, frame=3), octaves=octs, S=S, want=("inplace",))["inplace"]
        for o in range(octs):
            rebuilt = np.concatenate([g["band"][2][o] for g in gathered], axis=1)
            assert rebuilt.shape == full[o].shape
            assert np.array_equal(rebuilt.view(np.uint32), full[o].view(np.uint32)), f"octave {o}"
        seen = sorted(f for g in gathered for f in g["frames"])
        assert seen == list(range(5))
        want = O.ref_build(pkg.synth.noise(270, 480, frame=4), octaves=5, S=S, want=("inplace",))["inplace"]
        got = [g["frames"][4] for g in gathered if 4 in g["frames"]][0]
        assert all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(got, want))
        open(os.path.join(tmp, "ok"), "w").write("ok")
    d.barrier()
    d.close()


def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").read_text() == "ok"


def test_rank_geometry_covers_every_workload():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    import bench
    pkg = entry.load_package()
    for wl, (H, W, octs, frames, part, _) in bench.WORKLOADS.items():
        for world in (1, 2, 4, 8):
            geo = [bench.rank_geometry(pkg, wl, world, r) for r in range(world)]
            if part == "rowband":
                assert sum(g[0] for g in geo) == H and all(g[2] == H for g in geo)
            elif part == "batch":
                assert sum(g[5] for g in geo) == frames and all(g[0] == H for g in geo)
            else:
                assert all(g[0] == H and g[5] == frames for g in geo)
