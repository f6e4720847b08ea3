"""Multi-GPU partitioner (host logic only; no device code, no collectives).

The reference's only distributed design is level-per-rank MPI with a root gather, one message per image
row (GaussDePyramid-MPI.h:265-335, mpitest.cpp:35-189).  On one 8xB200 box the path shards without any
data-path communication in REF mode:

* BATCH    frames are independent          -> contiguous frame ranges per rank (:func:`shard_frames`)
* ROWBAND  one huge frame, pointwise math  -> contiguous row bands whose boundaries are multiples of
           2^(octaves-1), so that ``r << o`` decimation (GuassDePyramid.h:80) stays band-local for every
           octave (:func:`band_rows`); each rank also needs only its slice of the row window.

CONV mode adds a per-level neighbour halo exchange on top of the same bands.
"""
from __future__ import annotations


def shard_frames(n_frames: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous [first, first+count) frame range of `rank`; the first n_frames % world ranks get one more."""
    if world < 1 or not 0 <= rank < world or n_frames < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(n_frames, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def band_rows(height: int, octaves: int, world: int, rank: int) -> tuple[int, int]:
    """Row band [row0, row0+rows) of `rank` for an image of `height` rows and `octaves` octaves.

    Boundaries are multiples of align = 2^(octaves-1); the aligned row blocks are dealt as evenly as
    possible and the last band also takes the unaligned remainder.  Ranks beyond the number of blocks
    get (height, 0) -- an empty band.
    """
    if world < 1 or not 0 <= rank < world or height < 1 or octaves < 1:
        raise ValueError("bad band request")
    align = 1 << (octaves - 1)
    blocks = max(height // align, 1)
    used = min(world, blocks)
    if rank >= used:
        return height, 0
    base, extra = divmod(blocks, used)
    b0 = rank * base + min(rank, extra)
    nb = base + (1 if rank < extra else 0)
    row0 = b0 * align
    row1 = height if rank == used - 1 else (b0 + nb) * align
    return row0, row1 - row0
