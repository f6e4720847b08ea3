"""Row-band halo exchange for CONV mode (host plumbing; the blur itself is csrc/conv_kernel.cuh).

A frame split into row bands (partition.band_rows) is built level by level: before the blur that produces
level s of octave o, every band needs R_s rows of the blur's INPUT (level s-1, or the raw frame for (0, 0))
from the band above and the band below -- its own first / last R_s rows go the other way.  That is one
neighbour send/recv pair per side per level (2 messages of R_s x pitch_o x 4 bytes per GPU per level), over
NVLink: torch.distributed P2P on NCCL between ranks (one process per GPU), or plain device-to-device copies
between handles that live in one process (which is also how the multi-band path is tested on a single GPU).

REF mode never needs this: the reference's "filter" is pointwise (halo radius 0, GuassDePyramid.h:122-131).
The reference's own distributed design moves whole levels row by row to a root rank instead
(GaussDePyramid-MPI.h:276-303); nothing is gathered here -- outputs stay on the GPU that produced them.
"""
from __future__ import annotations


class _DevMem:
    """Zero-copy view of library-owned device memory for torch (CUDA array interface)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def device_view(ptr: int, nbytes: int, device):
    import torch
    return torch.as_tensor(_DevMem(ptr, nbytes), device=device)


def level_schedule(ss):
    """(octave, level) in build order."""
    return [(o, s) for o in range(ss.octaves) for s in range(ss.levels)]


class LocalExchanger:
    """Bands held by several handles of ONE process (same or different GPUs): neighbour copies."""

    def __init__(self, handles, devices=None):
        import torch
        self.hs = list(handles)
        self.devs = devices or [torch.device("cuda", torch.cuda.current_device())] * len(self.hs)

    def exchange(self, octave: int, level: int, frame: int = 0) -> None:
        if self.hs[0].halo_rows(octave, level) == 0 or len(self.hs) < 2:
            return
        ptrs = [h.halo_ptrs(octave, level, frame) for h in self.hs]
        for i in range(len(self.hs) - 1):           # band i (above) <-> band i+1 (below)
            up, dn = ptrs[i], ptrs[i + 1]
            n = up[4]
            device_view(dn[2], n, self.devs[i + 1]).copy_(device_view(up[1], n, self.devs[i]))   # my last rows -> its recv_up
            device_view(up[3], n, self.devs[i]).copy_(device_view(dn[0], n, self.devs[i + 1]))   # its first rows -> my recv_down

    def build(self, frame: int = 0) -> None:
        for o, s in level_schedule(self.hs[0]):
            self.exchange(o, s, frame)
            for h in self.hs:
                h.conv_step(o, s, frame)


class DistExchanger:
    """One band per rank of a torch.distributed (NCCL) group; ranks are ordered top to bottom."""

    def __init__(self, ss, rank: int, world: int, device, group=None):
        self.ss, self.rank, self.world, self.device, self.group = ss, rank, world, device, group

    def exchange(self, octave: int, level: int, frame: int = 0) -> None:
        import torch.distributed as dist
        if self.world < 2 or self.ss.halo_rows(octave, level) == 0:
            return
        send_up, send_dn, recv_up, recv_dn, n = self.ss.halo_ptrs(octave, level, frame)
        ops = []
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, device_view(send_up, n, self.device), self.rank - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, device_view(recv_up, n, self.device), self.rank - 1, self.group))
        if self.rank < self.world - 1:
            ops.append(dist.P2POp(dist.isend, device_view(send_dn, n, self.device), self.rank + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, device_view(recv_dn, n, self.device), self.rank + 1, self.group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    def build(self, frame: int = 0) -> None:
        for o, s in level_schedule(self.ss):
            self.exchange(o, s, frame)
            self.ss.conv_step(o, s, frame)


class LocalPeerLink:
    """Bands held by several handles of ONE process read each other's planes in place (no copies): every handle
    attaches its neighbours, then all bands issue the same level steps in lockstep.  On a single GPU the handles
    must share one stream (the progress-counter waits are then already satisfied when they run)."""

    def __init__(self, handles):
        self.hs = list(handles)
        for i, h in enumerate(self.hs):
            if i > 0:
                h.peer_attach_local(0, self.hs[i - 1])
            if i + 1 < len(self.hs):
                h.peer_attach_local(1, self.hs[i + 1])

    def build(self, frame: int = 0) -> None:
        for o, s in level_schedule(self.hs[0]):
            for h in self.hs:
                h.conv_step(o, s, frame)


class PeerExchanger:
    """One band per rank: neighbours' planes are mapped through CUDA IPC and read inside the blur kernel over
    NVLink; signal / wait kernels on a counter in peer memory keep the ranks in step.  torch.distributed only
    carries the 1 KB IPC blobs at start-up -- there is no per-level host-side communication at all."""

    def __init__(self, ss, rank: int, world: int, group=None):
        import torch.distributed as dist
        self.ss, self.rank, self.world = ss, rank, world
        blobs = [None] * world
        dist.all_gather_object(blobs, ss.ipc_export(), group=group)
        if rank > 0:
            ss.ipc_attach(0, blobs[rank - 1])
        if rank < world - 1:
            ss.ipc_attach(1, blobs[rank + 1])
        dist.barrier(group=group)

    def build(self, frame: int = 0) -> None:
        """One library call: octaves on concurrent streams, per-octave progress counters between the bands."""
        self.ss.build(frame)

    def build_stepwise(self, frame: int = 0) -> None:
        for o, s in level_schedule(self.ss):
            self.ss.conv_step(o, s, frame)
