"""Multi-rank plumbing: one subdomain (decomposePar `simple` block) per rank/GPU, solids replicated on every
rank, ONE all-reduce of the per-solid force/torque per step replacing the reference's 2N Foam::reduce calls
(reference src/solidcloud.cpp:427-431).  No halo exchange: every per-cell result depends only on the cell's
own geometry and U."""
from __future__ import annotations

import numpy as np

from .cases import decompose_simple


def block_extent(rank: int, world: int, n):
    """(lo_index[3], size[3]) of rank's block of an n[0] x n[1] x n[2] cell box under the `simple` split
    (x fastest, like OpenFOAM's simpleGeomDecomp)."""
    px, py, pz = decompose_simple(world)
    n = (n, n, n) if np.isscalar(n) else tuple(n)
    ix, iy, iz = rank % px, (rank // px) % py, rank // (px * py)
    size = (n[0] // px, n[1] // py, n[2] // pz)
    return (ix * size[0], iy * size[1], iz * size[2]), size


def local_to_global_cells(rank: int, world: int, n) -> np.ndarray:
    """cellProcAddressing of the block: global cell id of every local cell (blockMesh numbering on both sides)."""
    n = (n, n, n) if np.isscalar(n) else tuple(n)
    lo, sz = block_extent(rank, world, n)
    i = np.arange(sz[0])[None, None, :] + lo[0]
    j = np.arange(sz[1])[None, :, None] + lo[1]
    k = np.arange(sz[2])[:, None, None] + lo[2]
    return (i + n[0] * (j + n[1] * k)).reshape(-1)


def init_library_comm(ctx):
    """Give `ctx` (a Context) the library's own NCCL communicator over the ranks of the initialised torch.distributed group:
    rank 0 draws the id, the group broadcasts the 128 bytes (this is the only use of torch.distributed on the data path — the
    per-step all-gather of the solid slices and the all-reduce of the force/torque sums then run inside libsdfibm_b200.so on the
    context stream).  A C++/OpenFOAM host does the same with Pstream / MPI_Bcast."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    uid = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0], rank, world)
    return world, rank


def allreduce_force_torque(ft):
    """Sum the per-rank partial (F, T)[n_solids, 6] over all ranks, in place.  `ft` is a torch tensor on the
    rank's device (NCCL) or on the CPU (gloo)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(ft, op=dist.ReduceOp.SUM)
    return ft


class ReplicatedSolids:
    """The replicated solid states on this rank's GPU, refreshed per step by ONE 1/N PCIe upload per rank plus an NCCL
    all-gather over NVLink, instead of N identical full uploads (every rank's host holds the same array, reference
    src/solidcloud.cpp: every rank integrates every solid)."""

    def __init__(self, n_solids: int, record_bytes: int, device):
        import torch
        import torch.distributed as dist

        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.n, self.rb = int(n_solids), int(record_bytes)
        self.per = (self.n + self.world - 1) // self.world            # records per rank slice (last slice padded)
        self.full = torch.empty(self.per * self.world * self.rb, dtype=torch.uint8, device=device)
        self.slice = torch.zeros(self.per * self.rb, dtype=torch.uint8, device=device)

    def refresh(self, solids: np.ndarray) -> int:
        """Upload this rank's slice of `solids` (the same array on every rank) and gather; returns the device pointer."""
        import torch
        import torch.distributed as dist

        raw = np.ascontiguousarray(solids).view(np.uint8).reshape(-1)
        lo, hi = self.rank * self.per * self.rb, min((self.rank + 1) * self.per, self.n) * self.rb
        dst = self.full if self.world == 1 else self.slice
        if hi > lo:   # asynchronous when `solids` lives in page-locked memory (capi.pinned_like), staged by torch otherwise
            dst[: hi - lo].copy_(torch.from_numpy(raw[lo:hi]), non_blocking=True)
        if self.world > 1:
            dist.all_gather_into_tensor(self.full, self.slice)
        return self.full.data_ptr()
