"""Z-slab partitioning of the sample grid across ranks (one process per GPU).

The reference parallelises over disjoint 64^3 chunks, each triangle being re-voxelized per chunk with a clip box
(src/obj2voxel.cpp:245-272, src/voxelization.cpp:440-444).  Across GPUs the same independence is used at slab
granularity: rank r owns sample-space z in [z_r, z_{r+1}), boundaries multiples of 64 (one reference chunk row), every
voxel has exactly one owner and no voxel data is ever exchanged.  The only collectives are the triangle broadcast from the
ingest rank and an all-reduce of per-slab counts (torch.distributed: NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np

CHUNK = 64  # reference CHUNK_SIZE, src/constants.hpp:10


def equal_slabs(sample_resolution, world_size):
    """Slab boundaries [z_0=0, ..., z_W=padded S] in whole 64-voxel chunk rows, as even as the row count allows.
    Grids smaller than 64*W fall back to 8-voxel (tile) rows."""
    unit = CHUNK if sample_resolution >= CHUNK * world_size else 8
    rows = -(-sample_resolution // unit)
    bounds = [((rows * r) // world_size) * unit for r in range(world_size + 1)]
    return bounds


def balanced_slabs(z_histogram, sample_resolution, world_size, unit=CHUNK):
    """Slab boundaries from a per-z-row work histogram (len = rows of `unit` voxels): prefix-sum split so every rank gets
    about the same work.  Boundaries stay multiples of `unit`."""
    hist = np.asarray(z_histogram, dtype=np.float64)
    rows = len(hist)
    total = hist.sum()
    if total <= 0 or rows < world_size:
        return equal_slabs(sample_resolution, world_size)
    prefix = np.concatenate([[0.0], np.cumsum(hist)])
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        row = int(np.searchsorted(prefix, target, side="left"))
        row = min(max(row, bounds[-1] // unit + 1), rows - (world_size - r))
        bounds.append(row * unit)
    bounds.append(rows * unit)
    return bounds


def z_row_histogram(verts_z_min, verts_z_max, transform_row, sample_resolution, unit=CHUNK):
    """Work estimate per z row from per-triangle z extents in model space (numpy float32 arrays) and the z row of the
    mesh transform (m20, m21, m22, t2 ignored except scale/offset for axis-aligned transforms)."""
    scale, offset = transform_row
    lo = np.clip(np.floor(verts_z_min * scale + offset), 0, sample_resolution - 1).astype(np.int64) // unit
    hi = np.clip(np.floor(verts_z_max * scale + offset), 0, sample_resolution - 1).astype(np.int64) // unit
    rows = -(-sample_resolution // unit)
    hist = np.zeros(rows + 1, dtype=np.float64)
    np.add.at(hist, lo, 1.0)
    np.add.at(hist, hi + 1, -1.0)
    return np.cumsum(hist)[:rows]


def my_slab(bounds, rank):
    """(z0, z1) of the rank's slab, or None when the rank owns nothing (more ranks than rows: equal_slabs(8, 2) ==
    [0, 0, 8]).  Such a rank skips the job: a (0, 0) pair means "whole grid" to the engine."""
    z0, z1 = int(bounds[rank]), int(bounds[rank + 1])
    return (z0, z1) if z1 > z0 else None


def broadcast_mesh(tensors, src=0):
    """Broadcast the triangle arrays from the ingest rank (torch.distributed must be initialised)."""
    import torch.distributed as dist

    for t in tensors:
        if t is not None:
            dist.broadcast(t, src=src)


def allreduce_counts(values, device):
    """Sum per-slab counters (voxels, contributions, ...) over ranks; returns python ints."""
    import torch
    import torch.distributed as dist

    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(x) for x in t.tolist()]
