"""Element -> GPU partition of the hot path (no collective on the path: condensed matrices return to the owning rank).

Mirrors the reference's own distribution of the element list over MPI ranks:
  * block split (ZOLTAN_LB = 0), src/modules/par_mesh.F90:66-82: the first `remainder` ranks get base+1 elements;
  * weighted split for hp meshes: contiguous blocks balanced by a per-element weight (the role OBJ_WEIGHT_DIM=1 plays
    in zoltan_wrapper.F90), here the algorithmic dense flops of the element.
One rank drives one GPU (rank r -> device r mod gpus_per_node).
"""
import numpy as np


def block_partition(nreles, num_procs):
    """subd_next(1:NRELES) of distr_mesh for ZOLTAN_LB=0 (0-based owner per element)."""
    base, rem = divmod(int(nreles), int(num_procs))
    owner = np.empty(nreles, dtype=np.int32)
    iel = 0
    for iproc in range(num_procs):
        size = base + 1 if iproc < rem else base
        owner[iel:iel + size] = iproc
        iel += size
    return owner


def block_range(nreles, num_procs, rank):
    """[first, last) of the contiguous block owned by `rank` under block_partition."""
    base, rem = divmod(int(nreles), int(num_procs))
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def weighted_partition(weights, num_procs):
    """Contiguous blocks with (nearly) equal total weight: element e goes to the rank whose weight interval contains the
    midpoint of e's cumulative-weight interval."""
    w = np.asarray(weights, dtype=np.float64)
    c = np.cumsum(w)
    mid = c - 0.5 * w
    owner = np.minimum((mid / c[-1] * num_procs).astype(np.int32), num_procs - 1)
    return owner


def elem_subd(owner, rank):
    """ELEM_SUBD(1:NRELES_SUBD) of a rank (data_structure3D.F90:182-191), 0-based global element indices."""
    return np.nonzero(np.asarray(owner) == rank)[0].astype(np.int64)


def device_of_rank(rank, gpus_per_node=8):
    return rank % gpus_per_node
