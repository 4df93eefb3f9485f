"""
Slab decomposition of the 2-D grid over the GPUs of one box.

The grid is cut along y (the slow index of the column-major arrays, so one ghost row of one
field is contiguous): rank r owns interior rows [j_offset, j_offset + ny_local) and keeps its
own (Nx, ny_local + 6, .) public arrays; x boundaries stay local; a periodic y boundary wraps
rank P-1 <-> rank 0.  libweno5mhd exchanges four ghost rows of the cell fields per RK stage
with NCCL send/recv (weno5_comm_init); this module is the host-side bookkeeping around it.
"""
import numpy as np

NBND = 3


def slab(ny, nranks, rank):
    """(j_offset, ny_local) of `rank`: even split, the first ny % nranks ranks get one more row."""
    if not (0 <= rank < nranks):
        raise ValueError(f"rank {rank} outside 0..{nranks - 1}")
    base, rem = divmod(ny, nranks)
    ny_local = base + (1 if rank < rem else 0)
    j_offset = rank * base + min(rank, rem)
    if ny_local < 8:
        raise ValueError(f"slab of {ny_local} rows is too thin (need >= 8): ny={ny}, nranks={nranks}")
    return j_offset, ny_local


def neighbours(rank, nranks, periodic):
    """(down, up) neighbour ranks along y, -1 where the slab touches a physical boundary."""
    down = rank - 1 if rank > 0 else (nranks - 1 if periodic and nranks > 1 else -1)
    up = rank + 1 if rank + 1 < nranks else (0 if periodic and nranks > 1 else -1)
    return down, up


def cut_slab(a, j_offset, ny_local, ghost=NBND, periodic=False):
    """Rows [j_offset - ghost, j_offset + ny_local + ghost) (interior numbering) of a global
    public array a[Nx, Ny(, C)] that itself carries NBND ghost rows.  With `periodic`, rows
    outside the global array wrap around the interior; otherwise they clamp (zero gradient)."""
    Ny = a.shape[1]
    ny = Ny - 2 * NBND
    j = np.arange(j_offset - ghost, j_offset + ny_local + ghost)
    if periodic:
        src = np.mod(j, ny) + NBND
    else:
        src = np.clip(j + NBND, 0, Ny - 1)
        # rows beyond the stored ghosts copy the last interior row (outflow)
        src = np.where(j < -NBND, NBND, src)
        src = np.where(j >= ny + NBND, Ny - NBND - 1, src)
    return np.asfortranarray(a[:, src])


def assemble(slabs, ghost=NBND):
    """Inverse of cut_slab for the interior rows: concatenate the slabs' owned rows and put
    NBND rows of the first/last slab around them."""
    parts = [s[:, ghost:s.shape[1] - ghost] for s in slabs]
    lo = slabs[0][:, ghost - NBND:ghost]
    hi = slabs[-1][:, slabs[-1].shape[1] - ghost:slabs[-1].shape[1] - ghost + NBND]
    return np.asfortranarray(np.concatenate([lo] + parts + [hi], axis=1))


def broadcast_unique_id(make_id, rank, dist=None):
    """Rank 0 calls make_id() (-> 128 bytes, weno5_comm_unique_id) and every rank receives it
    through torch.distributed (any backend); without `dist` (single process) it is returned
    as is."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return make_id()
    obj = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    return obj[0]
