"""Host-side helpers for the multi-GPU slab decomposition along x (one process / handle per GPU)."""
import numpy as np


def balanced_slabs(N, ranks, lo=0.0, hi=1.0, ghost=1):
    """Cuts x-planes [0, N) into `ranks` slabs holding equal shares of a block of particles that
    is uniform in x over [lo, hi] (the outer slabs also take the empty planes up to the walls)."""
    cuts = [0] + [int(round(N * (lo + (hi - lo) * r / ranks))) for r in range(1, ranks)] + [N]
    slabs = [(cuts[r], cuts[r + 1]) for r in range(ranks)]
    for b, e in slabs:
        if e - b < 2 + 2 * ghost:
            raise ValueError(f"slab [{b},{e}) thinner than 2 + 2*ghost planes")
    return slabs


def base_node_x(x, N):
    """Base node along x exactly as the kernels compute it (f32 arithmetic, C truncation, clamp)."""
    dx = np.float32(1.0 / N)
    dx_inv = np.float32(1.0 / np.float64(dx))
    g = np.asarray(x, np.float32) * dx_inv
    return np.clip((g - np.float32(0.5)).astype(np.int32), 0, N - 1)


def owner(x, N, slabs):
    """Rank owning each particle: the slab containing its base node."""
    b = base_node_x(x, N)
    ends = np.array([e for _, e in slabs])
    return np.searchsorted(ends, b, side="right").astype(np.int32)


def share_unique_id(dist, rank, make_id, src=0):
    """Rank `src` creates the NCCL unique id (mpm_comm_unique_id) and broadcasts it through the
    caller's torch.distributed process group (any backend: nccl on GPUs, gloo in CPU tests)."""
    box = [make_id() if rank == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]
