"""Row sharding of the resolvent-norm grid over one-process-per-GPU ranks.

Grid points are independent given (T, q0): there is no data-path collective.  T is broadcast once
(NCCL over NVLink), the computed rows are dealt cyclically (iteration counts are spatially correlated, so
contiguous blocks would unbalance the ranks) and the sigma_min tiles are gathered on rank 0.  The functions
work on whatever backend the process group uses (`nccl` on the GPU box, `gloo` in the CPU tests)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def rows_for_rank(ly: int, rank: int, world: int) -> np.ndarray:
    """Global row indices owned by `rank`: rank, rank+world, ..."""
    return np.arange(rank, ly, world)


def broadcast_matrix(T: torch.Tensor, src: int = 0) -> torch.Tensor:
    """Broadcast the (complex128, column-major-as-transposed-contiguous) matrix storage from `src`."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        flat = torch.view_as_real(T).reshape(-1)
        dist.broadcast(flat, src=src)
    return T


def gather_rows(Z_local: torch.Tensor, ly: int, lx: int, dst: int = 0):
    """Gather per-rank row blocks (rows_r x lx, any float/int dtype) into the full (ly, lx) array on `dst`.
    Returns the assembled tensor on dst, None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return Z_local
    world, rank = dist.get_world_size(), dist.get_rank()
    rmax = (ly + world - 1) // world
    pad = torch.zeros((rmax, lx), dtype=Z_local.dtype, device=Z_local.device)
    pad[:Z_local.shape[0]] = Z_local
    if rank == dst:
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.gather(pad, parts, dst=dst)
        Z = torch.empty((ly, lx), dtype=Z_local.dtype, device=Z_local.device)
        for r in range(world):
            rows = rows_for_rank(ly, r, world)
            Z[torch.as_tensor(rows, device=Z.device)] = parts[r][:len(rows)]
        return Z
    dist.gather(pad, None, dst=dst)
    return None
