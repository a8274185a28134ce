"""Host-side helpers of the two multi-GPU modes.  One process (or thread) per GPU.

* node-batch mode: boxes are independent -> split the batch contiguously across ranks, replicate the matrix,
  no collective in the loop (``shard_boxes``).
* row-partition mode: every rank loads ITS block of rows (balanced by nnz, ``partition_rows``) and a replica of
  the box; ``GpuBoundEngine.comm_init`` attaches an NCCL communicator and every Jacobi round merges the
  candidate bounds with an all-reduce (MAX on lb, MIN on ub) over NVLink.  The merged result is bitwise
  independent of the number of ranks.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from .instances import LinearRows


def shard_boxes(n_boxes: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [b0, b1) slice of the box batch owned by ``rank``."""
    per = (n_boxes + world - 1) // world
    return min(n_boxes, rank * per), min(n_boxes, (rank + 1) * per)


def row_partition_bounds(row_ptr: np.ndarray, world: int) -> np.ndarray:
    """Row boundaries [world+1] of a block partition balanced by nnz (each block gets ~nnz/world entries)."""
    m = len(row_ptr) - 1
    nnz = int(row_ptr[-1])
    cuts = [0]
    for r in range(1, world):
        target = nnz * r // world
        cuts.append(int(np.searchsorted(row_ptr, target, side="left")))
    cuts.append(m)
    cuts = np.maximum.accumulate(np.minimum(np.asarray(cuts, np.int64), m))
    return cuts


def partition_rows(inst: LinearRows, world: int) -> List[LinearRows]:
    """Split the rows of ``inst`` into ``world`` blocks (all n columns each)."""
    cuts = row_partition_bounds(inst.row_ptr, world)
    blocks = []
    for r in range(world):
        i0, i1 = int(cuts[r]), int(cuts[r + 1])
        e0, e1 = int(inst.row_ptr[i0]), int(inst.row_ptr[i1])
        blocks.append(LinearRows(
            m=i1 - i0, n=inst.n, row_ptr=(inst.row_ptr[i0:i1 + 1] - e0).astype(np.int32), col=inst.col[e0:e1],
            val=inst.val[e0:e1], row_lb=inst.row_lb[i0:i1], row_ub=inst.row_ub[i0:i1], var_type=inst.var_type,
            lb=inst.lb, ub=inst.ub, row_active=None if inst.row_active is None else inst.row_active[i0:i1],
            name=f"{inst.name}[rows {i0}:{i1}]", xstar=inst.xstar))
    return blocks
