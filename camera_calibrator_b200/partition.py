"""Host-side sharding of the hot path over ranks (one process per GPU).

* one large calibration: contiguous, balanced ranges of views (frames) per rank; every rank
  keeps its views' observations, poses and pose blocks resident, the intrinsics (camera
  extrinsics) are replicated, and each LM iteration sums the packed reduced system over ranks
  with one small allreduce;
* batch of independent calibrations: contiguous ranges of problems per rank, no communication.
"""
from __future__ import annotations

import numpy as np


def split_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """[a, b) of `n` items owned by `rank` out of `world`: sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return n * rank // world, n * (rank + 1) // world


def split_views_by_observations(view_offsets, world: int, rank: int) -> tuple[int, int]:
    """Contiguous view range balanced by OBSERVATION count (ragged views)."""
    off = np.asarray(view_offsets, dtype=np.int64)
    n_views = len(off) - 1
    total = int(off[-1] - off[0])
    cuts = [int(np.searchsorted(off - off[0], total * r / world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, n_views
    for i in range(1, world + 1):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts[rank], cuts[rank + 1]


def slice_views(arrays, view_offsets, a: int, b: int):
    """Observation arrays and re-based offsets of views [a, b)."""
    off = np.asarray(view_offsets, dtype=np.int64)
    lo, hi = int(off[a]), int(off[b])
    return [np.asarray(x)[lo:hi] for x in arrays], off[a:b + 1] - off[a]
