"""Picture-level data parallelism: IDR pictures are independent, so the
G GPUs of one box each take a disjoint subset and no data ever crosses GPUs.  Picture i goes to
rank i mod G; results are put back in picture order on the host."""
from __future__ import annotations

import numpy as np


def shard_indices(n_pics: int, world: int, rank: int) -> np.ndarray:
    """Indices of the pictures rank `rank` reconstructs (round robin, like the frame selection
    of demuxer/filter.c feeds one decoder, here it feeds G contexts)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside world")
    return np.arange(rank, n_pics, world, dtype=np.int64)


def take_pictures(soa, idx: np.ndarray):
    """Sub-batch of a Soa holding the pictures `idx` (in that order)."""
    from .synth import Soa
    n = soa.n_mbs
    rows = (idx[:, None] * n + np.arange(n)[None, :]).reshape(-1)
    return Soa(soa.width_mbs, soa.height_mbs, len(idx), soa.mb_kind[rows], soa.i16_mode[rows], soa.chroma_mode[rows],
               soa.qp_y[rows], soa.cbp[rows], soa.luma_modes[rows], soa.coeff[rows],
               soa.lists4x4, soa.lists8x8, soa.cb_qp_offset, soa.cr_qp_offset)


def merge_results(n_pics: int, world: int, per_rank: list[np.ndarray]) -> np.ndarray:
    """Inverse of shard_indices: per_rank[r][k] is the result of picture shard_indices(...)[k]."""
    first = next(a for a in per_rank if len(a))
    out = np.empty((n_pics,) + first.shape[1:], first.dtype)
    for r, a in enumerate(per_rank):
        out[shard_indices(n_pics, world, r)] = a
    return out
