"""Sharding one reference batch over several GPUs / ranks.

Reads are independent, the index is replicated, so a batch is cut into
contiguous per-rank ranges and the per-rank outputs are concatenated in input
order.  The only batch-level state of the reference is the max_gapo clamp taken
from the batch's longest read (bwtaln.c:89-92): every shard must be run with
the WHOLE batch's max_len (engine knob `batch_max_len`).  No collective is
needed on the data path; `gather_in_order` is only the host-side merge.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, world: int, rank: int):
    """Contiguous range [lo, hi) of rank `rank` among `world` for n reads (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def take_shard(lens: np.ndarray, offs: np.ndarray, codes: np.ndarray, lo: int, hi: int):
    """Sub-batch [lo, hi) with offsets rebased to its own code buffer."""
    if hi <= lo:
        return lens[:0].copy(), offs[:0].copy(), codes[:0].copy()
    start = int(offs[lo])
    end = int(offs[hi]) if hi < len(offs) else len(codes)
    return lens[lo:hi].copy(), (offs[lo:hi] - start).copy(), codes[start:end].copy()


def gather_in_order(parts):
    """parts: list over ranks of (n_aln, records) -> concatenation in rank (= input) order."""
    n_aln = np.concatenate([p[0] for p in parts]) if parts else np.empty(0, np.int32)
    recs = np.concatenate([p[1] for p in parts]) if parts else np.empty(0)
    return n_aln, recs
