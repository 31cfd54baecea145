""".bwt / .rbwt files in the reference's on-disk layout.

Reference: bwt_restore_bwt bwtio.c:51-70, bwt_dump_bwt bwtio.c:7-15, layout
macros bwt.h:56-63, occ interleaving bwtmisc.c:122-144.  File = u32 primary,
u32 L2[1..4], then bwt_size u32 words: per 128 bases a block of 4 cumulative
count words followed by 8 words of 16 bases (2 bits, MSB first); the last block
is truncated and one final 4-word count block follows.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Bwt:
    primary: int
    L2: np.ndarray      # uint32[5], L2[0] == 0
    seq_len: int
    bwt: np.ndarray     # uint32[bwt_size]

    @property
    def bwt_size(self) -> int:
        return int(self.bwt.shape[0])


def bwt_restore_bwt(path: str) -> Bwt:
    raw = np.fromfile(path, dtype=np.uint32)
    if raw.size < 5:
        raise IOError(f"[bwt_restore_bwt] '{path}' is not a .bwt file")
    L2 = np.zeros(5, dtype=np.uint32)
    L2[1:] = raw[1:5]
    return Bwt(primary=int(raw[0]), L2=L2, seq_len=int(L2[4]), bwt=np.ascontiguousarray(raw[5:]))


def bwt_dump_bwt(path: str, b: Bwt) -> None:
    with open(path, "wb") as f:
        np.array([b.primary], dtype=np.uint32).tofile(f)
        np.asarray(b.L2[1:5], dtype=np.uint32).tofile(f)
        np.asarray(b.bwt, dtype=np.uint32).tofile(f)


def expected_words(seq_len: int) -> int:
    """Word count of the payload for a text of seq_len bases."""
    return (seq_len + 15) // 16 + 4 * ((seq_len + 127) // 128 + 1)


@dataclass
class Sa:
    """Sampled suffix array as bwt_restore_sa leaves it (bwtio.c:29-49): sa[0] = 0xffffffff, sa[j] = SA(j * sa_intv)."""
    primary: int
    L2: np.ndarray
    seq_len: int
    sa_intv: int
    sa: np.ndarray      # uint32[n_sa]


def bwt_restore_sa(path: str) -> Sa:
    raw = np.fromfile(path, dtype=np.uint32)
    L2 = np.zeros(5, dtype=np.uint32)
    L2[1:] = raw[1:5]
    sa_intv, seq_len = int(raw[5]), int(raw[6])
    n_sa = (seq_len + sa_intv) // sa_intv
    sa = np.empty(n_sa, dtype=np.uint32)
    sa[0] = 0xFFFFFFFF
    sa[1:] = raw[7:7 + n_sa - 1]
    return Sa(primary=int(raw[0]), L2=L2, seq_len=seq_len, sa_intv=sa_intv, sa=sa)


def bwt_dump_sa(path: str, s: Sa) -> None:
    """bwt_dump_sa (bwtio.c:17-27)."""
    with open(path, "wb") as f:
        np.array([s.primary], dtype=np.uint32).tofile(f)
        np.asarray(s.L2[1:5], dtype=np.uint32).tofile(f)
        np.array([s.sa_intv, s.seq_len], dtype=np.uint32).tofile(f)
        np.asarray(s.sa[1:], dtype=np.uint32).tofile(f)
