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
