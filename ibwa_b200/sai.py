""".sai wire format (bit-exact with the reference).

Reference: writer bwtaln.c:192,227-231; record bwt_aln1_t bwtaln.h:34-38
(u32 {n_mm:8,n_gapo:8,n_gape:8,a:1}, u32 k, u32 l, i32 score = 16 bytes);
readers bwase.c:660,674-684 and saiset.c:15-32,45-61.
"""
from __future__ import annotations

import numpy as np

from .opts import GapOpt

ALN_DTYPE = np.dtype([("packed", "<u4"), ("k", "<u4"), ("l", "<u4"), ("score", "<i4")])
assert ALN_DTYPE.itemsize == 16


def write_header(f, opt: GapOpt) -> None:
    f.write(opt.header_bytes())


def write_batch(f, n_aln: np.ndarray, records: np.ndarray) -> None:
    """Append one batch: per read `i32 n_aln` followed by its records."""
    n_aln = np.asarray(n_aln, dtype=np.int32)
    records = np.asarray(records, dtype=ALN_DTYPE)
    n = len(n_aln)
    total = int(n_aln.sum())
    assert total == len(records)
    out = np.empty(n + 4 * total, dtype=np.uint32)
    starts = np.zeros(n, dtype=np.int64)
    if n:
        np.cumsum(n_aln[:-1], out=starts[1:])
    pos = np.arange(n, dtype=np.int64) + 4 * starts      # word offset of each n_aln
    out[pos] = n_aln.view(np.uint32)
    if total:
        read_of = np.repeat(np.arange(n, dtype=np.int64), n_aln)
        rec_word = 4 * np.arange(total, dtype=np.int64) + read_of + 1
        rec = records.view(np.uint32).reshape(total, 4)
        for j in range(4):
            out[rec_word + j] = rec[:, j]
    f.write(out.tobytes())


def read_sai(path: str):
    """Returns (GapOpt header, n_aln int32[n], records ALN_DTYPE[total])."""
    raw = np.fromfile(path, dtype=np.uint8)
    opt = GapOpt.from_header(raw[:64].tobytes())
    words = raw[64:].view(np.uint32)
    n_aln = []
    recs = []
    p = 0
    nw = len(words)
    while p < nw:
        n = int(words[p].view(np.int32) if hasattr(words[p], "view") else words[p])
        p += 1
        n_aln.append(n)
        if n:
            recs.append(words[p:p + 4 * n])
            p += 4 * n
    rec = np.concatenate(recs).view(ALN_DTYPE) if recs else np.empty(0, dtype=ALN_DTYPE)
    return opt, np.array(n_aln, dtype=np.int32), rec


def unpack(records: np.ndarray):
    p = records["packed"]
    return dict(n_mm=p & 255, n_gapo=(p >> 8) & 255, n_gape=(p >> 16) & 255, a=(p >> 24) & 1,
                k=records["k"], l=records["l"], score=records["score"])
