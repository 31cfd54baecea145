"""Read input -> packed batches: host-side mirror of bwa_read_seq.

Reference: bwa_read_seq bwaseqio.c:145-208 (FASTA/FASTQ, plain or gzip), parser
kseq.h:150-194, nt4 table bntseq.c:39-56, quality trimming bwa_trim_read
bwaseqio.c:74-87, barcode strip :164-177, Illumina 1.3 offset :158-159.

The packed form handed to the engine is the read in sequencing orientation as
nt4 codes after trimming: `seq` (reversed) and `rseq` (reverse-complement, or
plain reverse with -c) of the reference are derived on the device.
"""
from __future__ import annotations

import gzip
from dataclasses import dataclass

import numpy as np

from .opts import BWA_MODE_IL13

BWA_MIN_RDLEN = 35

NT4 = np.full(256, 4, dtype=np.uint8)
for _ch, _v in zip(b"ACGTacgt", [0, 1, 2, 3, 0, 1, 2, 3]):
    NT4[_ch] = _v
NT4[ord("-")] = 5


@dataclass
class ReadBatch:
    lens: np.ndarray     # int32[n]   length after trimming
    offs: np.ndarray     # int64[n]   offset of each read in `codes`
    codes: np.ndarray    # uint8[sum(full_len)] nt4 codes, sequencing orientation
    names: list

    def __len__(self):
        return len(self.lens)


def _open(path: str):
    f = open(path, "rb")
    magic = f.read(2)
    f.close()
    return gzip.open(path, "rb") if magic == b"\x1f\x8b" else open(path, "rb")


def _records(data: bytes):
    """kseq_read semantics (kseq.h:150-194) over a whole buffer."""
    n = len(data)
    p = 0
    last = 0
    while True:
        if last == 0:
            while p < n and data[p] not in (62, 64):   # '>' '@'
                p += 1
            if p >= n:
                return
            p += 1
        # name: up to first whitespace
        q = p
        while q < n and not chr(data[q]).isspace():
            q += 1
        name = data[p:q]
        if q >= n:
            if q == p:
                return
            delim = -1
        else:
            delim = data[q]
        p = q + 1
        if delim != 10 and delim != -1:
            e = data.find(b"\n", p)
            p = n if e < 0 else e + 1
        seq = bytearray()
        c = -1
        while p < n:
            c = data[p]
            p += 1
            if c in (62, 43, 64):   # '>' '+' '@'
                break
            if 33 <= c <= 126:
                seq.append(c)
            c = -1
        last = c if c in (62, 64) else 0
        if c != 43:
            yield name, bytes(seq), None
            if p >= n and c == -1:
                return
            continue
        e = data.find(b"\n", p)
        if e < 0:
            return                    # -2: truncated
        p = e + 1
        qual = bytearray()
        while p < n and len(qual) < len(seq):
            c = data[p]
            p += 1
            if 33 <= c <= 127:
                qual.append(c)
        if p < n:
            p += 1                    # the parser consumes one more character
        last = 0
        if len(qual) != len(seq):
            return                    # -2: truncated quality; reading stops silently
        yield name, bytes(seq), bytes(qual)


def trim_len(trim_qual: int, qual: np.ndarray) -> int:
    """bwa_trim_read (bwaseqio.c:74-87): kept length."""
    L = len(qual)
    if trim_qual < 1:
        return L
    s = 0
    best = 0
    best_l = L - 1
    for l in range(L - 1, BWA_MIN_RDLEN - 2, -1):
        s += trim_qual - (int(qual[l]) - 33)
        if s < 0:
            break
        if s > best:
            best, best_l = s, l
    return best_l + 1


def read_batches(path: str, mode: int, trim_qual: int, n_needed: int = 0x40000):
    """Yields ReadBatch objects of up to n_needed reads (bwtaln.c:193)."""
    with _open(path) as f:
        data = f.read()
    is_64 = bool(mode & BWA_MODE_IL13)
    l_bc = (mode >> 24) & 0xFF
    lens, offs, chunks, names = [], [], [], []
    total = 0
    for name, seq, qual in _records(data):
        if qual is not None and is_64:
            qual = bytes((c - 31) & 0xFF for c in qual)
        if len(seq) <= l_bc:
            continue
        if l_bc:
            seq = seq[l_bc:]
            qual = qual[l_bc:] if qual is not None else None
        codes = NT4[np.frombuffer(seq, dtype=np.uint8)]
        L = len(codes)
        if qual is not None and trim_qual >= 1:
            L = trim_len(trim_qual, np.frombuffer(qual, dtype=np.uint8))
        nm = name.decode("latin-1")
        if len(nm) > 2 and nm[-2] == "/" and nm[-1] in "12":
            nm = nm[:-2]
        lens.append(L)
        offs.append(total)
        chunks.append(codes)
        names.append(nm)
        total += len(codes)
        if len(lens) == n_needed:
            yield ReadBatch(np.array(lens, np.int32), np.array(offs, np.int64), np.concatenate(chunks), names)
            lens, offs, chunks, names, total = [], [], [], [], 0
    if lens:
        yield ReadBatch(np.array(lens, np.int32), np.array(offs, np.int64), np.concatenate(chunks), names)
