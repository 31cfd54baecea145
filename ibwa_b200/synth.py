"""Deterministic synthetic references and reads.

Test/bench data only — nothing here is on the product path.  The generators are
numpy so the same inputs can be rebuilt on the GPU box (no network, no files
shipped) from a seed.

Read models
-----------
default : uniform start; per-base substitution p=0.01 to a different base; 2 % of
          reads carry one indel (ins/del 50/50, length U{1,2,3}, position
          U[10, len-10)); 50 % reverse-complemented.
stress  : per-base substitution 0.02; deletion-open 0.001 and insertion-open
          0.001 per base with geometric extension p=0.3 (config 3).
"""
from __future__ import annotations

import numpy as np

_COMP = np.array([3, 2, 1, 0, 4, 5], dtype=np.uint8)
_NT = np.frombuffer(b"ACGTN-", dtype=np.uint8)


def random_genome(n: int, seed: int) -> np.ndarray:
    """i.i.d. uniform ACGT as nt4 codes (uint8, 0..3)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty(n, dtype=np.uint8)
    step = 10_000_000
    for s in range(0, n, step):
        e = min(n, s + step)
        out[s:e] = rng.integers(0, 4, size=e - s, dtype=np.uint8)
    return out


def repeat_rich_genome(n: int, seed: int) -> np.ndarray:
    """Random sequence with embedded repeats (SURVEY.md §4 fixture F2): diverged
    copies of a 300-bp unit, (AC)n, A-run, (GATTACA)n."""
    rng = np.random.Generator(np.random.PCG64(seed))
    g = rng.integers(0, 4, size=n, dtype=np.uint8)
    unit = rng.integers(0, 4, size=300, dtype=np.uint8)
    pos = 1000
    for _ in range(40):
        if pos + 300 >= n:
            break
        cp = unit.copy()
        mut = rng.random(300) < 0.02
        cp[mut] = (cp[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
        g[pos:pos + 300] = cp
        pos += 300 + int(rng.integers(50, 2000))
    def put(at, arr):
        if at + len(arr) < n:
            g[at:at + len(arr)] = arr
    put(n // 2, np.tile(np.array([0, 1], dtype=np.uint8), 200))            # (AC)n
    put(n // 2 + 2000, np.zeros(400, dtype=np.uint8))                       # A x 400
    put(n // 2 + 4000, np.tile(np.array([2, 0, 3, 3, 0, 1, 0], dtype=np.uint8), 60))  # (GATTACA)n
    return g


def write_fasta(path: str, genome: np.ndarray, contig_len: int | None = None, prefix: str = "chr") -> None:
    n = len(genome)
    contig_len = contig_len or n
    with open(path, "wb") as f:
        ci = 0
        for s in range(0, n, contig_len):
            ci += 1
            f.write(b">" + prefix.encode() + str(ci).encode() + b"\n")
            seq = _NT[genome[s:min(n, s + contig_len)]]
            m = len(seq)
            full = m // 100 * 100
            if full:
                lines = np.empty((full // 100, 101), dtype=np.uint8)
                lines[:, :100] = seq[:full].reshape(-1, 100)
                lines[:, 100] = 10
                f.write(lines.tobytes())
            if m > full:
                f.write(seq[full:].tobytes() + b"\n")


def revcomp(a: np.ndarray) -> np.ndarray:
    return _COMP[a[::-1]]


def simulate_reads(genome: np.ndarray, n_reads: int, length: int, seed: int, model: str = "default",
                   n_frac: float = 0.0):
    """Returns list of nt4 uint8 arrays (reads in sequencing orientation)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = len(genome)
    reads = []
    for _ in range(n_reads):
        span = length + 16
        start = int(rng.integers(0, n - span))
        src = genome[start:start + span]
        if model == "default":
            r = src[:length].copy()
            if rng.random() < 0.02:
                il = int(rng.integers(1, 4))
                p = int(rng.integers(10, max(11, length - 10)))
                if rng.random() < 0.5:   # insertion in the read
                    ins = rng.integers(0, 4, size=il, dtype=np.uint8)
                    r = np.concatenate([src[:p], ins, src[p:]])[:length]
                else:                     # deletion from the read
                    r = np.concatenate([src[:p], src[p + il:]])[:length]
            sub = rng.random(length) < 0.01
            ns = int(sub.sum())
            if ns:
                r[sub] = (r[sub] + rng.integers(1, 4, size=ns, dtype=np.uint8)) & 3
        elif model == "stress":
            out = []
            j = 0
            while len(out) < length and j < span:
                u = rng.random()
                if u < 0.001:             # deletion: skip reference bases
                    j += 1
                    while rng.random() < 0.3:
                        j += 1
                    continue
                if u < 0.002:             # insertion: extra read bases
                    out.append(int(rng.integers(0, 4)))
                    while rng.random() < 0.3:
                        out.append(int(rng.integers(0, 4)))
                    continue
                b = int(src[j]); j += 1
                if rng.random() < 0.02:
                    b = (b + int(rng.integers(1, 4))) & 3
                out.append(b)
            while len(out) < length:
                out.append(int(rng.integers(0, 4)))
            r = np.array(out[:length], dtype=np.uint8)
        else:
            raise ValueError(model)
        if n_frac > 0:
            nm = rng.random(length) < n_frac
            r[nm] = 4
        if rng.random() < 0.5:
            r = revcomp(r)
        reads.append(np.ascontiguousarray(r))
    return reads


def simulate_reads_fast(genome: np.ndarray, n_reads: int, length: int, seed: int) -> np.ndarray:
    """Vectorised 'default' model for large counts: returns (n_reads, length) uint8."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = len(genome)
    starts = rng.integers(0, n - length - 8, size=n_reads)
    idx = starts[:, None] + np.arange(length)[None, :]
    # single indel on 2 % of reads: shift the tail of the index map
    has = rng.random(n_reads) < 0.02
    il = rng.integers(1, 4, size=n_reads)
    p = rng.integers(10, max(11, length - 10), size=n_reads)
    isdel = rng.random(n_reads) < 0.5
    col = np.arange(length)[None, :]
    tail = col >= p[:, None]
    dshift = np.where(has & isdel, il, 0)[:, None] * tail
    idx = idx + dshift
    r = genome[idx]
    insm = (has & ~isdel)[:, None] & tail & (col < (p + il)[:, None])
    ishift = np.where(has & ~isdel, il, 0)[:, None] * (col >= (p + il)[:, None])
    r = genome[idx - ishift]
    ni = int(insm.sum())
    if ni:
        r[insm] = rng.integers(0, 4, size=ni, dtype=np.uint8)
    sub = rng.random((n_reads, length)) < 0.01
    ns = int(sub.sum())
    r[sub] = (r[sub] + rng.integers(1, 4, size=ns, dtype=np.uint8)) & 3
    rc = rng.random(n_reads) < 0.5
    r[rc] = _COMP[r[rc][:, ::-1]]
    return np.ascontiguousarray(r)


def write_fastq(path: str, reads, names=None, lower_frac: float = 0.0, seed: int = 0) -> None:
    rng = np.random.Generator(np.random.PCG64(seed))
    with open(path, "wb") as f:
        for i, r in enumerate(reads):
            s = _NT[np.asarray(r)].tobytes()
            if lower_frac and rng.random() < lower_frac:
                s = s.lower()
            name = names[i] if names is not None else f"r{i}"
            f.write(b"@" + name.encode() + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")
