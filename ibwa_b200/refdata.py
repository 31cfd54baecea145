"""Reference-side index files and paired / alt-contig workloads for BASELINE configs 4 and 5 (test / bench data).

Nothing here is on the product path.  The downstream consumers of the engine's `.sai` — the UNCHANGED reference
`ibwa samse` / `ibwa sampe -R` — open more files than `aln` does: `<prefix>.ann/.amb/.pac` (bntseq.c:60-158,
166-254) and `<prefix>.sa/.rsa` (bwtio.c:17-49).  For synthetic genomes too large to index with the reference
binary in the time a GPU box has, these writers produce them from the text, byte for byte as `ibwa index` would
(tests/test_refdata.py compares every file with the reference's output).

  write_pac_ann_amb   bns_fasta2bntseq (bntseq.c:166-254) for N-free contigs
  write_index         .bwt .rbwt .sa .rsa (+ the above) from a text with the numpy or the torch builder
  make_alt_contigs    SURVEY §8d ALT set: 2 kbp windows of the primary with a SNP every 300 bp and alternately a
                      5-bp insertion (1000M5I1000M) or a 7-bp deletion (1000M7D993M), plus the `.remap` entries
                      (bwaremap.cpp:42-132)
  synth_pairs         SURVEY §8d paired model: insert ~ N(400, 40), mate reverse-complemented, ends swapped 50 %
"""
from __future__ import annotations

import os

import numpy as np

from . import fmbuild
from .bwtio import bwt_dump_bwt, bwt_dump_sa

_NT = np.frombuffer(b"ACGT", dtype=np.uint8)


# ------------------------------------------------------------------ files ----

def pac_bytes(text) -> bytes:
    """The `.pac` payload (bntseq.c:223-246): 2 bits per base, first base in the two highest bits of a byte; then a
    zero byte when the length is a multiple of four, then one byte holding length % 4."""
    try:
        import torch
        if isinstance(text, torch.Tensor):
            n = text.numel()
            pad = (-n) % 4
            t = torch.cat([text, torch.zeros(pad, dtype=torch.uint8, device=text.device)]) if pad else text
            q = t.reshape(-1, 4)
            packed = ((q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]).to(torch.uint8).cpu().numpy()
            tail = (b"\0" if n % 4 == 0 else b"") + bytes([n % 4])
            return packed.tobytes() + tail
    except ImportError:
        pass
    t = np.ascontiguousarray(text, dtype=np.uint8)
    n = len(t)
    pad = (-n) % 4
    if pad:
        t = np.concatenate([t, np.zeros(pad, np.uint8)])
    q = t.reshape(-1, 4)
    packed = (q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]
    return packed.astype(np.uint8).tobytes() + (b"\0" if n % 4 == 0 else b"") + bytes([n % 4])


def write_pac_ann_amb(prefix: str, text, names, lens) -> None:
    """`.pac`, `.ann`, `.amb` of N-free contigs `names` / `lens` laid end to end in `text` (bns_dump, bntseq.c:60-93;
    seed 11, no FASTA comment -> annotation "(null)", no holes)."""
    total = int(sum(lens))
    assert total == (text.numel() if hasattr(text, "numel") else len(text))
    with open(prefix + ".pac", "wb") as f:
        f.write(pac_bytes(text))
    with open(prefix + ".ann", "w") as f:
        f.write(f"{total} {len(names)} 11\n")
        off = 0
        for nm, ln in zip(names, lens):
            f.write(f"0 {nm} (null)\n{off} {int(ln)} 0\n")
            off += int(ln)
    with open(prefix + ".amb", "w") as f:
        f.write(f"{total} {len(names)} 0\n")


def write_fasta_contigs(path: str, text: np.ndarray, names, lens) -> None:
    with open(path, "wb") as f:
        off = 0
        for nm, ln in zip(names, lens):
            f.write(b">" + nm.encode() + b"\n")
            seq = _NT[text[off:off + ln]]
            for s in range(0, ln, 100):
                f.write(seq[s:s + 100].tobytes() + b"\n")
            off += ln


def write_index(prefix: str, text, names, lens, sa_intv: int = 32, with_sa: bool = True):
    """All files `aln`, `samse` and `sampe` open for `prefix` (what `ibwa index` leaves, minus `.rpac`).
    text: numpy uint8 (prefix-doubling builder, any text) or a torch tensor (31-mer radix builder, near-random
    texts: bench sizes).  Returns (bwt, rbwt)."""
    is_torch = hasattr(text, "numel")
    if is_torch:
        import torch
        rev = torch.flip(text, dims=[0])
        if with_sa:
            bwt, sa = fmbuild.build_bwt_torch(text, sa_intv=sa_intv)
            torch.cuda.empty_cache() if text.is_cuda else None
            rbwt, rsa = fmbuild.build_bwt_torch(rev, sa_intv=sa_intv)
        else:
            bwt, rbwt = fmbuild.build_bwt_torch(text), fmbuild.build_bwt_torch(rev)
        del rev
        torch.cuda.empty_cache() if text.is_cuda else None
    else:
        t = np.ascontiguousarray(text, dtype=np.uint8)
        if with_sa:
            bwt, sa = fmbuild.build_bwt_sa_numpy(t, sa_intv)
            rbwt, rsa = fmbuild.build_bwt_sa_numpy(np.ascontiguousarray(t[::-1]), sa_intv)
        else:
            bwt, rbwt = fmbuild.build_index_numpy(t)
    bwt_dump_bwt(prefix + ".bwt", bwt)
    bwt_dump_bwt(prefix + ".rbwt", rbwt)
    if with_sa:
        bwt_dump_sa(prefix + ".sa", sa)
        bwt_dump_sa(prefix + ".rsa", rsa)
    write_pac_ann_amb(prefix, text, names, lens)
    return bwt, rbwt


# ------------------------------------------------------------ ALT contigs ----

def make_alt_contigs(fetch, contig_names, contig_lens, n_alt: int = 200, seed: int = 20260105):
    """SURVEY §8d ALT set.  fetch(lo, hi) -> numpy uint8 codes of the primary text [lo, hi).
    Returns (alt_text uint8, alt_names, alt_lens, remap_text): alt j is a 2000-bp window of a primary contig with a
    SNP every 300 bp and, alternately, 5 inserted bases after its first 1000 (cigar 1000M5I1000M, as the remap file
    reads it: alt -> primary) or 7 primary bases skipped there (1000M7D993M)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    starts = np.concatenate([[0], np.cumsum(contig_lens)[:-1]]).astype(np.int64)
    seqs, names, lens, remap = [], [], [], []
    for j in range(n_alt):
        ci = int(rng.integers(0, len(contig_names)))
        s = int(rng.integers(0, int(contig_lens[ci]) - 2100))
        win = fetch(int(starts[ci]) + s, int(starts[ci]) + s + 2000).copy()
        snp = np.arange(150, 2000, 300)
        win[snp] = (win[snp] + 1) & 3
        if j % 2 == 0:
            a = np.concatenate([win[:1000], rng.integers(0, 4, size=5, dtype=np.uint8), win[1000:]])
            cigar = "1000M5I1000M"
        else:
            a = np.concatenate([win[:1000], win[1007:]])
            cigar = "1000M7D993M"
        seqs.append(a)
        names.append(f"alt{j}")
        lens.append(len(a))
        remap.append(f">alt{j}-{contig_names[ci]}|{s + 1}|{s + 2000}\n{cigar}\n")   # 1-based inclusive
    return np.concatenate(seqs), names, lens, "".join(remap)


# ------------------------------------------------------------------ pairs ----

def _reads_at(text, start, length: int, g):
    """Default error model for forward-strand reads starting at `start` (torch int64 [m])."""
    import torch
    dev = text.device
    m = start.numel()
    col = torch.arange(length, device=dev)[None, :]
    has = torch.rand(m, device=dev, generator=g) < 0.02
    il = torch.randint(1, 4, (m,), device=dev, generator=g)
    p = torch.randint(10, max(11, length - 10), (m,), device=dev, generator=g)
    isdel = torch.rand(m, device=dev, generator=g) < 0.5
    tail = col >= p[:, None]
    dshift = torch.where(has & isdel, il, torch.zeros_like(il))[:, None] * tail
    ins_len = torch.where(has & ~isdel, il, torch.zeros_like(il))
    ishift = ins_len[:, None] * (col >= (p + il)[:, None])
    idx = (start[:, None] + col + dshift - ishift).clamp_(0, text.numel() - 1)
    r = text[idx]
    insm = (has & ~isdel)[:, None] & tail & (col < (p + il)[:, None])
    rnd = torch.randint(0, 4, (m, length), dtype=torch.uint8, device=dev, generator=g)
    r = torch.where(insm, rnd, r)
    sub = torch.rand((m, length), device=dev, generator=g) < 0.01
    inc = torch.randint(1, 4, (m, length), dtype=torch.uint8, device=dev, generator=g)
    return torch.where(sub, (r + inc) & 3, r)


def synth_pairs(text, n_pairs: int, length: int, seed: int, alt_text=None, alt_lens=None, alt_frac: float = 0.0):
    """Paired reads (torch uint8 [n_pairs, length] x 2): fragment length ~ N(400, 40) clipped to [220, 600], read 1 =
    left end forward, read 2 = right end reverse-complemented, the two swapped for half of the pairs.  With
    alt_text (the ALT contigs laid end to end, lengths alt_lens), a fraction alt_frac of the pairs is drawn from
    inside one ALT contig instead of the primary `text`."""
    import torch
    dev = text.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    r1 = torch.empty((n_pairs, length), dtype=torch.uint8, device=dev)
    r2 = torch.empty((n_pairs, length), dtype=torch.uint8, device=dev)
    n = text.numel()
    if alt_text is not None:
        a_lens = torch.as_tensor(np.asarray(alt_lens), dtype=torch.int64, device=dev)
        a_off = torch.cumsum(a_lens, 0) - a_lens
    CH = 1 << 20
    for s in range(0, n_pairs, CH):
        m = min(CH, n_pairs - s)
        isz = (400 + 40 * torch.randn(m, device=dev, generator=g)).long().clamp_(220, 600)
        st = (torch.rand(m, device=dev, generator=g, dtype=torch.float64) * (n - 700)).long()
        left = _reads_at(text, st, length, g)
        right = _reads_at(text, st + isz - length, length, g)
        if alt_text is not None and alt_frac > 0:
            from_alt = torch.rand(m, device=dev, generator=g) < alt_frac
            cj = torch.randint(0, a_lens.numel(), (m,), device=dev, generator=g)
            room = (a_lens[cj] - isz - 1).clamp_min(1)
            ast = a_off[cj] + (torch.rand(m, device=dev, generator=g, dtype=torch.float64) * room).long()
            aleft = _reads_at(alt_text, ast, length, g)
            aright = _reads_at(alt_text, ast + isz - length, length, g)
            left = torch.where(from_alt[:, None], aleft, left)
            right = torch.where(from_alt[:, None], aright, right)
        right = 3 - torch.flip(right, dims=[1])
        swap = torch.rand(m, device=dev, generator=g) < 0.5
        r1[s:s + m] = torch.where(swap[:, None], right, left)
        r2[s:s + m] = torch.where(swap[:, None], left, right)
    return r1, r2


def write_fastq_pairs(prefix: str, r1: np.ndarray, r2: np.ndarray):
    """<prefix>_1.fq / <prefix>_2.fq with names p<i>/1, p<i>/2 (bwa_read_seq strips the suffix, bwaseqio.c:23-32)."""
    paths = []
    nt = np.frombuffer(b"ACGTN-", dtype=np.uint8)
    for tag, reads in (("1", r1), ("2", r2)):
        path = f"{prefix}_{tag}.fq"
        n, L = reads.shape
        qual = b"I" * L
        with open(path, "wb") as f:
            for s in range(0, n, 200000):
                blk = nt[reads[s:s + 200000]]
                f.write(b"".join(b"@p%d/%s\n" % (s + i, tag.encode()) + blk[i].tobytes() + b"\n+\n" + qual + b"\n"
                                 for i in range(len(blk))))
        paths.append(path)
    return paths


def md5_file(path: str) -> str:
    import hashlib
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def exists_all(prefix: str, exts=(".bwt", ".rbwt", ".sa", ".rsa", ".pac", ".ann", ".amb")) -> bool:
    return all(os.path.exists(prefix + e) for e in exts)
