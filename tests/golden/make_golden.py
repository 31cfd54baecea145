#!/usr/bin/env python
"""Regenerates the golden fixtures in this directory by running the UNMODIFIED
reference binary (oracle/_ref/ibwa, built from /root/reference by
oracle/Makefile.ref).  Run in the build container only; the outputs are
committed so the GPU box (which has no /root/reference) can use them.

  python tests/golden/make_golden.py

Fixtures (SURVEY.md §4 F1-F4 at reduced size):
  g1.fa.gz, g1.bwt, g1.rbwt ........ 120 kbp repeat-rich reference and its index
  g1_reads.fq.gz ................... mixed-length reads, N, indels, lower case, odd reads
  g1_short.fq.gz ................... reads < 38 bp (batch-level max_gapo clamp, bwtaln.c:91-92)
  g1_<tag>.sai ..................... reference `aln` output for the option sets in CASES
  maxdiff_table.txt ................ bwa_cal_maxdiff(17..250) as printed by bwtaln.c:317-324
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from ibwa_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "ibwa")

CASES = {
    "default": ([], "g1_reads"),
    "m200": (["-m", "200"], "g1_reads"),
    "N_n2": (["-N", "-n", "2"], "g1_reads"),
    "R2": (["-R", "2"], "g1_reads"),
    "o3": (["-o", "3"], "g1_reads"),
    "q20": (["-q", "20"], "g1_reads"),
    "L_e3": (["-L", "-o", "2", "-e", "3"], "g1_reads"),
    "stress": (["-n", "4", "-o", "2", "-e", "10", "-l", "32", "-k", "2"], "g1_reads"),
    "n001": (["-n", "0.01"], "g1_reads"),
    "MOE": (["-M", "2", "-O", "5", "-E", "2"], "g1_reads"),
    "i0_d3": (["-i", "0", "-d", "3", "-o", "2"], "g1_reads"),
    "l20_k1": (["-l", "20", "-k", "1"], "g1_reads"),
    "c": (["-c"], "g1_reads"),
    "short_o3": (["-o", "3"], "g1_short"),
    "short_default": ([], "g1_short"),
}


def make_reads(g, rng):
    reads, quals = [], []
    lens = [20, 32, 33, 36, 50, 75, 100, 125, 150, 200, 250]
    n = len(g)
    for i in range(1300):
        L = lens[i % len(lens)]
        if i % 3 == 0:   # bias towards the repeat-rich regions
            start = int(rng.integers(1000, 40000))
        elif i % 7 == 0:
            start = int(n // 2 + rng.integers(-200, 4500))
        else:
            start = int(rng.integers(0, n - L - 8))
        start = min(start, n - L - 8)
        src = g[start:start + L + 8]
        r = src[:L].copy()
        if rng.random() < 0.10 and L >= 30:
            p = int(rng.integers(6, L - 6))
            if rng.random() < 0.5:
                r = np.concatenate([src[:p], rng.integers(0, 4, size=2, dtype=np.uint8), src[p:]])[:L]
            else:
                r = np.concatenate([src[:p], src[p + 2:]])[:L]
        sub = rng.random(L) < 0.012
        r[sub] = (r[sub] + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) & 3
        nm = rng.random(L) < 0.005
        r[nm] = 4
        if rng.random() < 0.5:
            r = synth.revcomp(r)
        reads.append(np.ascontiguousarray(r))
        if i % 4 == 0:   # decaying qualities so that -q trims something
            q = np.clip(40 - (np.arange(L) * 45 // max(L, 1)) + rng.integers(-6, 7, size=L), 2, 40)
        else:
            q = np.full(L, 40)
        quals.append((q + 33).astype(np.uint8))
    odd = [np.full(60, 4, np.uint8), np.zeros(70, np.uint8), np.zeros(1, np.uint8) + 2,
           np.array([0, 1, 2, 3] * 10 + [5] + [0, 1, 2, 3] * 5, np.uint8),
           np.tile(np.array([0, 1], np.uint8), 40), np.tile(np.array([2, 0, 3, 3, 0, 1, 0], np.uint8), 12)]
    for r in odd:
        reads.append(r)
        quals.append(np.full(len(r), 73, np.uint8))
    return reads, quals


def write_fq(path, reads, quals, rng, lower_every=11):
    nt = np.frombuffer(b"ACGTN-", dtype=np.uint8)
    with gzip.open(path, "wb", compresslevel=9) as f:
        for i, (r, q) in enumerate(zip(reads, quals)):
            s = nt[r].tobytes()
            if i % lower_every == 0:
                s = s.lower()
            f.write(b"@r%d/1\n" % i + s + b"\n+\n" + q.tobytes() + b"\n")


def main():
    assert os.path.exists(REF), "build the reference first: make -C oracle ref"
    rng = np.random.Generator(np.random.PCG64(20260118))
    g = synth.repeat_rich_genome(120_000, 20260117)
    tmp = tempfile.mkdtemp(prefix="golden_")
    fa = os.path.join(tmp, "g1.fa")
    synth.write_fasta(fa, g)
    subprocess.check_call([REF, "index", "-a", "is", fa], stderr=subprocess.DEVNULL)
    shutil.copy(fa + ".bwt", os.path.join(HERE, "g1.bwt"))
    shutil.copy(fa + ".rbwt", os.path.join(HERE, "g1.rbwt"))
    with open(fa, "rb") as fi, gzip.open(os.path.join(HERE, "g1.fa.gz"), "wb", compresslevel=9) as fo:
        fo.write(fi.read())

    reads, quals = make_reads(g, rng)
    write_fq(os.path.join(HERE, "g1_reads.fq.gz"), reads, quals, rng)
    short = [(r[:int(rng.integers(18, 38))], q) for r, q in zip(reads[:400], quals[:400])]
    write_fq(os.path.join(HERE, "g1_short.fq.gz"), [r for r, _ in short], [q[:len(r)] for r, q in short], rng)

    for tag, (args, fq) in CASES.items():
        out = os.path.join(HERE, f"g1_{tag}.sai")
        with open(out, "wb") as fo:
            subprocess.check_call([REF, "aln"] + args + [fa, os.path.join(HERE, fq + ".fq.gz")], stdout=fo,
                                  stderr=subprocess.DEVNULL)
        print(tag, os.path.getsize(out))

    # max_diff table printed by the reference at start-up (bwtaln.c:317-324)
    p = subprocess.run([REF, "aln", fa, os.path.join(HERE, "g1_short.fq.gz")], stdout=subprocess.DEVNULL,
                       stderr=subprocess.PIPE)
    lines = [l for l in p.stderr.decode().splitlines() if l.startswith("[bwa_aln] ")]
    with open(os.path.join(HERE, "maxdiff_table.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
