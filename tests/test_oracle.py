"""Pins the CPU restatement (oracle/alnoracle.c) to the unmodified reference:
byte-for-byte against .sai files the reference binary produced (tests/golden)."""
import io
import os

import numpy as np
import pytest

from ibwa_b200 import parse_aln_args, sai, seqio
from ibwa_b200.opts import bwa_cal_maxdiff
from oracle import pyoracle
from cases import CASES


def oracle_sai_bytes(bwt, rbwt, args, fq):
    opt, _, _, _ = parse_aln_args(args + ["prefix", fq])
    ob, orb = pyoracle.as_orc_bwt(bwt), pyoracle.as_orc_bwt(rbwt)
    buf = io.BytesIO()
    sai.write_header(buf, opt)
    stats = None
    for batch in seqio.read_batches(fq, opt.mode, opt.trim_qual):
        n_aln, rec, stats = pyoracle.aln_batch(ob, orb, batch.lens, batch.offs, batch.codes, opt.to_c())
        sai.write_batch(buf, n_aln, rec)
    return buf.getvalue(), stats


@pytest.mark.parametrize("tag", sorted(CASES))
def test_oracle_matches_reference_sai(tag, golden_dir, g1_index):
    args, fq = CASES[tag]
    got, stats = oracle_sai_bytes(g1_index[0], g1_index[1], args, os.path.join(golden_dir, fq + ".fq.gz"))
    want = open(os.path.join(golden_dir, f"g1_{tag}.sai"), "rb").read()
    assert len(got) == len(want)
    assert got == want
    # the "last_diff_pos is inherited from the parent" restatement
    assert stats["inherit_violations"] == 0


def test_fixture_exercises_hard_paths(golden_dir):
    """The golden set must reach the paths a random genome does not."""
    _, n_aln, rec = sai.read_sai(os.path.join(golden_dir, "g1_default.sai"))
    assert n_aln.max() >= 8                       # repeats: many records per read
    assert (n_aln == 0).sum() > 0                 # unalignable reads
    u = sai.unpack(rec)
    assert (u["n_gapo"] > 0).any() and (u["n_mm"] > 0).any() and (u["a"] == 1).any() and (u["a"] == 0).any()
    d = open(os.path.join(golden_dir, "g1_default.sai"), "rb").read()
    m = open(os.path.join(golden_dir, "g1_m200.sai"), "rb").read()
    assert d[64:] != m[64:]                       # -m 200: max_entries cutoff fires


def test_maxdiff_table(golden_dir):
    want = {}
    for line in open(os.path.join(golden_dir, "maxdiff_table.txt")):
        parts = line.split()
        want[int(parts[1][:-2])] = int(parts[-1])
    got = {}
    k = 0
    for i in range(17, 251):
        v = bwa_cal_maxdiff(i, 0.02, 0.04)
        assert v == pyoracle.lib().orc_cal_maxdiff(i, 0.02, 0.04)
        if v != k:
            got[i] = v
        k = v
    assert got == want


def test_occ_against_naive(g1_index):
    bwt = g1_index[0]
    ob = pyoracle.as_orc_bwt(bwt)
    n = bwt.seq_len
    # decode the BWT string from the on-disk layout
    words = bwt.bwt
    nblk = (n + 127) // 128
    sym = np.empty(nblk * 128, dtype=np.uint8)
    for b in range(nblk):
        w = words[b * 12 + 4: b * 12 + 12]
        w = np.pad(w, (0, 8 - len(w)))
        shifts = (30 - 2 * np.arange(16)).astype(np.uint32)
        sym[b * 128:(b + 1) * 128] = ((w[:, None] >> shifts[None, :]) & 3).reshape(-1)
    sym = sym[:n]
    cum = np.zeros((4, n + 1), dtype=np.int64)
    for c in range(4):
        cum[c, 1:] = np.cumsum(sym == c)
    rng = np.random.default_rng(5)
    ks = list(rng.integers(0, n + 1, size=300)) + [0, 1, n, n - 1, bwt.primary, bwt.primary - 1, bwt.primary + 1,
                                                      127, 128, 129, 0xFFFFFFFF]
    for k in ks:
        k = int(k)
        if k == 0xFFFFFFFF:
            expect = [0, 0, 0, 0]
        else:
            kk = k - 1 if k >= bwt.primary else k      # row -> position in the $-less string
            expect = [int(cum[c, kk + 1]) for c in range(4)]
        assert list(pyoracle.occ4(ob, k)) == expect
        for c in range(4):
            assert pyoracle.occ(ob, k, c) == expect[c]
