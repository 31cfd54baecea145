"""The synthetic-index builders reproduce the reference's .bwt/.rbwt bytes."""
import gzip
import os

import numpy as np
import pytest

from ibwa_b200 import bwt_restore_bwt, fmbuild, seqio, synth
from ibwa_b200.bwtio import expected_words
from oracle import pyoracle


def golden_text(golden_dir):
    txt = gzip.open(os.path.join(golden_dir, "g1.fa.gz")).read().split(b"\n", 1)[1].replace(b"\n", b"")
    return seqio.NT4[np.frombuffer(txt, dtype=np.uint8)]


def same(a, b):
    return a.primary == b.primary and np.array_equal(a.L2, b.L2) and np.array_equal(a.bwt, b.bwt)


def test_numpy_builder_matches_reference_index(golden_dir, g1_index):
    t = golden_text(golden_dir)
    b, rb = fmbuild.build_index_numpy(t)
    assert same(b, g1_index[0]) and same(rb, g1_index[1])
    assert b.bwt_size == expected_words(len(t))


@pytest.mark.parametrize("n", [1000, 4096 * 128, 300001])
def test_torch_builder_matches_numpy(n):
    import torch
    t = synth.random_genome(n, 11 + n)
    t[-40:] = 0                       # an A-run into the end of the text: short suffixes tie with padding
    if n > 5000:
        t[2000:2050] = t[100:150]     # a 50-mer repeat: equal 31-mer keys
    a = fmbuild.build_bwt_numpy(t)
    b = fmbuild.build_bwt_torch(torch.from_numpy(t), chunk=1 << 16)
    assert same(a, b)


def test_builders_against_fresh_reference_run(tmp_path):
    """Where the reference binary exists: `ibwa index` on a fresh 1 Mbp random text."""
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref/ibwa not present")
    import torch
    t = synth.random_genome(1_000_003, 3)
    fa = str(tmp_path / "x.fa")
    synth.write_fasta(fa, t)
    pyoracle.run_ref(["index", "-a", "is", fa])
    ref_b, ref_rb = bwt_restore_bwt(fa + ".bwt"), bwt_restore_bwt(fa + ".rbwt")
    b, rb = fmbuild.build_index_torch(torch.from_numpy(t))
    assert same(b, ref_b) and same(rb, ref_rb)


def _plant_families(t, rng, n_fam, copies=40, unit=300, div=0.02):
    """40-copy families of diverged 300-bp units at random places of t"""
    for _ in range(n_fam):
        u = rng.integers(0, 4, size=unit, dtype=np.uint8)
        for _ in range(copies):
            cp = u.copy()
            mut = rng.random(unit) < div
            cp[mut] = (cp[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
            at = int(rng.integers(0, len(t) - unit))
            t[at:at + unit] = cp
    return t


@pytest.mark.parametrize("n,n_fam", [(400_000, 6), (150_000, 12)])
def test_torch_builder_on_repeat_rich_text(n, n_fam):
    """Repeat families, (AC)n, A x 400, (GATTACA)n and exact 2 kbp duplications: ties among 31-mer keys by the
    thousand, resolved on the device round by round (fmbuild._fix_ties); bytes equal to the general builder's."""
    import torch
    rng = np.random.default_rng(n)
    t = _plant_families(synth.repeat_rich_genome(n, 5 + n), rng, n_fam)
    t[5000:7000] = t[90_000:92_000]            # an exact 2 kbp duplicate: 65 rounds deep
    t[-300:] = t[20_000:20_300]                # a repeat running into the end of the text (host path)
    a, sa_a = fmbuild.build_bwt_sa_numpy(t, 32)
    b, sa_b = fmbuild.build_bwt_torch(torch.from_numpy(t), chunk=1 << 16, sa_intv=32)
    assert same(a, b) and np.array_equal(sa_a.sa, sa_b.sa)
    ra = fmbuild.build_bwt_numpy(np.ascontiguousarray(t[::-1]))
    rb = fmbuild.build_bwt_torch(torch.from_numpy(np.ascontiguousarray(t[::-1])), chunk=1 << 16)
    assert same(ra, rb)
