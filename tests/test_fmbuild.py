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
