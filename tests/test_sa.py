"""Scope row N2 — SA row -> text position (bwt_sa, bwt.c:69-79).  Oracle restatement against a naive
suffix array and the reference's .sa files; the device code on the CPU harness; the kernel on the GPU."""
import os

import numpy as np
import pytest

from ibwa_b200 import bwt_restore_sa, fmbuild, synth
from oracle import pyoracle


@pytest.fixture(scope="module")
def small():
    t = synth.repeat_rich_genome(40_000, 77)
    bwt, sa = fmbuild.build_bwt_sa_numpy(t)
    rbwt, rsa = fmbuild.build_bwt_sa_numpy(np.ascontiguousarray(t[::-1]))
    full = np.concatenate([[len(t)], fmbuild._suffix_array_numpy(t)]).astype(np.uint64)
    return t, bwt, sa, rbwt, rsa, full


def test_oracle_bwt_sa_equals_naive_suffix_array(small):
    t, bwt, sa, _, _, full = small
    rows = np.concatenate([np.arange(0, 200), np.arange(len(t) - 200, len(t) + 1),
                           np.random.default_rng(1).integers(0, len(t) + 1, size=3000), [bwt.primary, bwt.primary - 1]])
    got = pyoracle.bwt_sa(pyoracle.as_orc_bwt(bwt), pyoracle.as_orc_sa(sa), rows)
    want = full[rows].astype(np.uint32)
    want[rows == 0] = 0xFFFFFFFF            # sa[0] is stored as -1 (bwt.c:66)
    assert np.array_equal(got, want)


def test_sa_samples_equal_reference_files(tmp_path):
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref/ibwa not present")
    t = synth.random_genome(100_003, 5)
    fa = str(tmp_path / "x.fa")
    synth.write_fasta(fa, t)
    pyoracle.run_ref(["index", "-a", "is", fa])
    _, sa = fmbuild.build_bwt_sa_numpy(t)
    _, rsa = fmbuild.build_bwt_sa_numpy(np.ascontiguousarray(t[::-1]))
    ref, refr = bwt_restore_sa(fa + ".sa"), bwt_restore_sa(fa + ".rsa")
    assert ref.sa_intv == 32 and np.array_equal(ref.sa, sa.sa) and np.array_equal(refr.sa, rsa.sa)
    assert ref.primary == sa.primary and ref.seq_len == len(t)


def test_device_code_on_cpu_harness(small):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from harness import pyharness
    t, bwt, sa, _, _, full = small
    rows = np.random.default_rng(2).integers(0, len(t) + 1, size=5000)
    rows = np.concatenate([rows, [0, 1, 63, 64, 65, len(t), bwt.primary, bwt.primary + 1]])
    got = pyharness.bwt_sa(bwt, sa, rows)
    want = pyoracle.bwt_sa(pyoracle.as_orc_bwt(bwt), pyoracle.as_orc_sa(sa), rows)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_gpu_bwt_sa_and_sa2seq():
    from ibwa_b200 import engine
    t = synth.random_genome(1_000_000, 3)
    bwt, sa = fmbuild.build_bwt_sa_numpy(t)
    rbwt, rsa = fmbuild.build_bwt_sa_numpy(np.ascontiguousarray(t[::-1]))
    rng = np.random.default_rng(4)
    n = 200_000
    rows = rng.integers(0, len(t) + 1, size=n).astype(np.uint32)
    ob, osa = pyoracle.as_orc_bwt(bwt), pyoracle.as_orc_sa(sa)
    orb, orsa = pyoracle.as_orc_bwt(rbwt), pyoracle.as_orc_sa(rsa)
    with engine.Engine(bwt, rbwt, 0) as e:
        e.load_sa(0, sa)
        e.load_sa(1, rsa)
        got0, got1 = e.bwt_sa(0, rows), e.bwt_sa(1, rows)
        strand = rng.integers(0, 2, size=n).astype(np.uint8)
        lens = rng.integers(20, 150, size=n).astype(np.int32)
        pos = e.sa2seq(strand, rows, lens)
    sub = slice(0, 20000)
    assert np.array_equal(got0[sub], pyoracle.bwt_sa(ob, osa, rows[sub]))
    assert np.array_equal(got1[sub], pyoracle.bwt_sa(orb, orsa, rows[sub]))
    # whole batch against the suffix array itself
    full = np.concatenate([[len(t)], fmbuild._suffix_array_numpy(t)]).astype(np.uint64)
    want = full[rows].astype(np.uint32)
    want[rows == 0] = 0xFFFFFFFF
    assert np.array_equal(got0, want)
    want_pos = np.where(strand != 0, got0.astype(np.uint64),
                        ((np.uint64(len(t)) - (got1.astype(np.uint64) + lens.astype(np.uint64))) & np.uint64(0xFFFFFFFF)))
    assert np.array_equal(pos, want_pos)


@pytest.mark.gpu
def test_full_pipeline_property_at_scale():
    """Size-independent property at a larger size (no oracle): reads cut from a 50 Mbp genome without
    errors must come back from aln with a score-0 record whose SA interval, mapped through row N2
    (bwt_sa / sa2seq), contains the position they were cut from — on the right strand."""
    import torch
    from ibwa_b200 import engine, gap_init_opt, sai
    n_g, n, L = 50_000_000, 200_000, 100
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(123)
    text = torch.randint(0, 4, (n_g,), dtype=torch.uint8, device=dev, generator=g)
    bwt, sa = fmbuild.build_bwt_torch(text, sa_intv=32)
    rbwt, rsa = fmbuild.build_bwt_torch(torch.flip(text, dims=[0]), sa_intv=32)
    start = torch.randint(0, n_g - L, (n,), device=dev, generator=g)
    reads = text[start[:, None] + torch.arange(L, device=dev)[None, :]]
    rc = torch.rand(n, device=dev, generator=g) < 0.5
    reads = torch.where(rc[:, None], 3 - torch.flip(reads, dims=[1]), reads)
    reads_h, start_h, rc_h = reads.cpu().numpy(), start.cpu().numpy(), rc.cpu().numpy()
    del text, reads, start
    torch.cuda.empty_cache()
    lens = np.full(n, L, np.int32)
    offs = np.arange(n, dtype=np.int64) * L
    with engine.Engine(bwt, rbwt, 0) as e:
        n_aln, rec = e.cal_sa_reg_gap(lens, offs, reads_h.reshape(-1), gap_init_opt())
        e.load_sa(0, sa)
        e.load_sa(1, rsa)
        assert (n_aln >= 1).all()
        first = np.concatenate([[0], np.cumsum(n_aln)[:-1]])
        u = sai.unpack(rec)
        # records of a read come out in non-decreasing score order (bwase.c:39-42 relies on it)
        read_of = np.repeat(np.arange(n), n_aln)
        same_read = read_of[1:] == read_of[:-1]
        assert (u["score"][1:][same_read] >= u["score"][:-1][same_read]).all()
        best = rec[first]
        bu = sai.unpack(best)
        assert (bu["score"] == 0).all() and (bu["n_mm"] == 0).all() and (bu["n_gapo"] == 0).all()
        assert np.array_equal(bu["a"].astype(bool), rc_h)        # strand 1 <=> the read was reverse-complemented
        width = bu["l"].astype(np.int64) - bu["k"].astype(np.int64) + 1
        assert (width >= 1).all() and (width == 1).mean() > 0.999     # 100-mers of a random 50 Mbp text are unique
        uniq = width == 1
        pos = e.sa2seq(bu["a"][uniq].astype(np.uint8), bu["k"][uniq], lens[uniq])
    assert np.array_equal(pos.astype(np.int64), start_h[uniq])
