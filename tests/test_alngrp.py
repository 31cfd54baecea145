"""Scope row N4: the per-read merge of several .sai streams (alngrp_create, saiset.c:45-78).

The oracle restatement and the product's device code (alngrp_core.cuh, compiled for the CPU by tests/harness)
against golden vectors produced by the reference's own saiset.c (tests/golden/make_alngrp_golden.py), against a
fresh run of that driver where oracle/_ref exists, and the CUDA kernel against the oracle on the GPU."""
import os
import sys
import tempfile

import numpy as np
import pytest

from ibwa_b200 import sai
from harness import pyharness
from oracle import pyoracle

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_alngrp_golden as gold  # noqa: E402


def load(ns, golden_dir):
    z = np.load(os.path.join(golden_dir, f"alngrp_{ns}.npz"))
    n_alns = [z[f"n_aln{s}"] for s in range(ns)]
    recs = [z[f"rec{s}"].view(sai.ALN_DTYPE) for s in range(ns)]
    return n_alns, recs, int(z["s_mm"]), z["out_n"], z["out_rec"].view(sai.ALN_DTYPE), z["out_db"]


def packed(out_off, out_n, rec, db):
    """(out_off, out_n, slots) -> groups concatenated in read order, the layout the reference driver dumps"""
    idx = np.concatenate([np.arange(o, o + c) for o, c in zip(out_off, out_n)]) if len(out_n) else np.empty(0, np.int64)
    idx = idx.astype(np.int64)
    return rec[idx], db[idx]


def same(got, want_n, want_rec, want_db):
    out_off, out_n, rec, db = got
    assert np.array_equal(out_n, want_n)
    r, d = packed(out_off, out_n, rec, db)
    assert r.tobytes() == want_rec.tobytes() and np.array_equal(d, want_db)


@pytest.mark.parametrize("ns", [1, 2, 3])
def test_oracle_matches_reference_golden(ns, golden_dir):
    n_alns, recs, s_mm, want_n, want_rec, want_db = load(ns, golden_dir)
    same(pyoracle.alngrp_merge(n_alns, recs, s_mm), want_n, want_rec, want_db)


@pytest.mark.parametrize("ns", [1, 2, 3])
def test_device_code_on_cpu_matches_reference_golden(ns, golden_dir):
    n_alns, recs, s_mm, want_n, want_rec, want_db = load(ns, golden_dir)
    same(pyharness.alngrp_merge(n_alns, recs, s_mm), want_n, want_rec, want_db)


@pytest.mark.skipif(not os.path.exists(gold.DUMP), reason="oracle/_ref/alngrp_dump not built (needs /root/reference)")
@pytest.mark.parametrize("ns,seed,s_mm", [(2, 11, 3), (4, 12, 5), (3, 13, 10 ** 6)])
def test_oracle_matches_fresh_reference_run(ns, seed, s_mm):
    n_alns, recs = gold.make_streams(ns, 60, seed)
    with tempfile.TemporaryDirectory() as tmp:
        want_n, want_rec, want_db = gold.run_reference(n_alns, recs, tmp, s_mm)
    same(pyoracle.alngrp_merge(n_alns, recs, s_mm), want_n, want_rec, want_db)
    same(pyharness.alngrp_merge(n_alns, recs, s_mm), want_n, want_rec, want_db)


def adversarial(n, kind, rng):
    if kind == "equal":
        return np.zeros(n, np.int32)
    if kind == "sorted":
        return np.arange(n, dtype=np.int32)
    if kind == "reversed":
        return np.arange(n, dtype=np.int32)[::-1].copy()
    if kind == "two":
        return rng.integers(0, 2, n).astype(np.int32)
    if kind == "organ":
        h = np.arange(n // 2, dtype=np.int32)
        return np.concatenate([h, h[::-1], np.zeros(n - 2 * (n // 2), np.int32)])
    return rng.integers(0, 1 << 20, n).astype(np.int32)


def one_group(scores, ns):
    """a single read whose group is `scores`, dealt round-robin... no: in runs, over ns streams"""
    cuts = np.linspace(0, len(scores), ns + 1).astype(int)
    n_alns, recs = [], []
    for s in range(ns):
        seg = scores[cuts[s]:cuts[s + 1]]
        r = np.zeros(len(seg), sai.ALN_DTYPE)
        r["score"] = seg
        r["k"] = np.arange(len(seg), dtype=np.uint32) + 10_000_000 * s
        n_alns.append(np.array([len(seg)], np.int32))
        recs.append(r)
    return n_alns, recs


@pytest.mark.parametrize("kind", ["equal", "sorted", "reversed", "two", "organ", "random"])
@pytest.mark.parametrize("n", [3, 16, 17, 18, 33, 1000, 50000])
def test_device_code_equals_oracle_on_hard_orders(kind, n):
    rng = np.random.default_rng(n)
    n_alns, recs = one_group(adversarial(n, kind, rng), 2)
    want = pyoracle.alngrp_merge(n_alns, recs, 1 << 30)
    got = pyharness.alngrp_merge(n_alns, recs, 1 << 30)
    assert np.array_equal(got[1], want[1]) and got[2].tobytes() == want[2].tobytes() and np.array_equal(got[3], want[3])
    sc = got[2]["score"]
    assert (np.diff(sc) >= 0).all() and np.array_equal(np.sort(sc), np.sort(np.concatenate([r["score"] for r in recs])))


@pytest.mark.gpu
def test_gpu_merge_matches_oracle_and_golden(golden_dir, g1_index):
    from ibwa_b200 import engine
    with engine.Engine(g1_index[0], g1_index[1], 0) as e:
        for ns in (1, 2, 3):
            n_alns, recs, s_mm, want_n, want_rec, want_db = load(ns, golden_dir)
            same(e.alngrp_merge(n_alns, recs, s_mm), want_n, want_rec, want_db)
        rng = np.random.default_rng(5)
        for kind in ("equal", "two", "organ", "random", "reversed"):
            n_alns, recs = one_group(adversarial(20000, kind, rng), 3)
            want = pyoracle.alngrp_merge(n_alns, recs, 7)
            got = e.alngrp_merge(n_alns, recs, 7)
            assert np.array_equal(got[1], want[1])
            a, b = packed(*got), packed(*want)
            assert a[0].tobytes() == b[0].tobytes() and np.array_equal(a[1], b[1])
        # many reads, ragged, with an empty stream
        n_alns, recs = gold.make_streams(4, 5000, 77)
        n_alns[2][:] = 0
        recs[2] = recs[2][:0]
        want = pyoracle.alngrp_merge(n_alns, recs, 3)
        got = e.alngrp_merge(n_alns, recs, 3)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        a, b = packed(*got), packed(*want)
        assert a[0].tobytes() == b[0].tobytes() and np.array_equal(a[1], b[1])


@pytest.mark.gpu
def test_gpu_merge_of_primary_and_alt_sai_against_reference_saiset(tmp_path):
    """BASELINE config 5 shape: the same reads aligned against a primary index and an index of alt contigs cut
    from it, the two .sai streams merged per read — engine (`b200aln_alngrp_merge`) against the reference's own
    saiset.c reading the same two files (oracle/_ref/alngrp_dump)."""
    from ibwa_b200 import engine, fmbuild, gap_init_opt, synth
    if not os.path.exists(gold.DUMP):
        pytest.skip("oracle/_ref/alngrp_dump not present")
    rng = np.random.default_rng(5)
    pri = synth.random_genome(300_000, 20260105)
    alts = []
    for j in range(150):
        s0 = int(rng.integers(0, len(pri) - 2100))
        a = pri[s0:s0 + 2000].copy()
        snp = np.arange(150, 2000, 300)
        a[snp] = (a[snp] + 1) & 3
        alts.append(a)
    alt = np.concatenate(alts)
    reads = np.concatenate([synth.simulate_reads_fast(pri, 3000, 100, 1), synth.simulate_reads_fast(alt, 3000, 100, 2)])
    n, L = reads.shape
    lens = np.full(n, L, np.int32)
    offs = np.arange(n, dtype=np.int64) * L
    opt = gap_init_opt()
    n_alns, recs, paths = [], [], []
    for tag, g in (("pri", pri), ("alt", alt)):
        bwt, _ = fmbuild.build_bwt_sa_numpy(g)
        rbwt, _ = fmbuild.build_bwt_sa_numpy(np.ascontiguousarray(g[::-1]))
        with engine.Engine(bwt, rbwt, 0) as e:
            na, rc = e.cal_sa_reg_gap(lens, offs, reads.reshape(-1), opt)
        n_alns.append(na)
        recs.append(rc)
        p = str(tmp_path / f"{tag}.sai")
        with open(p, "wb") as f:
            sai.write_header(f, opt)
            sai.write_batch(f, na, rc)
        paths.append(p)
    import subprocess
    out = subprocess.run([gold.DUMP, str(n)] + paths, stdout=subprocess.PIPE, check=True).stdout
    w = np.frombuffer(out, dtype=np.uint32)
    want_n = np.zeros(n, np.int32)
    want_db, want_rec, p = [], [], 0
    for r in range(n):
        c = int(w[p]); p += 1
        want_n[r] = c
        blk = w[p:p + 5 * c].reshape(c, 5); p += 5 * c
        want_db.append(blk[:, 0]); want_rec.append(blk[:, 1:])
    want_db = np.concatenate(want_db)
    want_rec = np.concatenate(want_rec).reshape(-1).view(sai.ALN_DTYPE)
    with engine.Engine(bwt, rbwt, 0) as e:
        got = e.alngrp_merge(n_alns, recs, opt.s_mm)
    same(got, want_n, want_rec, want_db)
    assert ((n_alns[0] > 0) & (n_alns[1] > 0)).mean() > 0.3      # many reads hit both indexes
    assert len(np.unique(want_db)) == 2
