"""The product's per-read state machines (aln_core.cuh), compiled for the CPU by
tests/harness, against the reference's golden .sai files.  This checks the
kernel LOGIC where no GPU exists; the -m gpu tests check the kernels proper."""
import io
import os

import numpy as np
import pytest

from ibwa_b200 import parse_aln_args, sai, seqio
from harness import pyharness
from cases import CASES


def harness_sai(bwt, rbwt, args, fq, **kw):
    opt, _, _, _ = parse_aln_args(args + ["prefix", fq])
    buf = io.BytesIO()
    sai.write_header(buf, opt)
    nov_total = 0
    for batch in seqio.read_batches(fq, opt.mode, opt.trim_qual):
        n_aln, rec, nov, _ = pyharness.aln_batch(bwt, rbwt, batch.lens, batch.offs, batch.codes, opt.to_c(), **kw)
        nov_total += nov
        sai.write_batch(buf, np.maximum(n_aln, 0), rec)
    return buf.getvalue(), nov_total


@pytest.mark.parametrize("tag", sorted(CASES))
def test_core_matches_reference(tag, golden_dir, g1_index):
    args, fq = CASES[tag]
    got, nov = harness_sai(g1_index[0], g1_index[1], args, os.path.join(golden_dir, fq + ".fq.gz"),
                           arena_cap=32000, rec_cap=4096, rounds=1)      # one pruned pop per step, like the fast kernel
    want = open(os.path.join(golden_dir, f"g1_{tag}.sai"), "rb").read()
    assert nov == 0
    assert got == want


@pytest.mark.parametrize("tag", sorted(CASES))
def test_core_wide_heads_variant(tag, golden_dir, g1_index):
    """The large-pass configuration (32-bit heads in memory, big arena) on every case."""
    args, fq = CASES[tag]
    got, nov = harness_sai(g1_index[0], g1_index[1], args, os.path.join(golden_dir, fq + ".fq.gz"),
                           arena_cap=1 << 22, rec_cap=4096, reuse=True)
    assert nov == 0 and got == open(os.path.join(golden_dir, f"g1_{tag}.sai"), "rb").read()


@pytest.mark.parametrize("tag", ["default", "stress", "m200"])
def test_core_free_list_variant(tag, golden_dir, g1_index):
    args, fq = CASES[tag]
    got, nov = harness_sai(g1_index[0], g1_index[1], args, os.path.join(golden_dir, fq + ".fq.gz"),
                           arena_cap=32000, rec_cap=4096, reuse=True)
    want = open(os.path.join(golden_dir, f"g1_{tag}.sai"), "rb").read()
    assert nov == 0 and got == want


def test_small_arena_flags_overflow(golden_dir, g1_index):
    """A read whose stack outgrows the fast arena must be flagged, never truncated."""
    args, fq = CASES["default"]
    opt, _, _, _ = parse_aln_args(args + ["p", "q"])
    batch = next(seqio.read_batches(os.path.join(golden_dir, fq + ".fq.gz"), opt.mode, opt.trim_qual))
    n_aln, rec, nov, _ = pyharness.aln_batch(g1_index[0], g1_index[1], batch.lens, batch.offs, batch.codes,
                                             opt.to_c(), arena_cap=256, rec_cap=4)
    assert nov > 0 and (n_aln < 0).sum() == nov


@pytest.mark.parametrize("tag", ["default", "L_e3", "stress"])
def test_two_pass_flow_is_exact(tag, golden_dir, g1_index):
    """Fast pass with a tiny arena / record slab, flagged reads re-run with rebuilt widths (the
    product's large pass): bytes must equal the reference's."""
    args, fq = CASES[tag]
    got, nov = harness_sai(g1_index[0], g1_index[1], args, os.path.join(golden_dir, fq + ".fq.gz"),
                           arena_cap=64, rec_cap=1, big_cap=1 << 22)
    assert nov > 0
    assert got == open(os.path.join(golden_dir, f"g1_{tag}.sai"), "rb").read()


@pytest.mark.parametrize("lut_k", [1, 3, 6, 9])
@pytest.mark.parametrize("tag", ["default", "stress", "N_n2", "short_o3", "m200", "c"])
def test_interval_table_is_exact(tag, lut_k, golden_dir, g1_index):
    """The path-k-mer interval table replaces occ lookups for the first lut_k levels; bytes must not change."""
    args, fq = CASES[tag]
    got, nov = harness_sai(g1_index[0], g1_index[1], args, os.path.join(golden_dir, fq + ".fq.gz"),
                           arena_cap=32000, rec_cap=4096, lut_k=lut_k)
    assert nov == 0
    assert got == open(os.path.join(golden_dir, f"g1_{tag}.sai"), "rb").read()


def _vs_oracle(bwt, rbwt, reads, args, **kw):
    from oracle import pyoracle
    opt, _, _, _ = parse_aln_args(args + ["p", "q"])
    lens = np.array([len(r) for r in reads], np.int32)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    codes = np.concatenate(reads) if len(reads) and lens.sum() else np.empty(0, np.uint8)
    o_n, o_rec, _ = pyoracle.aln_batch(pyoracle.as_orc_bwt(bwt), pyoracle.as_orc_bwt(rbwt), lens, offs, codes, opt.to_c())
    h_n, h_rec, nov, _ = pyharness.aln_batch(bwt, rbwt, lens, offs, codes, opt.to_c(), **kw)
    assert nov == 0
    assert np.array_equal(o_n, h_n) and o_rec.tobytes() == h_rec.tobytes()
    return o_n


@pytest.mark.parametrize("tag", ["stress", "N_n2", "short_o3", "short_default"])
def test_q16_width_records_on_golden(tag, golden_dir, g1_index):
    """The fast pass's 16-bit width records (aln_core.cuh QF<16>) on the fixtures whose options allow them for the
    whole batch (max_diff < 7, max_seed_diff < 3): bytes must equal the reference's."""
    args, fq = CASES[tag]
    got, nov = harness_sai(g1_index[0], g1_index[1], args, os.path.join(golden_dir, fq + ".fq.gz"),
                           arena_cap=32000, rec_cap=4096, rounds=1, q16=True, lut_k=3)
    assert nov == 0 and pyharness.lib().hh_last_q16() == 1
    assert got == open(os.path.join(golden_dir, f"g1_{tag}.sai"), "rb").read()


@pytest.mark.parametrize("args", [[], ["-o", "3"], ["-L", "-o", "2", "-e", "3"], ["-R", "2"], ["-m", "200"],
                                  ["-l", "20", "-k", "1"], ["-i", "0", "-d", "3", "-o", "2"], ["-c"], ["-n", "6", "-k", "2"]])
def test_q16_width_records_vs_oracle(args, golden_dir, g1_index):
    """Default-style option sets on the repeat-rich fixture's reads of at most 150 bp (so that max_diff stays below
    7 and the 16-bit records are what runs): gap_shadow's edits, the seed fields and the saturating bid fields
    against the oracle, on the bump arena, the free-list arena and through the two-pass flow."""
    opt, _, _, _ = parse_aln_args(args + ["p", "q"])
    batch = next(seqio.read_batches(os.path.join(golden_dir, "g1_reads.fq.gz"), opt.mode, opt.trim_qual))
    reads = [batch.codes[o:o + l] for o, l in zip(batch.offs, batch.lens) if l <= 150][:1500]
    assert len(reads) > 800
    _vs_oracle(g1_index[0], g1_index[1], reads, args, arena_cap=1 << 20, rec_cap=4096, q16=True, rounds=1, lut_k=4)
    assert pyharness.lib().hh_last_q16() == 1
    _vs_oracle(g1_index[0], g1_index[1], reads[:400], args, arena_cap=1 << 20, rec_cap=4096, q16=True, reuse=True)
    from oracle import pyoracle
    lens = np.array([len(r) for r in reads[:300]], np.int32)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    codes = np.concatenate(reads[:300])
    o_n, o_rec, _ = pyoracle.aln_batch(pyoracle.as_orc_bwt(g1_index[0]), pyoracle.as_orc_bwt(g1_index[1]), lens, offs,
                                       codes, opt.to_c())
    h_n, h_rec, nov, _ = pyharness.aln_batch(g1_index[0], g1_index[1], lens, offs, codes, opt.to_c(), arena_cap=48,
                                             rec_cap=2, big_cap=1 << 21, q16=True)
    assert nov > 0 and np.array_equal(o_n, h_n) and o_rec.tobytes() == h_rec.tobytes()


@pytest.mark.parametrize("tag,kw", [("default", dict(arena_cap=32000, rec_cap=4096, rounds=1, lut_k=4)),
                                    ("stress", dict(arena_cap=32000, rec_cap=4096, q16=True, lut_k=2)),
                                    ("m200", dict(arena_cap=1 << 22, rec_cap=4096, reuse=True)),
                                    ("N_n2", dict(arena_cap=32000, rec_cap=4096, reuse=True, q16=True)),
                                    ("L_e3", dict(arena_cap=48, rec_cap=2, big_cap=1 << 21))])
def test_bounds_checked_state_machines(tag, kw, golden_dir, g1_index):
    """The state machines compiled with -DB2_CHECKED (every arena slot, bucket, width-record position, table and
    block address tested before use; a violation aborts the process) over the golden set's main variants."""
    args, fq = CASES[tag]
    got, _ = harness_sai(g1_index[0], g1_index[1], args, os.path.join(golden_dir, fq + ".fq.gz"), checked=True, **kw)
    assert got == open(os.path.join(golden_dir, f"g1_{tag}.sai"), "rb").read()


@pytest.mark.parametrize("tag,kw", [("default", dict(arena_cap=32000, rec_cap=4096, rounds=1, suspend_every=1)),
                                    ("stress", dict(arena_cap=32000, rec_cap=4096, q16=True, lut_k=3, suspend_every=7)),
                                    ("m200", dict(arena_cap=32000, rec_cap=4096, reuse=True, suspend_every=3)),
                                    ("N_n2", dict(arena_cap=1 << 22, rec_cap=4096, reuse=True, suspend_every=50)),
                                    ("short_o3", dict(arena_cap=32000, rec_cap=4096, q16=True, suspend_every=2, checked=True))])
def test_parked_and_resumed_lanes_are_exact(tag, kw, golden_dir, g1_index):
    """A search interrupted between any two steps and continued by another lane from the saved words (the kernel
    parks the stragglers of a draining launch this way, SearchLane::save_state / load_state) gives the same bytes."""
    args, fq = CASES[tag]
    got, nov = harness_sai(g1_index[0], g1_index[1], args, os.path.join(golden_dir, fq + ".fq.gz"), **kw)
    assert nov == 0
    assert got == open(os.path.join(golden_dir, f"g1_{tag}.sai"), "rb").read()


def test_long_reads_and_wide_score_ranges(golden_dir, g1_index):
    """1 kbp and 3 kbp reads: max_diff 23 / 75, i.e. 143 and 275 score buckets (the second needs the wide
    heads); also an empty read, which BAM input can deliver (the reference then reports the whole index)."""
    import gzip
    txt = gzip.open(os.path.join(golden_dir, "g1.fa.gz")).read().split(b"\n", 1)[1].replace(b"\n", b"")
    g = seqio.NT4[np.frombuffer(txt, dtype=np.uint8)]
    rng = np.random.default_rng(3)
    reads = []
    for L in (1000, 1000, 3000, 3000, 0, 7):
        s = int(rng.integers(50_000, len(g) - L - 1))
        r = g[s:s + L].copy()
        if L > 100:
            sub = rng.random(L) < 0.004
            r[sub] = (r[sub] + 1) & 3
        reads.append(r)
    n_aln = _vs_oracle(g1_index[0], g1_index[1], reads[:2], [], arena_cap=32000, rec_cap=64, lut_k=5)
    assert (n_aln > 0).all()
    _vs_oracle(g1_index[0], g1_index[1], reads, [], arena_cap=1 << 20, rec_cap=64, lut_k=5)
    _vs_oracle(g1_index[0], g1_index[1], reads, ["-n", "9", "-o", "2"], arena_cap=1 << 20, rec_cap=64)


@pytest.mark.parametrize("args", [["-M", "3", "-O", "3", "-E", "2"], ["-M", "4", "-O", "4", "-E", "4", "-o", "2"],
                                  ["-M", "1", "-O", "2", "-E", "1", "-n", "3"]])
def test_equal_penalties_share_a_bucket(args, golden_dir, g1_index):
    """With s_mm == s_gapo (or s_gape) the gap group and the mismatch group of one expansion land in the same
    bucket: the record is linked twice into it, mismatches on top; other unusual penalty mixes for good measure.
    Checked against the oracle, on the bump arena, the free-list arena and the 32-bit heads."""
    opt, _, _, _ = parse_aln_args(args + ["p", "q"])
    batch = next(seqio.read_batches(os.path.join(golden_dir, "g1_reads.fq.gz"), opt.mode, opt.trim_qual))
    reads = [batch.codes[o:o + l] for o, l in zip(batch.offs[:500], batch.lens[:500])]
    _vs_oracle(g1_index[0], g1_index[1], reads, args, arena_cap=1 << 22, rec_cap=4096, reuse=True, lut_k=4)
    _vs_oracle(g1_index[0], g1_index[1], reads[:200], args, arena_cap=1 << 22, rec_cap=4096, lut_k=2)


def _random_option_sets(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        a = []
        if rng.random() < 0.7:
            a += ["-n", str(int(rng.integers(0, 7)))] if rng.random() < 0.6 else ["-n", f"{rng.choice([0.01, 0.04, 0.1, 0.2])}"]
        if rng.random() < 0.6:
            a += ["-o", str(int(rng.integers(0, 4)))]
        if rng.random() < 0.5:
            a += ["-e", str(int(rng.integers(-1, 8)))]
        if rng.random() < 0.4:
            a += ["-i", str(int(rng.integers(0, 8)))]
        if rng.random() < 0.3:
            a += ["-d", str(int(rng.integers(0, 20)))]
        if rng.random() < 0.5:
            a += ["-l", str(int(rng.choice([10, 20, 32, 50, 1000])))]
        if rng.random() < 0.5:
            a += ["-k", str(int(rng.integers(0, 4)))]
        if rng.random() < 0.3:
            a += ["-m", str(int(rng.choice([50, 500, 5000, 200000])))]
        if rng.random() < 0.5:
            a += ["-M", str(int(rng.integers(1, 6)))]
        if rng.random() < 0.5:
            a += ["-O", str(int(rng.integers(1, 13)))]
        if rng.random() < 0.5:
            a += ["-E", str(int(rng.integers(1, 6)))]
        if rng.random() < 0.3:
            a += ["-R", str(int(rng.integers(1, 40)))]
        if rng.random() < 0.15:
            a += ["-N"]
        if rng.random() < 0.25:
            a += ["-L"]
        if rng.random() < 0.2:
            a += ["-c"]
        out.append(a)
    return out


@pytest.mark.parametrize("k,args", list(enumerate(_random_option_sets(24, 20261018))))
def test_random_option_sets_against_the_oracle(k, args, golden_dir, g1_index):
    """The expansion-record state machine under option mixes nobody chose by hand (penalties that make buckets
    coincide, tiny and huge budgets, seeding on and off, -N, -L, -c): device code on the CPU == oracle, on the
    bump arena, the free-list arena and through the two-pass flow."""
    opt, _, _, _ = parse_aln_args(args + ["p", "q"])
    fq = "g1_short.fq.gz" if k % 4 == 3 else "g1_reads.fq.gz"
    batch = next(seqio.read_batches(os.path.join(golden_dir, fq), opt.mode, opt.trim_qual))
    lo = (k * 97) % max(1, len(batch.lens) - 160)
    reads = [batch.codes[o:o + l] for o, l in zip(batch.offs[lo:lo + 160], batch.lens[lo:lo + 160])]
    _vs_oracle(g1_index[0], g1_index[1], reads, args, arena_cap=1 << 21, rec_cap=1 << 14, reuse=bool(k & 1), lut_k=k % 5,
               rounds=(k // 2) % 3,     # 0 = unlimited, 1 = the fast kernel's setting, 2
               q16=k % 3 != 1)          # 16-bit width records wherever the option set allows them
    if k % 3 == 0:
        from oracle import pyoracle
        lens = np.array([len(r) for r in reads], np.int32)
        offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
        codes = np.concatenate(reads)
        o_n, o_rec, _ = pyoracle.aln_batch(pyoracle.as_orc_bwt(g1_index[0]), pyoracle.as_orc_bwt(g1_index[1]), lens, offs,
                                           codes, opt.to_c())
        h_n, h_rec, nov, _ = pyharness.aln_batch(g1_index[0], g1_index[1], lens, offs, codes, opt.to_c(), arena_cap=48,
                                                 rec_cap=2, big_cap=1 << 21)
        assert np.array_equal(o_n, h_n) and o_rec.tobytes() == h_rec.tobytes()
