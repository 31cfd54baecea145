"""Parity tests proper: the CUDA path, called through the C ABI, against the
reference's golden .sai files, the CPU oracle and (where oracle/_ref travelled
to the box) the unmodified reference binary.  Bit-exact: integer / byte work."""
import ctypes
import io
import os
import subprocess

import numpy as np
import pytest

from ibwa_b200 import engine, fmbuild, gap_init_opt, parse_aln_args, sai, seqio, synth
from ibwa_b200.bwtio import bwt_dump_bwt
from oracle import pyoracle
from cases import CASES

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def eng(g1_index):
    e = engine.Engine(g1_index[0], g1_index[1], 0)
    yield e
    e.close()


def engine_sai(e, args, fq):
    opt, _, _, _ = parse_aln_args(args + ["prefix", fq])
    buf = io.BytesIO()
    sai.write_header(buf, opt)
    for batch in seqio.read_batches(fq, opt.mode, opt.trim_qual):
        n_aln, rec = e.cal_sa_reg_gap(batch.lens, batch.offs, batch.codes, opt)
        sai.write_batch(buf, n_aln, rec)
    return buf.getvalue()


@pytest.mark.parametrize("tag", sorted(CASES))
def test_golden_sai(tag, eng, golden_dir):
    args, fq = CASES[tag]
    got = engine_sai(eng, args, os.path.join(golden_dir, fq + ".fq.gz"))
    want = open(os.path.join(golden_dir, f"g1_{tag}.sai"), "rb").read()
    assert got == want


@pytest.mark.parametrize("tag", ["default", "N_n2", "q20", "short_o3"])
@pytest.mark.parametrize("pin", [False, True], ids=["pageable", "page_locked"])
def test_sai_bytes_formatted_on_the_device(tag, pin, eng, golden_dir):
    """b200aln_batch_sai: the batch as the bytes of the reference's fwrite loop (bwtaln.c:227-231), from pageable input
    arrays (through the context's staging buffer) and from arrays page-locked with b200aln_pin"""
    args, fq = CASES[tag]
    path = os.path.join(golden_dir, fq + ".fq.gz")
    opt, _, _, _ = parse_aln_args(args + ["prefix", path])
    buf = io.BytesIO()
    sai.write_header(buf, opt)
    for batch in seqio.read_batches(path, opt.mode, opt.trim_qual):
        buf.write(eng.cal_sa_reg_gap_sai(batch.lens, batch.offs, batch.codes, opt, pin=pin))
    assert buf.getvalue() == open(os.path.join(golden_dir, f"g1_{tag}.sai"), "rb").read()


@pytest.mark.parametrize("tag", ["default", "stress", "N_n2"])
def test_overflow_path_is_exact(tag, g1_index, golden_dir):
    """Tiny fast arenas force most reads through the large-arena pass; bytes must not change."""
    args, fq = CASES[tag]
    with engine.Engine(g1_index[0], g1_index[1], 0) as e:
        e.set("arena_cap", 64)
        e.set("rec_cap", 1)
        got = engine_sai(e, args, os.path.join(golden_dir, fq + ".fq.gz"))
        assert e.stats()["overflow_reads"] > 0
    assert got == open(os.path.join(golden_dir, f"g1_{tag}.sai"), "rb").read()


@pytest.mark.parametrize("tag", sorted(CASES))
@pytest.mark.parametrize("knobs", [dict(susp=31, susp_min=0), dict(susp=8, susp_min=64, search_block=32), dict(susp=0),
                                   dict(susp=16, search_blocks_per_sm=1)],
                         ids=["park_at_once", "park_sparse_warps", "never_park", "park_small_grid"])
def test_parked_searches_are_exact(tag, knobs, g1_index, golden_dir):
    """Stragglers of a draining launch are parked (SearchLane::save_state, bucket heads, open group) and resumed in
    dense warps by follow-up launches: whatever the threshold and however many rounds, the bytes are the reference's."""
    args, fq = CASES[tag]
    with engine.Engine(g1_index[0], g1_index[1], 0) as e:
        for k, v in knobs.items():
            e.set(k, v)
        got = engine_sai(e, args, os.path.join(golden_dir, fq + ".fq.gz"))
        launches = e.stats()["kernel_launches"]
    assert got == open(os.path.join(golden_dir, f"g1_{tag}.sai"), "rb").read()
    if knobs.get("susp") == 31:
        assert launches > 8           # width, order x 2, fast pass, scan x 3, compact + at least one resume launch


@pytest.mark.parametrize("tag", ["default", "stress"])
def test_all_three_passes_are_exact(tag, g1_index, golden_dir):
    """Tiny fast AND middle capacities: some reads need the wide pass (32-bit heads in memory)."""
    args, fq = CASES[tag]
    with engine.Engine(g1_index[0], g1_index[1], 0) as e:
        e.set("arena_cap", 64)
        e.set("rec_cap", 1)
        e.set("arena_cap_mid", 256)
        e.set("rec_cap_mid", 3)
        got = engine_sai(e, args, os.path.join(golden_dir, fq + ".fq.gz"))
    assert got == open(os.path.join(golden_dir, f"g1_{tag}.sai"), "rb").read()


@pytest.mark.parametrize("tag", ["default", "q20", "stress"])
def test_cli_binary(tag, golden_dir, tmp_path):
    """The `b200aln aln` command line: same options in, same .sai bytes out (bwtaln.c:243-328)."""
    args, fq = CASES[tag]
    prefix = str(tmp_path / "g1")
    os.symlink(os.path.join(golden_dir, "g1.bwt"), prefix + ".bwt")
    os.symlink(os.path.join(golden_dir, "g1.rbwt"), prefix + ".rbwt")
    out = str(tmp_path / "out.sai")
    exe = os.path.join(ROOT, "ibwa_b200", "b200aln")
    subprocess.check_call([exe, "aln"] + args + ["-f", out, prefix, os.path.join(golden_dir, fq + ".fq.gz")],
                          stderr=subprocess.DEVNULL)
    assert open(out, "rb").read() == open(os.path.join(golden_dir, f"g1_{tag}.sai"), "rb").read()
    # stdout variant
    p = subprocess.run([exe, "aln"] + args + [prefix, os.path.join(golden_dir, fq + ".fq.gz")],
                       stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True)
    assert p.stdout == open(os.path.join(golden_dir, f"g1_{tag}.sai"), "rb").read()


@pytest.mark.parametrize("gz", [True, False], ids=["gzip_stream", "plain_stream"])
def test_cli_reads_from_stdin(gz, golden_dir, tmp_path):
    """`aln <prefix> -` reads the FASTQ from standard input (utils.c:56-66: xzopen("-") = gzdopen(stdin), which
    takes plain and gzip streams alike); the reference binary fed the same pipe writes the same bytes."""
    import gzip
    prefix = str(tmp_path / "g1")
    os.symlink(os.path.join(golden_dir, "g1.bwt"), prefix + ".bwt")
    os.symlink(os.path.join(golden_dir, "g1.rbwt"), prefix + ".rbwt")
    data = open(os.path.join(golden_dir, "g1_reads.fq.gz"), "rb").read()
    if not gz:
        data = gzip.decompress(data)
    exe = os.path.join(ROOT, "ibwa_b200", "b200aln")
    p = subprocess.run([exe, "aln", prefix, "-"], input=data, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True)
    assert p.stdout == open(os.path.join(golden_dir, "g1_default.sai"), "rb").read()
    if pyoracle.have_ref():
        q = subprocess.run([pyoracle.REF_BIN, "aln", prefix, "-"], input=data, stdout=subprocess.PIPE,
                           stderr=subprocess.DEVNULL, check=True)
        assert q.stdout == p.stdout


@pytest.fixture(scope="module")
def rand_index():
    g = synth.random_genome(2_000_000, 20260101)
    bwt, rbwt = fmbuild.build_index_numpy(g)
    return g, bwt, rbwt


@pytest.mark.parametrize("model,args,length,n", [("default", [], 100, 20000),
                                                 ("stress", ["-n", "4", "-o", "2", "-e", "10", "-l", "32", "-k", "2"], 150, 4000)])
def test_random_genome_vs_oracle(rand_index, model, args, length, n):
    g, bwt, rbwt = rand_index
    if model == "default":
        reads = synth.simulate_reads_fast(g, n, length, 20260112)
    else:
        reads = np.stack(synth.simulate_reads(g, n, length, 20260103, model="stress"))
    lens = np.full(n, length, np.int32)
    offs = np.arange(n, dtype=np.int64) * length
    opt, _, _, _ = parse_aln_args(args + ["p", "q"])
    with engine.Engine(bwt, rbwt, 0) as e:
        e.set("count", 1)      # fast pass with its pop counter
        n_aln, rec = e.cal_sa_reg_gap(lens, offs, reads.reshape(-1), opt)
        st = e.stats()
        e.set("count", 0)      # and the product configuration (no counters) on the same reads
        n_aln2, rec2 = e.cal_sa_reg_gap(lens, offs, reads.reshape(-1), opt)
        assert np.array_equal(n_aln, n_aln2) and rec.tobytes() == rec2.tobytes()
    o_n, o_rec, ost = pyoracle.aln_batch(pyoracle.as_orc_bwt(bwt), pyoracle.as_orc_bwt(rbwt), lens, offs,
                                         reads.reshape(-1), opt.to_c())
    assert np.array_equal(n_aln, o_n)
    assert rec.tobytes() == o_rec.tobytes()
    if st["overflow_reads"] == 0:
        assert st["pops"] == ost["pops"], (st, ost)      # same search, pop for pop
    else:
        assert st["pops"] > ost["pops"], (st, ost)       # flagged reads are searched twice
    assert (n_aln > 0).mean() > (0.8 if model == "default" else 0.3)


def test_device_resident_entry_point(rand_index):
    import torch
    g, bwt, rbwt = rand_index
    n, length = 5000, 100
    reads = synth.simulate_reads_fast(g, n, length, 5)
    lens = np.full(n, length, np.int32)
    offs = np.arange(n, dtype=np.int64) * length
    opt = gap_init_opt()
    with engine.Engine(bwt, rbwt, 0) as e:
        n_aln, rec = e.cal_sa_reg_gap(lens, offs, reads.reshape(-1), opt)
        d_l, d_o, d_c = (torch.from_numpy(x).cuda() for x in (lens, offs, reads.reshape(-1)))
        torch.cuda.synchronize()
        pn, pr, total = e.batch_device(d_l.data_ptr(), d_o.data_ptr(), d_c.data_ptr(), n, length, opt)
        assert total == len(rec)
        out_n = torch.empty(n, dtype=torch.int32, device="cuda")
        out_r = torch.empty(total * 4, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        import ibwa_b200.devcopy as devcopy
        devcopy.d2d(out_n.data_ptr(), pn, n * 4)
        devcopy.d2d(out_r.data_ptr(), pr, total * 16)
        torch.cuda.synchronize()
        assert np.array_equal(out_n.cpu().numpy(), n_aln)
        assert out_r.cpu().numpy().tobytes() == rec.tobytes()


def test_chunk_pipeline_inside_one_call(rand_index):
    """A call cut into chunks that run on the context's sibling contexts (own stream, pinned staging, in-order
    assembly): host-buffer and device-resident entry points give what the single launch gives."""
    import torch
    import ibwa_b200.devcopy as devcopy
    g, bwt, rbwt = rand_index
    n, length = 20_000, 100
    reads = synth.simulate_reads_fast(g, n, length, 7)
    lens = np.full(n, length, np.int32)
    offs = np.arange(n, dtype=np.int64) * length
    opt = gap_init_opt()
    with engine.Engine(bwt, rbwt, 0) as e:
        e.set("slots", 1)
        n_aln, rec = e.cal_sa_reg_gap(lens, offs, reads.reshape(-1), opt)
        e.set("slots", 3)
        e.set("chunk_reads", 3000)                  # 7 chunks on 3 contexts
        e.set("chunk_reads_device", 3000)
        n2, rec2 = e.cal_sa_reg_gap(lens, offs, reads.reshape(-1), opt)
        assert np.array_equal(n_aln, n2) and rec.tobytes() == rec2.tobytes()
        d_l, d_o, d_c = (torch.from_numpy(x).cuda() for x in (lens, offs, reads.reshape(-1)))
        torch.cuda.synchronize()
        pn, pr, total = e.batch_device(d_l.data_ptr(), d_o.data_ptr(), d_c.data_ptr(), n, length, opt)
        assert total == len(rec)
        out_n = torch.empty(n, dtype=torch.int32, device="cuda")
        out_r = torch.empty(total * 4, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        devcopy.d2d(out_n.data_ptr(), pn, n * 4)
        devcopy.d2d(out_r.data_ptr(), pr, total * 16)
        torch.cuda.synchronize()
        assert np.array_equal(out_n.cpu().numpy(), n_aln) and out_r.cpu().numpy().tobytes() == rec.tobytes()


def test_first_batches_of_fresh_contexts_in_flight_together(rand_index):
    """Six fresh contexts on one index, every one's FIRST batch (the one that sets up its parking ring) in flight at
    the same time with parking on: the persistent blocks of six launches leave no room on the SMs, which is when a
    ring cleared on the wrong stream used to lose parked searches.  Every context must return what one batch on
    its own returns.  Repeated, since the interleaving is the point."""
    import threading
    g, bwt, rbwt = rand_index
    n, length = 150_000, 100
    reads = synth.simulate_reads_fast(g, n, length, 11)
    lens = np.full(n, length, np.int32)
    offs = np.arange(n, dtype=np.int64) * length
    opt = gap_init_opt()
    with engine.Engine(bwt, rbwt, 0) as e:
        e.set("susp", 0)
        want_n, want_rec = e.cal_sa_reg_gap(lens, offs, reads.reshape(-1), opt)
        for rep in range(3):
            clones = [e.clone() for _ in range(6)]
            got = [None] * len(clones)

            def run(i):
                clones[i].set("susp", 16)  # park whenever a warp is down to 16 searches, whatever else is in flight
                got[i] = clones[i].cal_sa_reg_gap(lens, offs, reads.reshape(-1), opt)

            th = [threading.Thread(target=run, args=(i,)) for i in range(len(clones))]
            for t in th:
                t.start()
            for t in th:
                t.join()
            for c in clones:
                c.close()
            for i, (n_aln, rec) in enumerate(got):
                assert np.array_equal(n_aln, want_n), (rep, i, int((n_aln != want_n).sum()))
                assert rec.tobytes() == want_rec.tobytes(), (rep, i)


def test_reference_binary_on_box(rand_index, tmp_path):
    """If the unmodified reference binary travelled with the repo, run it here on fresh inputs."""
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref/ibwa not present")
    g, bwt, rbwt = rand_index
    prefix = str(tmp_path / "r2m")
    bwt_dump_bwt(prefix + ".bwt", bwt)
    bwt_dump_bwt(prefix + ".rbwt", rbwt)
    reads = synth.simulate_reads_fast(g, 8000, 100, 99)
    fq = str(tmp_path / "r.fq")
    synth.write_fastq(fq, list(reads))
    ref_out = str(tmp_path / "ref.sai")
    pyoracle.run_ref(["aln", "-t", "4", prefix, fq], stdout_path=ref_out)
    out = str(tmp_path / "gpu.sai")
    engine.bwa_aln_core(prefix, fq, gap_init_opt(), out, 0)
    a, b = open(ref_out, "rb").read(), open(out, "rb").read()
    assert a[:52] == b[:52] and a[56:] == b[56:]      # byte 52..55 = n_threads

def test_cli_merges_reference_batches_only_when_allowed(rand_index, tmp_path):
    """The CLI driver hands several 0x40000-read reference batches to the GPU as one launch, but only batches
    that agree on the batch-level max_gapo clamp (bwtaln.c:89-92).  Input: one whole batch of reads < 38 bp
    (with -o 3 the clamp bites), then 100 bp reads, then a short tail — against the unmodified reference
    binary, with the default (up to eight reference batches per launch), no merging, up to 64 merged with one launch
    in flight, and pageable parse buffers."""
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref/ibwa not present")
    g, bwt, rbwt = rand_index
    prefix = str(tmp_path / "r2m")
    bwt_dump_bwt(prefix + ".bwt", bwt)
    bwt_dump_bwt(prefix + ".rbwt", rbwt)
    rng = np.random.default_rng(8)
    nt = np.frombuffer(b"ACGT", dtype=np.uint8)
    fq = str(tmp_path / "mixed.fq")
    with open(fq, "wb") as f:
        def emit(count, lo, hi, tag):
            lens = rng.integers(lo, hi + 1, size=count)
            starts = rng.integers(0, len(g) - 300, size=count)
            out = []
            for i in range(count):
                r = g[starts[i]:starts[i] + lens[i]].copy()
                if i % 3 == 0:
                    r[int(rng.integers(0, lens[i]))] ^= 1
                out.append(b"@" + tag + b"%d\n" % i + nt[r].tobytes() + b"\n+\n" + b"I" * int(lens[i]) + b"\n")
            f.write(b"".join(out))
        emit(0x40000, 20, 37, b"s")        # exactly one reference batch of short reads
        emit(0x40000 + 5000, 100, 100, b"l")   # one full batch and the start of the next ...
        emit(3000, 25, 36, b"t")           # ... which ends with short reads (max_len 100: no clamp)
    args = ["-o", "3"]
    ref_out = str(tmp_path / "ref.sai")
    pyoracle.run_ref(["aln", "-t", str(os.cpu_count() or 4)] + args + [prefix, fq], stdout_path=ref_out)
    want = open(ref_out, "rb").read()
    exe = os.path.join(ROOT, "ibwa_b200", "b200aln")
    for env in ({}, {"B200ALN_MERGE": "1"}, {"B200ALN_INFLIGHT": "1", "B200ALN_MERGE": "64"}, {"B200ALN_NO_PIN": "1", "B200ALN_MERGE": "2"}):
        out = str(tmp_path / "gpu.sai")
        subprocess.check_call([exe, "aln"] + args + ["-f", out, prefix, fq], stderr=subprocess.DEVNULL,
                              env=dict(os.environ, **env))
        got = open(out, "rb").read()
        assert len(got) == len(want) and got[:52] == want[:52] and got[56:] == want[56:], env


def test_cli_parse_arrays_grow_while_page_locked(rand_index, tmp_path):
    """150 bp reads outgrow the driver's parse arrays (reserved for 104 bases per read), so the vectors move while
    their old blocks are page-locked: the allocator unlocks them first.  Small reference batches (the test hook
    B200ALN_BATCH_READS; every batch has the same longest read, so the clamp is the same) make units of up to eight
    batches that all have to grow; the output equals one batch through the operator."""
    g, bwt, rbwt = rand_index
    prefix = str(tmp_path / "grow")
    bwt_dump_bwt(prefix + ".bwt", bwt)
    bwt_dump_bwt(prefix + ".rbwt", rbwt)
    n, length = 120_000, 150
    reads = synth.simulate_reads_fast(g, n, length, 13)
    fq = str(tmp_path / "long.fq")
    synth.write_fastq(fq, reads)
    opt = gap_init_opt()
    with engine.Engine(bwt, rbwt, 0) as e:
        n_aln, rec = e.cal_sa_reg_gap(np.full(n, length, np.int32), np.arange(n, dtype=np.int64) * length, reads.reshape(-1), opt)
    buf = io.BytesIO()
    sai.write_header(buf, opt)
    sai.write_batch(buf, n_aln, rec)
    want = buf.getvalue()
    exe = os.path.join(ROOT, "ibwa_b200", "b200aln")
    for env in ({"B200ALN_BATCH_READS": "5000"}, {"B200ALN_BATCH_READS": "7000", "B200ALN_INFLIGHT": "2"}):
        out = str(tmp_path / "gpu.sai")
        subprocess.check_call([exe, "aln", "-f", out, prefix, fq], stderr=subprocess.DEVNULL, env=dict(os.environ, **env))
        assert open(out, "rb").read() == want, env


def _ref_index(tmp_path, genome, name, contig_len=None):
    """Index a synthetic genome with the unmodified reference (`ibwa index -a is`)."""
    fa = str(tmp_path / f"{name}.fa")
    synth.write_fasta(fa, genome, contig_len=contig_len)
    pyoracle.run_ref(["index", "-a", "is", fa])
    return fa


def test_paired_end_and_downstream_sam(tmp_path):
    """BASELINE config 4 at reduced size: aln on both mates, .sai identity, and equality of the SAM that
    the UNCHANGED reference `sampe -R` / `samse` produce from the engine's .sai and from its own."""
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref/ibwa not present")
    g = synth.random_genome(300_000, 20260104)
    fa = _ref_index(tmp_path, g, "pe")
    rng = np.random.default_rng(4)
    n, L = 3000, 100
    starts = rng.integers(0, len(g) - 700, size=n)
    isz = np.clip(rng.normal(400, 40, size=n).astype(int), 220, 600)
    r1 = np.stack([g[s:s + L] for s in starts]).copy()
    r2 = np.stack([synth.revcomp(g[s + i - L:s + i]) for s, i in zip(starts, isz)]).copy()
    for r in (r1, r2):
        sub = rng.random(r.shape) < 0.01
        r[sub] = (r[sub] + rng.integers(1, 4, size=int(sub.sum()))) & 3
    names = [f"p{i}" for i in range(n)]
    fq1, fq2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    synth.write_fastq(fq1, list(r1), names=[x + "/1" for x in names])
    synth.write_fastq(fq2, list(r2), names=[x + "/2" for x in names])
    sais = {}
    for tag, fq in (("1", fq1), ("2", fq2)):
        ref_sai, gpu_sai = str(tmp_path / f"ref{tag}.sai"), str(tmp_path / f"gpu{tag}.sai")
        pyoracle.run_ref(["aln", fa, fq], stdout_path=ref_sai)
        engine.bwa_aln_core(fa, fq, gap_init_opt(), gpu_sai, 0)
        assert open(ref_sai, "rb").read() == open(gpu_sai, "rb").read()
        sais[tag] = (ref_sai, gpu_sai)
    ref_sam, gpu_sam = str(tmp_path / "ref.sam"), str(tmp_path / "gpu.sam")
    pyoracle.run_ref(["sampe", "-R", fa, sais["1"][0], sais["2"][0], fq1, fq2], stdout_path=ref_sam)
    pyoracle.run_ref(["sampe", "-R", fa, sais["1"][1], sais["2"][1], fq1, fq2], stdout_path=gpu_sam)
    a, b = open(ref_sam, "rb").read(), open(gpu_sam, "rb").read()
    assert a == b and a.count(b"\n") > 2 * n
    se_ref, se_gpu = str(tmp_path / "se_ref.sam"), str(tmp_path / "se_gpu.sam")
    pyoracle.run_ref(["samse", fa, sais["1"][0], fq1], stdout_path=se_ref)
    pyoracle.run_ref(["samse", fa, sais["1"][1], fq1], stdout_path=se_gpu)
    assert open(se_ref, "rb").read() == open(se_gpu, "rb").read()


def test_multi_contig_alt_index(tmp_path):
    """BASELINE config 5 at reduced size: aln against a many-contig index (primary + 200 alt contigs cut from
    it with SNPs / indels).  The concatenated-contig BWT is consumed like any other; .sai must be identical."""
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref/ibwa not present")
    rng = np.random.default_rng(5)
    pri = synth.random_genome(400_000, 20260105)
    alts = []
    for j in range(200):
        s = int(rng.integers(0, len(pri) - 2100))
        a = pri[s:s + 2000].copy()
        snp = np.arange(150, 2000, 300)
        a[snp] = (a[snp] + 1) & 3
        a = np.concatenate([a[:1000], rng.integers(0, 4, size=5, dtype=np.uint8), a[1000:]]) if j % 2 == 0 \
            else np.concatenate([a[:1000], a[1007:]])
        alts.append(a)
    genome = np.concatenate([pri] + alts)
    fa = str(tmp_path / "alt.fa")
    with open(fa, "wb") as f:      # contigs of different lengths
        nt = np.frombuffer(b"ACGT", dtype=np.uint8)
        f.write(b">chr1\n" + nt[pri].tobytes() + b"\n")
        for j, a in enumerate(alts):
            f.write(b">alt%d\n" % j + nt[a].tobytes() + b"\n")
    pyoracle.run_ref(["index", "-a", "is", fa])
    reads = synth.simulate_reads_fast(genome, 6000, 100, 20260105)
    fq = str(tmp_path / "r.fq")
    synth.write_fastq(fq, list(reads))
    ref_sai, gpu_sai = str(tmp_path / "ref.sai"), str(tmp_path / "gpu.sai")
    pyoracle.run_ref(["aln", "-t", "3", fa, fq], stdout_path=ref_sai)
    engine.bwa_aln_core(fa, fq, gap_init_opt(), gpu_sai, 0)
    a, b = open(ref_sai, "rb").read(), open(gpu_sai, "rb").read()
    assert a[:52] == b[:52] and a[56:] == b[56:]
    _, n_aln, _ = sai.read_sai(gpu_sai)
    assert (n_aln >= 2).mean() > 0.05          # reads from duplicated (alt) sequence have several records


def test_cli_bam_input(tmp_path, golden_dir):
    """`b200aln aln -b -1 ...` on a BAM file against the reference binary (bwaseqio.c:89-141)."""
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref/ibwa not present")
    import gzip as gz
    from test_bam import make_bam
    txt = gz.open(os.path.join(golden_dir, "g1.fa.gz")).read().split(b"\n", 1)[1].replace(b"\n", b"")
    genome = seqio.NT4[np.frombuffer(txt, dtype=np.uint8)]
    bam = str(tmp_path / "in.bam")
    make_bam(bam, genome, np.random.default_rng(11))
    prefix = str(tmp_path / "g1")
    os.symlink(os.path.join(golden_dir, "g1.bwt"), prefix + ".bwt")
    os.symlink(os.path.join(golden_dir, "g1.rbwt"), prefix + ".rbwt")
    exe = os.path.join(ROOT, "ibwa_b200", "b200aln")
    for flags in (["-b"], ["-b", "-1", "-q", "15"]):
        ref_out, out = str(tmp_path / "ref.sai"), str(tmp_path / "gpu.sai")
        pyoracle.run_ref(["aln"] + flags + [prefix, bam], stdout_path=ref_out)
        subprocess.check_call([exe, "aln"] + flags + ["-f", out, prefix, bam], stderr=subprocess.DEVNULL)
        assert open(out, "rb").read() == open(ref_out, "rb").read()


def test_long_reads_wide_heads_and_empty_read(g1_index, golden_dir):
    """3 kbp reads (275 score buckets: 32-bit heads in global memory on the fast pass), 1 kbp reads, an empty
    read and a 7-bp read in one batch, against the oracle."""
    import gzip as gz
    txt = gz.open(os.path.join(golden_dir, "g1.fa.gz")).read().split(b"\n", 1)[1].replace(b"\n", b"")
    g = seqio.NT4[np.frombuffer(txt, dtype=np.uint8)]
    rng = np.random.default_rng(3)
    reads = []
    for L in (1000, 1000, 3000, 3000, 0, 7) * 20:
        s = int(rng.integers(50_000, len(g) - L - 1))
        r = g[s:s + L].copy()
        if L > 100:
            sub = rng.random(L) < 0.004
            r[sub] = (r[sub] + 1) & 3
        reads.append(r)
    lens = np.array([len(r) for r in reads], np.int32)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    codes = np.concatenate(reads)
    for args in ([], ["-n", "9", "-o", "2"]):
        opt, _, _, _ = parse_aln_args(args + ["p", "q"])
        with engine.Engine(g1_index[0], g1_index[1], 0) as e:
            n_aln, rec = e.cal_sa_reg_gap(lens, offs, codes, opt)
        o_n, o_rec, _ = pyoracle.aln_batch(pyoracle.as_orc_bwt(g1_index[0]), pyoracle.as_orc_bwt(g1_index[1]), lens, offs,
                                           codes, opt.to_c())
        assert np.array_equal(n_aln, o_n) and rec.tobytes() == o_rec.tobytes()


def test_random_option_sets_on_the_gpu(g1_index, golden_dir):
    """The option mixes of tests/test_core_logic.py::test_random_option_sets_against_the_oracle through the CUDA
    kernels (all three passes reachable: small fast arena), records compared with the oracle."""
    from test_core_logic import _random_option_sets
    with engine.Engine(g1_index[0], g1_index[1], 0) as e:
        e.set("arena_cap", 512)
        for k, args in enumerate(_random_option_sets(24, 20261018)):
            if k % 2:           # every other set: the CPU oracle is what takes the time here
                continue
            opt, _, _, _ = parse_aln_args(args + ["p", "q"])
            fq = "g1_short.fq.gz" if k % 4 == 2 else "g1_reads.fq.gz"
            batch = next(seqio.read_batches(os.path.join(golden_dir, fq), opt.mode, opt.trim_qual))
            n = min(250, len(batch.lens))
            lens, offs = batch.lens[:n], batch.offs[:n]
            codes = batch.codes[:int(offs[n - 1] + lens[n - 1])]
            n_aln, rec = e.cal_sa_reg_gap(lens, offs, codes, opt)
            o_n, o_rec, _ = pyoracle.aln_batch(pyoracle.as_orc_bwt(g1_index[0]), pyoracle.as_orc_bwt(g1_index[1]), lens, offs,
                                               codes, opt.to_c())
            assert np.array_equal(n_aln, o_n), args
            assert rec.tobytes() == o_rec.tobytes(), args
