"""BASELINE configs 4 and 5 at a hundred times the size of the round-1 tests, and the next-row pieces composed:

  config 4  300 k pairs (2 x 300 k x 100 bp) against a 40 Mbp two-contig genome: both mates through the engine's
            own driver (parse + search + write), `.sai` identity with `ibwa aln`, and byte equality of the SAM that
            the UNCHANGED `ibwa sampe -R` writes from the engine's and from the reference's `.sai`
            (bwape.c:603-657)
  config 5  the same against a dbset: primary + 200 ALT contigs with a `.remap` file, 25 % of the pairs drawn from
            ALT sequence: four `.sai` streams, `sampe -R <pri> .. <alt> ..`, SAM equality including the ZR:Z lines
            (bwape.c:548-581,634-657, dbset.c:82-173, bwaremap.cpp:42-132)
  N4 -> N2  two engine `.sai` streams merged on the GPU (alngrp_create, saiset.c:45-78), the unique best hits
            mapped to positions on the GPU (bwtdb_sa2seq, dbset.c:240-245) and compared with the positions the
            reference's `samse` prints for the same reads (bwa_cal_pac_pos, bwase.c:128-161)
Index files come from ibwa_b200.refdata (byte-identical to `ibwa index`, tests/test_refdata.py)."""
import os

import numpy as np
import pytest

from ibwa_b200 import engine, gap_init_opt, refdata, sai
from ibwa_b200.bwtio import bwt_restore_bwt, bwt_restore_sa
from oracle import pyoracle

pytestmark = pytest.mark.gpu
NPROC = os.cpu_count() or 1


def _need_ref():
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref/ibwa not present")


def _genome(n, seed):
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return torch.randint(0, 4, (n,), dtype=torch.uint8, device="cuda", generator=g)


def _same_sai(a, b):
    x, y = open(a, "rb").read(), open(b, "rb").read()
    return len(x) == len(y) and x[:52] == y[:52] and x[56:] == y[56:]       # byte 52..55 = n_threads


def _aln_both(prefix, fqs, tmp_path, tag):
    """engine driver and `ibwa aln -t N` on every FASTQ: returns ([engine .sai], [reference .sai])"""
    e_sais, r_sais = [], []
    for j, fq in enumerate(fqs):
        e, r = str(tmp_path / f"{tag}_e{j}.sai"), str(tmp_path / f"{tag}_r{j}.sai")
        engine.bwa_aln_core(prefix, fq, gap_init_opt(), e, 0)
        pyoracle.run_ref(["aln", "-t", str(NPROC), prefix, fq], stdout_path=r)
        assert _same_sai(e, r), (prefix, fq)
        e_sais.append(e)
        r_sais.append(r)
    return e_sais, r_sais


@pytest.mark.parametrize("with_alt", [False, True], ids=["config4_paired", "config5_dbset_remap"])
def test_paired_configs_at_scale(with_alt, tmp_path):
    _need_ref()
    import torch
    n_pairs, L = 300_000, 100
    names, lens = ["chr1", "chr2"], [25_000_000, 15_000_000]
    text = _genome(sum(lens), 20260104 + with_alt)
    pri = str(tmp_path / "pri")
    refdata.write_index(pri, text, names, lens)
    kw = {}
    alt = None
    if with_alt:
        alt_text, a_names, a_lens, remap = refdata.make_alt_contigs(lambda lo, hi: text[lo:hi].cpu().numpy(), names, lens, 200)
        alt = str(tmp_path / "alt")
        refdata.write_index(alt, alt_text, a_names, a_lens)
        open(alt + ".remap", "w").write(remap)
        kw = dict(alt_text=torch.from_numpy(alt_text).cuda(), alt_lens=a_lens, alt_frac=0.25)
    r1, r2 = refdata.synth_pairs(text, n_pairs, L, 20260104, **kw)
    del text
    torch.cuda.empty_cache()
    fqs = refdata.write_fastq_pairs(str(tmp_path / "r"), r1.cpu().numpy(), r2.cpu().numpy())
    e_pri, r_pri = _aln_both(pri, fqs, tmp_path, "pri")
    e_alt, r_alt = _aln_both(alt, fqs, tmp_path, "alt") if with_alt else ([], [])
    sams = []
    for tag, sp, sa in (("engine", e_pri, e_alt), ("reference", r_pri, r_alt)):
        out = str(tmp_path / f"{tag}.sam")
        cmd = ["sampe", "-R", "-t", str(NPROC), pri, sp[0], sp[1], fqs[0], fqs[1]] + ([alt, sa[0], sa[1]] if with_alt else [])
        pyoracle.run_ref(cmd, stdout_path=out)
        sams.append(out)
    assert refdata.md5_file(sams[0]) == refdata.md5_file(sams[1])
    body = [ln for ln in open(sams[0], "rb") if not ln.startswith(b"@")]
    assert len(body) == 2 * n_pairs
    mapped = sum(1 for ln in body if not int(ln.split(b"\t")[1]) & 4)
    assert mapped > 0.99 * len(body)
    zr = sum(1 for ln in body if b"ZR:Z" in ln)
    assert (zr > 0.03 * len(body)) if with_alt else zr == 0      # alignments translated from ALT to primary coordinates


def test_sai_merge_then_positions_match_samse(tmp_path):
    _need_ref()
    import torch
    names, lens = ["chr1", "chr2", "chr3"], [1_200_000, 700_000, 100_000]
    text = _genome(sum(lens), 20260106)
    pri = str(tmp_path / "pri")
    refdata.write_index(pri, text, names, lens)
    alt_text, a_names, a_lens, _ = refdata.make_alt_contigs(lambda lo, hi: text[lo:hi].cpu().numpy(), names, lens, 200)
    alt = str(tmp_path / "alt")
    refdata.write_index(alt, alt_text, a_names, a_lens)
    n, L = 20_000, 100
    r1, _ = refdata.synth_pairs(text, n, L, 20260106, alt_text=torch.from_numpy(alt_text).cuda(), alt_lens=a_lens, alt_frac=0.3)
    reads = r1.cpu().numpy()
    del text
    fq = str(tmp_path / "r.fq")
    import bench
    bench.write_fastq(fq, reads)
    opt = gap_init_opt()
    lens_a, offs_a = np.full(n, L, np.int32), np.arange(n, dtype=np.int64) * L
    streams, engines = [], []
    for prefix in (pri, alt):
        e = engine.Engine(bwt_restore_bwt(prefix + ".bwt"), bwt_restore_bwt(prefix + ".rbwt"), 0)
        e.load_sa(0, bwt_restore_sa(prefix + ".sa"))
        e.load_sa(1, bwt_restore_sa(prefix + ".rsa"))
        streams.append(e.cal_sa_reg_gap(lens_a, offs_a, reads.reshape(-1), opt))
        engines.append(e)
    # N4 on the GPU: per read the records of both streams, sorted by score, cut at best + s_mm
    out_off, out_n, rec, db = engines[0].alngrp_merge([s[0] for s in streams], [s[1] for s in streams], opt.s_mm)
    # what samse prints for each index on its own stream (positions of unique, gap-free best hits)
    sam_pos = []
    for j, prefix in enumerate((pri, alt)):
        sp = str(tmp_path / f"s{j}.sai")
        with open(sp, "wb") as f:
            sai.write_header(f, opt)
            sai.write_batch(f, streams[j][0], streams[j][1])
        out = str(tmp_path / f"s{j}.sam")
        pyoracle.run_ref(["samse", prefix, sp, fq], stdout_path=out)
        off, running, pos = {}, 0, []
        for ln in open(out):
            if ln.startswith("@SQ"):                       # contigs in index order: offset in the concatenated text
                f_ = dict(x.split(":", 1) for x in ln.rstrip().split("\t")[1:])
                off[f_["SN"]] = running
                running += int(f_["LN"])
                continue
            if ln.startswith("@"):
                continue
            t = ln.split("\t")
            pos.append((int(t[1]), t[2], int(t[3]), t[5]))
        sam_pos.append((off, pos))
    checked = [0, 0]
    for j in (0, 1):
        off, pos = sam_pos[j]
        first = out_off[out_n > 0]
        ridx = np.nonzero(out_n > 0)[0]
        top = rec[first]
        p = top["packed"]
        nxt_worse = np.ones(len(first), bool)
        many = out_n[ridx] > 1
        nxt_worse[many] = rec[first[many] + 1]["score"] > top["score"][many]
        sel = (db[first] == j) & (top["k"] == top["l"]) & (((p >> 8) & 0xffff) == 0) & nxt_worse
        rows, strand = top["k"][sel], ((p >> 24) & 1)[sel].astype(np.uint8)
        got = engines[j].sa2seq(strand, rows, np.full(int(sel.sum()), L, np.int32))
        for r, g, st in zip(ridx[sel], got, strand):
            flag, rname, spos, cigar = pos[r]
            if flag & 4 or cigar != f"{L}M":
                continue
            assert off[rname] + spos - 1 == int(g) and bool(flag & 16) == bool(st), (j, r, pos[r], int(g))
            checked[j] += 1
    for e in engines:
        e.close()
    assert checked[0] > 0.5 * n and checked[1] > 0.03 * n, checked
