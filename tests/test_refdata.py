"""The reference-side index files and the config-4/5 workload generators (ibwa_b200/refdata.py, test / bench
data) against the UNMODIFIED reference binary: every file `ibwa index` writes, and the dbset + `.remap` flow of
`ibwa sampe -R <pri> ... <alt> ...` (bwape.c:548-581,634-657, dbset.c:82-173, bwaremap.cpp:42-132) on them."""
import os

import numpy as np
import pytest

from ibwa_b200 import refdata, synth
from oracle import pyoracle


def _need_ref():
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref/ibwa not present")


def test_index_files_equal_the_reference(tmp_path):
    _need_ref()
    names, lens = ["chrA", "chrB", "chrC"], [30000, 17001, 12999]      # total % 4 == 0: the extra zero byte of .pac
    g = synth.random_genome(sum(lens), 31)
    ours = str(tmp_path / "ours")
    refdata.write_index(ours, g, names, lens)
    fa = str(tmp_path / "ref.fa")
    refdata.write_fasta_contigs(fa, g, names, lens)
    pyoracle.run_ref(["index", "-a", "is", fa])
    for ext in (".pac", ".ann", ".amb", ".bwt", ".rbwt", ".sa", ".rsa"):
        assert open(ours + ext, "rb").read() == open(fa + ext, "rb").read(), ext
    # and a length that is not a multiple of four
    g2 = g[:59998]
    refdata.write_pac_ann_amb(ours + "2", g2, ["x"], [len(g2)])
    fa2 = str(tmp_path / "ref2.fa")
    refdata.write_fasta_contigs(fa2, g2, ["x"], [len(g2)])
    pyoracle.run_ref(["fa2pac", fa2])
    assert open(ours + "2.pac", "rb").read() == open(fa2 + ".pac", "rb").read()
    # the torch packer gives the same bytes
    import torch
    assert refdata.pac_bytes(torch.from_numpy(g2)) == refdata.pac_bytes(g2)


def test_alt_remap_flow_is_accepted_by_the_reference(tmp_path):
    """Config 5 on the reference alone (no GPU): primary + 200 ALT contigs + .remap, 25 % of the pairs from ALT
    sequence -> `sampe -R pri ... alt ...` emits remapped alignments (ZR:Z tags)."""
    _need_ref()
    import torch
    names, lens = ["chr1", "chr2"], [200_000, 150_000]
    g = synth.random_genome(sum(lens), 20260105)
    pri = str(tmp_path / "pri")
    refdata.write_index(pri, g, names, lens)
    alt_text, a_names, a_lens, remap = refdata.make_alt_contigs(lambda lo, hi: g[lo:hi], names, lens, 200)
    assert len(a_names) == 200 and sorted(set(a_lens)) == [1993, 2005]
    alt = str(tmp_path / "alt")
    refdata.write_index(alt, alt_text, a_names, a_lens)
    open(alt + ".remap", "w").write(remap)
    r1, r2 = refdata.synth_pairs(torch.from_numpy(g), 2000, 100, 20260105, torch.from_numpy(alt_text), a_lens, 0.25)
    fq1, fq2 = refdata.write_fastq_pairs(str(tmp_path / "r"), r1.numpy(), r2.numpy())
    sais = []
    for prefix in (pri, alt):
        for fq in (fq1, fq2):
            out = str(tmp_path / (os.path.basename(prefix) + os.path.basename(fq) + ".sai"))
            pyoracle.run_ref(["aln", "-t", "4", prefix, fq], stdout_path=out)
            sais.append(out)
    sam = str(tmp_path / "out.sam")
    pyoracle.run_ref(["sampe", "-R", pri, sais[0], sais[1], fq1, fq2, alt, sais[2], sais[3]], stdout_path=sam)
    lines = [ln for ln in open(sam) if not ln.startswith("@")]
    assert len(lines) == 4000
    mapped = sum(1 for ln in lines if not int(ln.split("\t")[1]) & 4)
    assert mapped > 3600
    assert sum("ZR:Z" in ln for ln in lines) > 150          # alignments translated from ALT to primary coordinates
    # pairs are proper: mates point at each other within the insert-size range
    isz = np.array([abs(int(ln.split("\t")[8])) for ln in lines if int(ln.split("\t")[1]) & 2])
    assert len(isz) > 3000 and 300 < np.median(isz) < 500
