"""BAM input (-b, -0/-1/-2): the native reader feeding the oracle must reproduce what the reference
binary writes for the same BAM (bwaseqio.c:89-141).  Runs without a GPU where oracle/_ref exists."""
import gzip
import io
import os
import struct

import numpy as np
import pytest

from ibwa_b200 import engine, parse_aln_args, sai
from oracle import pyoracle

NT16 = {0: 1, 1: 2, 2: 4, 3: 8, 4: 15}


def bam_record(name, codes, quals, flag):
    l = len(codes)
    seq = bytearray((l + 1) // 2)
    for i, c in enumerate(codes):
        seq[i // 2] |= NT16[int(c)] << (4 * (1 - i % 2))
    qn = name.encode() + b"\0"
    core = struct.pack("<iiIIiiii", -1, -1, (4680 << 16) | (0 << 8) | len(qn), (flag << 16) | 0, l, -1, -1, 0)
    body = core + qn + bytes(seq) + bytes(int(q) for q in quals)
    return struct.pack("<i", len(body)) + body


def make_bam(path, genome, rng, n=400, bgzf=False):
    out = io.BytesIO()
    text = b"@HD\tVN:1.0\n"
    out.write(b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", 1))
    out.write(struct.pack("<i", 5) + b"chr1\0" + struct.pack("<i", len(genome)))
    for i in range(n):
        L = int(rng.choice([36, 50, 75, 100, 101]))
        s = int(rng.integers(0, len(genome) - L))
        r = genome[s:s + L].copy()
        sub = rng.random(L) < 0.01
        r[sub] = (r[sub] + 1) & 3
        if i % 50 == 0:
            r[3] = 4
        q = np.clip(38 - np.arange(L) * 40 // L + rng.integers(-5, 6, size=L), 2, 41)
        kind = i % 3                       # read1 / read2 / single-end
        flag = {0: 1 | 64, 1: 1 | 128, 2: 0}[kind]
        if rng.random() < 0.5:             # stored reverse-complemented with the reverse flag
            flag |= 16
            r = np.array([3 - c if c < 4 else c for c in r[::-1]], dtype=np.uint8)
            q = q[::-1]
        out.write(bam_record(f"q{i}", r, q, flag))
    if bgzf:      # the real container: BGZF blocks (inflated block-parallel by the native reader)
        from test_reader import bgzf_bytes
        with open(path, "wb") as f:
            f.write(bgzf_bytes(out.getvalue(), block=3000))
    else:         # any gzip stream is accepted, like the reference (gzread)
        with gzip.open(path, "wb") as f:
            f.write(out.getvalue())


@pytest.mark.parametrize("bgzf", [False, True])
@pytest.mark.parametrize("flags", [["-b"], ["-b", "-1"], ["-b", "-2"], ["-b", "-0"], ["-b", "-1", "-2"],
                                   ["-b", "-q", "15"]])
def test_bam_reader_matches_reference(tmp_path, golden_dir, g1_index, flags, bgzf):
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref/ibwa not present")
    import gzip as gz
    from ibwa_b200 import seqio
    txt = gz.open(os.path.join(golden_dir, "g1.fa.gz")).read().split(b"\n", 1)[1].replace(b"\n", b"")
    genome = seqio.NT4[np.frombuffer(txt, dtype=np.uint8)]
    bam = str(tmp_path / "in.bam")
    make_bam(bam, genome, np.random.default_rng(7), bgzf=bgzf)
    prefix = str(tmp_path / "g1")
    os.symlink(os.path.join(golden_dir, "g1.bwt"), prefix + ".bwt")
    os.symlink(os.path.join(golden_dir, "g1.rbwt"), prefix + ".rbwt")
    ref_sai = str(tmp_path / "ref.sai")
    pyoracle.run_ref(["aln"] + flags + [prefix, bam], stdout_path=ref_sai)
    opt, _, _, _ = parse_aln_args(flags + ["p", "q"])
    buf = io.BytesIO()
    sai.write_header(buf, opt)
    ob, orb = pyoracle.as_orc_bwt(g1_index[0]), pyoracle.as_orc_bwt(g1_index[1])
    n_reads = 0
    for lens, offs, codes in engine.read_batches_native(bam, opt.mode, opt.trim_qual):
        n_aln, rec, _ = pyoracle.aln_batch(ob, orb, lens, offs, codes, opt.to_c())
        sai.write_batch(buf, n_aln, rec)
        n_reads += len(lens)
    assert n_reads > 100
    assert buf.getvalue() == open(ref_sai, "rb").read()
