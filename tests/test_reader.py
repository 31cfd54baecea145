"""The native FASTA/FASTQ reader (b200aln_reader_*, fast path + exact state machine) against the Python
mirror of the reference parser (ibwa_b200/seqio.py, itself pinned by the golden .sai tests: -q, lower case,
odd reads) on ordinary and hostile inputs.  No GPU needed."""
import gzip
import os

import numpy as np
import pytest

from ibwa_b200 import engine, seqio


def both(path, mode=3, trim_qual=0, n_needed=0x40000):
    a = [(b.lens, b.offs, b.codes) for b in seqio.read_batches(path, mode, trim_qual, n_needed)]
    b = list(engine.read_batches_native(path, mode, trim_qual, n_needed))
    assert len(a) == len(b)
    for (l1, o1, c1), (l2, o2, c2) in zip(a, b):
        assert np.array_equal(l1, l2) and np.array_equal(o1, o2) and np.array_equal(c1, c2)
    return sum(len(x[0]) for x in a)


@pytest.mark.parametrize("fq,tq", [("g1_reads.fq.gz", 0), ("g1_reads.fq.gz", 20), ("g1_short.fq.gz", 0)])
def test_golden_inputs(golden_dir, fq, tq):
    assert both(os.path.join(golden_dir, fq), trim_qual=tq) > 100


def test_batch_boundaries(golden_dir):
    assert both(os.path.join(golden_dir, "g1_reads.fq.gz"), n_needed=257) == 1306


HOSTILE = {
    "multiline_fasta": b">a desc\nACGT\nACGTNN\n>b\nGGG\n\n>c\n",
    "crlf": b"@r1\r\nACGT\r\n+\r\nIIII\r\n@r2\r\nGGCC\r\n+\r\nIIII\r\n",
    "at_in_quality": b"@r1\nACGTACGT\n+\n@@@@IIII\n@r2\nTTTT\n+r2\n@III\n",
    "truncated_quality": b"@r1\nACGT\n+\nIIII\n@r2\nACGTAC\n+\nII",
    "no_trailing_newline": b"@r1\nACGT\n+\nIIII\n@r2\nAC\n+\nII",
    "lower_and_dash": b"@x/1\nacgtn-ACGT\n+\nIIIIIIIIII\n",
    "junk_before_header": b"\n\n  garbage\n@r1 comment here\nACGT\n+anything\nIIII\n",
    "multiline_fastq": b"@r1\nACGT\nACGT\n+\nIIII\nIIII\n@r2\nAA\n+\nII\n",
    "empty_sequence": b"@r1\n\n+\n\n@r2\nACGT\n+\nIIII\n",
    "plus_in_header": b"@r+1\nACGT\n+\nII+I\n",
}


def _guess_traps():
    """Runs of ordinary 100-bp records (long enough for the structural scan, which guesses each sequence line's
    length from the previous record) with records in between that make the guess land on a newline although
    the line is shorter, or that differ in other ways only the character checks / the exact parser can see."""
    rng = np.random.default_rng(12)
    nt = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = []

    def rec(i, L, plus=b"+", qual=None, seq=None):
        sq = nt[rng.integers(0, 4, size=L)].tobytes() if seq is None else seq
        return b"@t%d\n" % i + sq + b"\n" + plus + b"\n" + (b"I" * L if qual is None else qual) + b"\n"

    for i in range(400):
        k = i % 40
        if k == 11:      # 48 + 1 + 3 + 48 = 100: the byte at seq + 100 is the newline after the quality string
            out.append(rec(i, 48, plus=b"+x"))
        elif k == 17:    # two-line sequence whose first line is 100 long: fine for the guess, not a 4-line record
            out.append(b"@t%d\n" % i + b"A" * 100 + b"\n" + b"C" * 20 + b"\n+\n" + b"I" * 120 + b"\n")
        elif k == 23:    # quality string with a newline inside the guessed region
            out.append(b"@t%d\n" % i + b"G" * 100 + b"\n+\n" + b"I" * 60 + b"\n" + b"I" * 40 + b"\n")
        elif k == 29:    # shorter read, then the ordinary length again
            out.append(rec(i, 36))
        elif k == 31:    # '>' inside the sequence line ends the sequence for the reference parser
            out.append(rec(i, 100, seq=b"ACGT" * 10 + b">" + b"ACGT" * 14 + b"ACG"))
        else:
            out.append(rec(i, 100))
    return b"".join(out)


HOSTILE["guess_traps"] = _guess_traps()


@pytest.mark.parametrize("name", sorted(HOSTILE))
@pytest.mark.parametrize("gz", [False, True])
def test_hostile_inputs(tmp_path, name, gz):
    p = str(tmp_path / (name + (".gz" if gz else "")))
    data = HOSTILE[name]
    with (gzip.open(p, "wb") if gz else open(p, "wb")) as f:
        f.write(data)
    both(p)


def test_barcode_and_il13(tmp_path):
    p = str(tmp_path / "bc.fq")
    with open(p, "wb") as f:
        f.write(b"@r1\nACGTACGTAC\n+\nhhhhhhhhhh\n@r2\nAC\n+\nhh\n@r3\nACGTA\n+\nhhhhh\n")
    both(p, mode=3 | (3 << 24))
    both(p, mode=3 | 0x200, trim_qual=5)


def test_large_file_spans_buffers(tmp_path):
    rng = np.random.default_rng(1)
    p = str(tmp_path / "big.fq")
    nt = np.frombuffer(b"ACGTN", dtype=np.uint8)
    with open(p, "wb") as f:
        for i in range(60000):
            L = int(rng.integers(30, 160))
            s = nt[rng.integers(0, 5, size=L)].tobytes()
            f.write(b"@read%d\n" % i + s + b"\n+\n" + b"I" * L + b"\n")
    assert both(p) == 60000


# (the reference itself crashes on a first record with an empty sequence: kseq.h:176 writes through a null buffer)
@pytest.mark.parametrize("name", sorted(set(HOSTILE) - {"empty_sequence"}))
def test_hostile_inputs_against_reference_binary(tmp_path, golden_dir, g1_index, name):
    """Pins the Python mirror of the parser itself: the reference binary's .sai on each hostile input
    (record count and per-read results) must equal the oracle run on what seqio parsed."""
    import io
    from ibwa_b200 import gap_init_opt, sai
    from oracle import pyoracle
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref/ibwa not present")
    prefix = str(tmp_path / "g1")
    os.symlink(os.path.join(golden_dir, "g1.bwt"), prefix + ".bwt")
    os.symlink(os.path.join(golden_dir, "g1.rbwt"), prefix + ".rbwt")
    p = str(tmp_path / (name + ".fq"))
    with open(p, "wb") as f:
        f.write(HOSTILE[name])
    ref_sai = str(tmp_path / "ref.sai")
    pyoracle.run_ref(["aln", prefix, p], stdout_path=ref_sai)
    opt = gap_init_opt()
    buf = io.BytesIO()
    sai.write_header(buf, opt)
    ob, orb = pyoracle.as_orc_bwt(g1_index[0]), pyoracle.as_orc_bwt(g1_index[1])
    for b in seqio.read_batches(p, opt.mode, opt.trim_qual):
        n_aln, rec, _ = pyoracle.aln_batch(ob, orb, b.lens, b.offs, b.codes, opt.to_c())
        sai.write_batch(buf, n_aln, rec)
    assert buf.getvalue() == open(ref_sai, "rb").read()


def test_parallel_path_falls_back_on_odd_records(tmp_path):
    """Long runs of ordinary records (parallel scan + convert) interrupted by records only the exact state
    machine may parse: multi-line, CRLF, '@'-leading quality, lower case, a FASTA record."""
    rng = np.random.default_rng(2)
    p = str(tmp_path / "mixed.fq")
    nt = np.frombuffer(b"ACGTN", dtype=np.uint8)
    odd = [b"@odd1\nACGT\nACGT\n+\nIIII\nIIII\n", b"@odd2\r\nACGTAC\r\n+\r\nIIIIII\r\n", b"@odd3\nACGTACGT\n+\n@@IIIIII\n",
           b">fa1 x\nacgtnACGT\n", b"@odd4 c\nAC GT\n+\nIIII\n"]
    with open(p, "wb") as f:
        for i in range(30000):
            if i % 4100 == 4099:
                f.write(odd[(i // 4100) % len(odd)])
            L = int(rng.integers(36, 120))
            s = nt[rng.integers(0, 5, size=L)].tobytes()
            q = bytes(rng.integers(35, 74, size=L).astype(np.uint8))
            f.write(b"@r%d/1\n" % i + s + b"\n+\n" + q + b"\n")
    assert both(p) > 30000
    assert both(p, trim_qual=25) > 30000
    assert both(p, mode=3 | 0x200, trim_qual=10, n_needed=9973) > 30000


def bgzf_bytes(data: bytes, block=60000, level=6) -> bytes:
    """`data` as a BGZF stream (SAM spec 4.1): gzip members of <= 64 KB that carry their size in a 'BC' extra
    field, then the empty end-of-file member."""
    import struct
    import zlib
    out = []
    for off in list(range(0, len(data), block)) + [None]:
        chunk = b"" if off is None else data[off:off + block]
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        comp = co.compress(chunk) + co.flush()
        bsize = 12 + 6 + len(comp) + 8
        out.append(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1) +
                   comp + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    return b"".join(out)


@pytest.mark.parametrize("block", [60000, 777, 65280])
def test_bgzf_fastq_is_inflated_block_parallel(tmp_path, block):
    """A bgzip-compressed FASTQ goes through the block-parallel inflater; the records must be the ones the
    zlib path (B200ALN_NO_BGZF) and the Python mirror deliver — across block and window boundaries."""
    rng = np.random.default_rng(3)
    nt = np.frombuffer(b"ACGTN", dtype=np.uint8)
    recs = []
    for i in range(30000):
        L = int(rng.integers(30, 160))
        recs.append(b"@read%d/1\n" % i + nt[rng.integers(0, 5, size=L)].tobytes() + b"\n+\n" + b"I" * L + b"\n")
    data = b"".join(recs) + HOSTILE["multiline_fastq"]
    p = str(tmp_path / "r.fq.gz")
    with open(p, "wb") as f:
        f.write(bgzf_bytes(data, block=block, level=1))
    assert both(p) == 30002
    assert both(p, n_needed=4099) == 30002
    os.environ["B200ALN_NO_BGZF"] = "1"
    try:
        assert both(p) == 30002
    finally:
        del os.environ["B200ALN_NO_BGZF"]


def test_parallel_scan_joins_only_matching_pieces(tmp_path):
    """Large enough for the parallel structural scan (one slice of the byte range per worker): the trap records
    of `guess_traps` every few hundred records — odd records end a worker's piece early, so the pieces after it
    must be dropped and scanned again — plus a quality line that looks like a header at a likely slice border."""
    unit = HOSTILE["guess_traps"]
    evil = b"@e\n" + b"ACGT" * 25 + b"\n+\n" + b"@" + b"I" * 99 + b"\n"      # quality starts with '@'
    data = b"".join(unit + evil * 3 for _ in range(60))
    p = str(tmp_path / "par.fq")
    with open(p, "wb") as f:
        f.write(data)
    assert len(data) > (4 << 20)
    assert both(p) == 60 * 403
    assert both(p, trim_qual=20) == 60 * 403


def _random_character_fastq(path, seed):
    """4-line records whose sequence and quality lines hold every byte the fast path has to classify: letters of both
    cases, '-', '.', digits, DEL, bytes >= 128, and now and then one of the characters the fast path must refuse
    ('>', '+', '@', a blank); lengths around the 32-byte vector width and the 35-base trimming floor"""
    rng = np.random.default_rng(seed)
    ok = np.array([c for c in range(33, 127) if c not in b">+@"], dtype=np.uint8)
    common = np.frombuffer(b"ACGTacgtNn-.", dtype=np.uint8)
    out = []
    for i in range(6000):
        L = int(rng.choice([1, 2, 15, 31, 32, 33, 34, 35, 36, 63, 64, 65, 95, 96, 97, 100, 127, 128, 129, 250]))
        s = np.where(rng.random(L) < 0.8, rng.choice(common, L), rng.choice(ok, L)).astype(np.uint8)
        q = rng.integers(33, 128, size=L).astype(np.uint8)            # 127 is a legal quality character (kseq.h:185)
        k = rng.random()
        if k < 0.02:
            s[int(rng.integers(0, L))] = int(rng.choice(np.frombuffer(b">+@ \t", dtype=np.uint8)))
        elif k < 0.04:
            s[int(rng.integers(0, L))] = int(rng.choice([127, 128, 200, 255, 1, 31]))
        elif k < 0.06:
            q[int(rng.integers(0, L))] = int(rng.choice([32, 128, 255, 10 if L > 1 else 32]))
        out.append(b"@r%d x\n" % i + s.tobytes() + b"\n+\n" + q.tobytes() + b"\n")
    with open(path, "wb") as f:
        f.write(b"".join(out))


@pytest.mark.parametrize("simd", [True, False], ids=["avx2", "table"])
def test_every_character_class_vectorised_and_not(tmp_path, simd):
    """the vectorised conversion of the fast path (nibble table + class checks) and the table version give what the
    reference's parser gives, on every byte value and on lengths around the vector width; each in a process of its
    own, since the choice is made once per process (B200ALN_NO_SIMD)"""
    import subprocess
    import sys
    p = str(tmp_path / "chars.fq")
    _random_character_fastq(p, 5)
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import test_reader as t; "
            "print(t.both(%r), t.both(%r, trim_qual=25), t.both(%r, mode=3 | 0x200, trim_qual=10))"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)), p, p, p))
    env = dict(os.environ, B200ALN_PAR_SCAN_MIN="4096")
    if not simd:
        env["B200ALN_NO_SIMD"] = "1"
    r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    counts = r.stdout.decode().split()
    assert len(counts) == 3 and int(counts[0]) > 1000


def _native(path, fast):
    """every batch of the native reader on `path`, with the gzip stream decoded by fast_inflate.h or by zlib"""
    if fast:
        os.environ.pop("B200ALN_NO_FAST_INFLATE", None)
    else:
        os.environ["B200ALN_NO_FAST_INFLATE"] = "1"
    try:
        return [(l.copy(), o.copy(), c.copy()) for l, o, c in engine.read_batches_native(path, 3, 0, 0x40000)]
    finally:
        os.environ.pop("B200ALN_NO_FAST_INFLATE", None)


def _same(a, b):
    return len(a) == len(b) and all(np.array_equal(x, y) for p, q in zip(a, b) for x, y in zip(p, q))


def test_gzip_streams_through_the_fast_decoder(tmp_path):
    """fast_inflate.h under the reader: a stream of several 16 MB chunks, several gzip members in one file, a header
    with a file name and an extra field, bytes after the last member, a wrong CRC, a wrong length, a truncated file —
    always what the zlib path (gzread) delivers, which is what the reference reads (utils.c:56-66)"""
    import zlib
    rng = np.random.default_rng(9)
    nt = np.frombuffer(b"ACGT", dtype=np.uint8)
    n = 110_000
    seqs = nt[rng.integers(0, 4, size=(n, 100))]
    quals = np.clip(rng.normal(70, 3, size=(n, 100)), 35, 74).astype(np.uint8)
    plain = b"".join(b"@q%d\n" % i + seqs[i].tobytes() + b"\n+\n" + quals[i].tobytes() + b"\n" for i in range(n))
    assert len(plain) > (20 << 20)

    def member(data, level=6, extra=False):
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        body = c.compress(data) + c.flush()
        flg = (4 | 8) if extra else 0
        hdr = b"\x1f\x8b\x08" + bytes([flg]) + b"\0\0\0\0\0\x03"
        if extra:
            hdr += b"\x06\x00XY\x02\x00ab" + b"reads.fq\0"
        return hdr + body + (zlib.crc32(data) & 0xffffffff).to_bytes(4, "little") + (len(data) & 0xffffffff).to_bytes(4, "little")

    cut = plain.index(b"\n@q60000\n") + 1
    whole = member(plain)
    cases = {
        "one_member": whole,
        "three_members_and_a_name": member(plain[:cut], 1, True) + member(b"") + member(plain[cut:], 9),
        "garbage_after": member(plain[:cut]) + b"\0\0\0\0 not a gzip member",
        "wrong_crc": whole[:-8] + bytes([whole[-8] ^ 1]) + whole[-7:],
        "wrong_length": whole[:-1] + bytes([whole[-1] ^ 1]),
        "truncated": whole[:len(whole) * 2 // 3],
        "damaged_in_the_middle": whole[:len(whole) // 2] + bytes([whole[len(whole) // 2] ^ 0x10]) + whole[len(whole) // 2 + 1:],
    }
    for name, blob in cases.items():
        p = str(tmp_path / (name + ".fq.gz"))
        with open(p, "wb") as f:
            f.write(blob)
        fast, slow = _native(p, True), _native(p, False)
        if name in ("wrong_crc", "wrong_length", "damaged_in_the_middle"):
            # gzread reports a failed check with the call in which it shows and drops that call's bytes; the calls are
            # 16 MB here and the fast path's chunks end a few hundred bytes earlier, so both lose the stream's last
            # chunk but not at the same byte (the reference, reading 4 KB at a time, loses less: bwaseqio / kseq.h)
            cf, cs = np.concatenate([b[2] for b in fast]), np.concatenate([b[2] for b in slow])
            m = min(len(cf), len(cs))
            assert m > 7_000_000 and np.array_equal(cf[:m], cs[:m]) and abs(len(cf) - len(cs)) < 100_000, name
            continue
        assert _same(fast, slow), name
        if name in ("one_member", "three_members_and_a_name"):
            assert sum(len(b[0]) for b in fast) == n, name
    # ... and the same file block-compressed (BGZF): every block through the fast decoder, then through zlib
    p = str(tmp_path / "blocks.fq.gz")
    with open(p, "wb") as f:
        f.write(bgzf_bytes(plain[:cut], block=60000))
    fast, slow = _native(p, True), _native(p, False)
    assert _same(fast, slow) and sum(len(b[0]) for b in fast) == 60000
