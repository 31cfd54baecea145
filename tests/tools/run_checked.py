#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Runs the golden option sets, the three-pass flow and a batch of random-genome reads through
whatever library B200ALN_LIB names (tests/test_gpu_checked.py points it at libb200aln_checked.so, the -DB2_CHECKED
build of the same sources) and compares with the reference's golden .sai bytes.  A bounds violation makes the
library abort; this script prints one line per case and exits non-zero on the first difference."""
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

from cases import CASES  # noqa: E402
from ibwa_b200 import bwt_restore_bwt, engine, parse_aln_args, sai, seqio  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def engine_sai(e, args, fq):
    opt, _, _, _ = parse_aln_args(args + ["prefix", fq])
    buf = io.BytesIO()
    sai.write_header(buf, opt)
    for batch in seqio.read_batches(fq, opt.mode, opt.trim_qual):
        n_aln, rec = e.cal_sa_reg_gap(batch.lens, batch.offs, batch.codes, opt)
        sai.write_batch(buf, n_aln, rec)
    return buf.getvalue()


def main():
    L = engine.load_library()
    ver = L.b200aln_version().decode()
    print("library:", engine.LIB_PATH, "|", ver)
    if "--expect-checked" in sys.argv and "B2_CHECKED" not in ver:
        print("not the bounds-checked build")
        return 2
    bwt, rbwt = bwt_restore_bwt(os.path.join(G, "g1.bwt")), bwt_restore_bwt(os.path.join(G, "g1.rbwt"))
    bad = 0
    # (knobs, cases): the fast pass alone; tiny arenas so that the middle and the wide pass run; one warp per block
    plans = [({}, sorted(CASES)),
             ({"arena_cap": 64, "rec_cap": 1}, ["default", "stress", "N_n2", "m200"]),
             ({"arena_cap": 64, "rec_cap": 1, "arena_cap_mid": 256, "rec_cap_mid": 3}, ["default", "stress"]),
             ({"search_block": 32, "q16": 0}, ["default", "stress", "short_o3"]),
             ({"lut_k": 3}, ["default", "L_e3"]),
             # parking: every warp gives up as soon as one of its lanes is done, 24 rounds of save / resume
             ({"susp": 31, "susp_min": 0}, ["default", "stress", "m200", "short_o3"]),
             ({"susp": 0}, ["default", "stress"])]
    for knobs, tags in plans:
        with engine.Engine(bwt, rbwt, 0) as e:
            for k, v in knobs.items():
                e.set(k, v)
            for tag in tags:
                args, fq = CASES[tag]
                got = engine_sai(e, args, os.path.join(G, fq + ".fq.gz"))
                ok = got == open(os.path.join(G, f"g1_{tag}.sai"), "rb").read()
                print(f"{'ok ' if ok else 'BAD'} {tag} {knobs}")
                bad += not ok
    # reads of one length on the same index (16-bit width records active), against the oracle port
    from oracle import pyoracle
    opt, _, _, _ = parse_aln_args(["p", "q"])
    batch = next(seqio.read_batches(os.path.join(G, "g1_reads.fq.gz"), opt.mode, opt.trim_qual))
    sel = [i for i, l in enumerate(batch.lens) if l <= 150]
    lens = batch.lens[sel].astype(np.int32)
    codes = np.concatenate([batch.codes[batch.offs[i]:batch.offs[i] + batch.lens[i]] for i in sel])
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    o_n, o_rec, _ = pyoracle.aln_batch(pyoracle.as_orc_bwt(bwt), pyoracle.as_orc_bwt(rbwt), lens, offs, codes, opt.to_c())
    with engine.Engine(bwt, rbwt, 0) as e:
        n_aln, rec = e.cal_sa_reg_gap(lens, offs, codes, opt)
    ok = bool(np.array_equal(n_aln, o_n) and rec.tobytes() == o_rec.tobytes())
    print(f"{'ok ' if ok else 'BAD'} reads <= 150 bp vs the oracle ({len(sel)} reads, 16-bit width records)")
    bad += not ok
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
