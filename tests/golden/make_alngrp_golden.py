#!/usr/bin/env python
"""Golden vectors for scope row N4 (multi-.sai merge): synthetic .sai streams run through the UNMODIFIED
reference functions saiset_create / alngrp_create (saiset.c:15-78) by oracle/_ref/alngrp_dump (our driver,
oracle/ref_shim/alngrp_dump.c, linked against the reference objects).  Run in the build container only:

  python tests/golden/make_alngrp_golden.py      ->  tests/golden/alngrp_{1,2,3}.npz
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from ibwa_b200 import gap_init_opt, sai  # noqa: E402

DUMP = os.path.join(ROOT, "oracle", "_ref", "alngrp_dump")


def make_streams(n_streams, n_reads, seed):
    """Group sizes from 0 to ~2000 per stream, few distinct scores (many ties), some already sorted /
    reversed / organ-pipe runs so that the quicksort, its <= 16 leftovers and the insertion pass all matter."""
    rng = np.random.default_rng(seed)
    n_alns, recs = [], []
    for s in range(n_streams):
        sizes = rng.choice([0, 0, 1, 1, 2, 3, 5, 8, 13, 17, 18, 40, 100], size=n_reads)
        big = rng.choice(n_reads, size=3, replace=False)
        sizes[big] = [300, 1000, 2000][: len(big)]
        total = int(sizes.sum())
        r = np.zeros(total, dtype=sai.ALN_DTYPE)
        r["packed"] = rng.integers(0, 1 << 25, total)
        r["k"] = np.arange(total, dtype=np.uint32) + 1_000_000 * s        # unique: makes the permutation visible
        r["l"] = r["k"] + rng.integers(0, 3, total).astype(np.uint32)
        score = rng.choice([0, 3, 6, 9, 11, 12, 14, 15, 18, 22], size=total).astype(np.int32)
        at = 0
        for i, c in enumerate(sizes):
            seg = score[at:at + c]
            kind = i % 5
            if kind == 1:
                seg.sort()
            elif kind == 2:
                seg[::-1].sort()
            elif kind == 3 and c > 2:
                seg.sort()
                seg[:] = np.concatenate([seg[::2], seg[1::2][::-1]])
            elif kind == 4:
                seg[:] = rng.integers(0, 2000, c)                         # nearly all distinct
            at += c
        r["score"] = score
        n_alns.append(sizes.astype(np.int32))
        recs.append(r)
    return n_alns, recs


def run_reference(n_alns, recs, tmp, s_mm):
    paths = []
    opt = gap_init_opt()
    opt.s_mm = s_mm                   # alngrp_create takes the cut from the first header (saiset.c:70)
    for s, (na, rc) in enumerate(zip(n_alns, recs)):
        p = os.path.join(tmp, f"s{s}.sai")
        with open(p, "wb") as f:
            sai.write_header(f, opt)
            sai.write_batch(f, na, rc)
        paths.append(p)
    out = subprocess.run([DUMP, str(len(n_alns[0]))] + paths, stdout=subprocess.PIPE, check=True).stdout
    w = np.frombuffer(out, dtype=np.uint32)
    n = len(n_alns[0])
    out_n = np.zeros(n, np.int32)
    db, rec = [], []
    p = 0
    for r in range(n):
        c = int(w[p]); p += 1
        out_n[r] = c
        blk = w[p:p + 5 * c].reshape(c, 5)
        p += 5 * c
        db.append(blk[:, 0].copy())
        rec.append(blk[:, 1:].copy())
    assert p == len(w)
    db = np.concatenate(db) if db else np.empty(0, np.uint32)
    rec = np.concatenate(rec).reshape(-1).view(sai.ALN_DTYPE) if rec else np.empty(0, sai.ALN_DTYPE)
    return out_n, rec, db


def main():
    for ns, s_mm in ((1, 3), (2, 3), (3, 100000)):   # the last keeps everything: the whole permutation is visible
        n_alns, recs = make_streams(ns, 120, 20260200 + ns)
        with tempfile.TemporaryDirectory() as tmp:
            out_n, rec, db = run_reference(n_alns, recs, tmp, s_mm)
        np.savez_compressed(os.path.join(HERE, f"alngrp_{ns}.npz"), s_mm=s_mm,
                            out_n=out_n, out_rec=rec, out_db=db,
                            **{f"n_aln{s}": n_alns[s] for s in range(ns)}, **{f"rec{s}": recs[s] for s in range(ns)})
        print(ns, "streams:", int(sum(a.sum() for a in n_alns)), "alignments in,", int(out_n.sum()), "out")


if __name__ == "__main__":
    main()
