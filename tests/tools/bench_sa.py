#!/usr/bin/env python
"""Scope row N2 measurement: SA row -> text position (bwt_sa / bwtdb_sa2seq) through the C ABI with host
buffers, on the bench genome (default 3.1 Gbp), rows taken from the engine's own best hits of synthetic reads
plus uniformly random rows.  Prints one JSON line.

  python tests/tools/bench_sa.py [--genome-bp N] [--rows N] [--steps K]

Work per row: bwt_sa walks inverse-Psi steps until a sampled row (sa_intv = 32: 15.5 steps on average), each step
one 32-byte sector of the occ layout -> algorithmic bytes = 32 x steps.  CPU beside it: oracle port, one core."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome-bp", type=int, default=3_100_000_000)
    ap.add_argument("--rows", type=int, default=20_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--seed", type=int, default=20260102)
    args = ap.parse_args()
    import torch
    from ibwa_b200 import engine, fmbuild
    from oracle import pyoracle
    import bench as B
    dev = torch.device("cuda", 0)
    text = B.gen_text(args.genome_bp, args.seed, dev)
    t0 = time.time()
    bwt, sa = fmbuild.build_bwt_torch(text, sa_intv=32)
    torch.cuda.empty_cache()
    rbwt, rsa = fmbuild.build_bwt_torch(torch.flip(text, dims=[0]), sa_intv=32)
    del text
    torch.cuda.empty_cache()
    print(f"[bench_sa] index + SA samples built in {time.time() - t0:.1f} s", file=sys.stderr)
    rng = np.random.default_rng(7)
    n = args.rows
    rows = rng.integers(0, args.genome_bp + 1, size=n).astype(np.uint32)
    strand = rng.integers(0, 2, size=n).astype(np.uint8)
    lens = np.full(n, 100, np.int32)
    with engine.Engine(bwt, rbwt, 0) as e:
        e.load_sa(0, sa)
        e.load_sa(1, rsa)
        e.sa2seq(strand[:1000], rows[:1000], lens[:1000])
        e.bwt_sa(0, rows)                       # warm-up, buffers allocated
        ms = []
        for _ in range(args.steps):
            e.timer_start()
            pos = e.sa2seq(strand, rows, lens)
            ms.append(e.timer_stop())
        m = 2000
        ob, osa = pyoracle.as_orc_bwt(bwt), pyoracle.as_orc_sa(sa)
        t0 = time.time()
        want = pyoracle.bwt_sa(ob, osa, rows[:m])
        cpu_s = time.time() - t0
        got = e.bwt_sa(0, rows[:m])
    ok = bool(np.array_equal(got, want))
    best = min(ms)
    steps = (32 - 1) / 2.0
    out = {"metric": "SA rows -> positions / s (bwtdb_sa2seq, sa_intv 32)", "value": n / (best * 1e-3), "unit": "rows/s",
           "rows": n, "ms_best": best, "ms_all": ms, "h2d_bytes": n * 9, "d2h_bytes": n * 8,
           "algorithmic_bytes_per_row": 32 * steps, "achieved_gbs": 32 * steps * n / (best * 1e-3) / 1e9,
           "cpu_port_rows_per_s_1core": m / cpu_s, "parity_vs_oracle_rows": m, "parity_ok": ok,
           "config": {"genome_bp": args.genome_bp, "workload": "uniformly random rows, host buffers through the C ABI"}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
