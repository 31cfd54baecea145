#!/usr/bin/env python
"""Repeats `b200aln aln` on one FASTQ (the cached 3.1 Gbp bench index) under different driver settings and checks that
every run writes the same bytes; on a difference it says which read differs and how.  Needs a GPU box.
usage: python tests/tools/cli_stress.py [n_reads] [n_runs]"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ENVS = [{}, {"B200ALN_MERGE": "1"}, {"B200ALN_MERGE": "16"}, {"B200ALN_INFLIGHT": "3"}, {"B200ALN_NO_PREALLOC": "1"},
        {"B200ALN_INFLIGHT": "6", "B200ALN_MERGE": "2"}, {"B200ALN_NO_PIN": "1"}, {"B200ALN_INFLIGHT": "1"}]
if os.environ.get("CLI_STRESS_ENVS"):  # e.g. "B200ALN_INFLIGHT=6,B200ALN_SET=susp=0;B200ALN_INFLIGHT=2" (runs cycle through them)
    ENVS = [dict(kv.split("=", 1) for kv in one.split(",") if kv) for one in os.environ["CLI_STRESS_ENVS"].split(";")]


def per_read(words):
    """n_aln of every read and the word where its part starts"""
    n_aln, start = [], []
    p, nw = 0, len(words)
    w = words.tolist()
    while p < nw:
        n_aln.append(w[p])
        start.append(p)
        p += 1 + 4 * w[p]
    return np.array(n_aln, dtype=np.int64), np.array(start, dtype=np.int64)


def read_index(words, upto):
    """index of the read whose part of the .sai stream holds word `upto`, and where that part starts"""
    p = r = 0
    while True:
        n = int(words[p])
        if p + 1 + 4 * n > upto:
            return r, p
        p += 1 + 4 * n
        r += 1


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    runs = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    genome_bp, seed = 3_100_000_000, 20260102
    dev = torch.device("cuda", 0)
    bwt, rbwt, text, prefix = bench.load_or_build_index(genome_bp, seed, dev, True)
    reads = bench.synth_reads_torch(text, n, 100, seed + 7).cpu().numpy()
    del text, bwt, rbwt
    torch.cuda.empty_cache()
    fq = "/tmp/cli_stress.fq"
    bench.write_fastq(fq, reads)
    os.sync()
    exe = os.path.join(ROOT, "ibwa_b200", "b200aln")
    first = None
    bad = 0
    for i in range(runs):
        env = ENVS[i % len(ENVS)]
        out = "/tmp/stress_first.sai" if first is None else "/tmp/stress_out.sai"
        subprocess.run([exe, "aln", "-f", out, prefix, fq], stderr=subprocess.DEVNULL, check=True, env=dict(os.environ, **env))
        a = np.fromfile(out, dtype=np.uint32)
        if first is None:
            first = a
            print(f"run {i} {env}: {a.nbytes} bytes", flush=True)
            continue
        if len(a) == len(first) and np.array_equal(a, first):
            print(f"run {i} {env}: same", flush=True)
            continue
        bad += 1
        m = min(len(a), len(first))
        d = np.flatnonzero(a[:m] != first[:m])
        print(f"run {i} {env}: DIFFERS, sizes {first.nbytes} / {a.nbytes}, {len(d)} differing words, first at word "
              f"{d[0] if len(d) else m}, last at {d[-1] if len(d) else m}", flush=True)
        if len(d):
            w0 = int(d[0])
            r, p = read_index(first[16:], w0 - 16)
            nf = int(first[16 + p])
            print(f"  read {r} (reference batch {r // 0x40000}, {r % 0x40000} into it): first run n_aln {nf}, records "
                  f"{first[16 + p + 1:16 + p + 1 + 4 * nf].tolist()}")
            print(f"  this run, same place: {a[16 + p:16 + p + 1 + 4 * max(nf, 2)].tolist()}")
            na, sa = per_read(first[16:])
            nb, sb = per_read(a[16:])
            if len(na) == len(nb):
                dn = np.flatnonzero(na != nb)
                print(f"  reads whose n_aln differs: {len(dn)}; first {dn[:12].tolist()}; by reference batch "
                      f"{dict(zip(*[x.tolist() for x in np.unique(dn // 0x40000, return_counts=True)]))}")
                print(f"  n_aln there, first run: {na[dn[:24]].tolist()}")
                print(f"  n_aln there, this run:  {nb[dn[:24]].tolist()}")
                same_n = np.flatnonzero(na == nb)
                # reads with equal n_aln but different records
                diff_rec = 0
                for r_ in same_n[:: max(1, len(same_n) // 200000)]:
                    if not np.array_equal(first[16 + sa[r_]:16 + sa[r_] + 1 + 4 * na[r_]], a[16 + sb[r_]:16 + sb[r_] + 1 + 4 * nb[r_]]):
                        diff_rec += 1
                print(f"  of a sample of reads with equal n_aln, records differ for {diff_rec}")
                if len(dn):
                    seqs = reads[dn[:6]]
                    print("  the first of them:", ["".join("ACGT"[c] for c in s_) for s_ in seqs])
            else:
                print(f"  read counts differ: {len(na)} / {len(nb)}")
        os.replace(out, f"/tmp/stress_bad_{i}.sai")
    print("runs that differ from the first:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
