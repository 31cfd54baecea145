#!/usr/bin/env python
"""Config-3 spot check on the GPU box: 150 bp stress reads (2 % substitutions + indels) with
`-n 4 -o 2 -e 10 -l 32 -k 2` against the cached 3.1 Gbp bench index; engine vs the reference binary.
usage: python tests/tools/stress_check.py [n_reads]   (run bench.py once before: it builds the index cache)"""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from ibwa_b200 import engine, parse_aln_args, sai, synth  # noqa: E402


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    genome_bp, seed = 3_100_000_000, 20260102
    dev = torch.device("cuda", 0)
    bwt, rbwt, text, prefix = bench.load_or_build_index(genome_bp, seed, dev, True)
    # stress reads need the text on the host only around the sampled loci: take a 64 Mbp window
    win = text[1_000_000_000:1_064_000_000].cpu().numpy()
    del text
    torch.cuda.empty_cache()
    reads = np.stack(synth.simulate_reads(win, n, 150, 20260103, model="stress"))
    args = ["-n", "4", "-o", "2", "-e", "10", "-l", "32", "-k", "2"]
    opt, _, _, _ = parse_aln_args(args + ["p", "q"])
    lens = np.full(n, 150, np.int32)
    offs = np.arange(n, dtype=np.int64) * 150
    with engine.Engine(bwt, rbwt, 0) as eng:
        eng.cal_sa_reg_gap(lens, offs, reads.reshape(-1), opt)
        t0 = time.perf_counter()
        n_aln, rec = eng.cal_sa_reg_gap(lens, offs, reads.reshape(-1), opt)
        dt = time.perf_counter() - t0
        st = eng.stats()
    print(f"engine: {n / dt:.0f} reads/s (batch of {n}), overflow_reads={st['overflow_reads']}, "
          f"pops/read={st['pops'] / n:.0f}, sectors/read={st['occ_lookups'] / n:.0f}, search {st['ms_search']:.1f} ms")
    fq = "/tmp/stress.fq"
    bench.write_fastq(fq, reads)
    t0 = time.perf_counter()
    with open("/tmp/stress_ref.sai", "wb") as fo:
        subprocess.run([bench.REF_BIN, "aln", "-t", str(os.cpu_count())] + args + [prefix, fq], stdout=fo,
                       stderr=subprocess.DEVNULL, check=True)
    print(f"reference -t {os.cpu_count()}: {n / (time.perf_counter() - t0):.0f} reads/s incl. index load")
    _, r_n, r_rec = sai.read_sai("/tmp/stress_ref.sai")
    same = np.array_equal(r_n, n_aln) and r_rec.tobytes() == rec.tobytes()
    print("identical to the reference binary:", same, "| aligned fraction", float((n_aln > 0).mean()),
          "| max n_aln", int(n_aln.max()))
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
