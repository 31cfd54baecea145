"""N>1 host logic on CPU: two gloo ranks each search their contiguous shard (kernel logic
compiled for the CPU by tests/harness), rank 0 merges in input order; bytes must equal the
reference's single-process .sai — including the batch-level max_gapo clamp (short reads, -o 3)."""
import io
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from ibwa_b200 import bwt_restore_bwt, parse_aln_args, sai, seqio, shard
from cases import CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_shard_ranges_cover_in_order():
    for n in (0, 1, 7, 100, 262144):
        for w in (1, 2, 3, 8):
            r = [shard.shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def _worker(rank, world, port, tag, out_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from harness import pyharness
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    args, fq = CASES[tag]
    opt, _, _, _ = parse_aln_args(args + ["p", "q"])
    bwt, rbwt = bwt_restore_bwt(os.path.join(GOLDEN, "g1.bwt")), bwt_restore_bwt(os.path.join(GOLDEN, "g1.rbwt"))
    batch = next(seqio.read_batches(os.path.join(GOLDEN, fq + ".fq.gz"), opt.mode, opt.trim_qual))
    lo, hi = shard.shard_range(len(batch), world, rank)
    lens, offs, codes = shard.take_shard(batch.lens, batch.offs, batch.codes, lo, hi)
    n_aln, rec, nov, _ = pyharness.aln_batch(bwt, rbwt, lens, offs, codes, opt.to_c(), arena_cap=32000, rec_cap=4096,
                                             batch_max_len=int(batch.lens.max()))
    assert nov == 0
    parts = [None] * world if rank == 0 else None
    dist.gather_object((n_aln, rec), parts, dst=0)
    if rank == 0:
        n_all, r_all = shard.gather_in_order(parts)
        buf = io.BytesIO()
        sai.write_header(buf, opt)
        sai.write_batch(buf, n_all, r_all)
        open(out_path, "wb").write(buf.getvalue())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("tag", ["default", "short_o3"])
def test_two_rank_sharding_matches_reference(tag, tmp_path):
    out = str(tmp_path / "merged.sai")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, tag, out), nprocs=2, join=True)
    assert open(out, "rb").read() == open(os.path.join(GOLDEN, f"g1_{tag}.sai"), "rb").read()


def test_shard_without_batch_max_len_differs_or_not():
    """Documents why batch_max_len exists: for the short-read fixture the clamp depends on max_len."""
    from ibwa_b200.opts import bwa_cal_maxdiff
    assert bwa_cal_maxdiff(37) == 2 and bwa_cal_maxdiff(38) == 3     # max_gapo=3 is clamped to 2 below 38 bp
