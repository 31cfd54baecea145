#!/usr/bin/env python
"""bench.py — `bwa aln` reads/s on B200 (BASELINE.json metric), one process per GPU.

Workload (config.workload): BASELINE.json configs[1] — N simulated 100 bp reads
vs a 3.1 Gbp synthetic genome, defaults (-n 0.04), index replicated on every
GPU, reads sharded (weak scaling: every rank aligns --reads reads of its own).
The genome is i.i.d. uniform ACGT generated on the GPU from a seed; its .bwt /
.rbwt are built on the GPU by ibwa_b200.fmbuild (bit-identical to what the
reference's `index` writes, tests/test_fmbuild.py) and cached under
/tmp/b200aln_bench so the reference arm and this arm read the same files.

A step = one pass of the hot path (bwa_cal_sa_reg_gap) over the rank's reads.
  value : reads resident in HBM before the timed region (b200aln_batch_device)
  e2e   : the same reads from pinned HOST buffers through b200aln_batch
          (H2D of reads and D2H of n_aln + records inside the timed region)
Timing: CUDA events on the engine's launch stream (b200aln_timer_*), barrier +
synchronize on both sides, max over ranks.  Index (3.1 GB) and per-read state
(> 10 GB) are far larger than L2, so no L2 flush is needed between steps.

`--impl reference` times the unmodified reference binary (oracle/_ref/ibwa aln
-t <all cores>) on a bounded sample of the same reads on the host cores.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CACHE_ROOT = os.environ.get("B200ALN_BENCH_CACHE", "/tmp/b200aln_bench")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ibwa")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------- data ------

def gen_text(genome_bp: int, seed: int, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    t = torch.empty(genome_bp, dtype=torch.uint8, device=device)
    step = 1 << 28
    for s in range(0, genome_bp, step):
        e = min(genome_bp, s + step)
        t[s:e] = torch.randint(0, 4, (e - s,), dtype=torch.uint8, device=device, generator=g)
    return t


def synth_reads_torch(text, n_reads: int, length: int, seed: int):
    """SURVEY.md §8d default read model on the GPU: uniform start, 1 % substitutions, 2 % of reads
    with one 1-3 bp indel, 50 % reverse-complemented.  Returns uint8 [n_reads, length] (nt4 codes)."""
    import torch
    dev = text.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n = text.numel()
    out = torch.empty((n_reads, length), dtype=torch.uint8, device=dev)
    col = torch.arange(length, device=dev)[None, :]
    CH = 1 << 20
    for s in range(0, n_reads, CH):
        m = min(CH, n_reads - s)
        start = torch.randint(0, n - length - 8, (m,), device=dev, generator=g)
        has = torch.rand(m, device=dev, generator=g) < 0.02
        il = torch.randint(1, 4, (m,), device=dev, generator=g)
        p = torch.randint(10, max(11, length - 10), (m,), device=dev, generator=g)
        isdel = torch.rand(m, device=dev, generator=g) < 0.5
        tail = col >= p[:, None]
        dshift = torch.where(has & isdel, il, torch.zeros_like(il))[:, None] * tail
        ins_len = torch.where(has & ~isdel, il, torch.zeros_like(il))
        ishift = ins_len[:, None] * (col >= (p + il)[:, None])
        idx = start[:, None] + col + dshift - ishift
        r = text[idx]
        insm = (has & ~isdel)[:, None] & tail & (col < (p + il)[:, None])
        rnd = torch.randint(0, 4, (m, length), dtype=torch.uint8, device=dev, generator=g)
        r = torch.where(insm, rnd, r)
        sub = torch.rand((m, length), device=dev, generator=g) < 0.01
        inc = torch.randint(1, 4, (m, length), dtype=torch.uint8, device=dev, generator=g)
        r = torch.where(sub, (r + inc) & 3, r)
        rc = torch.rand(m, device=dev, generator=g) < 0.5
        rcv = 3 - torch.flip(r, dims=[1])
        r = torch.where(rc[:, None], rcv, r)
        out[s:s + m] = r
    return out


def synth_stress_reads_torch(text, n_reads: int, length: int, seed: int):
    """SURVEY.md §8d stress read model (config 3) on the GPU: per base 2 % substitutions, deletion-open
    0.001 and insertion-open 0.001 with geometric extension p = 0.3; 50 % reverse-complemented."""
    import torch
    dev = text.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n = text.numel()
    out = torch.zeros((n_reads, length), dtype=torch.uint8, device=dev)
    j = torch.randint(0, n - 2 * length - 64, (n_reads,), device=dev, generator=g)      # source pointer
    p = torch.zeros(n_reads, dtype=torch.int64, device=dev)                             # bases emitted
    ins_left = torch.zeros(n_reads, dtype=torch.int64, device=dev)
    rows = torch.arange(n_reads, device=dev)
    log03 = float(np.log(0.3))
    for _ in range(length * 2 + 64):
        active = p < length
        if not bool(active.any()):
            break
        u = torch.rand(n_reads, device=dev, generator=g)
        geo = torch.floor(torch.log(torch.rand(n_reads, device=dev, generator=g).clamp_min(1e-12)) / log03).long()
        inserting = ins_left > 0
        is_del = active & ~inserting & (u < 0.001)
        is_ins = active & ~inserting & (u >= 0.001) & (u < 0.002)
        emit_src = active & ~inserting & (u >= 0.002)
        j = j + torch.where(is_del, 1 + geo, torch.zeros_like(geo))
        ins_left = torch.where(is_ins, 1 + geo, ins_left)
        inserting = ins_left > 0
        base = text[j.clamp_max(n - 1)]
        sub = torch.rand(n_reads, device=dev, generator=g) < 0.02
        inc = torch.randint(1, 4, (n_reads,), dtype=torch.uint8, device=dev, generator=g)
        base = torch.where(sub, (base + inc) & 3, base)
        rnd = torch.randint(0, 4, (n_reads,), dtype=torch.uint8, device=dev, generator=g)
        emit_ins = active & inserting
        val = torch.where(emit_ins, rnd, base)
        do = emit_src | emit_ins
        idx = p.clamp_max(length - 1)
        cur = out[rows, idx]
        out[rows, idx] = torch.where(do, val, cur)
        p = p + do.long()
        j = j + emit_src.long()
        ins_left = ins_left - emit_ins.long()
    rc = torch.rand(n_reads, device=dev, generator=g) < 0.5
    rcv = 3 - torch.flip(out, dims=[1])
    return torch.where(rc[:, None], rcv, out)


def cache_dir(genome_bp: int, seed: int) -> str:
    return os.path.join(CACHE_ROOT, f"g{genome_bp}_s{seed}")


def load_or_build_index(genome_bp: int, seed: int, device, is_writer: bool):
    """Returns (bwt, rbwt, text).  The text is always regenerated from the seed on the GPU."""
    import torch
    from ibwa_b200 import fmbuild
    from ibwa_b200.bwtio import bwt_dump_bwt, bwt_restore_bwt
    d = cache_dir(genome_bp, seed)
    prefix = os.path.join(d, "ref")
    text = gen_text(genome_bp, seed, device)
    done = os.path.join(d, "DONE")
    if os.path.exists(done):
        return bwt_restore_bwt(prefix + ".bwt"), bwt_restore_bwt(prefix + ".rbwt"), text, prefix
    t0 = time.time()
    bwt = fmbuild.build_bwt_torch(text)
    torch.cuda.empty_cache()
    rbwt = fmbuild.build_bwt_torch(torch.flip(text, dims=[0]))
    torch.cuda.empty_cache()
    log(f"[bench] built .bwt/.rbwt for {genome_bp} bp on the GPU in {time.time() - t0:.1f} s")
    if is_writer:
        os.makedirs(d, exist_ok=True)
        bwt_dump_bwt(prefix + ".bwt", bwt)
        bwt_dump_bwt(prefix + ".rbwt", rbwt)
        open(done, "w").write("ok\n")
    return bwt, rbwt, text, prefix


def write_fastq(path: str, reads: np.ndarray) -> None:
    nt = np.frombuffer(b"ACGTN-", dtype=np.uint8)
    n, L = reads.shape
    width = len(f"@r{n}") + 1
    with open(path, "wb") as f:
        CH = 200000
        for s in range(0, n, CH):
            blk = reads[s:s + CH]
            lines = []
            seqs = nt[blk]
            for i in range(len(blk)):
                lines.append(b"@r%d\n" % (s + i) + seqs[i].tobytes() + b"\n+\n" + b"I" * L + b"\n")
            f.write(b"".join(lines))
    del width


# -------------------------------------------------------------- clocks -------

class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            log(f"[bench] NVML unavailable: {e}")

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------ reference (CPU) ------

def time_reference(prefix: str, reads: np.ndarray, threads: int, work_dir: str, target_s: float = 15.0,
                   max_reads: int = 2_000_000, aln_args=()):
    """Times `ibwa aln -t threads` on a bounded sample; returns dict(reads/s, sample, sai path, ...)."""
    os.makedirs(work_dir, exist_ok=True)
    empty = os.path.join(work_dir, "empty.fq")
    open(empty, "w").close()

    def run(fq, out):
        t0 = time.perf_counter()
        with open(out, "wb") as fo:
            subprocess.run([REF_BIN, "aln", "-t", str(threads)] + list(aln_args) + [prefix, fq], stdout=fo,
                           stderr=subprocess.DEVNULL, check=True)
        return time.perf_counter() - t0

    run(empty, os.path.join(work_dir, "empty.sai"))          # page cache warm
    t_load = run(empty, os.path.join(work_dir, "empty.sai"))
    n0 = min(len(reads), 20000)
    fq0 = os.path.join(work_dir, "probe.fq")
    write_fastq(fq0, reads[:n0])
    t_probe = run(fq0, os.path.join(work_dir, "probe.sai"))
    rate = n0 / max(t_probe - t_load, 1e-3)
    n1 = int(min(len(reads), max_reads, max(n0, rate * target_s)))
    fq1 = os.path.join(work_dir, "sample.fq")
    write_fastq(fq1, reads[:n1])
    sai1 = os.path.join(work_dir, "sample.sai")
    t1 = run(fq1, sai1)
    return {"reads_per_s": n1 / max(t1 - t_load, 1e-3), "n": n1, "wall_s": t1, "index_load_s": t_load,
            "threads": threads, "sai": sai1, "fq": fq1}


def time_port(bwt, rbwt, reads: np.ndarray, opt, n: int = 3000):
    """Single-thread CPU restatement (oracle port) on a small sample; also returns its lookup counters."""
    from oracle import pyoracle
    n = min(n, len(reads))
    L = reads.shape[1]
    lens = np.full(n, L, np.int32)
    offs = np.arange(n, dtype=np.int64) * L
    ob, orb = pyoracle.as_orc_bwt(bwt), pyoracle.as_orc_bwt(rbwt)
    t0 = time.perf_counter()
    n_aln, rec, st = pyoracle.aln_batch(ob, orb, lens, offs, reads[:n].reshape(-1), opt.to_c())
    dt = time.perf_counter() - t0
    return {"reads_per_s": n / dt, "n": n, "stats": st, "n_aln": n_aln, "rec": rec}


# ----------------------------------------------------------------- main ------

def _claim_stdout():
    """Libraries (NCCL's version banner, for one) print to stdout; the contract is ONE JSON line there.
    Keep the real stdout for that line and send everything else written to fd 1 to stderr."""
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(os.dup(2), "w", buffering=1)
    return os.fdopen(real, "w", buffering=1)


def main():
    json_out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-bp", type=int, default=3_100_000_000)
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step")
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--seed", type=int, default=20260102)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--model", default="default", choices=["default", "stress"],
                    help="read model (SURVEY §8d); 'stress' + --read-len 150 + --aln-args = BASELINE configs[2]")
    ap.add_argument("--aln-args", default="", help="aln options for the run, e.g. '-n 4 -o 2 -e 10 -l 32 -k 2'")
    ap.add_argument("--in-flight", type=int, default=3,
                    help="batches in flight in the e2e measurement (contexts sharing the device index)")
    ap.add_argument("--set", action="append", default=[], help="engine knob key=value")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    import torch
    import torch.distributed as dist

    if args.impl == "reference" and rank != 0:
        return 0
    use_dist = world > 1 and args.impl == "ours"
    if not torch.cuda.is_available():
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "no CUDA device to synthesise the index with"}),
                  file=json_out, flush=True)
            return 0
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if use_dist:
        dist.init_process_group("nccl", device_id=dev)

    from ibwa_b200 import engine, gap_init_opt

    opt = gap_init_opt()
    aln_args = args.aln_args.split()
    if aln_args:
        from ibwa_b200 import parse_aln_args
        opt, _, _, _ = parse_aln_args(aln_args + ["prefix", "reads"])
    workload = (f"bwa aln {args.reads} simulated {args.read_len}bp reads/GPU ({args.model} model) vs "
                f"{args.genome_bp / 1e9:.2f} Gbp synthetic genome, {args.aln_args or 'defaults (-n 0.04)'}, "
                f"index replicated")
    config = {"workload": workload, "genome_bp": args.genome_bp, "reads_per_gpu": args.reads,
              "read_len": args.read_len, "parallelism": f"replicated index, reads sharded x{world}",
              "l2": "inputs larger than L2 (index 3.1 GB + per-read state); no flush"}

    # ---- inputs -----------------------------------------------------------
    t_setup = time.time()
    if use_dist and local_rank != 0:
        dist.barrier()                      # rank 0 builds and writes the cache first
    bwt, rbwt, text, prefix = load_or_build_index(args.genome_bp, args.seed, dev, is_writer=(local_rank == 0))
    if use_dist and local_rank == 0:
        dist.barrier()
    synth_fn = synth_reads_torch if args.model == "default" else synth_stress_reads_torch
    reads_d = synth_fn(text, args.reads, args.read_len, args.seed + 1000 + rank)
    del text
    torch.cuda.empty_cache()
    log(f"[bench r{rank}] inputs ready in {time.time() - t_setup:.1f} s")

    nproc = os.cpu_count() or 1
    work_dir = os.path.join(cache_dir(args.genome_bp, args.seed), f"work_r{rank}")

    # ---- reference arm ------------------------------------------------------
    if args.impl == "reference":
        sample_h = reads_d[: min(args.reads, 2_000_000)].cpu().numpy()
        del reads_d
        torch.cuda.empty_cache()
        if os.path.exists(REF_BIN):
            kind = "reference"
            vals = []
            res = None
            for i in range(args.warmup + args.steps):
                res = time_reference(prefix, sample_h, nproc, work_dir, target_s=8.0, aln_args=aln_args)
                if i >= args.warmup:
                    vals.append(res["reads_per_s"])
                if i == 0 and args.warmup > 1:
                    pass
            v = float(np.mean(vals))
            sample = (f"{res['n']} of the workload's reads per step, `ibwa aln -t {nproc}`, wall minus "
                      f"{res['index_load_s']:.2f} s index load")
            ms = 1e3 * res["n"] / v
        else:
            kind = "port"
            pr = time_port(bwt, rbwt, sample_h, opt, n=5000)
            v, nproc, ms = pr["reads_per_s"], 1, 1e3 * pr["n"] / pr["reads_per_s"]
            sample = f"{pr['n']} reads, single-thread CPU restatement (oracle port)"
        line = {"metric": "bwa aln reads/sec (100bp, 3.1Gbp ref)", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": config,
                "impl": "reference",
                "cpu_baseline": {"value": v, "unit": "reads/s", "cores": nproc, "kind": kind, "sample": sample},
                "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), file=json_out, flush=True)
        return 0

    # ---- our arm ------------------------------------------------------------
    eng = engine.Engine(bwt, rbwt, local_rank)
    for kv in args.set:
        k, v = kv.split("=")
        eng.set(k, int(v))
    n, L = args.reads, args.read_len
    d_lens = torch.full((n,), L, dtype=torch.int32, device=dev)
    d_offs = torch.arange(n, dtype=torch.int64, device=dev) * L
    d_codes = reads_d.reshape(-1)
    # pinned host copies for the end-to-end path
    h_lens = torch.full((n,), L, dtype=torch.int32).pin_memory()
    h_offs = (torch.arange(n, dtype=torch.int64) * L).pin_memory()
    h_codes = torch.empty(n * L, dtype=torch.uint8).pin_memory()
    h_codes.copy_(d_codes)
    h_naln = torch.empty(n, dtype=torch.int32).pin_memory()
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
            torch.cuda.synchronize()

    def step_device():
        eng.batch_device(d_lens.data_ptr(), d_offs.data_ptr(), d_codes.data_ptr(), n, L, opt)
        return eng.stats()

    # end-to-end: --in-flight batches in flight (engine + clones sharing the device index, one host thread each),
    # so the H2D / D2H copies of one step overlap the kernels of the other (double buffering)
    K = max(1, args.in_flight)
    engines = [eng] + [eng.clone() for _ in range(K - 1)]
    outs = [h_naln] + [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(K - 1)]

    def step_e2e(which=0):
        e, out = engines[which], outs[which]
        _, total = e.batch_pinned(h_lens.data_ptr(), h_offs.data_ptr(), h_codes.data_ptr(), n, opt, out.data_ptr())
        return total, e.stats()

    def timed_e2e(steps):
        results = [None] * steps

        def worker(which):
            for i in range(which, steps, K):
                results[i] = step_e2e(which)

        barrier()
        eng.timer_start()
        th = [threading.Thread(target=worker, args=(w,)) for w in range(K)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        ms = eng.timer_stop()
        barrier()
        if use_dist:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, results

    def timed(fn, steps):
        barrier()
        eng.timer_start()
        stats = []
        for _ in range(steps):
            stats.append(fn())
        ms = eng.timer_stop()
        barrier()
        if use_dist:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, stats

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev, st_dev = timed(step_device, args.steps)
    clocks = sampler.stop()
    for w in range(K):          # every in-flight engine allocates its buffers outside the timed region
        step_e2e(w)
    ms_e2e, st_e2e = timed_e2e(args.steps)

    eng.set("count", 1)         # one untimed step with the pop / sector counters compiled in
    counted = step_device()
    eng.set("count", 0)

    total_reads = world * n * args.steps
    value = total_reads / (ms_dev * 1e-3)
    e2e_value = total_reads / (ms_e2e * 1e-3)
    last = st_dev[-1]
    total_rec = st_e2e[-1][0]
    launches = int(sum(s["kernel_launches"] for s in st_dev))

    out = {"metric": "bwa aln reads/sec (100bp, 3.1Gbp ref)", "value": value, "unit": "reads/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
           "config": config, "clocks": clocks,
           "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(n * 12 + n * L),
                   "d2h_bytes_per_step": int(n * 4 + total_rec * 16), "ms_per_step": ms_e2e / args.steps,
                   "in_flight": K},
           "gpu_launches": launches,
           "kernel_ms": {k: float(np.mean([s[k] for s in st_dev])) for k in
                         ("ms_width", "ms_search", "ms_compact", "ms_total")},
           "overflow_reads_per_step": int(last["overflow_reads"]),
           "device_pops_per_read": counted["pops"] / n, "device_sectors_per_read": counted["occ_lookups"] / n}

    if rank == 0:
        # roofline: algorithmic bytes = 32 B x occ lookups of the REFERENCE algorithm (oracle-counted on a sample)
        sample_h = reads_d[: min(n, 2_000_000)].cpu().numpy()
        port = time_port(bwt, rbwt, sample_h, opt, n=3000)
        lookups_per_read = port["stats"]["lookups"] / port["n"]
        bytes_per_read = 32.0 * lookups_per_read
        ms_search = out["kernel_ms"]["ms_search"]
        achieved = bytes_per_read * n / (ms_search * 1e-3) / 1e9
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, which = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, which = 6650.0, "fallback (B200_PROFILING.md)"
        traffic = None                      # DRAM bytes of the dominant kernel per launch, from the committed ncu capture
        tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
        if os.path.exists(tpath):
            traffic = float(json.load(open(tpath))["k_search"]["dram_bytes_per_read"]) * n
        sector_roof = eng.sector_roofline(1 << 28, 3)
        pair_roof = eng.sector_roofline(1 << 27, -3)
        out["roofline"] = {"bound": "hbm", "kernel": "k_search", "achieved": achieved, "peak": peak, "unit": "GB/s",
                           "frac": achieved / peak, "traffic": traffic, "peak_source": which,
                           "traffic_source": "profiles/r1_traffic.json (ncu dram__bytes_read+write per read x reads)",
                           "algorithmic_bytes": bytes_per_read * n,
                           "algorithmic_bytes_per_read": bytes_per_read,
                           "oracle_lookups_per_read": lookups_per_read,
                           "oracle_pops_per_read": port["stats"]["pops"] / port["n"],
                           "random_sector_roof_gbs": sector_roof, "random_64B_pair_roof_gbs": pair_roof,
                           "frac_of_random_sector_roof": achieved / sector_roof if sector_roof else None,
                           "occ_sectors_per_s": counted["occ_lookups"] / (ms_search * 1e-3)}
        if traffic and sector_roof:
            # the memory system's limit for this access pattern is a transaction rate: DRAM moves the
            # kernel's traffic in 64-byte transactions; the gather roof counts random 32-byte requests per second
            tx = traffic / 64.0
            out["roofline"]["dram_transactions_per_read"] = tx / n
            out["roofline"]["dram_transactions_per_s"] = tx / (ms_search * 1e-3)
            out["roofline"]["random_requests_roof_per_s"] = sector_roof * 1e9 / 32.0
            out["roofline"]["frac_of_transaction_roof"] = (tx / (ms_search * 1e-3)) / (sector_roof * 1e9 / 32.0)
        # parity of this very run against the oracle port on the sample
        m = port["n"]
        n_aln_d, rec_d = eng.cal_sa_reg_gap(np.full(m, L, np.int32), np.arange(m, dtype=np.int64) * L,
                                            sample_h[:m].reshape(-1), opt)
        parity = {"oracle_port_reads": m,
                  "oracle_port_identical": bool(np.array_equal(n_aln_d, port["n_aln"]) and
                                                rec_d.tobytes() == port["rec"].tobytes())}
        if world == 1 and not args.no_cpu_baseline:
            if os.path.exists(REF_BIN):
                res = time_reference(prefix, sample_h, nproc, work_dir, target_s=15.0, aln_args=aln_args)
                out["cpu_baseline"] = {"value": res["reads_per_s"], "unit": "reads/s", "cores": nproc,
                                       "kind": "reference",
                                       "sample": f"{res['n']} reads of the workload, `ibwa aln -t {nproc}`, wall "
                                                 f"{res['wall_s']:.2f} s minus {res['index_load_s']:.2f} s index load"}
                from ibwa_b200 import sai
                _, r_n, r_rec = sai.read_sai(res["sai"])
                r_n = r_n[:res["n"]]
                k = res["n"]
                g_n, g_rec = eng.cal_sa_reg_gap(np.full(k, L, np.int32), np.arange(k, dtype=np.int64) * L,
                                                sample_h[:k].reshape(-1), opt)
                parity["reference_binary_reads"] = k
                parity["reference_binary_identical"] = bool(np.array_equal(g_n, r_n) and
                                                            g_rec.tobytes() == r_rec.tobytes())
            else:
                out["cpu_baseline"] = {"value": port["reads_per_s"], "unit": "reads/s", "cores": 1, "kind": "port",
                                       "sample": f"{port['n']} reads, single-thread oracle port"}
        out["parity"] = parity
        print(json.dumps(out), file=json_out, flush=True)
    for e in engines[1:]:
        e.close()
    eng.close()
    if use_dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
