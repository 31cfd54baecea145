#!/usr/bin/env python
"""bench.py — `bwa aln` reads/s on B200 (BASELINE.json metric), one process per GPU.

Workloads (--config, BASELINE.json `configs`, numbered from 1):
  1  100 k x 100 bp vs 5 Mbp, defaults                       (the reference's own CPU-sized case)
  2  10 M x 100 bp vs 3.1 Gbp, defaults                      (the headline; default)
  3  stress: 150 bp reads, -n 4 -o 2 -e 10 -l 32 -k 2, 2 % substitutions + indels, 3.1 Gbp
  4  paired end: 2 x 5 M x 100 bp, both mates through the engine; parity = .sai identity and the SAM of the
     UNCHANGED `ibwa sampe -R` on engine vs reference .sai
  5  dbset: primary + 200 ALT contigs with a .remap file, 25 % of the pairs from ALT sequence; both mates against
     both indexes; parity = four .sai and `sampe -R <pri> .. <alt> ..` SAM (ZR:Z tags) equality
The genome is i.i.d. uniform ACGT generated on the GPU from a seed; its index files are built on the GPU by
ibwa_b200.fmbuild / refdata (byte-identical to what the reference's `index` writes, tests/test_fmbuild.py,
tests/test_refdata.py) and cached under /tmp/b200aln_bench so that the reference binary reads the same files.

A step = one pass of the hot path (bwa_cal_sa_reg_gap) over the step's reads (every stream of the config).
  --scaling strong (default): the step's --reads are cut into contiguous per-rank shards
  --scaling weak            : every rank aligns --reads reads of its own
  value : reads resident in HBM before the timed region (b200aln_batch_device)
  e2e   : the same reads from pinned HOST buffers through b200aln_batch
          (H2D of reads and D2H of n_aln + records inside the timed region)
Both are throughputs of K steps with --in-flight steps in flight on contexts that share the device index (the
engine's double buffering); `sequential` holds the same steps one at a time on one context, which is also where
the per-kernel times of the roofline come from.  Timing: CUDA events on the engine's launch stream
(b200aln_timer_*), barrier + synchronize on both sides, max over ranks.  Index (3.1 GB) and per-read state
(> 10 GB) are far larger than L2, so no L2 flush is needed between steps.

`--impl reference` times the unmodified reference binary (oracle/_ref/ibwa aln -t <all cores>) on a bounded
sample of the same reads on the host cores.
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
    if REPEATS:
        plant_repeats(t, REPEATS, seed)
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


REPEATS = 0.0      # --repeats: fraction of the genome's bases planted as 40-copy families (set by main)


def cache_dir(genome_bp: int, seed: int) -> str:
    return os.path.join(CACHE_ROOT, f"g{genome_bp}_s{seed}" + (f"_rep{REPEATS:g}" if REPEATS else ""))


def plant_repeats(text, frac: float, seed: int):
    """SURVEY.md §4 F2 at scale: families of 40 copies of a random 300-bp unit, every copy diverged by 2 %
    substitutions, at random places, until `frac` of the bases lie in them (real genomes' repeats are what an
    i.i.d. text lacks: many near-equal hits per read, deep ties in the suffix order)."""
    import torch
    dev = text.device
    n = text.numel()
    g = torch.Generator(device=dev)
    g.manual_seed(seed + 77)
    n_fam = int(frac * n / (40 * 300))
    ar = torch.arange(300, device=dev)
    for s in range(0, n_fam, 1024):
        f = min(1024, n_fam - s)
        unit = torch.randint(0, 4, (f, 1, 300), dtype=torch.uint8, device=dev, generator=g).expand(f, 40, 300)
        mut = torch.rand((f, 40, 300), device=dev, generator=g) < 0.02
        inc = torch.randint(1, 4, (f, 40, 300), dtype=torch.uint8, device=dev, generator=g)
        cp = torch.where(mut, (unit + inc) & 3, unit).reshape(-1, 300)
        at = torch.randint(0, n - 300, (f * 40,), device=dev, generator=g)
        text[(at[:, None] + ar[None, :]).reshape(-1)] = cp.reshape(-1)
    return text


def contig_layout(genome_bp: int):
    """contigs chr1.. of 100 Mbp (a contig must stay below 2^31 bp, bntseq.h:42); the text is one stream."""
    step = 100_000_000
    lens = [step] * (genome_bp // step)
    if genome_bp % step:
        lens.append(genome_bp % step)
    return [f"chr{i + 1}" for i in range(len(lens))], lens


def load_or_build_index(genome_bp: int, seed: int, device, is_writer: bool, with_sa: bool = False):
    """Returns (bwt, rbwt, text, prefix).  The text is always regenerated from the seed on the GPU.  with_sa: also
    the files the downstream `sampe` opens (.sa .rsa .pac .ann .amb), written by ibwa_b200.refdata."""
    import torch
    from ibwa_b200 import fmbuild, refdata
    from ibwa_b200.bwtio import bwt_dump_bwt, bwt_dump_sa, bwt_restore_bwt
    d = cache_dir(genome_bp, seed)
    prefix = os.path.join(d, "ref")
    text = gen_text(genome_bp, seed, device)
    done = os.path.join(d, "DONE_SA" if with_sa else "DONE")
    if os.path.exists(done) or (not with_sa and os.path.exists(os.path.join(d, "DONE_SA"))):
        return bwt_restore_bwt(prefix + ".bwt"), bwt_restore_bwt(prefix + ".rbwt"), text, prefix
    t0 = time.time()
    if with_sa:
        bwt, sa = fmbuild.build_bwt_torch(text, sa_intv=32)
        torch.cuda.empty_cache()
        rbwt, rsa = fmbuild.build_bwt_torch(torch.flip(text, dims=[0]), sa_intv=32)
    else:
        bwt = fmbuild.build_bwt_torch(text)
        torch.cuda.empty_cache()
        rbwt = fmbuild.build_bwt_torch(torch.flip(text, dims=[0]))
    torch.cuda.empty_cache()
    log(f"[bench] built .bwt/.rbwt{'/.sa/.rsa' if with_sa else ''} for {genome_bp} bp on the GPU in {time.time() - t0:.1f} s")
    if is_writer:
        os.makedirs(d, exist_ok=True)
        bwt_dump_bwt(prefix + ".bwt", bwt)
        bwt_dump_bwt(prefix + ".rbwt", rbwt)
        if with_sa:
            bwt_dump_sa(prefix + ".sa", sa)
            bwt_dump_sa(prefix + ".rsa", rsa)
            names, lens = contig_layout(genome_bp)
            refdata.write_pac_ann_amb(prefix, text, names, lens)
        open(done, "w").write("ok\n")
    return bwt, rbwt, text, prefix


def load_or_build_alt(text, genome_bp: int, seed: int, is_writer: bool):
    """Config 5's ALT index: 200 contigs cut from the primary (refdata.make_alt_contigs), all index files and the
    .remap file (bwaremap.cpp:42-132).  Returns (bwt, rbwt, alt_text uint8 numpy, alt_lens, prefix)."""
    from ibwa_b200 import refdata
    from ibwa_b200.bwtio import bwt_restore_bwt
    d = cache_dir(genome_bp, seed)
    prefix = os.path.join(d, "alt")
    names, lens = contig_layout(genome_bp)
    alt_text, a_names, a_lens, remap = refdata.make_alt_contigs(
        lambda lo, hi: text[lo:hi].cpu().numpy(), names, lens, 200, 20260105)
    if not os.path.exists(prefix + ".DONE"):
        if is_writer:
            os.makedirs(d, exist_ok=True)
            refdata.write_index(prefix, alt_text, a_names, a_lens)
            open(prefix + ".remap", "w").write(remap)
            open(prefix + ".DONE", "w").write("ok\n")
    return bwt_restore_bwt(prefix + ".bwt"), bwt_restore_bwt(prefix + ".rbwt"), alt_text, a_lens, prefix


def write_fastq(path: str, reads: np.ndarray, name_fmt: bytes = b"@r%d\n") -> None:
    nt = np.frombuffer(b"ACGTN-", dtype=np.uint8)
    n, L = reads.shape
    tail = b"\n+\n" + b"I" * L + b"\n"
    with open(path, "wb") as f:
        CH = 200000
        for s in range(0, n, CH):
            seqs = nt[reads[s:s + CH]]
            f.write(b"".join(name_fmt % (s + i) + seqs[i].tobytes() + tail for i in range(len(seqs))))


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

def _ref_aln(prefix, fq, out, threads, aln_args):
    t0 = time.perf_counter()
    with open(out, "wb") as fo:
        subprocess.run([REF_BIN, "aln", "-t", str(threads)] + list(aln_args) + [prefix, fq], stdout=fo,
                       stderr=subprocess.DEVNULL, check=True)
    return time.perf_counter() - t0


def time_reference(streams, threads: int, work_dir: str, target_s: float = 15.0, max_reads: int = 2_000_000,
                   aln_args=(), tag: str = "sample"):
    """Times `ibwa aln -t threads` on a bounded sample of every stream of the workload.
    streams: list of dict(prefix=, reads=ndarray[n, L], name_fmt=bytes, first=bool) — `first` marks the streams
    that carry distinct reads (cfg 5 aligns the same reads against two indexes: they count once).
    Returns reads/s (distinct reads / summed wall minus index load), the sample size and the .sai / .fq paths."""
    os.makedirs(work_dir, exist_ok=True)
    empty = os.path.join(work_dir, "empty.fq")
    open(empty, "w").close()
    loads = {}
    for st in streams:
        if st["prefix"] not in loads:
            _ref_aln(st["prefix"], empty, os.path.join(work_dir, "empty.sai"), threads, aln_args)   # page cache warm
            loads[st["prefix"]] = _ref_aln(st["prefix"], empty, os.path.join(work_dir, "empty.sai"), threads, aln_args)
    s0 = streams[0]
    n_all = len(s0["reads"])
    n0 = min(n_all, 20000)
    fq0 = os.path.join(work_dir, "probe.fq")
    write_fastq(fq0, s0["reads"][:n0], s0["name_fmt"])
    t_probe = _ref_aln(s0["prefix"], fq0, os.path.join(work_dir, "probe.sai"), threads, aln_args)
    rate = n0 / max(t_probe - loads[s0["prefix"]], 1e-3)
    n1 = int(min(n_all, max_reads, max(n0, rate * target_s / len(streams))))
    wall = net = 0.0
    sais, fqs = [], []
    for i, st in enumerate(streams):
        fq = os.path.join(work_dir, f"{tag}_{i}.fq")
        if st.get("fq_of") is not None:
            fq = fqs[st["fq_of"]]                                   # same reads, another index
        else:
            write_fastq(fq, st["reads"][:n1], st["name_fmt"])
        sai_path = os.path.join(work_dir, f"{tag}_{i}.sai")
        t = _ref_aln(st["prefix"], fq, sai_path, threads, aln_args)
        wall += t
        net += t - loads[st["prefix"]]
        sais.append(sai_path)
        fqs.append(fq)
    distinct = n1 * sum(1 for st in streams if st.get("first", True))
    return {"reads_per_s": distinct / max(net, 1e-3), "n": n1, "distinct_reads": distinct, "wall_s": wall,
            "index_load_s": sum(loads.values()), "threads": threads, "sais": sais, "fqs": fqs}


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

CONFIGS = {
    1: dict(genome_bp=5_000_000, reads=100_000, read_len=100, model="default", aln_args="", pairs=False, alt=False,
            seed=20260101, name="configs[0]: 100 k x 100 bp vs 5 Mbp, defaults"),
    2: dict(genome_bp=3_100_000_000, reads=10_000_000, read_len=100, model="default", aln_args="", pairs=False,
            alt=False, seed=20260102, name="configs[1]: 10 M x 100 bp vs 3.1 Gbp, defaults"),
    3: dict(genome_bp=3_100_000_000, reads=2_000_000, read_len=150, model="stress",
            aln_args="-n 4 -o 2 -e 10 -l 32 -k 2", pairs=False, alt=False, seed=20260102,
            name="configs[2]: gapped stress, 150 bp, 2 % substitutions + indels vs 3.1 Gbp"),
    4: dict(genome_bp=3_100_000_000, reads=10_000_000, read_len=100, model="default", aln_args="", pairs=True,
            alt=False, seed=20260102, name="configs[3]: paired end 2 x 5 M x 100 bp vs 3.1 Gbp, .sai -> sampe -R"),
    5: dict(genome_bp=3_100_000_000, reads=10_000_000, read_len=100, model="default", aln_args="", pairs=True,
            alt=True, seed=20260102,
            name="configs[4]: dbset primary 3.1 Gbp + 200 ALT contigs (.remap), 2 x 5 M x 100 bp, 25 % of pairs from ALT"),
}
METRIC = "bwa aln reads/sec (100bp, 3.1Gbp ref)"


def _claim_stdout():
    """Libraries (NCCL's version banner, for one) print to stdout; the contract is ONE JSON line there.
    Keep the real stdout for that line and send everything else written to fd 1 to stderr."""
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(os.dup(2), "w", buffering=1)
    return os.fdopen(real, "w", buffering=1)


def _md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def _sampe(pri, sais_pri, fqs, alt, sais_alt, out, threads):
    cmd = [REF_BIN, "sampe", "-R", "-t", str(threads), pri, sais_pri[0], sais_pri[1], fqs[0], fqs[1]]
    if alt:
        cmd += [alt, sais_alt[0], sais_alt[1]]
    t0 = time.perf_counter()
    with open(out, "wb") as fo:
        subprocess.run(cmd, stdout=fo, stderr=subprocess.DEVNULL, check=True)
    return time.perf_counter() - t0


def main():
    json_out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configs, from 1")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: --reads per step in total, sharded over the ranks; weak: --reads per rank")
    ap.add_argument("--genome-bp", type=int, default=None)
    ap.add_argument("--reads", type=int, default=None, help="reads per step (strong: all ranks together)")
    ap.add_argument("--read-len", type=int, default=None)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--model", default=None, choices=["default", "stress"], help="read model (SURVEY 8d)")
    ap.add_argument("--aln-args", default=None, help="aln options for the run, e.g. '-n 4 -o 2 -e 10 -l 32 -k 2'")
    ap.add_argument("--in-flight", type=int, default=None,
                    help="steps in flight (contexts sharing the device index) in the value / e2e measurements; default "
                         "3 for >= 8 M reads per GPU and step, 4 from 4 M, else 6 (small launches need a deeper pipeline: "
                         "the engine then parks the stragglers of a draining launch, DESIGN.md)")
    ap.add_argument("--parity-pairs", type=int, default=200_000, help="configs 4/5: pairs through sampe -R")
    ap.add_argument("--set", action="append", default=[], help="engine knob key=value")
    ap.add_argument("--repeats", type=float, default=0.0,
                    help="fraction of the genome's bases in 40-copy families of diverged 300-bp units (0 = i.i.d. text)")
    args = ap.parse_args()
    global REPEATS
    REPEATS = args.repeats
    cfg = dict(CONFIGS[args.config])
    for k in ("genome_bp", "reads", "read_len", "seed", "model", "aln_args"):
        v = getattr(args, k)
        if v is not None:
            cfg[k] = v

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

    from ibwa_b200 import engine, gap_init_opt, refdata, sai
    from ibwa_b200.shard import shard_range

    opt = gap_init_opt()
    aln_args = cfg["aln_args"].split()
    if aln_args:
        from ibwa_b200 import parse_aln_args
        opt, _, _, _ = parse_aln_args(aln_args + ["prefix", "reads"])
    L = cfg["read_len"]
    n_step = cfg["reads"]                         # distinct reads of one step, all ranks (strong) / this rank (weak)
    if cfg["pairs"]:
        n_step -= n_step % 2
    weak = args.scaling == "weak" and world > 1
    unit = 2 if cfg["pairs"] else 1               # sharding unit: a read, or a pair (mates stay together)
    lo, hi = (0, n_step // unit) if weak or world == 1 else shard_range(n_step // unit, world, rank)
    config = {"workload": f"{cfg['name']} [{cfg['aln_args'] or 'defaults (-n 0.04)'}; {cfg['model']} read model; "
                          f"{n_step} reads per step {'per GPU' if weak else 'in total'}; index replicated]",
              "config_no": args.config, "genome_bp": cfg["genome_bp"], "repeats": args.repeats, "reads_per_step": n_step,
              "reads_per_gpu_per_step": (hi - lo) * unit, "read_len": L,
              "parallelism": f"replicated index, reads sharded x{world} ({'weak' if weak else 'strong'}), no collective",
              "l2": "inputs larger than L2 (index 3.1 GB + per-read state); no flush"}

    # ---- inputs -----------------------------------------------------------
    t_setup = time.time()
    need_sa = cfg["pairs"] and os.path.exists(REF_BIN) and world == 1 and args.impl == "ours" and not args.no_cpu_baseline
    if use_dist and local_rank != 0:
        dist.barrier()                      # rank 0 builds and writes the cache first
    bwt, rbwt, text, prefix = load_or_build_index(cfg["genome_bp"], cfg["seed"], dev, is_writer=(local_rank == 0),
                                                  with_sa=need_sa)
    alt = None
    if cfg["alt"]:
        a_bwt, a_rbwt, alt_text, alt_lens, alt_prefix = load_or_build_alt(text, cfg["genome_bp"], cfg["seed"],
                                                                          is_writer=(local_rank == 0))
        alt = dict(bwt=a_bwt, rbwt=a_rbwt, prefix=alt_prefix)
    if use_dist and local_rank == 0:
        dist.barrier()
    rseed = cfg["seed"] + 1000 + (rank if weak else 0)
    if cfg["pairs"]:
        n_pairs = n_step // 2
        kw = {}
        if cfg["alt"]:
            kw = dict(alt_text=torch.from_numpy(alt_text).to(dev), alt_lens=alt_lens, alt_frac=0.25)
        r1, r2 = refdata.synth_pairs(text, n_pairs, L, rseed, **kw)
        sets = [r1[lo:hi].contiguous(), r2[lo:hi].contiguous()]
        del r1, r2
        fmts = [b"@p%d/1\n", b"@p%d/2\n"]
    else:
        synth_fn = synth_reads_torch if cfg["model"] == "default" else synth_stress_reads_torch
        allr = synth_fn(text, n_step, L, rseed)
        sets = [allr[lo:hi].contiguous()]
        del allr
        fmts = [b"@r%d\n"]
    del text
    torch.cuda.empty_cache()
    # streams of one step: (index, read set)
    streams = [("pri", i) for i in range(len(sets))] + ([("alt", i) for i in range(len(sets))] if alt else [])
    n_rank = sum(int(s.shape[0]) for s in sets)   # distinct reads this rank aligns per step
    log(f"[bench r{rank}] inputs ready in {time.time() - t_setup:.1f} s ({n_rank} reads per step on this rank)")

    nproc = os.cpu_count() or 1
    work_dir = os.path.join(cache_dir(cfg["genome_bp"], cfg["seed"]), f"work_c{args.config}_r{rank}")

    def ref_streams(samples):
        out = [dict(prefix=prefix, reads=s, name_fmt=f, first=True) for s, f in zip(samples, fmts)]
        if alt:
            out += [dict(prefix=alt["prefix"], reads=s, name_fmt=f, first=False, fq_of=i)
                    for i, (s, f) in enumerate(zip(samples, fmts))]
        return out

    # ---- reference arm ------------------------------------------------------
    if args.impl == "reference":
        samples = [s[: min(int(s.shape[0]), 2_000_000)].cpu().numpy() for s in sets]
        del sets
        torch.cuda.empty_cache()
        if os.path.exists(REF_BIN):
            kind = "reference"
            vals = []
            res = None
            for i in range(args.warmup + args.steps):
                res = time_reference(ref_streams(samples), nproc, work_dir, target_s=8.0, aln_args=aln_args)
                if i >= args.warmup:
                    vals.append(res["reads_per_s"])
            v = float(np.mean(vals))
            sample = (f"{res['distinct_reads']} of the workload's reads per step ({len(streams)} `ibwa aln -t {nproc}` "
                      f"runs), wall minus {res['index_load_s']:.2f} s index load")
            ms = 1e3 * res["distinct_reads"] / v
        else:
            kind = "port"
            pr = time_port(bwt, rbwt, samples[0], opt, n=5000)
            v, nproc, ms = pr["reads_per_s"], 1, 1e3 * pr["n"] / pr["reads_per_s"]
            sample = f"{pr['n']} reads, single-thread CPU restatement (oracle port)"
        line = {"metric": METRIC, "value": v, "unit": "reads/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": config, "impl": "reference",
                "cpu_baseline": {"value": v, "unit": "reads/s", "cores": nproc, "kind": kind, "sample": sample},
                "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), file=json_out, flush=True)
        return 0

    # ---- our arm ------------------------------------------------------------
    if args.in_flight is None:
        # config 3: its re-run arenas are large; config 5 runs two indexes, each with its own contexts
        args.in_flight = 3 if n_rank >= 8_000_000 or args.config == 3 else 4 if n_rank >= 4_000_000 else 8 if not cfg["alt"] and 1_000_000 <= n_rank < 2_000_000 else 6
    K = max(1, args.in_flight)
    eng = engine.Engine(bwt, rbwt, local_rank)
    knobs = [kv.split("=") for kv in args.set]
    for k, v in knobs:
        eng.set(k, int(v))
    eng.set("slots", 1)                     # steps in flight come from the clones below, not from chunks of a call
    engs = {"pri": [eng] + [eng.clone() for _ in range(K - 1)]}
    if alt:
        a_eng = engine.Engine(alt["bwt"], alt["rbwt"], local_rank)
        for k, v in knobs:
            a_eng.set(k, int(v))
        a_eng.set("slots", 1)
        engs["alt"] = [a_eng] + [a_eng.clone() for _ in range(K - 1)]

    dsets, hsets = [], []
    for s in sets:
        n = int(s.shape[0])
        d = dict(n=n, lens=torch.full((n,), L, dtype=torch.int32, device=dev),
                 offs=torch.arange(n, dtype=torch.int64, device=dev) * L, codes=s.reshape(-1))
        h = dict(lens=torch.full((n,), L, dtype=torch.int32).pin_memory(),
                 offs=(torch.arange(n, dtype=torch.int64) * L).pin_memory(),
                 codes=torch.empty(n * L, dtype=torch.uint8).pin_memory())
        h["codes"].copy_(d["codes"])
        dsets.append(d)
        hsets.append(h)
    h_out = [[torch.empty(dsets[i]["n"], dtype=torch.int32).pin_memory() for (_, i) in streams] for _ in range(K)]
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
            torch.cuda.synchronize()

    def step_device(w=0):
        out = []
        for (ix, i) in streams:
            d = dsets[i]
            e = engs[ix][w]
            if d["n"]:
                e.batch_device(d["lens"].data_ptr(), d["offs"].data_ptr(), d["codes"].data_ptr(), d["n"], L, opt)
            out.append(e.stats())
        return out

    def step_e2e(w=0):
        out = []
        for j, (ix, i) in enumerate(streams):
            h = hsets[i]
            e = engs[ix][w]
            total = 0
            if dsets[i]["n"]:
                _, total = e.batch_pinned(h["lens"].data_ptr(), h["offs"].data_ptr(), h["codes"].data_ptr(),
                                          dsets[i]["n"], opt, h_out[w][j].data_ptr())
            out.append((total, e.stats()))
        return out

    def timed(fn, steps, k):
        """`steps` steps, k in flight (host thread w runs steps w, w + k, ...); ms by CUDA events, max over ranks"""
        results = [None] * steps
        done_at = [0.0] * steps

        def worker(w):
            for i in range(w, steps, k):
                results[i] = fn(w)
                done_at[i] = time.perf_counter()

        barrier()
        eng.timer_start()
        if k == 1:
            worker(0)
        else:
            th = [threading.Thread(target=worker, args=(w,)) for w in range(k)]
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
        timed.done_at = sorted(done_at)
        return ms, results

    def steady_ms(done):
        """mean spacing of step completions over the middle half of a pipelined run (host clock): what the pipeline
        sustains between its fill and its drain — informational, the timed region is what `value` is"""
        if len(done) < 12:
            return None
        a, b = len(done) // 4, len(done) - len(done) // 4
        return 1e3 * float(done[b - 1] - done[a]) / (b - 1 - a)

    for w in range(K):                      # every context allocates its buffers outside the timed regions
        for _ in range(max(1, (args.warmup + K - 1) // K)):
            step_device(w)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_seq, st_seq = timed(step_device, args.steps, 1)
    ms_dev, _ = timed(step_device, args.steps, K) if K > 1 else (ms_seq, None)
    steady_dev = steady_ms(timed.done_at)
    clocks = sampler.stop()
    for w in range(K):
        step_e2e(w)
    ms_e2e, st_e2e = timed(step_e2e, args.steps, K)

    free_b, total_b = torch.cuda.mem_get_info(dev)
    hbm_in_use = (total_b - free_b) / 2**30     # index, tables and every in-flight context's per-read state and arenas
    for e in engs["pri"][:1]:
        e.set("count", 1)                   # one untimed step with the pop / sector counters compiled in
    counted = step_device(0)
    engs["pri"][0].set("count", 0)

    n_glob = (world * n_rank) if weak else n_step          # distinct reads of one step, all ranks
    total_reads = n_glob * args.steps
    value = total_reads / (ms_dev * 1e-3)
    e2e_value = total_reads / (ms_e2e * 1e-3)
    total_rec = sum(t for (t, _) in st_e2e[-1])
    launches = int(sum(s["kernel_launches"] for step in st_seq for s in step))
    pri_ix = [j for j, (ix, _) in enumerate(streams) if ix == "pri"]
    n_pri = sum(dsets[i]["n"] for (ix, i) in streams if ix == "pri")

    def kms(key):                           # per step: summed over the step's streams
        return [float(sum(s[key] for s in step)) for step in st_seq]

    step_ms = kms("ms_total")
    h2d = int(sum((12 + L) * dsets[i]["n"] for (_, i) in streams))
    out = {"metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
           "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "u32",
           "data": "synthetic", "config": config, "clocks": clocks, "in_flight": K,
           "steady_state": None if not steady_dev else {
               "ms_between_steps": steady_dev, "reads_per_s_this_rank": n_rank / (steady_dev * 1e-3),
               "note": "mean spacing of step completions over the middle half of the pipelined run on this rank: the "
                       "rate between the pipeline's fill and drain (informational)"},
           "sequential": {"value": total_reads / (ms_seq * 1e-3), "ms_per_step": ms_seq / args.steps,
                          "best_step_ms": float(np.min(step_ms)), "median_step_ms": float(np.median(step_ms)),
                          "note": "the same steps one at a time on one context (rank 0's device times)"},
           "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": int(4 * sum(dsets[i]["n"] for (_, i) in streams) + total_rec * 16),
                   "ms_per_step": ms_e2e / args.steps, "in_flight": K},
           "gpu_launches": launches,
           "kernel_ms": {k: float(np.mean(kms(k))) for k in ("ms_width", "ms_search", "ms_compact", "ms_total")},
           "overflow_reads_per_step": int(sum(s["overflow_reads"] for s in st_seq[-1])),
           "hbm_in_use_gib": hbm_in_use,
           "device_pops_per_read": sum(counted[j]["pops"] for j in pri_ix) / max(n_pri, 1),
           "device_sectors_per_read": sum(counted[j]["occ_lookups"] for j in pri_ix) / max(n_pri, 1)}

    ok = True
    if rank == 0:
        # roofline: algorithmic bytes = 32 B x occ lookups of the REFERENCE algorithm (oracle-counted on a sample)
        samples = [s[: min(int(s.shape[0]), 2_000_000)].cpu().numpy() for s in sets]
        port = time_port(bwt, rbwt, samples[0], opt, n=3000)
        lookups_per_read = port["stats"]["lookups"] / port["n"]
        bytes_per_read = 32.0 * lookups_per_read
        ms_search = float(np.mean([sum(step[j]["ms_search"] for j in pri_ix) for step in st_seq]))
        achieved = bytes_per_read * n_pri / (ms_search * 1e-3) / 1e9
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, which = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, which = 6650.0, "fallback (B200_PROFILING.md)"
        traffic = traffic_src = None        # DRAM bytes of the dominant kernel per launch, from the committed ncu capture
        tpath = os.path.join(ROOT, "profiles", "r2_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath)).get(f"config{args.config}")
            if tj:
                traffic = float(tj["k_search"]["dram_bytes_per_read"]) * n_pri
                traffic_src = (f"profiles/r2_traffic.json: ncu dram__bytes_read+write of k_search per read, captured "
                               f"at {tj['capture_reads']} reads per launch, x the reads of this launch")
        sector_roof = eng.sector_roofline(1 << 28, 3)
        pair_roof = eng.sector_roofline(1 << 27, -3)
        out["roofline"] = {"bound": "hbm", "kernel": "k_search", "achieved": achieved, "peak": peak, "unit": "GB/s",
                           "frac": achieved / peak, "traffic": traffic, "peak_source": which,
                           "traffic_source": traffic_src,
                           "algorithmic_bytes": bytes_per_read * n_pri,
                           "algorithmic_bytes_per_read": bytes_per_read,
                           "oracle_lookups_per_read": lookups_per_read,
                           "oracle_pops_per_read": port["stats"]["pops"] / port["n"],
                           "kernel_ms_per_launch": ms_search, "reads_per_launch": n_pri,
                           "random_sector_roof_gbs": sector_roof, "random_64B_pair_roof_gbs": pair_roof,
                           "random_sector_roof_note": "k_sector_gather over BOTH device indexes (3.1 GB)",
                           "frac_of_random_sector_roof": achieved / sector_roof if sector_roof else None,
                           "occ_sectors_per_s": sum(counted[j]["occ_lookups"] for j in pri_ix) / (ms_search * 1e-3)}
        if traffic and sector_roof:
            # What binds this access pattern is a transaction rate, not bytes: DRAM moves the kernel's
            # traffic in 64-byte transactions, and the gather microbenchmark counts how many random ones it serves.
            tj = json.load(open(tpath))[f"config{args.config}"]["k_search"]
            rd_tx = tj["dram_read_bytes_per_read"] / 64.0
            wr_tx = tj["dram_write_bytes_per_read"] / 64.0
            rate = n_pri / (ms_search * 1e-3)
            out["roofline"].update({
                "dram_read_transactions_per_read": rd_tx, "dram_write_transactions_per_read": wr_tx,
                "dram_read_transactions_per_s": rd_tx * rate,
                "random_requests_roof_per_s": sector_roof * 1e9 / 32.0,
                "frac_of_transaction_roof_reads": rd_tx * rate / (sector_roof * 1e9 / 32.0),
                "frac_of_transaction_roof_reads_and_writebacks": (rd_tx + wr_tx) * rate / (sector_roof * 1e9 / 32.0)})
        # parity of this very run against the oracle port on the sample
        m = port["n"]
        n_aln_d, rec_d = eng.cal_sa_reg_gap(np.full(m, L, np.int32), np.arange(m, dtype=np.int64) * L,
                                            samples[0][:m].reshape(-1), opt)
        parity = {"oracle_port_reads": m,
                  "oracle_port_identical": bool(np.array_equal(n_aln_d, port["n_aln"]) and
                                                rec_d.tobytes() == port["rec"].tobytes())}
        ok = parity["oracle_port_identical"]
        if world == 1 and not args.no_cpu_baseline:
            if os.path.exists(REF_BIN):
                cap = args.parity_pairs if cfg["pairs"] else 2_000_000
                res = time_reference(ref_streams(samples), nproc, work_dir, target_s=15.0, max_reads=cap,
                                     aln_args=aln_args)
                out["cpu_baseline"] = {"value": res["reads_per_s"], "unit": "reads/s", "cores": nproc,
                                       "kind": "reference",
                                       "sample": f"{res['distinct_reads']} reads of the workload, {len(streams)} x "
                                                 f"`ibwa aln -t {nproc}`, wall {res['wall_s']:.2f} s minus "
                                                 f"{res['index_load_s']:.2f} s index load"}
                t1 = time_reference(ref_streams(samples)[:1], 1, work_dir, target_s=6.0, max_reads=200_000,
                                    aln_args=aln_args, tag="t1")
                out["cpu_baseline"]["t1_value"] = t1["reads_per_s"]
                out["cpu_baseline"]["t1_sample"] = f"{t1['n']} reads, `ibwa aln -t 1`"
                out["cpu_baseline"]["ideal_all_core_bound"] = t1["reads_per_s"] * nproc
                k = res["n"]
                same = True
                eng_sais = []
                for j, (ix, i) in enumerate(streams):
                    e = engs[ix][0]
                    g_n, g_rec = e.cal_sa_reg_gap(np.full(k, L, np.int32), np.arange(k, dtype=np.int64) * L,
                                                  samples[i][:k].reshape(-1), opt)
                    _, r_n, r_rec = sai.read_sai(res["sais"][j])
                    same = same and bool(np.array_equal(g_n, r_n[:k]) and g_rec.tobytes() == r_rec.tobytes())
                    if cfg["pairs"]:
                        p = os.path.join(work_dir, f"engine_{j}.sai")
                        with open(p, "wb") as f:
                            sai.write_header(f, opt)
                            sai.write_batch(f, g_n, g_rec)
                        eng_sais.append(p)
                parity["reference_binary_reads"] = k * len(streams)
                parity["reference_binary_identical"] = same
                # the same reads once more with parking forced on (budget parking, heavy warps, drain parking, resume
                # launches): in the timed runs it switches itself on only while four or more batches are in flight
                eng.set("susp", 16)
                p_n, p_rec = eng.cal_sa_reg_gap(np.full(k, L, np.int32), np.arange(k, dtype=np.int64) * L,
                                                samples[0][:k].reshape(-1), opt)
                eng.set("susp", -16)
                _, r_n, r_rec = sai.read_sai(res["sais"][0])
                parity["parked_identical"] = bool(np.array_equal(p_n, r_n[:k]) and p_rec.tobytes() == r_rec.tobytes())
                same = same and parity["parked_identical"]
                ok = ok and same
                if cfg["pairs"]:            # downstream: the unchanged `sampe -R` on engine vs reference .sai
                    np_ = len(sets)
                    sam_e, sam_r = os.path.join(work_dir, "engine.sam"), os.path.join(work_dir, "reference.sam")
                    t_s = _sampe(prefix, eng_sais[:np_], res["fqs"][:np_], alt["prefix"] if alt else None,
                                 eng_sais[np_:], sam_e, nproc)
                    _sampe(prefix, res["sais"][:np_], res["fqs"][:np_], alt["prefix"] if alt else None,
                           res["sais"][np_:], sam_r, nproc)
                    md_e, md_r = _md5(sam_e), _md5(sam_r)
                    body = [ln for ln in open(sam_e, "rb") if not ln.startswith(b"@")]
                    parity["sampe"] = {"pairs": k, "sam_lines": len(body), "sam_md5_engine": md_e,
                                       "sam_md5_reference": md_r, "sam_identical": md_e == md_r,
                                       "mapped": sum(1 for ln in body if not int(ln.split(b"\t")[1]) & 4),
                                       "zr_tags": sum(1 for ln in body if b"ZR:Z" in ln), "sampe_wall_s": t_s,
                                       "command": "ibwa sampe -R -t N <pri> e1.sai e2.sai r1.fq r2.fq"
                                                  + (" <alt> a1.sai a2.sai (with <alt>.remap)" if alt else "")}
                    ok = ok and md_e == md_r and len(body) == 2 * k
            else:
                out["cpu_baseline"] = {"value": port["reads_per_s"], "unit": "reads/s", "cores": 1, "kind": "port",
                                       "sample": f"{port['n']} reads, single-thread oracle port"}
        out["parity"] = parity
        if not ok:                          # a wrong result has no throughput
            out["value"] = None
            out["e2e"]["value"] = None
            out["error"] = "parity failed: engine output differs from the checker's"
        print(json.dumps(out), file=json_out, flush=True)
    for lst in engs.values():
        for e in lst[1:]:
            e.close()
        lst[0].close()
    if use_dist:
        dist.destroy_process_group()
    return 0 if ok else 3


if __name__ == "__main__":
    sys.exit(main())
