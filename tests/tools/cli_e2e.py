#!/usr/bin/env python
"""Process-seam end-to-end check on the GPU box: `b200aln aln` vs `ibwa aln -t <cores>` on the same FASTQ
and the cached 3.1 Gbp bench index (run bench.py once before).  Wall clock includes index load, parsing,
the search and writing the .sai; the outputs must be byte-identical except header bytes 52..55.
usage: python tests/tools/cli_e2e.py [n_reads] [ref_reads]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    n_ref = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000
    genome_bp, seed = 3_100_000_000, 20260102
    dev = torch.device("cuda", 0)
    bwt, rbwt, text, prefix = bench.load_or_build_index(genome_bp, seed, dev, True)
    reads = bench.synth_reads_torch(text, n, 100, seed + 7).cpu().numpy()
    del text, bwt, rbwt
    torch.cuda.empty_cache()
    fq, fq_ref = "/tmp/cli_e2e.fq", "/tmp/cli_e2e_ref.fq"
    bench.write_fastq(fq, reads)
    bench.write_fastq(fq_ref, reads[:n_ref])
    os.sync()  # the write-back of 2 GB of dirty pages would otherwise run next to the first timed process
    exe = os.path.join(ROOT, "ibwa_b200", "b200aln")
    cores = os.cpu_count()
    def run_cli(path, env):
        """wall clock of the whole process, and the driver's own timeline (B200ALN_TRACE): seconds until the index is
        resident on the device, seconds from there until the output is closed"""
        t0 = time.perf_counter()
        p = subprocess.run([exe, "aln", "-f", "/tmp/gpu_out.sai", prefix, path], stderr=subprocess.PIPE, check=True,
                           env=dict(os.environ, B200ALN_TRACE="1", **env))
        dt = time.perf_counter() - t0
        t_load = t_done = t_parsed = None
        for line in p.stderr.decode(errors="replace").splitlines():
            if line.startswith("[trace]"):
                try:
                    sec = float(line.split()[1])
                except ValueError:
                    continue
                if "index resident" in line:
                    t_load = sec
                elif "output closed" in line:
                    t_done = sec
                elif "parsed reads" in line:
                    t_parsed = sec
        if os.environ.get("CLI_E2E_TRACE"):  # the driver's whole timeline of this run, for a look at the stages
            with open(os.environ["CLI_E2E_TRACE"], "ab") as f:
                f.write(b"==== " + path.encode() + b" " + repr(env).encode() + b"\n" + p.stderr)
        return dt, t_load, t_done, t_parsed

    # the reader alone (parse only, results dropped), for comparison with the whole pipeline
    import ctypes
    from ibwa_b200 import engine
    L = engine.load_library()
    for rep in range(2):
        t0 = time.perf_counter()
        r = L.b200aln_reader_open(fq.encode(), 0)
        pl, po, pc, nb = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
        tot = 0
        while True:
            k = L.b200aln_reader_next(r, 0x40000, 0, 0, ctypes.byref(pl), ctypes.byref(po), ctypes.byref(pc), ctypes.byref(nb))
            if k == 0:
                break
            tot += k
        L.b200aln_reader_close(r)
        dt = time.perf_counter() - t0
        print(f"reader alone ({tot} reads, {cores} cores): {dt:.2f} s = {tot / dt / 1e6:.2f} M reads/s", flush=True)

    # gzip input: the first 2 M reads, compressed like most FASTQ files are (gzip -6); reader alone with the decoder of
    # this library and with zlib, then the whole command line
    n_gz = min(n, 2_000_000)
    fq_gz = "/tmp/cli_e2e_gz.fq.gz"
    bench.write_fastq("/tmp/cli_e2e_gz.fq", reads[:n_gz])
    subprocess.run("gzip -6 -c /tmp/cli_e2e_gz.fq > " + fq_gz, shell=True, check=True)
    os.sync()
    for env_name, tag in ((None, "fast_inflate.h"), ("B200ALN_NO_FAST_INFLATE", "zlib")):
        if env_name:
            os.environ[env_name] = "1"
        t0 = time.perf_counter()
        r = L.b200aln_reader_open(fq_gz.encode(), 0)
        tot = 0
        while True:
            k = L.b200aln_reader_next(r, 0x40000, 0, 0, ctypes.byref(pl), ctypes.byref(po), ctypes.byref(pc), ctypes.byref(nb))
            if k == 0:
                break
            tot += k
        L.b200aln_reader_close(r)
        dt = time.perf_counter() - t0
        if env_name:
            del os.environ[env_name]
        print(f"reader alone, gzip input ({tot} reads, {tag}): {dt:.2f} s = {tot / dt / 1e6:.2f} M reads/s", flush=True)
    dt, t_load, t_done, t_parsed = run_cli(fq_gz, {})
    print(f"b200aln aln (gzip input, {n_gz} reads): {dt:.2f} s wall incl. index load ({t_load:.2f} s); parse + search + write "
          f"{t_done - t_load:.2f} s = {n_gz / (t_done - t_load) / 1e6:.2f} M reads/s", flush=True)

    for label, path, cnt in (("full", fq, n), ("ref-sized", fq_ref, n_ref)):
        full_md5 = set()
        for env in ({}, {}, {"B200ALN_NO_PREALLOC": "1"}, {"B200ALN_MERGE": "16"}, {"B200ALN_INFLIGHT": "3"},
                    {"B200ALN_MERGE": "1"}) if label == "full" else ({},):
            dt, t_load, t_done, t_parsed = run_cli(path, env)
            if label == "ref-sized":
                os.replace("/tmp/gpu_out.sai", "/tmp/gpu_ref-sized.sai")
            else:
                md5 = subprocess.run(["md5sum", "/tmp/gpu_out.sai"], stdout=subprocess.PIPE, check=True).stdout.split()[0]
                full_md5.add(md5)
                print("  md5", md5.decode(), os.path.getsize("/tmp/gpu_out.sai"))
                if len(full_md5) == 1 and not os.path.exists("/tmp/gpu_first.sai"):
                    os.replace("/tmp/gpu_out.sai", "/tmp/gpu_first.sai")
                elif os.path.exists("/tmp/gpu_out.sai"):
                    import numpy as np
                    a = np.fromfile("/tmp/gpu_first.sai", dtype=np.uint8)
                    b = np.fromfile("/tmp/gpu_out.sai", dtype=np.uint8)
                    m = min(len(a), len(b))
                    d = np.flatnonzero(a[:m] != b[:m])
                    if len(d) or len(a) != len(b):
                        print(f"  DIFFERS from the first run: sizes {len(a)} / {len(b)}, {len(d)} differing bytes, first at "
                              f"{d[0] if len(d) else m}, last at {d[-1] if len(d) else m}", flush=True)
            print(f"b200aln aln ({label}: {cnt} reads{', ' + str(env) if env else ''}): {dt:.2f} s wall = "
                  f"{cnt / dt / 1e6:.2f} M reads/s incl. index load ({t_load:.2f} s); parse + search + write "
                  f"{t_done - t_load:.2f} s = {cnt / (t_done - t_load) / 1e6:.2f} M reads/s (last batch parsed "
                  f"{t_parsed - t_load:.2f} s after the index was resident)", flush=True)
        if label == "full":
            print("all runs of the full input wrote the same bytes:", len(full_md5) == 1)
    # a longer input (the same reads four times over): what the pipeline does once it is full
    big = "/tmp/cli_e2e_big.fq"
    with open(big, "wb") as fo:
        for _ in range(4):
            subprocess.run(["cat", fq], stdout=fo, check=True)
    os.sync()
    for env in ({}, {"B200ALN_MERGE": "16"}):
        if os.environ.get("CLI_E2E_QUICK"):
            break
        dt, t_load, t_done, t_parsed = run_cli(big, env)
        print(f"b200aln aln (4 x full: {4 * n} reads{', ' + str(env) if env else ''}): {dt:.2f} s wall = "
              f"{4 * n / dt / 1e6:.2f} M reads/s incl. index load ({t_load:.2f} s); parse + search + write "
              f"{t_done - t_load:.2f} s = {4 * n / (t_done - t_load) / 1e6:.2f} M reads/s", flush=True)
    os.remove(big)
    t0 = time.perf_counter()
    with open("/tmp/ref.sai", "wb") as fo:
        subprocess.run([bench.REF_BIN, "aln", "-t", str(cores), prefix, fq_ref], stdout=fo, stderr=subprocess.DEVNULL,
                       check=True)
    dt = time.perf_counter() - t0
    print(f"ibwa aln -t {cores} ({n_ref} reads): {dt:.2f} s wall = {n_ref / dt / 1e6:.3f} M reads/s incl. index load")
    a, b = open("/tmp/ref.sai", "rb").read(), open("/tmp/gpu_ref-sized.sai", "rb").read()
    same = a[:52] == b[:52] and a[56:] == b[56:]
    print("byte-identical .sai (except n_threads):", same)
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
