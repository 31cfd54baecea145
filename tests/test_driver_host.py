"""The host driver (b200aln_aln_core in ibwa_b200/csrc/aln_host.cpp: parse units, launches cut at the batch-level
clamp, one worker per (GPU, slot), output in input order) in a container without a GPU: the driver's own object code
linked with tests/harness/stub_device.cpp, which computes a batch with the CPU build of the product's state machines
instead of the kernels.  What is checked is the driver's logic; the kernels are checked by the -m gpu tests."""
import ctypes
import gzip
import io
import os
import subprocess

import numpy as np
import pytest

from ibwa_b200 import parse_aln_args, sai, seqio
from harness import pyharness
from cases import CASES

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "harness", "libdriverstub.so")
SRCS = [os.path.join(ROOT, "ibwa_b200", "csrc", "aln_host.cpp"), os.path.join(HERE, "harness", "stub_device.cpp"),
        os.path.join(HERE, "harness", "host_harness.cpp")]
DEPS = SRCS + [os.path.join(ROOT, "ibwa_b200", "csrc", f) for f in ("aln_core.cuh", "host_params.h", "fm_layout.cuh")] + \
    [os.path.join(ROOT, "include", "b200aln.h")]


@pytest.fixture(scope="module")
def stub():
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in DEPS):
        subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-DB200ALN_COUNTERS", "-DB2_EARLY_Q",
                               "-o", LIB] + SRCS + ["-lz", "-lpthread"])
    L = ctypes.CDLL(LIB)
    L.b200aln_aln_core.restype = ctypes.c_int64
    L.b200aln_aln_core.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    return L


def run_driver(L, prefix, fq, args, tmp_path, device, **env):
    opt, _, _, _ = parse_aln_args(args + ["prefix", fq])
    oc = opt.to_c()
    out = str(tmp_path / "out.sai")
    keys = ("B200ALN_BATCH_READS", "B200ALN_MERGE", "B200ALN_INFLIGHT", "B200ALN_STUB_DEVICES")
    saved = {k: os.environ.get(k) for k in keys}
    try:
        for k in keys:
            os.environ.pop(k, None)
        for k, v in env.items():
            os.environ[k] = str(v)
        fd = os.open(out, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
        try:
            n = L.b200aln_aln_core(prefix.encode(), fq.encode(), ctypes.byref(oc), fd, device)
        finally:
            os.close(fd)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return n, open(out, "rb").read()


def expected_by_batches(g1_index, args, fq, batch_reads):
    """what bwa_aln_core writes when it works in batches of batch_reads reads (bwtaln.c:193-232): every batch on its own"""
    opt, _, _, _ = parse_aln_args(args + ["prefix", fq])
    buf = io.BytesIO()
    sai.write_header(buf, opt)
    n = 0
    for batch in seqio.read_batches(fq, opt.mode, opt.trim_qual, n_needed=batch_reads):
        n_aln, rec, nov, _ = pyharness.aln_batch(g1_index[0], g1_index[1], batch.lens, batch.offs, batch.codes, opt.to_c(),
                                                 arena_cap=32000, rec_cap=4096, reuse=True, big_cap=1 << 21)
        assert (n_aln >= 0).all()
        sai.write_batch(buf, n_aln, rec)
        n += len(batch.lens)
    return n, buf.getvalue()


@pytest.mark.parametrize("tag", ["default", "q20", "short_o3"])
def test_driver_one_reference_batch_matches_golden(stub, tag, golden_dir, tmp_path):
    args, fq = CASES[tag]
    n, got = run_driver(stub, os.path.join(golden_dir, "g1"), os.path.join(golden_dir, fq + ".fq.gz"), args, tmp_path, 0)
    want = open(os.path.join(golden_dir, f"g1_{tag}.sai"), "rb").read()
    assert got == want


@pytest.mark.parametrize("env", [
    dict(B200ALN_BATCH_READS=100, B200ALN_MERGE=1, B200ALN_INFLIGHT=1),
    dict(B200ALN_BATCH_READS=100, B200ALN_MERGE=8, B200ALN_INFLIGHT=4),
    dict(B200ALN_BATCH_READS=64, B200ALN_MERGE=3, B200ALN_INFLIGHT=2),
    dict(B200ALN_BATCH_READS=37, B200ALN_MERGE=64, B200ALN_INFLIGHT=8),
])
def test_driver_batches_launches_and_order(stub, env, golden_dir, g1_index, tmp_path):
    """mixed read lengths (20 - 250 bp): consecutive batches disagree on the max_gapo clamp again and again, so units are
    cut into several launches; whatever the unit size and the number of workers, the output is batch-by-batch"""
    fq = os.path.join(golden_dir, "g1_reads.fq.gz")
    n_want, want = expected_by_batches(g1_index, [], fq, env["B200ALN_BATCH_READS"])
    n, got = run_driver(stub, os.path.join(golden_dir, "g1"), fq, [], tmp_path, 0, **env)
    assert n == n_want == 1306
    assert got == want


def test_driver_every_visible_device(stub, golden_dir, g1_index, tmp_path):
    """device = -1: launches go round the workers of all devices (three here), output still in input order"""
    fq = os.path.join(golden_dir, "g1_reads.fq.gz")
    _, want = expected_by_batches(g1_index, ["-o", "2"], fq, 50)
    n, got = run_driver(stub, os.path.join(golden_dir, "g1"), fq, ["-o", "2"], tmp_path, -1, B200ALN_BATCH_READS=50,
                        B200ALN_MERGE=2, B200ALN_INFLIGHT=2, B200ALN_STUB_DEVICES=3)
    assert n == 1306 and got == want


def test_driver_plain_file_and_empty_input(stub, golden_dir, g1_index, tmp_path):
    fq = str(tmp_path / "reads.fq")
    with open(fq, "wb") as f:
        f.write(gzip.open(os.path.join(golden_dir, "g1_short.fq.gz")).read())
    _, want = expected_by_batches(g1_index, [], fq, 128)
    n, got = run_driver(stub, os.path.join(golden_dir, "g1"), fq, [], tmp_path, 0, B200ALN_BATCH_READS=128)
    assert n == 400 and got == want
    empty = str(tmp_path / "empty.fq")
    open(empty, "wb").close()
    n, got = run_driver(stub, os.path.join(golden_dir, "g1"), empty, [], tmp_path, 0)
    hdr = io.BytesIO()
    sai.write_header(hdr, parse_aln_args(["prefix", empty])[0])
    assert n == 0 and got == hdr.getvalue()
