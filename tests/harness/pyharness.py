"""ctypes binding of tests/harness/libhostharness.so: the product's per-read state
machines (ibwa_b200/csrc/aln_core.cuh) compiled for the CPU.  Test-only."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libhostharness.so")
SRCS = [os.path.join(HERE, "host_harness.cpp"), os.path.join(ROOT, "ibwa_b200", "csrc", "aln_core.cuh"),
        os.path.join(ROOT, "ibwa_b200", "csrc", "alngrp_core.cuh"),
        os.path.join(ROOT, "ibwa_b200", "csrc", "fm_layout.cuh"), os.path.join(ROOT, "ibwa_b200", "csrc", "host_params.h")]
ALN_DTYPE = np.dtype([("packed", "<u4"), ("k", "<u4"), ("l", "<u4"), ("score", "<i4")])


class BwtView(ctypes.Structure):
    _fields_ = [("primary", ctypes.c_uint32), ("L2", ctypes.c_uint32 * 5), ("seq_len", ctypes.c_uint32),
                ("bwt_size", ctypes.c_uint64), ("bwt", ctypes.c_void_p)]


def view(b):
    v = BwtView()
    v.primary = b.primary
    for i in range(5):
        v.L2[i] = int(b.L2[i])
    v.seq_len = b.seq_len
    arr = np.ascontiguousarray(b.bwt, dtype=np.uint32)
    v._keep = arr
    v.bwt_size = arr.shape[0]
    v.bwt = arr.ctypes.data
    return v


def build():
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in SRCS):
        subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-DB200ALN_COUNTERS", "-DB2_EARLY_Q", "-o", LIB,
                               SRCS[0]])
    return LIB


LIB_CHECKED = os.path.join(HERE, "libhostharness_checked.so")


def build_checked():
    """The same sources with -DB2_CHECKED (aln_core.cuh): every derived index is tested, a violation aborts."""
    if not os.path.exists(LIB_CHECKED) or any(os.path.getmtime(s) > os.path.getmtime(LIB_CHECKED) for s in SRCS):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-DB200ALN_COUNTERS", "-DB2_EARLY_Q", "-DB2_CHECKED",
                               "-o", LIB_CHECKED, SRCS[0]])
    return LIB_CHECKED


_lib = None
_lib_checked = None


def lib(checked=False):
    global _lib, _lib_checked
    if checked:
        if _lib_checked is None:
            _lib_checked = ctypes.CDLL(build_checked())
            _lib_checked.hh_aln_batch.restype = ctypes.c_int64
        return _lib_checked
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.hh_aln_batch.restype = ctypes.c_int64
    return _lib


def aln_batch(bwt, rbwt, lens, offs, codes, opt_c, arena_cap=1 << 16, rec_cap=256, reuse=False, big_cap=0, batch_max_len=0,
              lut_k=0, rounds=0, q16=False, checked=False, suspend_every=0):
    L = lib(checked)
    L.hh_set_suspend_every(ctypes.c_int(suspend_every))   # > 0: park / resume the lane every so many steps
    L.hh_set_q16(ctypes.c_int(int(q16)))    # the product's fast pass: 16-bit width records when the options allow
    L.hh_set_lut_k(ctypes.c_int(lut_k))
    L.hh_set_rounds(ctypes.c_int(rounds))   # 0 = unlimited; the fast kernel runs with 1
    lens = np.ascontiguousarray(lens, np.int32)
    offs = np.ascontiguousarray(offs, np.int64)
    codes = np.ascontiguousarray(codes, np.uint8)
    n = len(lens)
    n_aln = np.zeros(n, np.int32)
    rec_p = ctypes.c_void_p()
    nov = ctypes.c_int64()
    counters = (ctypes.c_uint64 * 2)()
    v0, v1 = view(bwt), view(rbwt)
    tot = L.hh_aln_batch(ctypes.byref(v0), ctypes.byref(v1), ctypes.c_int(n), ctypes.c_void_p(lens.ctypes.data),
                         ctypes.c_void_p(offs.ctypes.data), ctypes.c_void_p(codes.ctypes.data), ctypes.byref(opt_c),
                         ctypes.c_uint32(arena_cap), ctypes.c_int(rec_cap), ctypes.c_int(int(reuse)), ctypes.c_uint32(big_cap), ctypes.c_int(batch_max_len),
                         ctypes.c_void_p(n_aln.ctypes.data), ctypes.byref(rec_p), ctypes.byref(nov), counters)
    rec = np.frombuffer((ctypes.c_uint8 * (16 * tot)).from_address(rec_p.value), dtype=ALN_DTYPE).copy() if tot \
        else np.empty(0, ALN_DTYPE)
    L.hh_free(rec_p)
    return n_aln, rec, nov.value, {"pops": counters[0], "sectors": counters[1]}


def bwt_sa(bwt, sa, rows):
    L = lib()
    rows = np.ascontiguousarray(rows, np.uint32)
    arr = np.ascontiguousarray(sa.sa, np.uint32)
    out = np.empty(len(rows), np.uint32)
    v = view(bwt)
    L.hh_bwt_sa(ctypes.byref(v), ctypes.c_void_p(arr.ctypes.data), ctypes.c_uint32(sa.sa_intv),
                ctypes.c_int64(len(rows)), ctypes.c_void_p(rows.ctypes.data), ctypes.c_void_p(out.ctypes.data))
    return out


def alngrp_merge(n_alns, recs, s_mm):
    """alngrp_core.cuh (scope row N4) on the CPU; same layout as Engine.alngrp_merge."""
    L = lib()
    L.hh_alngrp_merge.restype = ctypes.c_int64
    ns, n = len(n_alns), len(n_alns[0])
    n_alns = [np.ascontiguousarray(a, dtype=np.int32) for a in n_alns]
    recs = [np.ascontiguousarray(r, dtype=ALN_DTYPE) for r in recs]
    total = int(sum(int(a.sum()) for a in n_alns))
    pn = (ctypes.c_void_p * ns)(*[a.ctypes.data for a in n_alns])
    pr = (ctypes.c_void_p * ns)(*[r.ctypes.data for r in recs])
    out_off = np.zeros(n, np.int64)
    out_n = np.zeros(n, np.int32)
    out_rec = np.zeros(max(total, 1), ALN_DTYPE)
    out_db = np.zeros(max(total, 1), np.uint32)
    got = L.hh_alngrp_merge(ctypes.c_int(ns), ctypes.c_int(n), pn, pr, ctypes.c_int(int(s_mm)),
                            ctypes.c_void_p(out_off.ctypes.data), ctypes.c_void_p(out_n.ctypes.data),
                            ctypes.c_void_p(out_rec.ctypes.data), ctypes.c_void_p(out_db.ctypes.data))
    assert got == total
    return out_off, out_n, out_rec[:total], out_db[:total]
