"""ctypes binding of oracle/libalnoracle.so (the CPU restatement) and helpers to
run the unmodified reference binary oracle/_ref/ibwa.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libalnoracle.so")
REF_BIN = os.path.join(HERE, "_ref", "ibwa")


class OrcBwt(ctypes.Structure):
    _fields_ = [("primary", ctypes.c_uint32), ("L2", ctypes.c_uint32 * 5), ("seq_len", ctypes.c_uint32),
                ("n_words", ctypes.c_uint64), ("bwt", ctypes.c_void_p)]


class OrcSa(ctypes.Structure):
    _fields_ = [("sa_intv", ctypes.c_int32), ("n_sa", ctypes.c_uint64), ("sa", ctypes.c_void_p)]


class OrcStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in
                ("pops", "pushes", "lookups", "occ1_calls", "occ4_calls", "hits", "stack_hwm", "cutoff_reads",
                 "inherit_violations")]


ALN_DTYPE = np.dtype([("packed", "<u4"), ("k", "<u4"), ("l", "<u4"), ("score", "<i4")])
WIDTH_DTYPE = np.dtype([("w", "<u4"), ("bid", "<i4")])


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
            os.path.join(HERE, "alnoracle.c")):
        subprocess.check_call(["make", "-C", HERE, "libalnoracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB_PATH)
        L.orc_occ.restype = ctypes.c_uint32
        L.orc_occ.argtypes = [ctypes.POINTER(OrcBwt), ctypes.c_uint32, ctypes.c_int]
        L.orc_occ4.argtypes = [ctypes.POINTER(OrcBwt), ctypes.c_uint32, ctypes.c_void_p]
        L.orc_cal_width.restype = ctypes.c_int
        L.orc_cal_width.argtypes = [ctypes.POINTER(OrcBwt), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_cal_maxdiff.restype = ctypes.c_int
        L.orc_cal_maxdiff.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double]
        L.orc_trim_len.restype = ctypes.c_int
        L.orc_trim_len.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.orc_aln_batch.restype = ctypes.c_int64
        L.orc_aln_batch.argtypes = [ctypes.POINTER(OrcBwt), ctypes.POINTER(OrcBwt), ctypes.c_int, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.POINTER(ctypes.c_void_p)]
        L.orc_free.argtypes = [ctypes.c_void_p]
        L.orc_bwt_sa.restype = ctypes.c_uint32
        L.orc_bwt_sa.argtypes = [ctypes.POINTER(OrcBwt), ctypes.POINTER(OrcSa), ctypes.c_uint32]
        L.orc_sa2seq.restype = ctypes.c_uint64
        L.orc_sa2seq.argtypes = [ctypes.POINTER(OrcBwt), ctypes.POINTER(OrcSa), ctypes.POINTER(OrcBwt),
                                 ctypes.POINTER(OrcSa), ctypes.c_int, ctypes.c_uint32, ctypes.c_int]
        L.orc_stats_get.argtypes = [ctypes.POINTER(OrcStats)]
        _lib = L
    return _lib


def as_orc_bwt(b) -> OrcBwt:
    """b: ibwa_b200.bwtio.Bwt (keeps a reference to the numpy payload)."""
    o = OrcBwt()
    o.primary = b.primary
    for i in range(5):
        o.L2[i] = int(b.L2[i])
    o.seq_len = b.seq_len
    o.n_words = b.bwt.shape[0]
    arr = np.ascontiguousarray(b.bwt, dtype=np.uint32)
    o._keep = arr
    o.bwt = arr.ctypes.data
    return o


def as_orc_sa(s) -> OrcSa:
    """s: ibwa_b200.bwtio.Sa"""
    o = OrcSa()
    o.sa_intv = s.sa_intv
    arr = np.ascontiguousarray(s.sa, dtype=np.uint32)
    o._keep = arr
    o.n_sa = arr.shape[0]
    o.sa = arr.ctypes.data
    return o


def bwt_sa(ob: OrcBwt, osa: OrcSa, rows) -> np.ndarray:
    L = lib()
    return np.array([L.orc_bwt_sa(ctypes.byref(ob), ctypes.byref(osa), int(k) & 0xFFFFFFFF) for k in rows],
                    dtype=np.uint32)


def alngrp_merge(n_alns, recs, s_mm: int):
    """orc_alngrp_merge: (out_off, out_n, records, dbidx) — the layout of Engine.alngrp_merge."""
    L = lib()
    L.orc_alngrp_merge.restype = ctypes.c_int64
    ns, n = len(n_alns), len(n_alns[0])
    n_alns = [np.ascontiguousarray(a, dtype=np.int32) for a in n_alns]
    recs = [np.ascontiguousarray(r, dtype=ALN_DTYPE) for r in recs]
    total = int(sum(int(a.sum()) for a in n_alns))
    pn = (ctypes.c_void_p * ns)(*[a.ctypes.data for a in n_alns])
    pr = (ctypes.c_void_p * ns)(*[r.ctypes.data for r in recs])
    out_off = np.zeros(n, dtype=np.int64)
    out_n = np.zeros(n, dtype=np.int32)
    out_rec = np.zeros(max(total, 1), dtype=ALN_DTYPE)
    out_db = np.zeros(max(total, 1), dtype=np.uint32)
    got = L.orc_alngrp_merge(ctypes.c_int(ns), ctypes.c_int(n), pn, pr, ctypes.c_int(int(s_mm)),
                             ctypes.c_void_p(out_off.ctypes.data), ctypes.c_void_p(out_n.ctypes.data),
                             ctypes.c_void_p(out_rec.ctypes.data), ctypes.c_void_p(out_db.ctypes.data))
    assert got == total
    return out_off, out_n, out_rec[:total], out_db[:total]


def occ(ob: OrcBwt, k: int, c: int) -> int:
    return lib().orc_occ(ctypes.byref(ob), k & 0xFFFFFFFF, c)


def occ4(ob: OrcBwt, k: int) -> np.ndarray:
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_occ4(ctypes.byref(ob), k & 0xFFFFFFFF, out.ctypes.data)
    return out


def cal_width(ob: OrcBwt, s: np.ndarray) -> np.ndarray:
    s = np.ascontiguousarray(s, dtype=np.uint8)
    out = np.zeros(len(s) + 1, dtype=WIDTH_DTYPE)
    lib().orc_cal_width(ctypes.byref(ob), len(s), s.ctypes.data, out.ctypes.data)
    return out


def aln_batch(ob: OrcBwt, orb: OrcBwt, lens: np.ndarray, offs: np.ndarray, fwd: np.ndarray, opt_c):
    """Returns (n_aln int32[n], records ALN_DTYPE[total], stats dict)."""
    L = lib()
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    offs = np.ascontiguousarray(offs, dtype=np.int64)
    fwd = np.ascontiguousarray(fwd, dtype=np.uint8)
    n = len(lens)
    n_aln = np.zeros(n, dtype=np.int32)
    rec_p = ctypes.c_void_p()
    L.orc_stats_reset()
    tot = L.orc_aln_batch(ctypes.byref(ob), ctypes.byref(orb), n, lens.ctypes.data, offs.ctypes.data,
                          fwd.ctypes.data, ctypes.byref(opt_c), n_aln.ctypes.data, ctypes.byref(rec_p))
    if tot:
        buf = (ctypes.c_uint8 * (16 * tot)).from_address(rec_p.value)
        rec = np.frombuffer(buf, dtype=ALN_DTYPE).copy()
    else:
        rec = np.empty(0, dtype=ALN_DTYPE)
    if rec_p.value:
        L.orc_free(rec_p)
    st = OrcStats()
    L.orc_stats_get(ctypes.byref(st))
    return n_aln, rec, {n: getattr(st, n) for n, _ in OrcStats._fields_}


def have_ref() -> bool:
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def run_ref(args, stdout_path=None, check=True, timeout=None):
    """Run the unmodified reference binary (oracle/_ref/ibwa)."""
    out = open(stdout_path, "wb") if stdout_path else subprocess.DEVNULL
    try:
        return subprocess.run([REF_BIN] + list(args), stdout=out, stderr=subprocess.DEVNULL, check=check,
                              timeout=timeout)
    finally:
        if stdout_path:
            out.close()
