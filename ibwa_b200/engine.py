"""ctypes binding of libb200aln.so — the host-side mirror of the reference's
`aln` interface on top of the CUDA engine.

  bwa_cal_sa_reg_gap (bwtaln.c:80-140)  -> Engine.cal_sa_reg_gap(batch, opt)
  bwa_aln_core       (bwtaln.c:173-241) -> bwa_aln_core(prefix, reads, opt, out)
  bwa_aln            (bwtaln.c:243-328) -> bwa_aln(argv)

There is no CPU fallback: if the shared library is missing, or no CUDA device
is present when a context is opened, this module raises / the library aborts.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .bwtio import Bwt, bwt_restore_bwt
from .opts import GapOpt, GapOptC, UsageError, parse_aln_args
from .sai import ALN_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200ALN_LIB", os.path.join(_HERE, "libb200aln.so"))   # override: kernel A/B experiments


class BwtView(ctypes.Structure):
    _fields_ = [("primary", ctypes.c_uint32), ("L2", ctypes.c_uint32 * 5), ("seq_len", ctypes.c_uint32),
                ("bwt_size", ctypes.c_uint64), ("bwt", ctypes.c_void_p)]


class Stats(ctypes.Structure):
    _fields_ = [("ms_h2d", ctypes.c_double), ("ms_width", ctypes.c_double), ("ms_search", ctypes.c_double),
                ("ms_compact", ctypes.c_double), ("ms_d2h", ctypes.c_double), ("ms_total", ctypes.c_double),
                ("kernel_launches", ctypes.c_uint64), ("overflow_reads", ctypes.c_uint64),
                ("pops", ctypes.c_uint64), ("occ_lookups", ctypes.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SaView(ctypes.Structure):
    _fields_ = [("primary", ctypes.c_uint32), ("seq_len", ctypes.c_uint32), ("sa_intv", ctypes.c_int32),
                ("n_sa", ctypes.c_uint64), ("sa", ctypes.c_void_p)]


class SeqLayout(ctypes.Structure):
    _fields_ = [(n, ctypes.c_size_t) for n in ("size", "off_name", "off_seq", "off_rseq", "off_qual", "off_lenword",
                                                 "off_n_aln", "off_aln", "off_sa", "off_c1c2")]


EXPORTS = ["b200aln_version", "b200aln_opt_init", "b200aln_cal_maxdiff", "b200aln_device_count", "b200aln_open",
           "b200aln_open_prefix", "b200aln_clone", "b200aln_close", "b200aln_batch", "b200aln_batch_sai", "b200aln_pin", "b200aln_unpin", "b200aln_prealloc", "b200aln_prealloc_release", "b200aln_batch_device", "b200aln_last_stats",
           "b200aln_set_int", "b200aln_timer_start", "b200aln_timer_stop", "b200aln_cal_sa_reg_gap", "b200aln_seq_layout", "b200aln_aln_core", "b200aln_aln_main", "b200aln_reader_open", "b200aln_reader_next", "b200aln_reader_close",
           "b200aln_sector_roofline", "b200aln_sa_load", "b200aln_bwt_sa", "b200aln_sa2seq", "b200aln_alngrp_merge", "b200aln_warm_device"]

_lib = None


def load_library():
    """Loads libb200aln.so (built in-tree by `make -C ibwa_b200` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C ibwa_b200` "
                           "(the engine has no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    L.b200aln_version.restype = ctypes.c_char_p
    L.b200aln_cal_maxdiff.restype = ctypes.c_int
    L.b200aln_cal_maxdiff.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double]
    L.b200aln_device_count.restype = ctypes.c_int
    L.b200aln_open.restype = ctypes.c_void_p
    L.b200aln_open.argtypes = [ctypes.POINTER(BwtView), ctypes.POINTER(BwtView), ctypes.c_int]
    L.b200aln_open_prefix.restype = ctypes.c_void_p
    L.b200aln_open_prefix.argtypes = [ctypes.c_char_p, ctypes.c_int]
    L.b200aln_clone.restype = ctypes.c_void_p
    L.b200aln_clone.argtypes = [ctypes.c_void_p]
    L.b200aln_close.argtypes = [ctypes.c_void_p]
    L.b200aln_batch.restype = ctypes.c_void_p
    L.b200aln_batch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                ctypes.POINTER(GapOptC), ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]
    L.b200aln_batch_sai.restype = ctypes.c_void_p
    L.b200aln_batch_sai.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.POINTER(GapOptC), ctypes.POINTER(ctypes.c_int64)]
    L.b200aln_pin.restype = ctypes.c_int
    L.b200aln_pin.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    L.b200aln_unpin.argtypes = [ctypes.c_void_p]
    L.b200aln_prealloc.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.b200aln_prealloc_release.argtypes = [ctypes.c_int]
    L.b200aln_batch_device.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.POINTER(GapOptC), ctypes.POINTER(ctypes.c_void_p),
                                       ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64)]
    L.b200aln_last_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(Stats)]
    L.b200aln_timer_start.argtypes = [ctypes.c_void_p]
    L.b200aln_timer_stop.restype = ctypes.c_double
    L.b200aln_timer_stop.argtypes = [ctypes.c_void_p]
    L.b200aln_set_int.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]
    L.b200aln_aln_core.restype = ctypes.c_int64
    L.b200aln_aln_core.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(GapOptC), ctypes.c_int,
                                   ctypes.c_int]
    L.b200aln_sector_roofline.restype = ctypes.c_double
    L.b200aln_sector_roofline.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int]
    L.b200aln_seq_layout.argtypes = [ctypes.POINTER(SeqLayout)]
    L.b200aln_sa_load.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(SaView)]
    L.b200aln_bwt_sa.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    L.b200aln_sa2seq.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                 ctypes.c_void_p]
    L.b200aln_alngrp_merge.restype = ctypes.c_int64
    L.b200aln_alngrp_merge.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.b200aln_reader_open.restype = ctypes.c_void_p
    L.b200aln_reader_open.argtypes = [ctypes.c_char_p, ctypes.c_int]
    L.b200aln_reader_next.restype = ctypes.c_int
    L.b200aln_reader_next.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                                      ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64)]
    L.b200aln_reader_close.argtypes = [ctypes.c_void_p]
    _lib = L
    return L


def _view(b: Bwt) -> BwtView:
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


class Engine:
    """A device-resident pair of FM-indexes plus search scratch on one GPU
    (what bwa_aln_core keeps in bwt[2], bwtaln.c:184-189)."""

    def __init__(self, bwt: Bwt, rbwt: Bwt, device: int = 0):
        self._L = load_library()
        v0, v1 = _view(bwt), _view(rbwt)
        self._ctx = self._L.b200aln_open(ctypes.byref(v0), ctypes.byref(v1), device)
        self.device = device
        self.seq_len = bwt.seq_len

    def clone(self) -> "Engine":
        """A sibling on the same GPU sharing the device index (own stream and buffers): lets a second
        batch be in flight from another host thread.  Close it before this engine."""
        other = object.__new__(Engine)
        other._L = self._L
        other._ctx = self._L.b200aln_clone(self._ctx)
        other.device = self.device
        other.seq_len = self.seq_len
        return other

    @classmethod
    def from_prefix(cls, prefix: str, device: int = 0) -> "Engine":
        return cls(bwt_restore_bwt(prefix + ".bwt"), bwt_restore_bwt(prefix + ".rbwt"), device)

    def set(self, key: str, value: int) -> None:
        self._L.b200aln_set_int(self._ctx, key.encode(), int(value))

    def close(self) -> None:
        if self._ctx:
            self._L.b200aln_close(self._ctx)
            self._ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def cal_sa_reg_gap(self, lens, offs, codes, opt: GapOpt):
        """One reference batch on host buffers -> (n_aln int32[n], records ALN_DTYPE[total])."""
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        n = len(lens)
        n_aln = np.zeros(n, dtype=np.int32)
        total = ctypes.c_int64()
        oc = opt.to_c()
        p = self._L.b200aln_batch(self._ctx, n, lens.ctypes.data, offs.ctypes.data, codes.ctypes.data,
                                  ctypes.byref(oc), n_aln.ctypes.data, ctypes.byref(total))
        t = total.value
        if t:
            rec = np.frombuffer((ctypes.c_uint8 * (16 * t)).from_address(p), dtype=ALN_DTYPE).copy()
        else:
            rec = np.empty(0, dtype=ALN_DTYPE)
        return n_aln, rec

    def cal_sa_reg_gap_sai(self, lens, offs, codes, opt: GapOpt, pin: bool = False) -> bytes:
        """One reference batch -> the bytes bwa_aln_core writes for it (bwtaln.c:227-231), formatted on the device
        (b200aln_batch_sai).  pin: page-lock the input arrays for the call (b200aln_pin / b200aln_unpin)."""
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        offs = np.ascontiguousarray(offs, dtype=np.int64)
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        nb = ctypes.c_int64()
        oc = opt.to_c()
        pinned = []
        if pin:
            for a in (lens, offs, codes):
                if a.nbytes and self._L.b200aln_pin(a.ctypes.data, a.nbytes) == 0:
                    pinned.append(a)
        try:
            p = self._L.b200aln_batch_sai(self._ctx, len(lens), lens.ctypes.data, offs.ctypes.data, codes.ctypes.data,
                                          ctypes.byref(oc), ctypes.byref(nb))
            return bytes((ctypes.c_uint8 * nb.value).from_address(p)) if nb.value else b""
        finally:
            for a in pinned:
                self._L.b200aln_unpin(a.ctypes.data)

    def batch_pinned(self, lens_ptr: int, offs_ptr: int, codes_ptr: int, n: int, opt: GapOpt, n_aln_ptr: int):
        """Same call on raw host pointers (pinned buffers owned by the caller); returns (records ptr, total)."""
        total = ctypes.c_int64()
        oc = opt.to_c()
        p = self._L.b200aln_batch(self._ctx, n, lens_ptr, offs_ptr, codes_ptr, ctypes.byref(oc), n_aln_ptr,
                                  ctypes.byref(total))
        return p, total.value

    def batch_device(self, d_lens: int, d_offs: int, d_codes: int, n: int, max_len: int, opt: GapOpt):
        """Device-resident inputs (raw device pointers) -> (d_n_aln ptr, d_recs ptr, total)."""
        oc = opt.to_c()
        pn, pr, total = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
        self._L.b200aln_batch_device(self._ctx, n, max_len, d_lens, d_offs, d_codes, ctypes.byref(oc),
                                     ctypes.byref(pn), ctypes.byref(pr), ctypes.byref(total))
        return pn.value, pr.value, total.value

    def load_sa(self, which: int, sa) -> None:
        """bwt_restore_sa onto the device; sa: ibwa_b200.bwtio.Sa, which = 0 (.sa) or 1 (.rsa)."""
        v = SaView()
        v.primary, v.seq_len, v.sa_intv = sa.primary, sa.seq_len, sa.sa_intv
        arr = np.ascontiguousarray(sa.sa, dtype=np.uint32)
        v.n_sa = arr.shape[0]
        v.sa = arr.ctypes.data
        self._L.b200aln_sa_load(self._ctx, which, ctypes.byref(v))

    def bwt_sa(self, which: int, rows) -> np.ndarray:
        """bwt_sa (bwt.c:69-79) for a batch of SA rows."""
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        out = np.empty(len(rows), dtype=np.uint32)
        self._L.b200aln_bwt_sa(self._ctx, which, len(rows), rows.ctypes.data, out.ctypes.data)
        return out

    def sa2seq(self, strand, rows, lens) -> np.ndarray:
        """bwtdb_sa2seq with offset 0 (dbset.c:240-245)."""
        strand = np.ascontiguousarray(strand, dtype=np.uint8)
        rows = np.ascontiguousarray(rows, dtype=np.uint32)
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        out = np.empty(len(rows), dtype=np.uint64)
        self._L.b200aln_sa2seq(self._ctx, len(rows), strand.ctypes.data, rows.ctypes.data, lens.ctypes.data,
                               out.ctypes.data)
        return out

    def alngrp_merge(self, n_alns, recs, s_mm: int):
        """alngrp_create (saiset.c:45-78) for a batch: n_alns[s] / recs[s] = stream s as cal_sa_reg_gap returns it.
        Returns (out_off int64[n], out_n int32[n], records ALN_DTYPE[total], dbidx uint32[total])."""
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
        got = self._L.b200aln_alngrp_merge(self._ctx, ns, n, pn, pr, int(s_mm), out_off.ctypes.data, out_n.ctypes.data,
                                           out_rec.ctypes.data, out_db.ctypes.data)
        assert got == total
        return out_off, out_n, out_rec[:total], out_db[:total]

    def timer_start(self) -> None:
        self._L.b200aln_timer_start(self._ctx)

    def timer_stop(self) -> float:
        """Milliseconds between timer_start and now, measured with CUDA events on the engine's stream."""
        return self._L.b200aln_timer_stop(self._ctx)

    def stats(self) -> dict:
        s = Stats()
        self._L.b200aln_last_stats(self._ctx, ctypes.byref(s))
        return s.as_dict()

    def sector_roofline(self, n_loads: int = 1 << 28, repeats: int = 3) -> float:
        return self._L.b200aln_sector_roofline(self._ctx, n_loads, repeats)


def read_batches_native(path: str, mode: int, trim_qual: int, n_needed: int = 0x40000):
    """bwa_read_seq through the native reader (b200aln_reader_*): yields (lens, offs, codes) numpy copies."""
    L = load_library()
    r = L.b200aln_reader_open(path.encode(), mode)
    try:
        while True:
            pl, po, pc, nb = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int64()
            n = L.b200aln_reader_next(r, n_needed, mode, trim_qual, ctypes.byref(pl), ctypes.byref(po),
                                      ctypes.byref(pc), ctypes.byref(nb))
            if n == 0:
                break
            lens = np.frombuffer((ctypes.c_int32 * n).from_address(pl.value), dtype=np.int32).copy()
            offs = np.frombuffer((ctypes.c_int64 * n).from_address(po.value), dtype=np.int64).copy()
            codes = np.frombuffer((ctypes.c_uint8 * nb.value).from_address(pc.value), dtype=np.uint8).copy() \
                if nb.value else np.empty(0, np.uint8)
            yield lens, offs, codes
    finally:
        L.b200aln_reader_close(r)


def bwa_aln_core(prefix: str, fn_fa: str, opt: GapOpt, out_path: str, device: int = 0) -> int:
    """bwa_aln_core (bwtaln.c:173-241) through the native driver: FASTA/FASTQ(.gz) in, .sai out."""
    L = load_library()
    fd = os.open(out_path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
    try:
        oc = opt.to_c()
        return L.b200aln_aln_core(prefix.encode(), fn_fa.encode(), ctypes.byref(oc), fd, device)
    finally:
        os.close(fd)


def bwa_aln(argv) -> int:
    """bwa_aln (bwtaln.c:243-328): `aln [options] <prefix> <in.fq>`; returns the exit status."""
    try:
        opt, prefix, reads, out = parse_aln_args(argv)
    except UsageError:
        return 1
    bwa_aln_core(prefix, reads, opt, out if out else "/dev/stdout")
    return 0
