"""gap_opt_t mirror and `aln` option parsing.

Reference: gap_opt_t bwtaln.h:105-115 (16 x 4 bytes, persisted verbatim as the
.sai header, bwtaln.c:192), defaults gap_init_opt bwtaln.c:21-37, option
parsing bwa_aln bwtaln.c:243-328, bwa_cal_maxdiff bwtaln.c:39-51.
"""
from __future__ import annotations

import ctypes
import getopt
import math
from dataclasses import dataclass

BWA_MODE_GAPE = 0x01
BWA_MODE_COMPREAD = 0x02
BWA_MODE_LOGGAP = 0x04
BWA_MODE_NONSTOP = 0x10
BWA_MODE_BAM = 0x20
BWA_MODE_BAM_SE = 0x40
BWA_MODE_BAM_READ1 = 0x80
BWA_MODE_BAM_READ2 = 0x100
BWA_MODE_IL13 = 0x200
BWA_AVG_ERR = 0.02
ALN_GETOPT = "n:o:e:i:d:l:k:cLR:m:t:NM:O:E:q:f:b012IB:"  # bwtaln.c:249


class GapOptC(ctypes.Structure):
    """C layout of gap_opt_t (bwtaln.h:105-115)."""
    _fields_ = [
        ("s_mm", ctypes.c_int32), ("s_gapo", ctypes.c_int32), ("s_gape", ctypes.c_int32),
        ("mode", ctypes.c_int32),
        ("indel_end_skip", ctypes.c_int32), ("max_del_occ", ctypes.c_int32), ("max_entries", ctypes.c_int32),
        ("fnr", ctypes.c_float),
        ("max_diff", ctypes.c_int32), ("max_gapo", ctypes.c_int32), ("max_gape", ctypes.c_int32),
        ("max_seed_diff", ctypes.c_int32), ("seed_len", ctypes.c_int32),
        ("n_threads", ctypes.c_int32),
        ("max_top2", ctypes.c_int32),
        ("trim_qual", ctypes.c_int32),
    ]


assert ctypes.sizeof(GapOptC) == 64


@dataclass
class GapOpt:
    s_mm: int = 3
    s_gapo: int = 11
    s_gape: int = 4
    mode: int = BWA_MODE_GAPE | BWA_MODE_COMPREAD
    indel_end_skip: int = 5
    max_del_occ: int = 10
    max_entries: int = 2000000
    fnr: float = 0.04
    max_diff: int = -1
    max_gapo: int = 1
    max_gape: int = 6
    max_seed_diff: int = 2
    seed_len: int = 32
    n_threads: int = 1
    max_top2: int = 30
    trim_qual: int = 0

    def to_c(self) -> GapOptC:
        c = GapOptC()
        for name, _ in GapOptC._fields_:
            setattr(c, name, getattr(self, name))
        return c

    def header_bytes(self) -> bytes:
        """The 64 bytes written at the head of a .sai (bwtaln.c:192)."""
        return bytes(self.to_c())

    @classmethod
    def from_header(cls, raw: bytes) -> "GapOpt":
        c = GapOptC.from_buffer_copy(raw[:64])
        return cls(**{name: getattr(c, name) for name, _ in GapOptC._fields_})


def gap_init_opt() -> GapOpt:
    """bwtaln.c:21-37."""
    return GapOpt()


def bwa_cal_maxdiff(l: int, err: float = BWA_AVG_ERR, thres: float = 0.04) -> int:
    """bwtaln.c:39-51 (32-bit wrapping factorial like the compiled reference)."""
    elambda = math.exp(-l * err)
    s = elambda
    y = 1.0
    x = 1
    for k in range(1, 1000):
        y *= l * err
        x = (x * k) & 0xFFFFFFFF
        xs = x - (1 << 32) if x & 0x80000000 else x
        s += (elambda * y / xs) if xs else math.inf
        if 1.0 - s < thres:
            return k
    return 2


class UsageError(Exception):
    pass


def parse_aln_args(argv):
    """Parse `aln` arguments exactly like bwa_aln (bwtaln.c:243-285).

    Returns (opt, prefix, reads_path, out_path_or_None).  Raises UsageError where
    the reference prints usage and returns 1.
    """
    opt = gap_init_opt()
    opte = -1
    out = None
    try:
        pairs, rest = getopt.gnu_getopt(list(argv), ALN_GETOPT)   # glibc getopt permutes: options may follow operands
    except getopt.GetoptError as e:  # reference: `default: return 1`
        raise UsageError(str(e))
    for flag, val in pairs:
        f = flag[1]
        if f == "n":
            if "." in val:
                opt.fnr, opt.max_diff = _atof(val), -1
            else:
                opt.max_diff, opt.fnr = _atoi(val), -1.0
        elif f == "o": opt.max_gapo = _atoi(val)
        elif f == "e": opte = _atoi(val)
        elif f == "M": opt.s_mm = _atoi(val)
        elif f == "O": opt.s_gapo = _atoi(val)
        elif f == "E": opt.s_gape = _atoi(val)
        elif f == "d": opt.max_del_occ = _atoi(val)
        elif f == "i": opt.indel_end_skip = _atoi(val)
        elif f == "l": opt.seed_len = _atoi(val)
        elif f == "k": opt.max_seed_diff = _atoi(val)
        elif f == "m": opt.max_entries = _atoi(val)
        elif f == "t": opt.n_threads = _atoi(val)
        elif f == "L": opt.mode |= BWA_MODE_LOGGAP
        elif f == "R": opt.max_top2 = _atoi(val)
        elif f == "q": opt.trim_qual = _atoi(val)
        elif f == "c": opt.mode &= ~BWA_MODE_COMPREAD
        elif f == "N":
            opt.mode |= BWA_MODE_NONSTOP
            opt.max_top2 = 0x7FFFFFFF
        elif f == "f": out = val
        elif f == "b": opt.mode |= BWA_MODE_BAM
        elif f == "0": opt.mode |= BWA_MODE_BAM_SE
        elif f == "1": opt.mode |= BWA_MODE_BAM_READ1
        elif f == "2": opt.mode |= BWA_MODE_BAM_READ2
        elif f == "I": opt.mode |= BWA_MODE_IL13
        elif f == "B": opt.mode |= _atoi(val) << 24
    if opte > 0:
        opt.max_gape = opte
        opt.mode &= ~BWA_MODE_GAPE
    if len(rest) < 2:
        raise UsageError("Usage:   bwa aln [options] <prefix> <in.fq>")
    return opt, rest[0], rest[1], out


def _atof(s: str) -> float:
    """C atof: the longest numeric prefix (sign, digits, fraction, exponent), 0.0 when there is none."""
    import re
    m = re.match(r"\s*[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)", s)
    return float(m.group(0)) if m else 0.0


def _atoi(s: str) -> int:
    """C atoi: leading whitespace, optional sign, digits; 0 when none."""
    s = s.lstrip()
    sign = 1
    i = 0
    if s[:1] in "+-":
        sign = -1 if s[0] == "-" else 1
        i = 1
    j = i
    while j < len(s) and s[j].isdigit():
        j += 1
    return sign * int(s[i:j]) if j > i else 0
