"""The C-ABI library loads and exports every symbol include/b200aln.h declares
(no compute: runs without a GPU)."""
import ctypes
import os
import re

import pytest

from ibwa_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(engine.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return engine.load_library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200aln.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200aln_[a-z0-9_]+)\s*\(", text)))


def test_exports_match_header(lib):
    syms = declared_symbols()
    assert set(syms) == set(engine.EXPORTS)
    for s in syms:
        assert hasattr(lib, s), s


def test_struct_sizes_and_defaults(lib):
    assert ctypes.sizeof(engine.GapOptC) == 64
    o = engine.GapOptC()
    lib.b200aln_opt_init(ctypes.byref(o))
    from ibwa_b200 import gap_init_opt
    assert bytes(o) == gap_init_opt().header_bytes()
    lay = engine.SeqLayout()
    lib.b200aln_seq_layout(ctypes.byref(lay))
    assert lay.size == 176 and lay.off_seq == 8 and lay.off_n_aln == 48 and lay.off_aln == 56
    assert lib.b200aln_cal_maxdiff(100, 0.02, 0.04) == 5
    assert b"sm_100a" in lib.b200aln_version()


def test_bwa_seq_layout_against_reference_header(tmp_path):
    """Where the reference tree is present, check the mirrored bwa_seq_t offsets against its header."""
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "bwtaln.h")):
        pytest.skip("reference tree not present")
    src = tmp_path / "off.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "bwtaln.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(bwa_seq_t), offsetof(bwa_seq_t,name),'
                   'offsetof(bwa_seq_t,seq),offsetof(bwa_seq_t,rseq),offsetof(bwa_seq_t,qual),offsetof(bwa_seq_t,n_aln),'
                   'offsetof(bwa_seq_t,aln),offsetof(bwa_seq_t,sa));return 0;}\n')
    exe = tmp_path / "off"
    import subprocess
    subprocess.check_call(["gcc", "-I", ref, "-o", str(exe), str(src)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    lay = engine.SeqLayout()
    engine.load_library().b200aln_seq_layout(ctypes.byref(lay))
    assert got == [lay.size, lay.off_name, lay.off_seq, lay.off_rseq, lay.off_qual, lay.off_n_aln, lay.off_aln,
                   lay.off_sa]


def test_python_option_parser_follows_glibc_getopt_and_atof():
    """bwa_aln parses with glibc getopt (options may follow operands) and atof / atoi (numeric prefix, 0 when there is
    none; bwtaln.c:249-284): the Python mirror used by tests and bench.py must read the same command lines alike."""
    from ibwa_b200 import parse_aln_args
    o, prefix, reads, _ = parse_aln_args(["pfx", "-n", "0.01x", "reads.fq", "-o", "2junk", "-e", "abc"])
    assert (prefix, reads) == ("pfx", "reads.fq")
    assert abs(o.fnr - 0.01) < 1e-9 and o.max_diff == -1 and o.max_gapo == 2
    assert o.max_gape == 6 and o.mode & 1            # -e 0 (atoi of "abc") leaves the defaults
    o, _, _, _ = parse_aln_args(["-n", "3", "p", "r"])
    assert o.max_diff == 3 and o.fnr == -1.0
