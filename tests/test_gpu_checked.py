"""The bounds-checked build (ibwa_b200/libb200aln_checked.so = the product sources with -DB2_CHECKED, see
aln_core.cuh) over the golden set, the multi-pass flows and the kernel variants: every arena slot, bucket number,
width-record position, record fill, interval-table and occ-block address the kernels derive from memory is tested
on the device; a violation aborts the run with the offending read.  Stands in for compute-sanitizer, which the GPU
pool does not allow."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKED = os.path.join(ROOT, "ibwa_b200", "libb200aln_checked.so")


def test_checked_build_runs_clean_and_exact():
    assert os.path.exists(CHECKED), "build it with `make -C ibwa_b200` (__graft_entry__.build())"
    env = dict(os.environ, B200ALN_LIB=CHECKED)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "run_checked.py"), "--expect-checked"],
                       env=env, capture_output=True, text=True, timeout=900)
    sys.stdout.write(p.stdout)
    sys.stderr.write(p.stderr[-3000:])
    assert p.returncode == 0, "bounds-checked run failed or differed from the reference"
    assert "BAD" not in p.stdout and "B2_CHECKED" in p.stdout
