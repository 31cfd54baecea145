"""The batch-operator seam on the reference's OWN structures.

oracle/_ref/ibwa_seam is the unmodified reference `ibwa` with exactly one function replaced at link time:
bwa_cal_sa_reg_gap (bwtaln.h:148) -> oracle/ref_shim/seam_cal_sa_reg_gap.c -> b200aln_cal_sa_reg_gap in
libb200aln.so.  So the reference's bwa_aln option parser, bwa_read_seq (bwaseqio.c:145-208), the fwrite loop
(bwtaln.c:227-231) and bwa_free_read_seq (bwaseqio.c:210-222) all run on bwa_seq_t arrays the engine filled."""
import os
import subprocess

import pytest

from oracle import pyoracle
from cases import CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEAM = os.path.join(ROOT, "oracle", "_ref", "ibwa_seam")


def _prefix(tmp_path, golden_dir):
    prefix = str(tmp_path / "g1")
    os.symlink(os.path.join(golden_dir, "g1.bwt"), prefix + ".bwt")
    os.symlink(os.path.join(golden_dir, "g1.rbwt"), prefix + ".rbwt")
    return prefix


def test_seam_binary_has_no_cpu_path(tmp_path, golden_dir):
    """Without a CUDA device the seam must die in b200aln_open — the replaced operator has no CPU fallback."""
    if not os.path.exists(SEAM):
        pytest.skip("oracle/_ref/ibwa_seam not built (needs /root/reference)")
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([SEAM, "aln", _prefix(tmp_path, golden_dir), os.path.join(golden_dir, "g1_short.fq.gz")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode != 0
    assert b"no CUDA device available" in p.stderr
    assert len(p.stdout) <= 64          # at most the header the reference wrote before the first batch


@pytest.mark.gpu
@pytest.mark.parametrize("tag,extra", [("default", []), ("q20", []), ("c", []), ("short_o3", []), ("stress", []),
                                       ("default", ["-t", "4"])])
def test_reference_driver_with_engine_operator(tag, extra, tmp_path, golden_dir):
    if not (os.path.exists(SEAM) and pyoracle.have_ref()):
        pytest.skip("oracle/_ref not present")
    args, fq = CASES[tag]
    fq = os.path.join(golden_dir, fq + ".fq.gz")
    prefix = _prefix(tmp_path, golden_dir)
    ref_out = str(tmp_path / "ref.sai")
    pyoracle.run_ref(["aln"] + args + extra + [prefix, fq], stdout_path=ref_out)
    p = subprocess.run([SEAM, "aln"] + args + extra + [prefix, fq], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       check=True)
    want = open(ref_out, "rb").read()
    assert p.stdout == want                          # header included: same -t on both sides
    assert want == open(os.path.join(golden_dir, f"g1_{tag}.sai"), "rb").read() or extra   # and both equal the fixture
    assert b"sequences have been processed" in p.stderr                                      # the reference's own loop ran
