"""b200-aln: B200-native engine for the `bwa aln` hot path of genome/ibwa.

Host-side mirror of the reference interface for that path:

  gap_opt_t / gap_init_opt / bwa_aln option parsing ... ibwa_b200.opts
  .bwt / .rbwt reader (bwt_restore_bwt) ................ ibwa_b200.bwtio
  .sai wire format ..................................... ibwa_b200.sai
  bwa_cal_sa_reg_gap / bwa_aln_core over the CUDA lib .. ibwa_b200.engine

The compute lives in ibwa_b200/csrc (CUDA, sm_100a) behind the C ABI declared
in include/b200aln.h.  There is no CPU fallback: importing ibwa_b200.engine and
opening a context fails loudly when libb200aln.so is missing.
"""
from .opts import GapOpt, gap_init_opt, parse_aln_args  # noqa: F401
from .bwtio import Bwt, Sa, bwt_restore_bwt, bwt_dump_bwt, bwt_restore_sa, bwt_dump_sa   # noqa: F401
from . import sai                                        # noqa: F401

__all__ = ["GapOpt", "gap_init_opt", "parse_aln_args", "Bwt", "bwt_restore_bwt", "bwt_dump_bwt", "sai"]
