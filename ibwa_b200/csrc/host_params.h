/*
 * host_params.h — host-side derivation of launch parameters from gap_opt_t,
 * i.e. what bwa_cal_sa_reg_gap does before its per-read loop (bwtaln.c:86-93)
 * and per read (bwtaln.c:125-126), plus bwa_cal_maxdiff (bwtaln.c:39-51).
 */
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../include/b200aln.h"
#include "aln_core.cuh"

namespace b2host {

/* bwtaln.c:39-51.  `x` is a 32-bit int in the reference; unsigned wrap-around
 * reproduces the compiled binary beyond 12!. */
inline int cal_maxdiff(int l, double err, double thres)
{
    double elambda = exp(-l * err), sum = elambda, y = 1.0;
    uint32_t x = 1;
    for (int k = 1; k < 1000; ++k) {
        y *= l * err;
        x *= (uint32_t)k;
        sum += elambda * y / (int32_t)x;
        if (1.0 - sum < thres) return k;
    }
    return 2;
}

[[noreturn]] inline void fatal(const char *func, const char *msg)
{ /* reference convention: utils.c:67-82 */
    fprintf(stderr, "[%s] %s Abort!\n", func, msg);
    abort();
}

/* Fills P (batch-level) and md[len] = per-read max_diff for len in [0, max_len]. */
inline void make_params(const b200aln_opt_t &o, int max_len, const int32_t *lens, int n_reads, b2::Params &P,
                        std::vector<int> &md)
{
    if (o.s_mm <= 0 || o.s_gapo <= 0 || o.s_gape <= 0)
        fatal("b200aln_batch", "non-positive penalties (-M/-O/-E <= 0) are not supported: the reference's result "
                               "then depends on stale stack slots (bwtgap.c:60).");
    if (max_len > 65535) fatal("b200aln_batch", "reads longer than 65535 bp are not supported (bwtgap.c:142).");
    int batch_max_diff = o.max_diff;
    if (o.fnr > 0.0f) batch_max_diff = cal_maxdiff(max_len, 0.02, o.fnr);
    int max_gapo = o.max_gapo;
    if (batch_max_diff < max_gapo) max_gapo = batch_max_diff; /* bwtaln.c:91-92 */
    P.s_mm = o.s_mm; P.s_gapo = o.s_gapo; P.s_gape = o.s_gape;
    P.mode = o.mode;
    P.indel_end_skip = o.indel_end_skip; P.max_del_occ = o.max_del_occ; P.max_entries = o.max_entries;
    P.max_gapo = max_gapo; P.max_gape = o.max_gape; P.max_seed_diff = o.max_seed_diff; P.seed_len = o.seed_len;
    P.max_top2 = o.max_top2;
    P.n_buckets = (batch_max_diff + 1) * o.s_mm + (max_gapo + 1) * o.s_gapo + (o.max_gape + 1) * o.s_gape;
    if (P.n_buckets < 1 || P.n_buckets > 2048)
        fatal("b200aln_batch", "score range outside [1,2048] (bwtgap.c:54 packs the score in 11 bits).");
    if (o.seed_len < 0) fatal("b200aln_batch", "negative seed length.");
    if (batch_max_diff >= 255 || o.max_diff >= 255 || o.max_seed_diff >= 31)
        fatal("b200aln_batch", "max_diff >= 255 or max_seed_diff >= 31 exceed the packed width-record fields.");
    md.assign((size_t)max_len + 1, o.max_diff);
    if (o.fnr > 0.0f) {
        if (max_len <= 1024) {
            for (int l = 0; l <= max_len; ++l) md[l] = cal_maxdiff(l, 0.02, o.fnr);
        } else {
            std::vector<char> seen((size_t)max_len + 1, 0);
            for (int r = 0; r < n_reads; ++r)
                if (!seen[lens[r]]) { seen[lens[r]] = 1; md[lens[r]] = cal_maxdiff(lens[r], 0.02, o.fnr); }
        }
    }
}

} // namespace b2host
