/*
 * alngrp_core.cuh — scope row N4: the per-read merge of several .sai streams.
 *
 * Reference behaviour being re-implemented (file:line under the reference):
 *   alngrp_create ............ saiset.c:45-78   concatenate the reads' alignments stream by stream; with more
 *                                               than one stream, sort by score and cut at best + s_mm
 *   ks_introsort(alignment) .. ksort.h:172-224  the sort it uses (comparator a.score < b.score, saiset.c:8-10)
 *   ks_combsort .............. ksort.h:150-171  its fallback when the depth budget runs out
 *   __ks_insertsort .......... ksort.h:142-149  the final pass
 * The reference's sort is NOT stable, and samse/sampe pick hits by position in the sorted group
 * (select_sai, saiset.c:85-112), so the order of equal scores has to come out the same: this file
 * performs the same sequence of comparisons and exchanges, on indexes instead of pointers.
 * Compiled by nvcc (k_alngrp in b200aln.cu) and by g++ (tests/harness).
 */
#pragma once
#include <stdint.h>
#include "aln_core.cuh"

namespace b2 {

/* the group being sorted: records and the stream (database) each came from, moved together */
struct GroupView {
    Rec *rec;
    uint32_t *db;
    B2_HD bool lt(int64_t i, int64_t j) const { return rec[i].score < rec[j].score; }
    B2_HD bool lt_key(int64_t i, int32_t key) const { return rec[i].score < key; }
    B2_HD bool key_lt(int32_t key, int64_t j) const { return key < rec[j].score; }
    B2_HD void swap(int64_t i, int64_t j)
    {
        const Rec r = rec[i]; rec[i] = rec[j]; rec[j] = r;
        const uint32_t d = db[i]; db[i] = db[j]; db[j] = d;
    }
};

/* ksort.h:142-149 on [s, t) */
B2_HD void grp_insertsort(GroupView &g, int64_t s, int64_t t)
{
    for (int64_t i = s + 1; i < t; ++i)
        for (int64_t j = i; j > s && g.lt(j, j - 1); --j) g.swap(j, j - 1);
}

/* ksort.h:150-171 on n elements starting at a */
B2_HD void grp_combsort(GroupView &g, int64_t a, int64_t n)
{
    const double shrink = 1.2473309501039786540366528676643;
    bool swapped;
    uint64_t gap = (uint64_t)n;
    do {
        if (gap > 2) {
            gap = (uint64_t)((double)gap / shrink);
            if (gap == 9 || gap == 10) gap = 11;
        }
        swapped = false;
        for (int64_t i = a; i < a + n - (int64_t)gap; ++i) {
            const int64_t j = i + (int64_t)gap;
            if (g.lt(j, i)) { g.swap(i, j); swapped = true; }
        }
    } while (swapped || gap > 2);
    if (gap != 1) grp_insertsort(g, a, a + n);
}

/* ksort.h:172-224: median-of-three quicksort that leaves runs of <= 16 unsorted, with a depth budget
 * of 2 * ceil(log2 n) (comb sort beyond it), then one insertion sort over everything */
B2_HD void grp_introsort(GroupView &g, int64_t n)
{
    if (n < 1) return;
    if (n == 2) {
        if (g.lt(1, 0)) g.swap(0, 1);
        return;
    }
    int d;
    for (d = 2; ((uint64_t)1 << d) < (uint64_t)n; ++d) {}
    /* the larger side of every split is deferred and the smaller one continued: at most log2 n deferred ranges */
    int64_t st_l[48], st_r[48];
    int st_d[48], top = 0;
    int64_t s = 0, t = n - 1;
    d <<= 1;
    for (;;) {
        if (s < t) {
            if (--d == 0) {
                grp_combsort(g, s, t - s + 1);
                t = s;
                continue;
            }
            int64_t i = s, j = t, k = i + ((j - i) >> 1) + 1;
            if (g.lt(k, i)) {
                if (g.lt(k, j)) k = j;
            } else k = g.lt(j, i) ? i : j;
            const int32_t pivot = g.rec[k].score; /* only the key of the pivot is ever compared */
            if (k != t) g.swap(k, t);
            for (;;) {
                do ++i; while (g.lt_key(i, pivot));
                do --j; while (i <= j && g.key_lt(pivot, j));
                if (j <= i) break;
                g.swap(i, j);
            }
            g.swap(i, t);
            if (i - s > t - i) {
                if (i - s > 16) { st_l[top] = s; st_r[top] = i - 1; st_d[top] = d; ++top; }
                s = t - i > 16 ? i + 1 : t;
            } else {
                if (t - i > 16) { st_l[top] = i + 1; st_r[top] = t; st_d[top] = d; ++top; }
                t = i - s > 16 ? i - 1 : s;
            }
        } else {
            if (top == 0) {
                grp_insertsort(g, 0, n);
                return;
            }
            --top;
            s = st_l[top]; t = st_r[top]; d = st_d[top];
        }
    }
}

/*
 * One read: gather its alignments from every stream (in stream order) into out_rec / out_db at `off`,
 * sort and cut when there is more than one stream (saiset.c:64-76).  Returns the size of the group.
 *   n_aln[s][r], rec_off[s][r] (exclusive prefix sums of n_aln[s]), recs[s]: the streams
 */
B2_HD int32_t alngrp_merge_one(int n_streams, int64_t r, const int32_t *const *n_aln, const int64_t *const *rec_off,
                               const Rec *const *recs, int s_mm, int64_t off, Rec *out_rec, uint32_t *out_db)
{
    int64_t n = 0;
    for (int s = 0; s < n_streams; ++s) {
        const int32_t c = n_aln[s][r];
        const Rec *src = recs[s] + rec_off[s][r];
        for (int32_t j = 0; j < c; ++j) {
            out_rec[off + n] = src[j];
            out_db[off + n] = (uint32_t)s;
            ++n;
        }
    }
    if (n_streams > 1 && n > 0) {
        GroupView g;
        g.rec = out_rec + off;
        g.db = out_db + off;
        grp_introsort(g, n);
        const int32_t best = g.rec[0].score;
        for (int64_t i = 0; i < n; ++i)
            if (g.rec[i].score > best + s_mm) { n = i; break; }
    }
    return (int32_t)n;
}

} // namespace b2
