/*
 * alnoracle.c — CPU restatement of the `bwa aln` hot path of genome/ibwa (BWA 0.5.9).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (ibwa_b200/, include/, the
 * CUDA library, the CLI) links, imports or executes this file.  It is used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the
 * checker.  Parity is PINNED: tests/test_oracle.py checks this restatement
 * byte-for-byte against .sai files produced by the unmodified reference binary
 * (oracle/_ref/ibwa, built by oracle/Makefile.ref) — committed as fixtures under
 * tests/golden/ by tests/golden/make_golden.py — and, where oracle/_ref exists,
 * against fresh runs of that binary.
 *
 * What is restated (all citations are into /root/reference):
 *   occ / occ4 on the on-disk BWT layout ........ bwt.c:81-214, bwt.h:56-63
 *   exact backward extension .................... bwt.c:235-250
 *   lower-bound array D(i) ("width") ............ bwtaln.c:54-78
 *   max_diff from read length ................... bwtaln.c:39-51
 *   bounded best-first gapped search ............ bwtgap.c:104-264
 *   bucketed LIFO with slot reuse ............... bwtgap.c:13-79 (incl. the
 *       `last_diff_pos` written only on diff pushes, bwtgap.c:60)
 *   per-batch driver ............................ bwtaln.c:80-140
 *   read orientation (reverse / rev-comp) ....... bwaseqio.c:55-72,189-192
 *   quality trimming ............................ bwaseqio.c:74-87
 *   SA row -> text position (row N2 of the scope) bwt.c:69-79, bwt.h:58-70, dbset.c:240-245
 *
 * Written from the behaviour described there, with its own data structures; it
 * additionally counts pops / occ lookups, which bench.py uses as the roofline
 * denominator.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_BLOCK 128u      /* bases per occ checkpoint (bwt.h:35 OCC_INTERVAL) */
#define ORC_BLOCK_WORDS 12u /* 4 count words + 8 words of 16 bases (bwtmisc.c:122-144) */

#define MODE_GAPE 0x01
#define MODE_COMPREAD 0x02
#define MODE_LOGGAP 0x04
#define MODE_NONSTOP 0x10

enum { ST_M = 0, ST_I = 1, ST_D = 2 };

typedef struct {
    uint32_t primary;
    uint32_t L2[5];
    uint32_t seq_len;
    uint64_t n_words;
    const uint32_t *bwt; /* reference .bwt payload, header stripped */
} orc_bwt_t;

typedef struct {
    uint32_t w;
    int32_t bid;
} orc_width_t;

typedef struct { /* == bwt_aln1_t, bwtaln.h:34-38 */
    uint32_t packed; /* n_mm | n_gapo<<8 | n_gape<<16 | a<<24 */
    uint32_t k, l;
    int32_t score;
} orc_aln_t;

typedef struct { /* == gap_opt_t, bwtaln.h:105-115 (64 bytes) */
    int32_t s_mm, s_gapo, s_gape;
    int32_t mode;
    int32_t indel_end_skip, max_del_occ, max_entries;
    float fnr;
    int32_t max_diff, max_gapo, max_gape;
    int32_t max_seed_diff, seed_len;
    int32_t n_threads;
    int32_t max_top2;
    int32_t trim_qual;
} orc_opt_t;

typedef struct {
    uint64_t pops, pushes, lookups, occ1_calls, occ4_calls, hits, stack_hwm, cutoff_reads;
    uint64_t inherit_violations; /* times the literal slot value differed from "inherit from parent" */
} orc_stats_t;

static orc_stats_t g_stats;

void orc_stats_reset(void) { memset(&g_stats, 0, sizeof g_stats); }
void orc_stats_get(orc_stats_t *out) { *out = g_stats; }

/* ---------------------------------------------------------------- occ ---- */

/* number of symbols equal to c among the first `n` (0..16) bases of a word,
 * bases stored MSB-first (bwt.h:58-60) */
static inline uint32_t count_in_word(uint32_t word, unsigned n, int c)
{
    uint32_t x, m;
    if (n == 0) return 0;
    x = word ^ (0x55555555u * (uint32_t)(3 - c)); /* symbol==c  ->  both bits set */
    m = x & (x >> 1) & 0x55555555u;
    if (n < 16) m &= ~((1u << (2 * (16 - n))) - 1u);
    return (uint32_t)__builtin_popcount(m);
}

/* occ(k, c): occurrences of c in BWT rows 0..k inclusive; k == (uint32_t)-1 -> 0.
 * Row numbering includes the sentinel row, which the stored string omits
 * (bwt.c:95-97). */
uint32_t orc_occ(const orc_bwt_t *b, uint32_t k, int c)
{
    const uint32_t *blk;
    uint32_t n, inblk, j;
    if (k == 0xffffffffu) return 0;
    if (k == b->seq_len) return b->L2[c + 1] - b->L2[c];
    if (k >= b->primary) --k;
    g_stats.lookups++;
    blk = b->bwt + (uint64_t)(k / ORC_BLOCK) * ORC_BLOCK_WORDS;
    n = blk[c];
    inblk = k % ORC_BLOCK + 1; /* bases of this block to count */
    for (j = 0; inblk > 0; ++j) {
        unsigned take = inblk > 16 ? 16 : inblk;
        n += count_in_word(blk[4 + j], take, c);
        inblk -= take;
    }
    return n;
}

void orc_occ4(const orc_bwt_t *b, uint32_t k, uint32_t cnt[4])
{
    const uint32_t *blk;
    uint32_t inblk, j;
    int c;
    if (k == 0xffffffffu) {
        cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0;
        return;
    }
    /* NB: unlike the single-symbol form, the reference's 4-symbol form has no
     * k == seq_len shortcut (bwt.c:157-174); the generic path gives the same
     * value because seq_len >= primary always holds. */
    if (k >= b->primary) --k;
    g_stats.lookups++;
    blk = b->bwt + (uint64_t)(k / ORC_BLOCK) * ORC_BLOCK_WORDS;
    for (c = 0; c < 4; ++c) cnt[c] = blk[c];
    inblk = k % ORC_BLOCK + 1;
    for (j = 0; inblk > 0; ++j) {
        unsigned take = inblk > 16 ? 16 : inblk;
        for (c = 0; c < 4; ++c) cnt[c] += count_in_word(blk[4 + j], take, c);
        inblk -= take;
    }
}

/* one backward-search step on symbol c; returns 0 if the interval empties */
static inline int step_exact(const orc_bwt_t *b, int c, uint32_t *k, uint32_t *l)
{
    uint32_t ok, ol;
    g_stats.occ1_calls++;
    ok = orc_occ(b, *k - 1, c);
    ol = (*k - 1 == *l) ? ok : orc_occ(b, *l, c);
    *k = b->L2[c] + ok + 1;
    *l = b->L2[c] + ol;
    return *k <= *l;
}

/* bwt.c:235-250 */
static int extend_exact(const orc_bwt_t *b, int len, const uint8_t *str, uint32_t *k0, uint32_t *l0)
{
    uint32_t k = *k0, l = *l0;
    int i;
    for (i = len - 1; i >= 0; --i) {
        if (str[i] > 3) return 0;
        if (!step_exact(b, str[i], &k, &l)) return 0;
    }
    *k0 = k;
    *l0 = l;
    return 1;
}

/* bwtaln.c:54-78 */
int orc_cal_width(const orc_bwt_t *b, int len, const uint8_t *str, orc_width_t *width)
{
    uint32_t k = 0, l = b->seq_len;
    int i, bid = 0;
    for (i = 0; i < len; ++i) {
        int c = str[i], alive = 0;
        if (c < 4) alive = step_exact(b, c, &k, &l);
        if (!alive) {
            k = 0;
            l = b->seq_len;
            ++bid;
        }
        width[i].w = l - k + 1;
        width[i].bid = bid;
    }
    width[len].w = 0;
    width[len].bid = ++bid;
    return bid;
}

/* bwtaln.c:39-51.  The factorial is a 32-bit int in the reference; unsigned
 * wrap-around here reproduces what the -O2 binary does past 12!. */
int orc_cal_maxdiff(int l, double err, double thres)
{
    double elambda = exp(-l * err), sum = elambda, y = 1.0;
    uint32_t fact = 1;
    int k;
    for (k = 1; k < 1000; ++k) {
        y *= l * err;
        fact *= (uint32_t)k;
        sum += elambda * y / (int32_t)fact;
        if (1.0 - sum < thres) return k;
    }
    return 2;
}

/* ------------------------------------------------------ bucketed stack ---- */

typedef struct {
    uint32_t k, l;
    int32_t i, ldp; /* ldp: last_diff_pos */
    uint8_t n_mm, n_gapo, n_gape, state, a;
    int32_t score;
    int32_t ldp_inherit; /* shadow value under the "inherit from parent" rule (stats only) */
} entry_t;

typedef struct {
    int n, cap;
    entry_t *e;
} bucket_t;

typedef struct {
    int n_buckets, best, n_entries;
    bucket_t *b;
} stack_t;

static stack_t *stack_new(int n_buckets)
{
    stack_t *s = (stack_t *)calloc(1, sizeof *s);
    int i;
    s->n_buckets = n_buckets;
    s->b = (bucket_t *)calloc((size_t)n_buckets, sizeof(bucket_t));
    for (i = 0; i < n_buckets; ++i) {
        s->b[i].cap = 4;
        s->b[i].e = (entry_t *)calloc(4, sizeof(entry_t)); /* zeroed: bwtgap.c:23 */
    }
    return s;
}

static void stack_free(stack_t *s)
{
    int i;
    for (i = 0; i < s->n_buckets; ++i) free(s->b[i].e);
    free(s->b);
    free(s);
}

static void stack_clear(stack_t *s)
{ /* counters only: slot contents survive from read to read (bwtgap.c:36-43) */
    int i;
    for (i = 0; i < s->n_buckets; ++i) s->b[i].n = 0;
    s->best = s->n_buckets;
    s->n_entries = 0;
}

static void stack_push(stack_t *s, const orc_opt_t *o, int a, int i, uint32_t k, uint32_t l, int n_mm, int n_gapo,
                       int n_gape, int state, int is_diff, int parent_ldp)
{
    int score = n_mm * o->s_mm + n_gapo * o->s_gapo + n_gape * o->s_gape;
    bucket_t *q = &s->b[score];
    entry_t *p;
    if (q->n == q->cap) {
        q->cap *= 2;
        q->e = (entry_t *)realloc(q->e, sizeof(entry_t) * (size_t)q->cap);
        /* the reference leaves the new half uninitialised; a non-diff push can
         * only land in the slot its parent just vacated (positive penalties),
         * so those bytes are never observed.  Zero them for determinism. */
        memset(q->e + q->cap / 2, 0, sizeof(entry_t) * (size_t)(q->cap / 2));
    }
    p = &q->e[q->n++];
    p->k = k; p->l = l; p->i = i; p->a = (uint8_t)a;
    p->n_mm = (uint8_t)n_mm; p->n_gapo = (uint8_t)n_gapo; p->n_gape = (uint8_t)n_gape;
    p->state = (uint8_t)state; p->score = score;
    if (is_diff) p->ldp = i; /* else: keep what the slot held (bwtgap.c:60) */
    p->ldp_inherit = is_diff ? i : parent_ldp;
    if (p->ldp != p->ldp_inherit) g_stats.inherit_violations++;
    s->n_entries++;
    if (score < s->best) s->best = score;
    g_stats.pushes++;
}

static entry_t stack_pop(stack_t *s)
{
    bucket_t *q = &s->b[s->best];
    entry_t e = q->e[--q->n];
    s->n_entries--;
    if (s->n_entries == 0) s->best = s->n_buckets;
    else if (q->n == 0) {
        int j = s->best + 1;
        while (j < s->n_buckets && s->b[j].n == 0) ++j;
        s->best = j;
    }
    return e;
}

/* --------------------------------------------------------- the search ---- */

static int ilog2_u32(uint32_t v)
{ /* floor(log2 v), 0 for v == 0 (bwtgap.c:93-102) */
    int c = 0;
    while (v > 1) { v >>= 1; ++c; }
    return c;
}

typedef struct {
    orc_aln_t *a;
    int n, cap;
} alnvec_t;

/* bwtgap.c:81-91 */
static void shadow(uint32_t x, uint32_t max, int last_diff_pos, orc_width_t *w)
{
    int i, j = 0;
    for (i = 0; i < last_diff_pos; ++i) {
        if (w[i].w > x) w[i].w -= x;
        else if (w[i].w == x) {
            w[i].bid = 1;
            w[i].w = max - (uint32_t)(++j);
        }
    }
}

/* bwtgap.c:104-264.  `o` carries the per-read max_diff / seed_len and the
 * batch-clamped max_gapo, exactly like `local_opt` in bwtaln.c:86-126. */
static void match_gap(const orc_bwt_t *const bw[2], int len, const uint8_t *const seq[2], orc_width_t *const w[2],
                      orc_width_t *const seed_w[2], int use_seed, const orc_opt_t *o, stack_t *st, alnvec_t *out)
{
    const int gape_mode = o->mode & MODE_GAPE, nonstop = o->mode & MODE_NONSTOP;
    int best_score = (o->max_diff + 1) * o->s_mm + (o->max_gapo + 1) * o->s_gapo + (o->max_gape + 1) * o->s_gape;
    int best_diff = o->max_diff + 1, max_diff = o->max_diff;
    int32_t best_cnt = 0; /* int in the reference; wraps like the -O2 binary */
    int j, n_N = 0;
    uint64_t hwm = 0;

    out->n = 0;
    for (j = 0; j < len; ++j) n_N += seq[0][j] > 3;
    if (n_N > max_diff) return;

    stack_clear(st);
    stack_push(st, o, 0, len, 0, bw[0]->seq_len, 0, 0, 0, ST_M, 0, 0);
    stack_push(st, o, 1, len, 0, bw[0]->seq_len, 0, 0, 0, ST_M, 0, 0);

    while (st->n_entries) {
        entry_t e;
        const orc_bwt_t *b;
        const uint8_t *str;
        orc_width_t *width;
        const orc_width_t *sw = 0;
        uint32_t k, l, ck[4], cl[4], occ;
        int a, i, m, m_seed = 0, hit = 0, allow_diff = 1, allow_M = 1, gaps;

        if ((uint64_t)st->n_entries > hwm) hwm = (uint64_t)st->n_entries;
        if (st->n_entries > o->max_entries) { g_stats.cutoff_reads++; break; }
        e = stack_pop(st);
        g_stats.pops++;
        k = e.k; l = e.l; a = e.a; i = e.i;
        if (!nonstop && e.score > best_score + o->s_mm) break;

        m = max_diff - e.n_mm - e.n_gapo - (gape_mode ? e.n_gape : 0);
        if (m < 0) continue;
        b = bw[1 - a]; str = seq[a]; width = w[a];
        if (use_seed) {
            sw = seed_w[a];
            m_seed = o->max_seed_diff - e.n_mm - e.n_gapo - (gape_mode ? e.n_gape : 0);
        }
        if (i > 0 && m < width[i - 1].bid) continue;

        if (i == 0) hit = 1;
        else if (m == 0 && (e.state == ST_M || gape_mode || e.n_gape == o->max_gape)) {
            if (!extend_exact(b, i, str, &k, &l)) continue;
            hit = 1;
        }

        if (hit) {
            int add = 1;
            if (out->n == 0) {
                best_score = e.score;
                best_diff = e.n_mm + e.n_gapo + (gape_mode ? e.n_gape : 0);
                if (!nonstop) max_diff = best_diff + 1 > o->max_diff ? o->max_diff : best_diff + 1;
            }
            if (e.score == best_score) best_cnt = (int32_t)((uint32_t)best_cnt + (l - k + 1));
            else if (best_cnt > o->max_top2) break;
            if (e.n_gapo)
                for (j = 0; j < out->n; ++j)
                    if (out->a[j].k == k && out->a[j].l == l) { add = 0; break; }
            if (add) {
                orc_aln_t *p;
                shadow(l - k + 1, b->seq_len, e.ldp, width);
                if (out->n == out->cap) {
                    out->cap = out->cap ? out->cap * 2 : 4;
                    out->a = (orc_aln_t *)realloc(out->a, sizeof(orc_aln_t) * (size_t)out->cap);
                }
                p = &out->a[out->n++];
                p->packed = (uint32_t)e.n_mm | (uint32_t)e.n_gapo << 8 | (uint32_t)e.n_gape << 16 | (uint32_t)a << 24;
                p->k = k; p->l = l; p->score = e.score;
                g_stats.hits++;
            }
            continue;
        }

        --i;
        g_stats.occ4_calls++;
        orc_occ4(b, k - 1, ck);
        if (k - 1 == l) memcpy(cl, ck, sizeof cl);
        else orc_occ4(b, l, cl);
        occ = l - k + 1;

        if (i > 0) {
            int ii = i - (len - o->seed_len);
            if (width[i - 1].bid > m - 1) allow_diff = 0;
            else if (width[i - 1].bid == m - 1 && width[i].bid == m - 1 && width[i - 1].w == width[i].w) allow_M = 0;
            if (use_seed && ii > 0) {
                if (sw[ii - 1].bid > m_seed - 1) allow_diff = 0;
                else if (sw[ii - 1].bid == m_seed - 1 && sw[ii].bid == m_seed - 1 && sw[ii - 1].w == sw[ii].w) allow_M = 0;
            }
        }

        gaps = (o->mode & MODE_LOGGAP) ? ilog2_u32((uint32_t)(e.n_gape + e.n_gapo)) / 2 + 1 : e.n_gapo + e.n_gape;
        if (allow_diff && i >= o->indel_end_skip + gaps && len - i >= o->indel_end_skip + gaps) {
            if (e.state == ST_M) {
                if (e.n_gapo < o->max_gapo) {
                    stack_push(st, o, a, i, k, l, e.n_mm, e.n_gapo + 1, e.n_gape, ST_I, 1, e.ldp);
                    for (j = 0; j < 4; ++j) {
                        uint32_t nk = b->L2[j] + ck[j] + 1, nl = b->L2[j] + cl[j];
                        if (nk <= nl) stack_push(st, o, a, i + 1, nk, nl, e.n_mm, e.n_gapo + 1, e.n_gape, ST_D, 1, e.ldp);
                    }
                }
            } else if (e.state == ST_I) {
                if (e.n_gape < o->max_gape)
                    stack_push(st, o, a, i, k, l, e.n_mm, e.n_gapo, e.n_gape + 1, ST_I, 1, e.ldp);
            } else {
                if (e.n_gape < o->max_gape && (e.n_gape + e.n_gapo < max_diff || occ < (uint32_t)o->max_del_occ))
                    for (j = 0; j < 4; ++j) {
                        uint32_t nk = b->L2[j] + ck[j] + 1, nl = b->L2[j] + cl[j];
                        if (nk <= nl) stack_push(st, o, a, i + 1, nk, nl, e.n_mm, e.n_gapo, e.n_gape + 1, ST_D, 1, e.ldp);
                    }
            }
        }

        if (allow_diff && allow_M) {
            for (j = 1; j <= 4; ++j) {
                int c = (str[i] + j) & 3, is_mm = (j != 4 || str[i] > 3);
                uint32_t nk = b->L2[c] + ck[c] + 1, nl = b->L2[c] + cl[c];
                if (nk <= nl) stack_push(st, o, a, i, nk, nl, e.n_mm + is_mm, e.n_gapo, e.n_gape, ST_M, is_mm, e.ldp);
            }
        } else if (str[i] < 4) {
            int c = str[i];
            uint32_t nk = b->L2[c] + ck[c] + 1, nl = b->L2[c] + cl[c];
            if (nk <= nl) stack_push(st, o, a, i, nk, nl, e.n_mm, e.n_gapo, e.n_gape, ST_M, 0, e.ldp);
        }
    }
    if (hwm > g_stats.stack_hwm) g_stats.stack_hwm = hwm;
}

/* ----------------------------------------------------------- the batch ---- */

/* bwaseqio.c:74-87: returns the kept length (BWA_MIN_RDLEN = 35) */
int orc_trim_len(int trim_qual, int len, const uint8_t *qual)
{
    int s = 0, l, best = 0, best_l = len - 1;
    if (trim_qual < 1 || qual == 0) return len;
    for (l = len - 1; l >= 35 - 1; --l) {
        s += trim_qual - ((int)qual[l] - 33);
        if (s < 0) break;
        if (s > best) { best = s; best_l = l; }
    }
    return best_l + 1;
}

/*
 * One reference batch (bwtaln.c:80-140).  Reads are given in sequencing
 * orientation as nt4 codes (0-3, >3 = ambiguous) already trimmed; read r
 * occupies fwd[off[r] .. off[r]+len[r]).
 * Output: n_aln[r] and a malloc'ed concatenation of the records in read order
 * (caller frees with orc_free).  Returns the total record count.
 */
int64_t orc_aln_batch(const orc_bwt_t *bwt, const orc_bwt_t *rbwt, int n_reads, const int32_t *len,
                      const int64_t *off, const uint8_t *fwd, const orc_opt_t *opt, int32_t *n_aln,
                      orc_aln_t **records)
{
    const orc_bwt_t *bw[2];
    orc_opt_t lo = *opt;
    stack_t *st;
    alnvec_t hits = {0, 0, 0};
    orc_aln_t *all = 0;
    int64_t n_all = 0, cap_all = 0;
    int r, max_len = 0;
    orc_width_t *w[2], *seed_w[2];
    uint8_t *seq[2];

    bw[0] = bwt; bw[1] = rbwt;
    for (r = 0; r < n_reads; ++r) if (len[r] > max_len) max_len = len[r];
    if (opt->fnr > 0.0f) lo.max_diff = orc_cal_maxdiff(max_len, 0.02, opt->fnr);
    if (lo.max_diff < lo.max_gapo) lo.max_gapo = lo.max_diff;
    st = stack_new((lo.max_diff + 1) * lo.s_mm + (lo.max_gapo + 1) * lo.s_gapo + (lo.max_gape + 1) * lo.s_gape);

    w[0] = (orc_width_t *)calloc((size_t)max_len + 1, sizeof(orc_width_t));
    w[1] = (orc_width_t *)calloc((size_t)max_len + 1, sizeof(orc_width_t));
    seed_w[0] = (orc_width_t *)calloc((size_t)opt->seed_len + 1, sizeof(orc_width_t));
    seed_w[1] = (orc_width_t *)calloc((size_t)opt->seed_len + 1, sizeof(orc_width_t));
    seq[0] = (uint8_t *)malloc((size_t)max_len + 1);
    seq[1] = (uint8_t *)malloc((size_t)max_len + 1);

    for (r = 0; r < n_reads; ++r) {
        const uint8_t *f = fwd + off[r];
        int L = len[r], j, use_seed;
        const uint8_t *cseq[2];
        for (j = 0; j < L; ++j) {
            uint8_t c = f[L - 1 - j];
            seq[0][j] = c; /* reversed */
            seq[1][j] = (opt->mode & MODE_COMPREAD) ? (c < 4 ? (uint8_t)(3 - c) : c) : c; /* reverse(-complement) */
        }
        cseq[0] = seq[0]; cseq[1] = seq[1];
        orc_cal_width(bw[0], L, seq[0], w[0]);
        orc_cal_width(bw[1], L, seq[1], w[1]);
        if (opt->fnr > 0.0f) lo.max_diff = orc_cal_maxdiff(L, 0.02, opt->fnr);
        use_seed = L > opt->seed_len;
        lo.seed_len = use_seed ? opt->seed_len : 0x7fffffff;
        if (use_seed) {
            orc_cal_width(bw[0], opt->seed_len, seq[0] + (L - opt->seed_len), seed_w[0]);
            orc_cal_width(bw[1], opt->seed_len, seq[1] + (L - opt->seed_len), seed_w[1]);
        }
        match_gap(bw, L, cseq, w, seed_w, use_seed, &lo, st, &hits);
        n_aln[r] = hits.n;
        if (n_all + hits.n > cap_all) {
            cap_all = (n_all + hits.n) * 2 + 1024;
            all = (orc_aln_t *)realloc(all, sizeof(orc_aln_t) * (size_t)cap_all);
        }
        if (hits.n) memcpy(all + n_all, hits.a, sizeof(orc_aln_t) * (size_t)hits.n);
        n_all += hits.n;
    }
    free(hits.a);
    free(seq[0]); free(seq[1]);
    free(w[0]); free(w[1]); free(seed_w[0]); free(seed_w[1]);
    stack_free(st);
    *records = all;
    return n_all;
}

void orc_free(void *p) { free(p); }

/* ------------------------------------------ SA row -> position (N2) ---- */

typedef struct { /* bwt_restore_sa, bwtio.c:29-49: sa[0] = -1, sa[j] = SA(j * sa_intv) */
    int32_t sa_intv;
    uint64_t n_sa;
    const uint32_t *sa;
} orc_sa_t;

/* character of the sentinel-free BWT string at index p (bwt.h:58-63) */
static int bwt_char(const orc_bwt_t *b, uint32_t p)
{
    const uint32_t *blk = b->bwt + (uint64_t)(p / ORC_BLOCK) * ORC_BLOCK_WORDS;
    uint32_t in = p % ORC_BLOCK;
    return (int)(blk[4 + in / 16] >> (2 * (15 - in % 16)) & 3u);
}

/* inverse Psi (bwt.h:66-70): the row of the suffix one position to the left */
static uint32_t inv_psi(const orc_bwt_t *b, uint32_t k)
{
    int c;
    if (k == b->primary) return 0;
    c = bwt_char(b, k < b->primary ? k : k - 1);
    return b->L2[c] + orc_occ(b, k, c);
}

/* bwt_sa (bwt.c:69-79): walk left until a sampled row; note sa[0] == (uint32_t)-1 */
uint32_t orc_bwt_sa(const orc_bwt_t *b, const orc_sa_t *s, uint32_t k)
{
    uint32_t steps = 0;
    while (k % (uint32_t)s->sa_intv != 0) {
        ++steps;
        k = inv_psi(b, k);
    }
    return steps + s->sa[k / (uint32_t)s->sa_intv];
}

/* bwtdb_sa2seq with offset 0 (dbset.c:240-245): strand != 0 uses (bwt, sa), else (rbwt, rsa) */
uint64_t orc_sa2seq(const orc_bwt_t *bwt, const orc_sa_t *sa, const orc_bwt_t *rbwt, const orc_sa_t *rsa, int strand,
                    uint32_t row, int seq_len)
{
    if (strand) return (uint64_t)orc_bwt_sa(bwt, sa, row);
    return (uint64_t)(uint32_t)(rbwt->seq_len - (orc_bwt_sa(rbwt, rsa, row) + (uint32_t)seq_len));
}

/* ---------------------------------- multi-.sai merge (N4) ---- */
/* TEST INFRASTRUCTURE.  alngrp_create (saiset.c:45-78) for one read: the alignments of every stream, stream by
 * stream; with more than one stream they are sorted with klib's introsort (ksort.h:172-224, comparator
 * a.score < b.score, saiset.c:8-10) and cut at the first score > best + s_mm.  The sort is not stable, so it
 * is restated step by step (same comparisons, same exchanges).  Pinned against the reference's own
 * saiset.c through oracle/_ref/alngrp_dump (tests/test_alngrp.py). */
typedef struct { orc_aln_t aln; uint32_t dbidx; } orc_galn_t;

#define G_LT(a, b) ((a).aln.score < (b).aln.score)

static void g_insertsort(orc_galn_t *s, orc_galn_t *t) /* ksort.h:142-149 */
{
    orc_galn_t *i, *j, tmp;
    for (i = s + 1; i < t; ++i)
        for (j = i; j > s && G_LT(*j, *(j - 1)); --j) { tmp = *j; *j = *(j - 1); *(j - 1) = tmp; }
}

static void g_combsort(size_t n, orc_galn_t *a) /* ksort.h:150-171 */
{
    const double shrink_factor = 1.2473309501039786540366528676643;
    int do_swap;
    size_t gap = n;
    orc_galn_t tmp, *i, *j;
    do {
        if (gap > 2) {
            gap = (size_t)(gap / shrink_factor);
            if (gap == 9 || gap == 10) gap = 11;
        }
        do_swap = 0;
        for (i = a; i < a + n - gap; ++i) {
            j = i + gap;
            if (G_LT(*j, *i)) { tmp = *i; *i = *j; *j = tmp; do_swap = 1; }
        }
    } while (do_swap || gap > 2);
    if (gap != 1) g_insertsort(a, a + n);
}

static void g_introsort(size_t n, orc_galn_t *a) /* ksort.h:172-224 */
{
    struct { orc_galn_t *left, *right; int depth; } stack[8 * 66 + 2], *top = stack;
    orc_galn_t rp, tmp, *s, *t, *i, *j, *k;
    int d;
    if (n < 1) return;
    if (n == 2) {
        if (G_LT(a[1], a[0])) { tmp = a[0]; a[0] = a[1]; a[1] = tmp; }
        return;
    }
    for (d = 2; 1ul << d < n; ++d) {}
    s = a; t = a + (n - 1); d <<= 1;
    for (;;) {
        if (s < t) {
            if (--d == 0) { g_combsort(t - s + 1, s); t = s; continue; }
            i = s; j = t; k = i + ((j - i) >> 1) + 1;
            if (G_LT(*k, *i)) { if (G_LT(*k, *j)) k = j; }
            else k = G_LT(*j, *i) ? i : j;
            rp = *k;
            if (k != t) { tmp = *k; *k = *t; *t = tmp; }
            for (;;) {
                do ++i; while (G_LT(*i, rp));
                do --j; while (i <= j && G_LT(rp, *j));
                if (j <= i) break;
                tmp = *i; *i = *j; *j = tmp;
            }
            tmp = *i; *i = *t; *t = tmp;
            if (i - s > t - i) {
                if (i - s > 16) { top->left = s; top->right = i - 1; top->depth = d; ++top; }
                s = t - i > 16 ? i + 1 : t;
            } else {
                if (t - i > 16) { top->left = i + 1; top->right = t; top->depth = d; ++top; }
                t = i - s > 16 ? i - 1 : s;
            }
        } else {
            if (top == stack) { g_insertsort(a, a + n); return; }
            --top; s = top->left; t = top->right; d = top->depth;
        }
    }
}

/* whole batch: n_aln[s][r], recs[s] packed in read order; out_* sized for the sum of all counts;
 * out_off[r] = first slot of read r (sum of the unmerged totals before it), out_n[r] = merged size */
int64_t orc_alngrp_merge(int n_streams, int n_reads, const int32_t *const *n_aln, const orc_aln_t *const *recs, int s_mm,
                         int64_t *out_off, int32_t *out_n, orc_aln_t *out_recs, uint32_t *out_dbidx)
{
    int64_t *pos = (int64_t *)calloc((size_t)n_streams, sizeof(int64_t)), at = 0;
    orc_galn_t *g = NULL;
    size_t cap = 0;
    for (int r = 0; r < n_reads; ++r) {
        size_t n = 0, i;
        for (int s = 0; s < n_streams; ++s) n += (size_t)n_aln[s][r];
        if (n > cap) { cap = n * 2 + 16; g = (orc_galn_t *)realloc(g, cap * sizeof *g); }
        n = 0;
        for (int s = 0; s < n_streams; ++s) /* saiset.c:51-62 */
            for (int j = 0; j < n_aln[s][r]; ++j) { g[n].aln = recs[s][pos[s]++]; g[n].dbidx = (uint32_t)s; ++n; }
        const size_t total = n;
        if (n_streams > 1 && n > 0) { /* saiset.c:64-76 */
            g_introsort(n, g);
            const int best = g[0].aln.score;
            for (i = 0; i < n; ++i)
                if (g[i].aln.score > best + s_mm) { n = i; break; }
        }
        out_off[r] = at;
        out_n[r] = (int32_t)n;
        for (i = 0; i < n; ++i) { out_recs[at + (int64_t)i] = g[i].aln; out_dbidx[at + (int64_t)i] = g[i].dbidx; }
        at += (int64_t)total;
    }
    free(g); free(pos);
    return at;
}
