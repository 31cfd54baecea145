/*
 * aln_core.cuh — per-read state machines of the `aln` hot path.
 *
 * Everything here is written once and compiled twice: by nvcc for sm_100a (the
 * product kernels in b200aln.cu) and by g++ for the logic tests in
 * tests/harness (which run the same state machines on the CPU next to the
 * oracle).  No warp intrinsics in this file; the warp-level plumbing (work
 * distribution, launches) lives in b200aln.cu.
 *
 * Reference behaviour being re-implemented (file:line under the reference):
 *   occ / 2occ / 2occ4 ............ bwt.c:81-214
 *   bwt_match_exact_alt ........... bwt.c:235-250
 *   bwt_cal_width ................. bwtaln.c:54-78
 *   bwt_match_gap ................. bwtgap.c:104-264
 *   gap_push / gap_pop / shadow ... bwtgap.c:45-91
 *
 * Device index layout (one 32-byte sector per occ lookup):
 *   block b covers bases [64b, 64b+64) of the sentinel-free BWT string:
 *     U4 #0 : L2[c] + (number of c before base 64b), c = A,C,G,T
 *     U4 #1 : lo0, lo1, hi0, hi1 — two bit planes, base j at bit j%32 of word j/32
 *   occ_excl(q) (count in [0,q)) therefore needs block q>>6 and q&63 mask bits.
 *   Row numbers map to q without special cases:
 *     lower end  occ(k-1): q = k - (k > primary)
 *     upper end  occ(l)  : q = l + 1 - (l >= primary)
 *   which also covers k-1 == -1 (q = 0) and l == seq_len.
 */
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#define B2_D __device__ __forceinline__
#else
#define B2_HD inline
#define B2_D inline
#endif

#if defined(B2_CHECKED) && !defined(__CUDACC__)
#include <stdio.h>
#include <stdlib.h>
#endif

namespace b2 {

/*
 * Bounds-checked build (-DB2_CHECKED; libb200aln_checked.so and the CPU harness's checked variant): every
 * index the state machines derive from data in memory — arena slots, bucket numbers, width-record positions,
 * record-slab fills, interval-table and occ-block addresses, work items — is tested before it is used.  Stands
 * in for compute-sanitizer, which the GPU pool does not allow.  On the device the first violation is recorded
 * (code, read row, the two values) and the lane ends with LANE_CHECK; the host reports it and aborts.  On the
 * CPU it aborts on the spot.  The product build compiles the checks away.
 */
enum { CHK_ARENA_SLOT = 1, CHK_BUCKET = 2, CHK_POP_EMPTY = 3, CHK_Q_POS = 4, CHK_REC_FILL = 5, CHK_LUT = 6,
       CHK_BLK = 7, CHK_MEMBER = 8, CHK_WORK = 9, CHK_ENTRIES = 10 };
#ifdef B2_CHECKED
#if defined(__CUDACC__)
__device__ unsigned int b2_check_rec[4]; /* code (0 = clean), row, a, b of the first violation */
__host__ __device__ __forceinline__ bool b2_check_fail(int code, uint32_t row, uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    if (atomicCAS(&b2_check_rec[0], 0u, (unsigned)code) == 0u) { b2_check_rec[1] = row; b2_check_rec[2] = a; b2_check_rec[3] = b; }
#else
    (void)code; (void)row; (void)a; (void)b; /* host pass of nvcc: never called */
#endif
    return false;
}
#else
inline bool b2_check_fail(int code, uint32_t row, uint32_t a, uint32_t b)
{
    fprintf(stderr, "[B2_CHECKED] violation %d at row %u: %u vs %u\n", code, row, a, b);
    abort();
}
#endif
#define B2_CHECK(cond, code, row, a, b) ((cond) || b2_check_fail((code), (uint32_t)(row), (uint32_t)(a), (uint32_t)(b)))
#else
#define B2_CHECK(cond, code, row, a, b) (true)
#endif

struct alignas(16) U4 {
    uint32_t x, y, z, w;
};

/* one occ block = one 32-byte sector: checkpoint counts + two bit planes of 64 bases */
struct alignas(32) OccBlk {
    U4 cnt;  /* L2[c] + occurrences of c before the block */
    U4 bits; /* lo0, lo1, hi0, hi1 */
};

/*
 * One stack record = the children of ONE expansion (bwtgap.c:216-258), 64 bytes = one aligned pair of
 * sectors (one DRAM transaction: L2 fills are 64 bytes wide).  The reference pushes the children of an
 * expansion as two runs of consecutive entries: the gap group (insertion + deletions, or one gap
 * extension) into bucket score + s_gapo / s_gape and the mismatch group into bucket score + s_mm.  All
 * of them are the parent's interval or one of the same four child intervals, so the record stores the
 * parent and the four children once and is linked into both buckets.
 *   w[0] link of the gap group   w[1] link of the mismatch group   (bucket lists; w[0] = free list)
 *   w[2] parent's position in the interval table (path word)
 *   w[3] i | base << 16 | parent state << 19 | strand << 21 | gap mask << 22 | mismatch mask << 27
 *        (masks: bit c = child c is a member, bit 4 of the gap mask = the insertion)
 *   w[4] parent's n_mm | n_gapo << 8 | n_gape << 16      w[5], w[6] parent's k, l
 *   w[7] set once one of the two groups has been taken (free-list arenas only)
 *   w[8..15] children: k0, l0, k1, l1, k2, l2, k3, l3
 * A bucket head or link is (slot << 1 | group), group 0 = gap, 1 = mismatch.
 */
struct alignas(64) StackRec {
    uint32_t w[16];
};

#if defined(__CUDA_ARCH__)
#ifdef B2_L2_HINTS
/* L2 eviction priorities: the 3 GB of occ blocks and the stack entries stream through (evict_first), so
 * that the small per-read width records (evict_last) survive in L2 between the steps of a search */
B2_D uint64_t pol_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
B2_D uint64_t pol_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
#endif
B2_D OccBlk ld_blk(const OccBlk *p)
{ /* read-only index data: one 256-bit load on the non-coherent path, not allocated in L1
     (no reuse; keeps L1 for the per-read width records that ARE re-read along a chain) */
    OccBlk r;
#ifdef B2_L2_HINTS
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(r.cnt.x), "=r"(r.cnt.y), "=r"(r.cnt.z), "=r"(r.cnt.w), "=r"(r.bits.x), "=r"(r.bits.y),
                   "=r"(r.bits.z), "=r"(r.bits.w)
                 : "l"(p), "l"(pol_evict_first()));
#else
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.cnt.x), "=r"(r.cnt.y), "=r"(r.cnt.z), "=r"(r.cnt.w), "=r"(r.bits.x), "=r"(r.bits.y),
                   "=r"(r.bits.z), "=r"(r.bits.w)
                 : "l"(p));
#endif
    return r;
}
B2_D void ld8cg(const uint32_t *p, uint32_t v[8])
{ /* stack records: written once, read at most twice -> L2 only, first in line for eviction */
#if defined(B2_L2_HINTS) && defined(B2_HINT_REC)
    asm volatile("ld.global.cg.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p), "l"(pol_evict_first()) : "memory");
#else
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p) : "memory");
#endif
}
B2_D void st8rec(uint32_t *p, const uint32_t v[8])
{ /* one sector of a stack record */
#if defined(B2_L2_HINTS) && defined(B2_HINT_REC)
    asm volatile("st.global.cg.L2::cache_hint.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8}, %9;"
                 :: "l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                    "l"(pol_evict_first()) : "memory");
#else
    asm volatile("st.global.cg.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
#endif
}
struct alignas(32) U8x { uint32_t v[8]; };
B2_D U8x ld_lut8(const void *p, bool keep = false)
{ /* the four children intervals of a node: one sector of the interval table.  keep: a shallow level, which
     all reads share and which should stay in L2; deep levels are touched once per chain */
    U8x r;
#if defined(B2_L2_HINTS) && defined(B2_HINT_LUT)
    if (keep)
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                     : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                       "=r"(r.v[7])
                     : "l"(p), "l"(pol_evict_last()));
    else
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                     : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                       "=r"(r.v[7])
                     : "l"(p), "l"(pol_evict_first()));
#else
    (void)keep;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                   "=r"(r.v[7])
                 : "l"(p));
#endif
    return r;
}
B2_D void ld_lut_pair(const uint32_t *p, uint32_t &k, uint32_t &l)
{ /* one (k, l) pair of the interval table */
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(k), "=r"(l) : "l"(p));
}
B2_D void st8(uint32_t *p, const uint32_t v[8])
{ /* eight consecutive words as one full 32-byte sector (p is 32-byte aligned) */
    asm volatile("st.global.cg.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
B2_D uint32_t ld_q(const uint32_t *p)
{ /* width records: re-read along a chain (8 per sector) -> keep them in L1 */
#if defined(B2_Q_EVICT_LAST) && defined(B2_L2_HINTS)
    uint32_t v;
    asm volatile("ld.global.L1::evict_last.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol_evict_last()) : "memory");
    return v;
#elif defined(B2_Q_EVICT_LAST)
    uint32_t v;
    asm volatile("ld.global.L1::evict_last.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#else
    return *p;
#endif
}
B2_D uint32_t ld_q(const uint16_t *p)
{ /* 16-bit width records (QF<16>), same cache policy */
    uint16_t v;
#if defined(B2_Q_EVICT_LAST) && defined(B2_L2_HINTS)
    asm volatile("ld.global.L1::evict_last.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(v) : "l"(p), "l"(pol_evict_last()) : "memory");
#elif defined(B2_Q_EVICT_LAST)
    asm volatile("ld.global.L1::evict_last.u16 %0, [%1];" : "=h"(v) : "l"(p) : "memory");
#else
    v = *p;
#endif
    return v;
}
B2_D void ld4(const uint32_t *p, uint32_t v[4])
{ /* four consecutive words of a 16-byte aligned row piece */
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "l"(p) : "memory");
}
B2_D void st4(uint32_t *p, const uint32_t v[4])
{
    asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
B2_D void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
B2_D void ld8(const uint32_t *p, uint32_t v[8])
{ /* eight consecutive words of a 32-byte aligned row */
    asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p) : "memory");
}
B2_D int popc32(uint32_t v) { return __popc(v); }
B2_D int ctz32(uint32_t v) { return __ffs((int)v) - 1; }
#else
inline void ld8(const uint32_t *p, uint32_t v[8]) { for (int i = 0; i < 8; ++i) v[i] = p[i]; }
inline void prefetch_l2(const void *) {}
inline uint32_t ld_q(const uint32_t *p) { return *p; }
inline uint32_t ld_q(const uint16_t *p) { return *p; }
inline void ld4(const uint32_t *p, uint32_t v[4]) { for (int i = 0; i < 4; ++i) v[i] = p[i]; }
inline void st4(uint32_t *p, const uint32_t v[4]) { for (int i = 0; i < 4; ++i) p[i] = v[i]; }
inline void st8(uint32_t *p, const uint32_t v[8]) { for (int i = 0; i < 8; ++i) p[i] = v[i]; }
inline OccBlk ld_blk(const OccBlk *p) { return *p; }
inline void ld8cg(const uint32_t *p, uint32_t v[8]) { for (int i = 0; i < 8; ++i) v[i] = p[i]; }
inline void st8rec(uint32_t *p, const uint32_t v[8]) { for (int i = 0; i < 8; ++i) p[i] = v[i]; }
struct alignas(32) U8x { uint32_t v[8]; };
inline U8x ld_lut8(const void *p, bool = false) { return *reinterpret_cast<const U8x *>(p); }
inline void ld_lut_pair(const uint32_t *p, uint32_t &k, uint32_t &l) { k = p[0]; l = p[1]; }
inline int popc32(uint32_t v) { return __builtin_popcount(v); }
inline int ctz32(uint32_t v) { return __builtin_ctz(v); }
#endif

enum { MODE_GAPE = 0x01, MODE_COMPREAD = 0x02, MODE_LOGGAP = 0x04, MODE_NONSTOP = 0x10 };
enum { ST_M = 0, ST_I = 1, ST_D = 2 };

struct FmView {
    const OccBlk *blk; /* one per 64 bases */
    uint32_t primary; /* row of the sentinel */
    uint32_t seq_len;
    /* Interval table of the first lut_k search levels (0 = none): level L (1..lut_k) holds, for every
     * L-mer X (the characters prepended so far, first one most significant), the SA interval
     * (k, l) of X as two u32, empty = (1, 0).  The four children X*4+c of a node are one 32-byte
     * sector, so an expansion above level lut_k costs ONE request instead of two occ sectors,
     * and the small top levels stay L2 resident. */
    const uint32_t *lut; /* shared by both indexes: level L = [index 0: 4^L pairs][index 1: 4^L pairs] */
    int lut_k;
    int lut_w;           /* which half of every level is this index's (0 = bwt, 1 = rbwt) */
};

#define B2_PATH_DEAD 31u /* depth field value of an entry that left the interval table */
/* first (k,l) pair of a level: (4^level - 4) / 3; (4^L - 1) / 3 is the bit pattern 0101..01 with L ones */
B2_HD uint64_t lut_level_off(int level) { return (0x5555555555555555ull >> (64 - 2 * level)) - 1u; }
/* pair index of node X of `level` for index half w.  The levels of BOTH indexes are interleaved so that
 * the small, hot top of the table is one contiguous range (pinned in L2 with an access-policy window). */
B2_HD uint64_t lut_pair(int w, int level, uint64_t X) { return 2u * lut_level_off(level) + ((uint64_t)w << (2 * level)) + X; }
B2_HD uint64_t lut_total_pairs(int lut_k) { return lut_k > 0 ? 2u * lut_level_off(lut_k + 1) : 0; } /* both indexes */
B2_HD uint32_t path_root() { return 0u; }
/* path of the child reached by prepending character c to a node with path p */
B2_HD uint32_t path_ext(uint32_t p, int c, int lut_k)
{
    const uint32_t d = p & 31u; /* B2_PATH_DEAD = 31 >= any lut_k, so one comparison covers both cases */
    if ((int)d + 1 >= lut_k) return B2_PATH_DEAD;
    return (d + 1u) | (((p >> 5) << 2 | (uint32_t)c) << 5);
}

/* Launch-constant search parameters: gap_opt_t after the batch-level clamps of
 * bwtaln.c:89-92 (max_gapo) — per-read max_diff comes from a table. */
struct Params {
    int s_mm, s_gapo, s_gape;
    int mode;
    int indel_end_skip, max_del_occ, max_entries;
    int max_gapo, max_gape, max_seed_diff, seed_len;
    int max_top2;
    int n_buckets;
};

struct Rec { /* == bwt_aln1_t */
    uint32_t packed, k, l;
    int32_t score;
};

/* everything a search lane needs that is constant for a launch.  The per-read buffers are addressed from
 * these bases with small per-lane indexes (row, slab, lane number), so that a lane carries no pointers: on
 * the device the bases stay in the kernel-parameter constant bank. */
struct SearchEnv {
    FmView fm[2]; /* fm[0] = bwt, fm[1] = rbwt */
    Params P;
    int prefetch_next; /* 1: prefetch the next pop candidate into L2 (helps when few, long reads are left) */
    uint32_t *Q;  /* width records: row 2 * read + strand, strideQ records each (QF<32> words or QF<16> halfwords) */
    uint32_t *W;  /* widths: row 2 * read + strand, strideW words each */
    int strideQ, strideW;
    Rec *recs;    /* record slabs, rec_cap each */
    int rec_cap;
    StackRec *ent; /* arenas, arena_cap records per lane */
    uint32_t arena_cap;
};

/* ---- packed per-position record Q[a][j] (built by the width pass) -------- */
/* Everything the search asks about position j of a strand in one record: base code, bid[j], bid[j-1],
 * w[j-1]==w[j], and the seed-width equivalents (bwtgap.c:155,205-213).  The bid fields saturate; every
 * comparison in the search is against m <= max_diff resp. m_seed <= max_seed_diff, so a field that
 * saturates above those bounds is exact.  Two formats (QF<bits>):
 *   QF<32>  bits 0-2 base | 3 eq | 4 seed eq | 5 seed active (ii>0) | 6-10 seed bid[ii] | 11-15 seed bid[ii-1]
 *           | 16-23 bid[j] | 24-31 bid[j-1]            exact for max_diff < 255, max_seed_diff < 31
 *   QF<16>  bits 0-2 base | 3 eq | 4 seed eq | 5 seed active | 6-7 seed bid[ii] | 8-9 seed bid[ii-1]
 *           | 10-12 bid[j] | 13-15 bid[j-1]            exact for max_diff < 7, max_seed_diff < 3
 * The 16-bit records halve the per-read state the fast pass re-reads along its chains (the rows of the
 * lanes in flight then stay in L2); the host picks the format per launch (b200aln.cu: q16_ok). */
typedef uint32_t QRec; /* a record of either format, widened */
template <int QB> struct QF;
template <> struct QF<32> {
    typedef uint32_t T;
    enum { PER_SECTOR = 8 };
    static B2_HD uint32_t pack(uint32_t base, uint32_t eq, uint32_t seq, uint32_t sact, uint32_t sb, uint32_t sbp,
                               uint32_t bid, uint32_t bidp)
    {
        return (base & 7u) | (eq & 1u) << 3 | (seq & 1u) << 4 | (sact & 1u) << 5 | (sb > 31u ? 31u : sb) << 6 |
               (sbp > 31u ? 31u : sbp) << 11 | (bid > 255u ? 255u : bid) << 16 | (bidp > 255u ? 255u : bidp) << 24;
    }
    static B2_HD int sbid(QRec q) { return (int)(q >> 6 & 31u); }
    static B2_HD int sbidp(QRec q) { return (int)(q >> 11 & 31u); }
    static B2_HD int bid(QRec q) { return (int)(q >> 16 & 255u); }
    static B2_HD int bidp(QRec q) { return (int)(q >> 24); }
    static B2_HD QRec with_bid(QRec q, uint32_t v) { return (q & ~(0xffu << 16)) | v << 16; }
    static B2_HD QRec with_prev(QRec q, uint32_t prev_bid, uint32_t eq)
    {
        return (q & ~(0xffu << 24 | 1u << 3)) | prev_bid << 24 | eq << 3;
    }
};
template <> struct QF<16> {
    typedef uint16_t T;
    enum { PER_SECTOR = 16 };
    static B2_HD uint32_t pack(uint32_t base, uint32_t eq, uint32_t seq, uint32_t sact, uint32_t sb, uint32_t sbp,
                               uint32_t bid, uint32_t bidp)
    {
        return (base & 7u) | (eq & 1u) << 3 | (seq & 1u) << 4 | (sact & 1u) << 5 | (sb > 3u ? 3u : sb) << 6 |
               (sbp > 3u ? 3u : sbp) << 8 | (bid > 7u ? 7u : bid) << 10 | (bidp > 7u ? 7u : bidp) << 13;
    }
    static B2_HD int sbid(QRec q) { return (int)(q >> 6 & 3u); }
    static B2_HD int sbidp(QRec q) { return (int)(q >> 8 & 3u); }
    static B2_HD int bid(QRec q) { return (int)(q >> 10 & 7u); }
    static B2_HD int bidp(QRec q) { return (int)(q >> 13 & 7u); }
    static B2_HD QRec with_bid(QRec q, uint32_t v) { return (q & ~(7u << 10)) | v << 10; }
    static B2_HD QRec with_prev(QRec q, uint32_t prev_bid, uint32_t eq)
    {
        return (q & ~(7u << 13 | 1u << 3)) | prev_bid << 13 | eq << 3;
    }
};
/* the low six bits are common to both formats */
B2_HD int q_base(QRec q) { return (int)(q & 7u); }
B2_HD int q_eq(QRec q) { return (int)(q >> 3 & 1u); }
B2_HD int q_seq(QRec q) { return (int)(q >> 4 & 1u); }
B2_HD int q_sact(QRec q) { return (int)(q >> 5 & 1u); }
/* eight consecutive records of a row (8-aligned position) <-> eight widened values */
B2_HD void q_load8(const uint32_t *row, int j0, uint32_t q[8]) { ld8(row + j0, q); }
B2_HD void q_store8(uint32_t *row, int j0, const uint32_t q[8]) { st8(row + j0, q); }
B2_HD void q_load8(const uint16_t *row, int j0, uint32_t q[8])
{
    uint32_t v[4];
    ld4(reinterpret_cast<const uint32_t *>(row + j0), v);
    for (int t = 0; t < 4; ++t) { q[2 * t] = v[t] & 0xffffu; q[2 * t + 1] = v[t] >> 16; }
}
B2_HD void q_store8(uint16_t *row, int j0, const uint32_t q[8])
{
    uint32_t v[4];
    for (int t = 0; t < 4; ++t) v[t] = (q[2 * t] & 0xffffu) | q[2 * t + 1] << 16;
    st4(reinterpret_cast<uint32_t *>(row + j0), v);
}

/* ------------------------------------------------------------- occ -------- */

/* v[c] for a runtime c without forcing the array into local memory */
B2_HD uint32_t pick4(const uint32_t v[4], int c)
{
    uint32_t lo = (c & 1) ? v[1] : v[0], hi = (c & 1) ? v[3] : v[2];
    return (c & 2) ? hi : lo;
}

B2_HD uint32_t q_lower(const FmView &f, uint32_t k) { return k - (k > f.primary ? 1u : 0u); }
B2_HD uint32_t q_upper(const FmView &f, uint32_t l) { return l + 1u - (l >= f.primary ? 1u : 0u); }

B2_HD void occ_count4(U4 c, U4 b, uint32_t r, uint32_t o[4])
{
    uint32_t m0 = r >= 32u ? 0xffffffffu : ((1u << r) - 1u);
    uint32_t m1 = r > 32u ? ((1u << (r - 32u)) - 1u) : 0u;
    uint32_t t = (uint32_t)(popc32(b.x & b.z & m0) + popc32(b.y & b.w & m1));
    uint32_t g = (uint32_t)(popc32(~b.x & b.z & m0) + popc32(~b.y & b.w & m1));
    uint32_t cc = (uint32_t)(popc32(b.x & ~b.z & m0) + popc32(b.y & ~b.w & m1));
    o[0] = c.x + (r - t - g - cc);
    o[1] = c.y + cc;
    o[2] = c.z + g;
    o[3] = c.w + t;
}

/* counts (with L2 pre-added) at both ends of an interval [k,l]; one sector when
 * both ends fall in the same 64-base block (cf. bwt.c:125,187) */
B2_HD void occ2x4(const FmView &f, uint32_t k, uint32_t l, uint32_t ck[4], uint32_t cl[4], uint32_t &n_sectors)
{
    uint32_t qa = q_lower(f, k), qb = q_upper(f, l);
    if (!B2_CHECK(qa <= f.seq_len && qb <= f.seq_len && k <= l, CHK_BLK, k, qa, qb)) { qa = qb = 0; }
    const OccBlk *pa = f.blk + (qa >> 6), *pb = f.blk + (qb >> 6);
    OccBlk ba = ld_blk(pa), bb = ba;
    n_sectors = 1;
    if (pa != pb) {
        bb = ld_blk(pb);
        n_sectors = 2;
    }
    occ_count4(ba.cnt, ba.bits, qa & 63u, ck);
    occ_count4(bb.cnt, bb.bits, qb & 63u, cl);
}

/* ---- SA row -> text position (scope row N2): bwt_sa / bwt_invPsi (bwt.c:69-79, bwt.h:66-70) -------------- */

/* inverse Psi: row of the suffix one position to the left.  One occ block gives both the BWT character at
 * the row and its count. */
B2_HD uint32_t inv_psi(const FmView &f, uint32_t k)
{
    if (k == f.primary) return 0;
    const uint32_t p = (k < f.primary ? k : k - 1u); /* index in the sentinel-free BWT string */
    const OccBlk b = ld_blk(f.blk + (p >> 6));
    const uint32_t j = p & 63u, wd = j >> 5, bit = j & 31u;
    const uint32_t lo = wd ? b.bits.y : b.bits.x, hi = wd ? b.bits.w : b.bits.z;
    const int c = (int)((lo >> bit & 1u) | (hi >> bit & 1u) << 1);
    /* occurrences of c among the first j + 1 bases of the block */
    const uint32_t sel_lo0 = (c & 1) ? b.bits.x : ~b.bits.x, sel_hi0 = (c & 2) ? b.bits.z : ~b.bits.z;
    const uint32_t sel_lo1 = (c & 1) ? b.bits.y : ~b.bits.y, sel_hi1 = (c & 2) ? b.bits.w : ~b.bits.w;
    const uint32_t m_in = bit == 31u ? 0xffffffffu : ((2u << bit) - 1u); /* bits 0..bit */
    const uint32_t m0 = wd ? 0xffffffffu : m_in, m1 = wd ? m_in : 0u;
    const uint32_t n = (uint32_t)(popc32(sel_lo0 & sel_hi0 & m0) + popc32(sel_lo1 & sel_hi1 & m1));
    const uint32_t base = c == 0 ? b.cnt.x : c == 1 ? b.cnt.y : c == 2 ? b.cnt.z : b.cnt.w; /* L2[c] + before block */
    return base + n;
}

/* bwt_sa: walk left until a sampled row (sa[0] == (uint32_t)-1 as in bwt_restore_sa, bwtio.c:45) */
B2_HD uint32_t sa_of_row(const FmView &f, const uint32_t *sa, uint32_t sa_intv, uint32_t k)
{
    uint32_t steps = 0;
    while (k % sa_intv != 0) {
        ++steps;
        k = inv_psi(f, k);
    }
    return steps + sa[k / sa_intv];
}

/* The SA intervals of the four one-character extensions of [k, l] (node `path` of index f):
 * nk[c] = L2[c] + occ(k-1, c) + 1, nl[c] = L2[c] + occ(l, c) (bwt.c:177-214, bwtgap.c:222-223),
 * read from the interval table while the node is inside it, from the occ blocks otherwise. */
B2_HD void children4(const FmView &f, uint32_t path, uint32_t k, uint32_t l, uint32_t nk[4], uint32_t nl[4],
                     uint32_t &n_sectors)
{
    const uint32_t d = path & 31u;
    if (d != B2_PATH_DEAD && (int)d < f.lut_k) {
        const uint32_t *p = f.lut + 2 * lut_pair(f.lut_w, (int)d + 1, (uint64_t)(path >> 5) << 2);
        if (!B2_CHECK((uint64_t)(path >> 5) < ((uint64_t)1 << (2 * d)) && lut_pair(f.lut_w, (int)d + 1, (uint64_t)(path >> 5) << 2) + 4 <= lut_total_pairs(f.lut_k),
                      CHK_LUT, d, path >> 5, f.lut_k)) { for (int c = 0; c < 4; ++c) { nk[c] = 1u; nl[c] = 0u; } n_sectors = 0; return; }
        const U8x v = ld_lut8(p, d < 11u); /* levels <= 11: 90 MB for both indexes */
        nk[0] = v.v[0]; nl[0] = v.v[1]; nk[1] = v.v[2]; nl[1] = v.v[3];
        nk[2] = v.v[4]; nl[2] = v.v[5]; nk[3] = v.v[6]; nl[3] = v.v[7];
        n_sectors = 1;
        return;
    }
    uint32_t ck[4], cl[4];
    occ2x4(f, k, l, ck, cl, n_sectors);
    for (int c = 0; c < 4; ++c) { nk[c] = ck[c] + 1u; nl[c] = cl[c]; }
}

/* Builds the four children of table node X of level `level` (level 0 = the root) into level + 1.
 * lut is the table being built (levels <= level complete); f.lut_k is ignored here, f.lut_w selects the half. */
B2_HD void lut_build_node(const FmView &f, uint32_t *lut, int level, uint64_t X)
{
    uint32_t k = 0, l = f.seq_len;
    if (level > 0) {
        const uint32_t *p = lut + 2 * lut_pair(f.lut_w, level, X);
        k = p[0];
        l = p[1];
    }
    uint32_t *o = lut + 2 * lut_pair(f.lut_w, level + 1, X << 2);
    if (k <= l) {
        uint32_t ck[4], cl[4], ns;
        occ2x4(f, k, l, ck, cl, ns);
        for (int c = 0; c < 4; ++c) { o[2 * c] = ck[c] + 1u; o[2 * c + 1] = cl[c]; }
    } else {
        for (int c = 0; c < 4; ++c) { o[2 * c] = 1u; o[2 * c + 1] = 0u; }
    }
}

/* ---------------------------------------------------------- width pass ---- */

/* One chain of bwt_cal_width (bwtaln.c:54-78): the interval, the restart counter and the chain's
 * position in the interval table.  A symbol is consumed in two halves — issue() starts the
 * memory reads, advance() uses them — so that two chains can overlap their latencies. */
struct WidthChain {
    uint32_t k, l, path;
    int bid;
    /* in flight */
    bool use_lut;
    OccBlk ba, bb;
    uint32_t qa, qb, lk, ll;
    B2_HD void reset(const FmView &f) { k = 0; l = f.seq_len; bid = 0; path = path_root(); }
    B2_HD void issue(const FmView &f, int c)
    {
        const uint32_t d = path & 31u;
        use_lut = d != B2_PATH_DEAD && (int)d < f.lut_k;
        if (use_lut) {
            const uint32_t *p = f.lut + 2 * lut_pair(f.lut_w, (int)d + 1, ((uint64_t)(path >> 5) << 2) + (uint32_t)c);
            lk = ld_q(p);
            ll = ld_q(p + 1);
        } else {
            qa = q_lower(f, k);
            qb = q_upper(f, l);
            const OccBlk *pa = f.blk + (qa >> 6), *pb = f.blk + (qb >> 6);
            ba = ld_blk(pa);
            bb = ba;
            if (pa != pb) bb = ld_blk(pb);
        }
    }
    /* finish symbol c (c > 3: ambiguous, nothing was issued); returns the width l-k+1 */
    B2_HD uint32_t advance(const FmView &f, int c)
    {
        bool alive = false;
        if (c < 4) {
            if (use_lut) {
                k = lk;
                l = ll;
            } else {
                uint32_t ck[4], cl[4];
                occ_count4(ba.cnt, ba.bits, qa & 63u, ck);
                occ_count4(bb.cnt, bb.bits, qb & 63u, cl);
                k = pick4(ck, c) + 1u;
                l = pick4(cl, c);
            }
            alive = k <= l;
            path = path_ext(path, c, f.lut_k);
        }
        if (!alive) {
            k = 0;
            l = f.seq_len;
            path = path_root();
            ++bid;
        }
        return l - k + 1u;
    }
};

/* symbol i of seq[a] as the reference stores it (bwaseqio.c:189-192):
 * seq[0] = reversed read, seq[1] = reverse-complement (plain reverse with -c) */
B2_HD int strand_sym(const uint8_t *fwd, int len, int a, int i, bool comp)
{
    int c = fwd[len - 1 - i];
    if (a == 1 && comp && c < 4) c = 3 - c;
    return c;
}

B2_HD int round_up8(int v) { return (v + 7) & ~7; }

/*
 * Width pass for one (read, strand a): fills W[0..len] (u32 widths) and the packed records
 * Q[0..len-1], using index fm[a] on seq[a] (bwtaln.c:123-130).  The seed chain
 * (bwt_cal_width on the last seed_len symbols, bwtaln.c:127-130) consumes the same symbols
 * as the tail of the main chain, so both advance in the same iteration: no scratch, and
 * their lookups overlap.  Rows are written eight words (one sector) at a time; W and Q rows
 * must be 32-byte aligned with a stride that is a multiple of 8.
 */
struct WidthOut {
    int n_amb; /* ambiguous symbols in the strand */
    int bid;   /* D(len - 1): the lower bound on the differences of the whole strand */
};
template <int QB>
B2_HD WidthOut width_pass(const FmView &f, const uint8_t *fwd, int len, int a, bool comp, int seed_len, uint32_t *W,
                          typename QF<QB>::T *Q)
{
    const bool use_seed = len > seed_len;
    const int shift = len - seed_len; /* ii = j - shift */
    int n_amb = 0;
    WidthChain m, s;
    m.reset(f);
    s.reset(f);
    uint32_t w_prev = 0, sw_prev = 0;
    int bid_prev = 0, sbid_prev = 0;
    for (int j0 = 0; j0 <= len; j0 += 8) {
        uint32_t qb[8], wb[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int t = 0; t < 8; ++t) {
            const int j = j0 + t;
            qb[t] = 0;
            wb[t] = 0;
            if (j < len) {
                const int c = strand_sym(fwd, len, a, j, comp);
                const bool seed_on = use_seed && j >= shift;
                n_amb += c > 3;
                if (c < 4) { /* start the reads of both chains before using either */
                    m.issue(f, c);
                    if (seed_on) s.issue(f, c);
                }
                const uint32_t w = m.advance(f, c);
                uint32_t sact = 0, sb2 = 0, sbp = 0, seq = 0;
                if (seed_on) {
                    const uint32_t sw = s.advance(f, c);
                    if (j > shift) { /* ii > 0 */
                        sact = 1;
                        sb2 = (uint32_t)s.bid;
                        sbp = (uint32_t)sbid_prev;
                        seq = sw == sw_prev;
                    }
                    sw_prev = sw;
                    sbid_prev = s.bid;
                }
                qb[t] = QF<QB>::pack((uint32_t)c, j > 0 && w == w_prev, seq, sact, sb2, sbp, (uint32_t)m.bid,
                                     (uint32_t)bid_prev);
                wb[t] = w;
                w_prev = w;
                bid_prev = m.bid;
            }
        }
        if (j0 < len) q_store8(Q, j0, qb);
        st8(W + j0, wb); /* includes the W[len] = 0 sentinel */
    }
    WidthOut o;
    o.n_amb = n_amb;
    o.bid = m.bid;
    return o;
}

/*
 * Work class of a read from its two strands' lower bounds D = width[len-1].bid (bwtaln.c:54-78) and its
 * max_diff: how far the search has to go correlates with min(D0, D1) — reads with an exact occurrence
 * (D = 0) all take the same short course, reads with D > max_diff are pruned at the root (bwtgap.c:155),
 * D == max_diff leaves no freedom, and in between the work grows with D.  Classes are processed from the
 * highest to the lowest: long searches start first, and the lanes of a warp work on reads of one class
 * (the north star's "bucketed by mismatch budget").  Only the ORDER of processing depends on it.
 */
#define B2_N_CLASSES 16
B2_HD int work_class(int d0, int d1, int max_diff)
{
    const int d = d0 < d1 ? d0 : d1;
    if (d > max_diff) return 0;
    if (d == 0) return 1;
    if (d == max_diff) return 2;
    return d + 2 > B2_N_CLASSES - 1 ? B2_N_CLASSES - 1 : d + 2;
}

/* gap_shadow (bwtgap.c:81-91) on the split representation (bid lives in Q, w in W), fused with the
 * refresh of the packed fields it invalidates (bid[j-1] and w[j-1]==w[j] of records 1..last_diff_pos).
 * Rows are walked one 32-byte sector at a time: W and Q rows are 32-byte aligned and padded to a
 * multiple of eight entries, so whole sectors can be read and written back. */
template <int QB>
B2_HD void shadow_update(uint32_t x, uint32_t max, int last_diff_pos, int len, uint32_t *W, typename QF<QB>::T *Q)
{
    if (last_diff_pos <= 0) return;
    const int last = last_diff_pos < len ? last_diff_pos : len - 1; /* last record whose packed fields can change */
    int j = 0;
    uint32_t prev_w = 0, prev_bid = 0;
    for (int i0 = 0; i0 <= last; i0 += 8) {
        uint32_t w8[8], q8[8];
        ld8(W + i0, w8);
        q_load8(Q, i0, q8);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int t = 0; t < 8; ++t) {
            const int i = i0 + t;
            uint32_t w = w8[t], q = q8[t];
            if (i < last_diff_pos) { /* bwtgap.c:84-89 */
                if (w > x) w -= x;
                else if (w == x) {
                    w = max - (uint32_t)(++j);
                    q = QF<QB>::with_bid(q, 1u); /* bid[i] = 1 */
                }
            }
            if (i >= 1 && i <= last) /* bid[i-1] and the equal-width flag as the search reads them */
                q = QF<QB>::with_prev(q, prev_bid, (uint32_t)(w == prev_w));
            prev_w = w;
            prev_bid = (uint32_t)QF<QB>::bid(q);
            w8[t] = w;
            q8[t] = q;
        }
        st8(W + i0, w8);
        q_store8(Q, i0, q8);
    }
}

/* --------------------------------------------------------------- stack ---- */

#define B2_NIL 0xffffffffu

#ifdef B2_DEBUG_COUNTS /* logic-test instrumentation (tests/harness only) */
extern uint64_t b2_dbg[16];
#define B2_DBG(i) (++b2_dbg[i])
#else
#define B2_DBG(i) ((void)0)
#endif

/* per-lane arena in global memory (SearchEnv::ent): 64-byte records, bump allocated (optionally with a
 * free list through w[0], REUSE) */

/* The open group of a lane — the members of the group popped last that have not been taken yet — lives
 * outside the lane's registers: words at p[w * stride] (a shared-memory column on the device, a plain
 * array in the CPU harness).  Words 0-7 the four child intervals, 8-9 the parent's interval, 10 its path. */
struct GroupStore {
    uint32_t *p;
    int stride;
    B2_HD uint32_t get(int w) const { return p[(size_t)w * stride]; }
    B2_HD void set(int w, uint32_t v) { p[(size_t)w * stride] = v; }
};
enum { OG_PK = 8, OG_PL = 9, OG_PATH = 10, OG_WORDS = 11 };
enum { GRP_NONE = 0, GRP_X = 1, GRP_G = 2, GRP_ROOT = 3 };

/* Bucket heads (top slot of each score bucket's linked list) and the set of
 * non-empty buckets.  Two storage policies; both keep the lane's scalar state
 * out of local memory (no dynamically indexed member arrays in SearchLane).
 *
 * HeadsStrided16: 16-bit heads at h[sc * stride] — the fast kernel points h at a
 *   shared-memory column (stride = threads per block, conflict free), the CPU
 *   harness at a plain array (stride 1).  A head is slot << 1 | group, so the arena holds at most
 *   32767 records (0xffff = empty).
 * HeadsWide32: 32-bit heads in global memory — the large-arena pass and exotic score
 *   ranges (up to 2048 buckets).
 * The lowest non-empty bucket is tracked by the lane like the reference does (bwtgap.c:63,73-78):
 * lowered on push, found by scanning upwards when a pop empties its bucket. */
struct HeadsStrided16 {
    uint16_t *h;
    int stride;
    static B2_HD uint32_t nil() { return 0xffffu; }
    B2_HD void clear(int nb)
    {
        for (int i = 0; i < nb; ++i) h[(size_t)i * stride] = 0xffffu;
    }
    B2_HD uint32_t get(int sc) const { return h[(size_t)sc * stride]; }
    B2_HD void set(int sc, uint32_t slot) { h[(size_t)sc * stride] = (uint16_t)slot; }
};

/* 32-bit heads at h[sc * stride]: the wide pass with its heads in shared memory (stride = threads per
 * block), or in global memory for score ranges that do not fit (stride 1) */
struct HeadsStrided32 {
    uint32_t *h;
    int stride;
    static B2_HD uint32_t nil() { return B2_NIL; }
    B2_HD void clear(int nb)
    {
        for (int i = 0; i < nb; ++i) h[(size_t)i * stride] = B2_NIL;
    }
    B2_HD uint32_t get(int sc) const { return h[(size_t)sc * stride]; }
    B2_HD void set(int sc, uint32_t slot) { h[(size_t)sc * stride] = slot; }
};

struct HeadsWide32 {
    uint32_t *h; /* [n_buckets] */
    static B2_HD uint32_t nil() { return B2_NIL; }
    B2_HD void clear(int nb)
    {
        for (int i = 0; i < nb; ++i) h[i] = B2_NIL;
    }
    B2_HD uint32_t get(int sc) const { return h[sc]; }
    B2_HD void set(int sc, uint32_t slot) { h[sc] = slot; }
};

#define B2_SAVE_WORDS 36 /* SearchLane::save_state: 33 words used, + the read and the work item of the kernel's lane */
enum LaneStatus { LANE_OK = 0, LANE_ARENA_FULL = 1, LANE_REC_FULL = 2, LANE_CHECK = 3 /* B2_CHECKED: a violation */ };

/*
 * One read's best-first search (bwt_match_gap), advanced one occ lookup per
 * step() so that the lanes of a warp stay in lock-step on the memory access.
 *
 * Exactness notes:
 *  - bucket order: lowest score first, LIFO inside a bucket, children taken in the reverse of the
 *    reference's push order (insertion, deletions 0..3, mismatches j=1..3(4), exact match last).
 *  - the exact-match child has its parent's score and is pushed last, so it is always the next pop;
 *    it is kept in registers ("held") instead of going through memory, but is still counted in
 *    n_entries so that the `n_entries > max_entries` cutoff (bwtgap.c:139) fires at the same pop.
 *  - the other children of an expansion go to memory as ONE record linked into (at most) two buckets
 *    (StackRec above).  When a group is popped it becomes the lane's open group and its members are
 *    taken one by one.  While members of a group of score B are left, B is the lowest non-empty
 *    bucket and everything pushed meanwhile scores more than B (positive penalties), so the next pop
 *    of the reference is always the next member of the open group: no other group can be opened
 *    before it is exhausted.  n_entries counts members, as the reference's stack does.
 *  - the two root entries (bwtgap.c:126-127) never touch memory: strand 1 starts as the held entry,
 *    strand 0 as a one-member open group (bucket 0 holds nothing else but held children).
 *  - last_diff_pos: inherited from the parent on non-diff pushes (the slot reuse of bwtgap.c:60),
 *    which requires positive penalties (checked by the host before launch).
 */
template <class Heads, bool REUSE, bool STATS = true, int QB = 32>
struct SearchLane {
    typedef QF<QB> Qf;
    typedef typename QF<QB>::T QT;
    /* constant per read */
    GroupStore gs;
    uint32_t lane_no; /* which arena */
    uint32_t row;     /* 2 * read: rows row (strand 0) and row + 1 (strand 1) of Q and W */
    uint32_t slab;    /* which record slab */
    int len, opt_max_diff;
    /* mutable */
    Heads bk;
    uint32_t top, free_head; /* bump pointer / free list */
    int best, n_mem;         /* lowest non-empty bucket (n_buckets when none); groups linked in memory */
    bool prefetch_next;      /* L2 prefetch of the next pop candidate (latency-bound passes) */
    int n_entries;           /* the reference's stack->n_entries (members in memory + open group + held) */
    int max_diff, best_score, best_diff, best_cnt, n_aln;
    int status;
    bool finished;
    /* current entry */
    bool have_cur, extending;
    bool qe; /* pq already holds the width record of position ci - 1: loaded at the end of the previous step, so
                that its latency runs under the rest of the warp's iteration instead of in front of the next lookup */
    uint32_t ck, cl;
    int ci, cldp, cmm, cgo, cge, cstate, ca, cscore;
    uint32_t cpath; /* position of the current entry in the interval table */
    /* open group: member mask (bits 0-3 children, bit 4 insertion / root) | kind << 5 | base << 7 | i << 16;
     * 0 = none.  Its members share score (cscore), strand (ca) and counters (cmm, cgo, cge) with the
     * entries being worked on in between, so those registers simply stay. */
    uint32_t og;
    uint32_t n_pops, n_lookups; /* instrumentation: pops and 32-byte sectors of this read */

    static B2_HD int score_of(const Params &P, int mm, int go, int ge)
    {
        return mm * P.s_mm + go * P.s_gapo + ge * P.s_gape;
    }

    /* E: the launch constants, passed by reference at every call so that on the device they stay
     * in the kernel-parameter constant bank instead of being re-loaded through a pointer */
    B2_HD QT *qrow(const SearchEnv &E, int a) const { return reinterpret_cast<QT *>(E.Q) + (size_t)(row + (uint32_t)a) * E.strideQ; }
    B2_HD StackRec *arena(const SearchEnv &E) const { return E.ent + (size_t)lane_no * E.arena_cap; }
    B2_HD Rec *records(const SearchEnv &E) const { return E.recs + (size_t)slab * E.rec_cap; }
    /* the width record of position p (0 <= p < len) of strand a */
    B2_HD QRec fetch_q(const SearchEnv &E, int a, int p) const
    {
        if (!B2_CHECK(p >= 0 && p < len && p < E.strideQ && (a == 0 || a == 1), CHK_Q_POS, row, p, len)) return 0;
        return ld_q(qrow(E, a) + p);
    }

    /* heads_clean: the caller has already emptied the bucket heads (the kernel does it with the whole warp) */
    B2_HD void begin(const SearchEnv &E, Heads heads_, GroupStore gs_, uint32_t lane_no_, uint32_t read, uint32_t slab_,
                     int len_, int max_diff_, int n_amb, bool heads_clean = false)
    {
        bk = heads_;
        gs = gs_;
        const Params *P = &E.P;
        const FmView *fm = E.fm;
        lane_no = lane_no_; row = 2u * read; slab = slab_;
        len = len_; opt_max_diff = max_diff_;
        max_diff = max_diff_;
        best_score = score_of(E.P, max_diff_ + 1, P->max_gapo + 1, P->max_gape + 1);
        best_diff = max_diff_ + 1;
        best_cnt = 0; n_aln = 0; status = LANE_OK;
        finished = false; have_cur = false; extending = false; qe = false;
        og = 0;
        pq = 0;
        top = 0; free_head = B2_NIL; n_entries = 0;
        best = P->n_buckets; n_mem = 0;
        prefetch_next = E.prefetch_next != 0;
        n_pops = n_lookups = 0;
        if (n_amb > max_diff_) { finished = true; return; } /* bwtgap.c:117-122 */
        if (!heads_clean) bk.clear(P->n_buckets);
        /* roots: strand 0 then strand 1 (bwtgap.c:126-127) -> strand 1 pops first */
        n_entries = 2;
        og = 16u | (uint32_t)GRP_ROOT << 5;
        ck = 0; cl = fm[0].seq_len; ci = len; cldp = 0; cmm = cgo = cge = 0; cstate = ST_M; ca = 1; cscore = 0;
        cpath = path_root();
        have_cur = true;
    }

    /* The children of the expansion of the current entry (position i, intervals nk4/nl4) that go to
     * memory: gap group `gmask` into bucket sg, mismatch group `xmask` into bucket sx. */
    B2_HD void push_groups(const SearchEnv &E, int i, int base, uint32_t gmask, int sg, uint32_t xmask, int sx,
                           const uint32_t nk4[4], const uint32_t nl4[4])
    {
        StackRec *ent = arena(E);
        uint32_t slot;
        if (REUSE && free_head != B2_NIL) {
            slot = free_head;
            free_head = ent[slot].w[0];
        } else {
            if (top >= E.arena_cap) { status = LANE_ARENA_FULL; finished = true; return; }
            slot = top++;
        }
        if (!B2_CHECK(slot < E.arena_cap && slot < top, CHK_ARENA_SLOT, row, slot, top) ||
            !B2_CHECK((!gmask || (sg >= 0 && sg < E.P.n_buckets)) && (!xmask || (sx >= 0 && sx < E.P.n_buckets)), CHK_BUCKET, row,
                      gmask ? sg : sx, E.P.n_buckets)) { status = LANE_CHECK; finished = true; return; }
        uint32_t h[8];
        h[0] = h[1] = Heads::nil();
        if (gmask) { /* gap group first: when both share a bucket the mismatches are on top (pushed later) */
            h[0] = bk.get(sg);
            bk.set(sg, slot << 1);
            best = sg < best ? sg : best;
            ++n_mem;
        }
        if (xmask) {
            h[1] = bk.get(sx);
            bk.set(sx, slot << 1 | 1u);
            best = sx < best ? sx : best;
            ++n_mem;
        }
        h[2] = cpath;
        h[3] = (uint32_t)i | (uint32_t)base << 16 | (uint32_t)cstate << 19 | (uint32_t)ca << 21 | gmask << 22 | xmask << 27;
        h[4] = (uint32_t)cmm | (uint32_t)cgo << 8 | (uint32_t)cge << 16;
        h[5] = ck; h[6] = cl; h[7] = 0;
        uint32_t *rec = ent[slot].w;
        st8rec(rec, h);
        if ((gmask & 15u) | xmask) { /* a lone insertion does not need the children's sector */
            uint32_t c8[8] = {nk4[0], nl4[0], nk4[1], nl4[1], nk4[2], nl4[2], nk4[3], nl4[3]};
            st8rec(rec + 8, c8);
        }
        n_entries += popc32(gmask) + popc32(xmask);
    }

    /* take the top group of the lowest bucket out of memory: it becomes the open group */
    B2_HD void pop_group(const SearchEnv &E)
    {
        B2_DBG(0);
        const int b = best;
        if (!B2_CHECK(b >= 0 && b < E.P.n_buckets && n_mem > 0, CHK_BUCKET, row, b, n_mem)) { status = LANE_CHECK; finished = true; og = 0; return; }
        const uint32_t ref = bk.get(b);
        const uint32_t slot = ref >> 1, which = ref & 1u;
        if (!B2_CHECK(ref != Heads::nil(), CHK_POP_EMPTY, row, b, ref) || !B2_CHECK(slot < top, CHK_ARENA_SLOT, row, slot, top)) {
            status = LANE_CHECK; finished = true; og = 0; return;
        }
        uint32_t h[8], c8[8];
        StackRec *ent = arena(E);
        const uint32_t *rec = ent[slot].w;
        ld8cg(rec, h);
        ld8cg(rec + 8, c8);
        const uint32_t prev = which ? h[1] : h[0];
        bk.set(b, prev);
        --n_mem;
        uint32_t next_top = prev;
        if (prev == Heads::nil()) { /* bucket emptied: next non-empty one upwards (bwtgap.c:73-78) */
            if (n_mem == 0) best = E.P.n_buckets;
            else {
                int nb = b + 1;
                while ((next_top = bk.get(nb)) == Heads::nil()) {
                    ++nb;
                    if (!B2_CHECK(nb < E.P.n_buckets, CHK_BUCKET, row, nb, n_mem)) { status = LANE_CHECK; finished = true; og = 0; return; }
                }
                best = nb;
            }
        }
        /* the record that will most likely be popped next: start bringing it into L2 while this
         * group is worked on (a hint only; pushes may still overtake it) */
        if (prefetch_next && n_mem > 0) prefetch_l2(ent + (next_top >> 1));
        const uint32_t info = h[3];
        const uint32_t gmask = info >> 22 & 31u, xmask = info >> 27 & 15u;
        if (REUSE) { /* the slot is free once both of its groups have been taken */
            if ((which ? gmask : xmask) == 0 || h[7] != 0) {
                ent[slot].w[0] = free_head;
                free_head = slot;
            } else ent[slot].w[7] = 1u;
        }
        const int pstate = (int)(info >> 19 & 3u);
        ca = (int)(info >> 21 & 1u);
        cscore = b;
        cmm = (int)(h[4] & 255u); cgo = (int)(h[4] >> 8 & 255u); cge = (int)(h[4] >> 16 & 255u);
        uint32_t mask;
        if (which) { ++cmm; mask = xmask; }
        else {
            if (pstate == ST_M) ++cgo; else ++cge;
            mask = gmask;
        }
        og = mask | (which ? (uint32_t)GRP_X : (uint32_t)GRP_G) << 5 | (info >> 16 & 7u) << 7 | (info & 0xffffu) << 16;
        for (int w = 0; w < 8; ++w) gs.set(w, c8[w]);
        gs.set(OG_PK, h[5]);
        gs.set(OG_PL, h[6]);
        gs.set(OG_PATH, h[2]);
    }

    /*
     * Next member of the open group, in the reverse of the reference's push order, becomes the current
     * entry (with pq = the width record of its position).  The mismatch members of a group share position
     * and counters, and so do its deletion members: the pruning tests of bwtgap.c:146-158 (m < 0,
     * m < width[i-1].bid) give the same answer for all of them.  When that answer is "skip", the whole run
     * is dropped at once — the reference would pop them one after the other and `continue` each time;
     * n_entries only falls meanwhile, so its cutoff test cannot fire in between.  Returns false when
     * nothing was taken (the members dropped are counted in n_entries / n_pops).
     */
    B2_HD bool take_member(const SearchEnv &E)
    {
        const Params *P = &E.P;
        const int K = E.fm[0].lut_k;
        const uint32_t kind = og >> 5 & 3u, mask = og & 31u;
        const int i = (int)(og >> 16), base = (int)(og >> 7 & 7u);
        if (kind == GRP_ROOT) { /* the strand-0 root */
            og = 0;
            ck = 0; cl = E.fm[0].seq_len; ci = len; cldp = 0; cmm = cgo = cge = 0; cstate = ST_M; ca = 0; cscore = 0;
            cpath = path_root();
            if (len > 0) pq = fetch_q(E, 0, len - 1);
            --n_entries;
            if (STATS) ++n_pops;
            return true;
        }
        if (!B2_CHECK(mask != 0 && i <= len && (kind == GRP_X || kind == GRP_G) && n_entries > 0, CHK_MEMBER, row, og, n_entries)) {
            status = LANE_CHECK; finished = true; og = 0; return false;
        }
        const uint32_t run = mask & 15u; /* members that are child intervals: mismatches, or deletions */
        const int m = max_diff - cmm - cgo - ((P->mode & MODE_GAPE) ? cge : 0);
        if (run) {
            const int pos = kind == GRP_X ? i : i + 1; /* their position; pos > 0 for deletions */
            if (pos > 0) pq = fetch_q(E, ca, pos - 1);
            const bool stop = !(P->mode & MODE_NONSTOP) && cscore > best_score + P->s_mm; /* bwtgap.c:143: the first one ends the search */
            if (!stop && (m < 0 || (pos > 0 && m < Qf::bid(pq)))) {
                const int n = popc32(run);
                n_entries -= n;
                if (STATS) n_pops += (uint32_t)n;
                og &= ~15u;
                if (!(og & 31u)) og = 0;
                B2_DBG(9);
                return false;
            }
            int c;
            if (kind == GRP_X) { /* pushed j = 1..3 (then j = 4 for an ambiguous base): (base + j) & 3 */
                c = base & 3;
                if (!(base > 3 && (mask >> c & 1u))) {
                    c = (base + 3) & 3;
                    if (!(mask >> c & 1u)) {
                        c = (base + 2) & 3;
                        if (!(mask >> c & 1u)) c = (base + 1) & 3;
                    }
                }
                ci = i; cldp = i; cstate = ST_M;
            } else { /* pushed insertion, deletions 0..3 */
                c = (mask & 8u) ? 3 : (mask & 4u) ? 2 : (mask & 2u) ? 1 : 0;
                ci = i + 1; cldp = i + 1; cstate = ST_D;
            }
            og &= ~(1u << c);
            if (!(og & 31u)) og = 0;
            ck = gs.get(2 * c); cl = gs.get(2 * c + 1);
            cpath = path_ext(gs.get(OG_PATH), c, K);
        } else { /* insertion / insertion extension: the parent's interval, one read base consumed */
            og = 0;
            ck = gs.get(OG_PK); cl = gs.get(OG_PL);
            cpath = gs.get(OG_PATH);
            ci = i; cldp = i; cstate = ST_I;
            if (i > 0) pq = fetch_q(E, ca, i - 1);
        }
        --n_entries;
        if (STATS) ++n_pops;
        return true;
    }

    /* hit bookkeeping, bwtgap.c:165-198; returns false when the search must stop */
    B2_HD bool on_hit(const SearchEnv &E)
    {
        const Params *P = &E.P;
        const FmView &f = E.fm[1 - ca];
        const bool gape_mode = P->mode & MODE_GAPE;
        if (n_aln == 0) {
            best_score = cscore;
            best_diff = cmm + cgo + (gape_mode ? cge : 0);
            if (!(P->mode & MODE_NONSTOP)) max_diff = best_diff + 1 > opt_max_diff ? opt_max_diff : best_diff + 1;
        }
        if (cscore == best_score) best_cnt = (int)((uint32_t)best_cnt + (cl - ck + 1u));
        else if (best_cnt > P->max_top2) return false;
        bool add = true;
        Rec *recs = records(E);
        if (cgo)
            for (int j = 0; j < n_aln; ++j)
                if (recs[j].k == ck && recs[j].l == cl) { add = false; break; }
        if (add) {
            shadow_update<QB>(cl - ck + 1u, f.seq_len, cldp, len, E.W + (size_t)(row + (uint32_t)ca) * E.strideW, qrow(E, ca));
            if (n_aln >= E.rec_cap) { status = LANE_REC_FULL; return false; }
            if (!B2_CHECK(n_aln >= 0 && cldp >= 0 && cldp <= len, CHK_REC_FILL, row, n_aln, cldp)) { status = LANE_CHECK; return false; }
            Rec r;
            r.packed = (uint32_t)cmm | (uint32_t)cgo << 8 | (uint32_t)cge << 16 | (uint32_t)ca << 24;
            r.k = ck; r.l = cl; r.score = cscore;
            recs[n_aln++] = r;
        }
        return true;
    }

    /*
     * One step = prepare() -> one occ lookup -> apply().  The three phases are separate so that
     * the kernel can re-converge the warp between them (all lanes issue their lookup loads
     * together); step() chains them for callers that do not care (CPU logic tests).
     */
    enum { NONE = -1, EXPAND = 0, EXTEND = 1, JUMP = 2 };
    QRec pq; /* width record of the position being worked on (prepare -> apply); JUMP: the target node's k-mer */
    int pm;  /* differences still allowed for the current entry; JUMP: the number of symbols it consumes */

    /*
     * bwt_match_exact_alt (bwt.c:235-250) only reports the final interval, or failure at the first ambiguous
     * symbol / empty interval.  While the chain is inside the interval table, the node reached after the next
     * n symbols is ONE table entry away (descendants of an empty node are stored empty), so n steps of the
     * chain cost one request and one warp iteration instead of n.  Sets up the jump (pq = k-mer of the target
     * node, pm = n) and returns JUMP; returns EXTEND when the chain is outside the table or only one level
     * is left (the ordinary step reads the children's sector then); returns NONE after ending a chain that
     * runs into an ambiguous symbol (bwt.c:241: the reference returns 0 there, whatever came before).
     */
    B2_HD int extend_mode(const SearchEnv &E)
    {
        const int K = E.fm[0].lut_k;
        const uint32_t d = cpath & 31u;
        int n = K - (int)d; /* B2_PATH_DEAD = 31 >= any K: n <= 0 */
        if (n > ci) n = ci;
        if (n < 2) {
            if (!qe) pq = fetch_q(E, ca, ci - 1);
            qe = false;
            return EXTEND;
        }
        qe = false;
        uint32_t X = cpath >> 5, amb = 0;
        for (int j = 1; j <= n; ++j) {
            const uint32_t b = (uint32_t)q_base(fetch_q(E, ca, ci - j));
            amb |= b >> 2;
            X = X << 2 | (b & 3u);
        }
        if (amb) { extending = false; return NONE; }
        pq = X;
        pm = n;
        return JUMP;
    }

    /* true when the next entry does not have to come from memory */
    B2_HD bool ready() const { return have_cur || extending || og != 0; }

    /*
     * Suspension (b200aln.cu: stragglers of a draining launch are parked and resumed in dense warps): the
     * lane's whole state between two steps, as B2_SAVE_WORDS words.  Everything else a search owns lives in
     * memory already (arena, width rows, record slab) and stays where it is — lane_no keeps naming the arena —
     * except the bucket heads and the open group, which the caller saves next to these words and puts back
     * (bk, gs are re-pointed by the caller after load_state).
     */
    B2_HD void save_state(uint32_t *d) const
    {
        d[0] = lane_no; d[1] = row; d[2] = slab; d[3] = (uint32_t)len; d[4] = (uint32_t)opt_max_diff;
        d[5] = top; d[6] = free_head; d[7] = (uint32_t)best; d[8] = (uint32_t)n_mem;
        d[9] = (uint32_t)prefetch_next | (uint32_t)finished << 1 | (uint32_t)have_cur << 2 | (uint32_t)extending << 3 | (uint32_t)qe << 4;
        d[10] = (uint32_t)n_entries; d[11] = (uint32_t)max_diff; d[12] = (uint32_t)best_score; d[13] = (uint32_t)best_diff;
        d[14] = (uint32_t)best_cnt; d[15] = (uint32_t)n_aln; d[16] = (uint32_t)status;
        d[17] = ck; d[18] = cl; d[19] = (uint32_t)ci; d[20] = (uint32_t)cldp; d[21] = (uint32_t)cmm; d[22] = (uint32_t)cgo;
        d[23] = (uint32_t)cge; d[24] = (uint32_t)cstate; d[25] = (uint32_t)ca; d[26] = (uint32_t)cscore; d[27] = cpath;
        d[28] = og; d[29] = n_pops; d[30] = n_lookups; d[31] = pq; d[32] = (uint32_t)pm;
    }
    B2_HD void load_state(const uint32_t *d)
    {
        lane_no = d[0]; row = d[1]; slab = d[2]; len = (int)d[3]; opt_max_diff = (int)d[4];
        top = d[5]; free_head = d[6]; best = (int)d[7]; n_mem = (int)d[8];
        prefetch_next = d[9] & 1u; finished = d[9] >> 1 & 1u; have_cur = d[9] >> 2 & 1u; extending = d[9] >> 3 & 1u; qe = d[9] >> 4 & 1u;
        n_entries = (int)d[10]; max_diff = (int)d[11]; best_score = (int)d[12]; best_diff = (int)d[13];
        best_cnt = (int)d[14]; n_aln = (int)d[15]; status = (int)d[16];
        ck = d[17]; cl = d[18]; ci = (int)d[19]; cldp = (int)d[20]; cmm = (int)d[21]; cgo = (int)d[22];
        cge = (int)d[23]; cstate = (int)d[24]; ca = (int)d[25]; cscore = (int)d[26]; cpath = d[27];
        og = d[28]; n_pops = d[29]; n_lookups = d[30]; pq = d[31]; pm = (int)d[32];
    }

    /* max_rounds: how many pruned pops (a run of group members, or a single entry failing the width bound) the
     * lane may go through before it gives the warp's iteration back without a lookup — the other lanes do not
     * wait while one lane prunes its way through a bucket (measured: 1 is best).
     * pop until something needs a lookup; returns its kind, or NONE when the search ended
     * (finished is set) or when the next entry has to come from memory and allow_pop is false:
     * the kernel lets the lanes of a warp take their memory pops in batches, so that the extra
     * dependent load and the pop code are paid once for several lanes instead of every iteration. */
    B2_HD int prepare(const SearchEnv &E, bool allow_pop = true, int max_rounds = 1 << 30)
    {
        const Params *P = &E.P;
        const bool gape_mode = P->mode & MODE_GAPE, nonstop = P->mode & MODE_NONSTOP;
        int round = 0;
        for (;;) {
            if (extending) {
                const int em = extend_mode(E);
                if (em != NONE) return em;
                continue; /* the chain met an ambiguous symbol */
            }
            if (n_entries == 0) { finished = true; return NONE; }
            if (!B2_CHECK(n_entries > 0, CHK_ENTRIES, row, n_entries, n_mem)) { status = LANE_CHECK; finished = true; return NONE; }
            if (n_entries > P->max_entries) { finished = true; return NONE; } /* bwtgap.c:139 */
            if (!have_cur) {
                if (og == 0) {
                    if (!allow_pop) return NONE;
                    pop_group(E);
#ifdef B2_CHECKED
                    if (finished) return NONE;
#endif
                }
                if (!take_member(E)) { /* a whole run pruned */
                    if (++round >= max_rounds) return NONE; /* the warp moves on; this lane goes on in the next iteration */
                    continue;
                }
            } else { /* held exact child: same accounting as a push followed by a pop */
                --n_entries;
                if (STATS) ++n_pops;
                have_cur = false;
                if (ci > 0 && !qe) pq = fetch_q(E, ca, ci - 1);
                qe = false;
            }
            if (!nonstop && cscore > best_score + P->s_mm) { finished = true; return NONE; }
            pm = max_diff - cmm - cgo - (gape_mode ? cge : 0);
            if (pm < 0) { B2_DBG(8); continue; }
            if (ci > 0 && pm < Qf::bid(pq)) { /* pq = record of position ci - 1 */
                if (++round >= max_rounds) return NONE;
                continue;
            }
            if (ci == 0) {
                if (!on_hit(E)) { finished = true; return NONE; }
                continue;
            }
            if (pm == 0 && (cstate == ST_M || gape_mode || cge == P->max_gape)) {
                B2_DBG(10);
                extending = true;
                const int em = extend_mode(E);
                if (em != NONE) return em;
                continue;
            }
            return EXPAND;
        }
    }

    /* the memory access of a step: the children of the current entry, or the table entry a JUMP lands on
     * (returned in nk4[0], nl4[0]) */
    B2_HD void lookup(const SearchEnv &E, int mode, uint32_t nk4[4], uint32_t nl4[4], uint32_t &ns) const
    {
        const FmView &f = E.fm[1 - ca];
        if (mode == JUMP) {
            if (!B2_CHECK((int)(cpath & 31u) + pm <= f.lut_k && (uint64_t)pq < ((uint64_t)1 << (2 * ((cpath & 31u) + (uint32_t)pm))), CHK_LUT, row,
                          (cpath & 31u) + (uint32_t)pm, pq)) { nk4[0] = 1u; nl4[0] = 0u; ns = 0; return; }
            ld_lut_pair(f.lut + 2 * lut_pair(f.lut_w, (int)(cpath & 31u) + pm, (uint64_t)pq), nk4[0], nl4[0]);
            ns = 1;
            return;
        }
        children4(f, cpath, ck, cl, nk4, nl4, ns);
    }

    /* consume the children intervals (children4) of the current entry [ck, cl] on fm[1 - ca] */
    B2_HD void apply(const SearchEnv &E, int mode, const uint32_t nk4[4], const uint32_t nl4[4], uint32_t ns)
    {
        const Params *P = &E.P;
        const int K = E.fm[0].lut_k; /* same for both indexes */
        const bool gape_mode = P->mode & MODE_GAPE;
        if (STATS) n_lookups += ns;
        if (mode == JUMP) { /* pm steps of bwt_match_exact_alt at once */
            B2_DBG(7);
            if (nk4[0] > nl4[0]) { extending = false; return; }
            const uint32_t d = (cpath & 31u) + (uint32_t)pm;
            ck = nk4[0];
            cl = nl4[0];
            cpath = (int)d >= K ? B2_PATH_DEAD : (d | pq << 5);
            ci -= pm;
            if (ci == 0) {
                extending = false;
                if (!on_hit(E)) finished = true;
            }
            return;
        }
        const QRec q = pq;
        const int m = pm;

        const int i = ci - 1;
        const int base = q_base(q);
        if (mode == EXTEND) { /* one step of bwt_match_exact_alt (bwt.c:235-250) */
            B2_DBG(5);
            if ((cpath & 31u) != B2_PATH_DEAD) B2_DBG(11);
            if (base > 3) { extending = false; return; }
            ck = pick4(nk4, base);
            cl = pick4(nl4, base);
            if (ck > cl) { extending = false; return; }
            cpath = path_ext(cpath, base, K);
            ci = i;
            if (ci == 0) {
                extending = false;
                if (!on_hit(E)) finished = true;
            }
#ifdef B2_EARLY_Q
            else if ((cpath & 31u) == B2_PATH_DEAD) { pq = fetch_q(E, ca, ci - 1); qe = true; } /* next chain step: outside the table */
#endif
            return;
        }

        B2_DBG(6);
        if ((cpath & 31u) != B2_PATH_DEAD) B2_DBG(13);
        const uint32_t occ = cl - ck + 1u;
        bool allow_diff = true, allow_M = true;
        if (i > 0) { /* bwtgap.c:205-214, written without short-circuits to keep the lanes together */
            const int bp = Qf::bidp(q), bd = Qf::bid(q);
            const bool d1 = bp > m - 1;
            const bool e1 = (bp == m - 1) & (bd == m - 1) & (q_eq(q) != 0);
            const int m_seed = P->max_seed_diff - cmm - cgo - (gape_mode ? cge : 0);
            const int sp = Qf::sbidp(q), sd = Qf::sbid(q);
            const bool sa = q_sact(q) != 0;
            const bool d2 = sa & (sp > m_seed - 1);
            const bool e2 = sa & (sp == m_seed - 1) & (sd == m_seed - 1) & (q_seq(q) != 0);
            allow_diff = !(d1 | d2);
            allow_M = !((!d1 & e1) | (!d2 & e2));
        }
        int gaps;
        if (P->mode & MODE_LOGGAP) {
            uint32_t v = (uint32_t)(cge + cgo);
            int lg = 0;
            while (v > 1u) { v >>= 1; ++lg; }
            gaps = lg / 2 + 1;
        } else gaps = cgo + cge;

        uint32_t live = 0; /* which of the four children exist */
        for (int j = 0; j < 4; ++j) live |= (nk4[j] <= nl4[j] ? 1u : 0u) << j;
        uint32_t gmask = 0, xmask = 0;
        int sg = 0;
        if (allow_diff && i >= P->indel_end_skip + gaps && len - i >= P->indel_end_skip + gaps) { /* bwtgap.c:217-243 */
            if (cstate == ST_M) {
                if (cgo < P->max_gapo) { gmask = 16u | live; sg = cscore + P->s_gapo; B2_DBG(1); }
            } else if (cstate == ST_I) {
                if (cge < P->max_gape) { gmask = 16u; sg = cscore + P->s_gape; }
            } else {
                if (cge < P->max_gape && (cge + cgo < max_diff || occ < (uint32_t)P->max_del_occ)) {
                    gmask = live;
                    sg = cscore + P->s_gape;
                }
            }
        }
        bool child = base < 4;
        if (allow_diff && allow_M) /* bwtgap.c:245-253; for an ambiguous base the j == 4 child is a mismatch too */
            xmask = base > 3 ? live : live & ~(1u << base);
        if (gmask | xmask) {
            push_groups(E, i, base, gmask, sg, xmask, cscore + P->s_mm, nk4, nl4);
            if (finished) return; /* arena overflow */
        }

        if (child) { /* exact-match child: held in registers, counted like a push */
            uint32_t nk = pick4(nk4, base), nl = pick4(nl4, base);
            if (nk <= nl) {
                ck = nk; cl = nl; ci = i; cstate = ST_M; /* counters, score, ldp, strand inherited */
                cpath = path_ext(cpath, base, K);
                have_cur = true;
                ++n_entries;
#ifdef B2_EARLY_Q
                if (ci > 0) { pq = fetch_q(E, ca, ci - 1); qe = true; } /* the next step's width record, early */
#endif
            }
        }
    }

    B2_HD void step(const SearchEnv &E, int max_rounds = 1 << 30)
    {
        const int mode = prepare(E, true, max_rounds);
        if (mode == NONE) return;
        uint32_t nk4[4], nl4[4], ns;
        lookup(E, mode, nk4, nl4, ns);
        apply(E, mode, nk4, nl4, ns);
    }
};

} // namespace b2
