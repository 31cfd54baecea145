/*
 * b200aln.cu — CUDA kernels (sm_100a) and the C ABI of include/b200aln.h.
 *
 * Data flow of one batch (replaces bwa_cal_sa_reg_gap, bwtaln.c:80-140):
 *   H2D (lens, offs, codes)                         pinned/pageable host -> HBM
 *   k_width   : one thread per (read, strand)       bwt_cal_width x4  -> W, Q
 *   k_search  : persistent lanes, one read per lane bwt_match_gap     -> record slabs
 *   k_search (large arena) for flagged reads        same state machine, free-list arena
 *   k_scan*   : exclusive sum of n_aln              -> record offsets
 *   k_compact : slabs -> packed records in read order
 *   D2H (n_aln, records)
 *
 * The per-read state machines are in aln_core.cuh (shared with the CPU logic
 * tests); this file owns memory, launches and warp-level work distribution.
 * There is no CPU fallback anywhere in this library.
 */
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200aln.h"
#include "aln_core.cuh"
#include "alngrp_core.cuh"
#include "fm_layout.cuh"
#include "host_params.h"

using namespace b2;

#define CK(call)                                                                                               \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess) {                                                                               \
            fprintf(stderr, "[b200aln] CUDA error %s at %s:%d (%s). Abort!\n", cudaGetErrorString(e_), __FILE__, \
                    __LINE__, #call);                                                                          \
            abort();                                                                                           \
        }                                                                                                      \
    } while (0)

static_assert(sizeof(b200aln_opt_t) == 64, "gap_opt_t is 64 bytes (bwtaln.h:105-115)");
static_assert(sizeof(b200aln_rec_t) == 16 && sizeof(Rec) == 16, "bwt_aln1_t is 16 bytes (bwtaln.h:34-38)");

/* ------------------------------------------------------------ kernels ---- */

__global__ void k_convert_index(RefBwt r, OccBlk *out, uint64_t nb)
{
    for (uint64_t b = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b < nb; b += (uint64_t)gridDim.x * blockDim.x)
        out[b] = fm_convert_block(r, b);
}

__global__ void k_lut_build(FmView f, uint32_t *lut, int level, uint64_t n_nodes)
{ /* children of every node of `level` -> level + 1 of the interval table */
    for (uint64_t X = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; X < n_nodes; X += (uint64_t)gridDim.x * blockDim.x)
        lut_build_node(f, lut, level, X);
}

__global__ void __launch_bounds__(256) k_bwt_sa(FmView f, const uint32_t *sa, uint32_t sa_intv, int64_t n,
                                                const uint32_t *rows, uint32_t *out)
{ /* bwt_sa (bwt.c:69-79) for a batch of rows: one row per thread, ~sa_intv/2 dependent sector reads each */
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = sa_of_row(f, sa, sa_intv, rows[i]);
}

__global__ void __launch_bounds__(256) k_sa2seq(FmView f0, const uint32_t *sa0, FmView f1, const uint32_t *sa1,
                                                uint32_t sa_intv0, uint32_t sa_intv1, int64_t n,
                                                const uint8_t *strand, const uint32_t *rows, const int32_t *lens,
                                                uint64_t *out)
{ /* bwtdb_sa2seq with offset 0 (dbset.c:240-245) */
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (strand[i]) out[i] = (uint64_t)sa_of_row(f0, sa0, sa_intv0, rows[i]);
        else out[i] = (uint64_t)(uint32_t)(f1.seq_len - (sa_of_row(f1, sa1, sa_intv1, rows[i]) + (uint32_t)lens[i]));
    }
}

struct WidthArgs {
    FmView fm[2];
    int n_reads;               /* number of work items */
    const int32_t *work_list;  /* null: work item w is read w */
    const int32_t *lens;
    const int64_t *offs;
    const uint8_t *codes;
    int comp, seed_len, strideQ, strideW;
    uint32_t *Q; /* QF<QB>::T records */
    uint32_t *W;
    int32_t *n_amb;
    uint8_t *dkey; /* [2 * read + strand]: the strand's lower bound on differences, saturated (null: not wanted) */
    int rows_by_work; /* rows of Q / W are numbered by work item (the re-run passes) instead of by read */
};

template <int QB>
__global__ void __launch_bounds__(128, 5) k_width(const __grid_constant__ WidthArgs A)
{
    typedef typename QF<QB>::T QT;
    const int nthreads = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t t = tid; t < 2 * (int64_t)A.n_reads; t += nthreads) {
        const int wi = (int)(t >> 1), a = (int)(t & 1);
        const int r = A.work_list ? A.work_list[wi] : wi;
        const size_t slot = (size_t)2 * (A.rows_by_work ? wi : r) + a;
        const WidthOut o = width_pass<QB>(A.fm[a], A.codes + A.offs[r], A.lens[r], a, A.comp != 0, A.seed_len,
                                          A.W + slot * A.strideW, reinterpret_cast<QT *>(A.Q) + slot * A.strideQ);
        if (a == 0) A.n_amb[r] = o.n_amb;
        if (A.dkey) A.dkey[(size_t)2 * r + a] = (uint8_t)(o.bid > 255 ? 255 : o.bid);
    }
}

/* Work order of the fast pass: a counting sort of the reads by work class (aln_core.cuh: work_class), highest
 * class first.  k_class_count: histogram; k_class_scatter: every block reserves its share of each class's
 * range and places its reads there (the order inside a class is whatever the atomics give: results do not
 * depend on the order reads are searched in). */
struct OrderArgs {
    int n_reads;
    const int32_t *lens;
    const int32_t *md;
    const uint8_t *dkey;
    unsigned int *cnt;  /* [B2_N_CLASSES] reads per class */
    unsigned int *fill; /* [B2_N_CLASSES] slots handed out so far */
    int32_t *order;
};
__device__ __forceinline__ int class_of_read(const OrderArgs &A, int r)
{
    return work_class(A.dkey[2 * (size_t)r], A.dkey[2 * (size_t)r + 1], A.md[A.lens[r]]);
}
__global__ void __launch_bounds__(256) k_class_count(const __grid_constant__ OrderArgs A)
{
    __shared__ unsigned int h[B2_N_CLASSES];
    if (threadIdx.x < B2_N_CLASSES) h[threadIdx.x] = 0;
    __syncthreads();
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < A.n_reads; r += gridDim.x * blockDim.x)
        atomicAdd(&h[class_of_read(A, r)], 1u);
    __syncthreads();
    if (threadIdx.x < B2_N_CLASSES && h[threadIdx.x]) atomicAdd(&A.cnt[threadIdx.x], h[threadIdx.x]);
}
__global__ void __launch_bounds__(256) k_class_scatter(const __grid_constant__ OrderArgs A)
{
    __shared__ unsigned int h[B2_N_CLASSES], base[B2_N_CLASSES];
    const int per = (A.n_reads + gridDim.x - 1) / gridDim.x; /* a contiguous slice per block */
    const int lo = blockIdx.x * per, hi = lo + per < A.n_reads ? lo + per : A.n_reads;
    if (threadIdx.x < B2_N_CLASSES) h[threadIdx.x] = 0;
    __syncthreads();
    for (int r = lo + threadIdx.x; r < hi; r += blockDim.x) atomicAdd(&h[class_of_read(A, r)], 1u);
    __syncthreads();
    if (threadIdx.x < B2_N_CLASSES) {
        unsigned int start = 0; /* classes above this one come first */
        for (int c = B2_N_CLASSES - 1; c > (int)threadIdx.x; --c) start += A.cnt[c];
        base[threadIdx.x] = start + (h[threadIdx.x] ? atomicAdd(&A.fill[threadIdx.x], h[threadIdx.x]) : 0u);
        h[threadIdx.x] = 0;
    }
    __syncthreads();
    for (int r = lo + threadIdx.x; r < hi; r += blockDim.x) {
        const int c = class_of_read(A, r);
        A.order[base[c] + atomicAdd(&h[c], 1u)] = r;
    }
}

#define WIDE_TAG 0x40000000

struct SearchArgs {
    SearchEnv env;
    int n_work;
    const int32_t *work_list; /* null: work item w is read w */
    const int32_t *lens;
    const int32_t *n_amb;
    const int32_t *md; /* max_diff by read length */
    int recs_by_work; /* slab index: work item (large pass) or read (fast pass) */
    int rows_by_work; /* rows of Q / W: by work item (the re-run passes) or by read */
    int32_t *n_aln;
    int32_t *over_slot; /* re-run passes: over_slot[r] = work item | slot_tag */
    int32_t slot_tag;   /* 0 = middle pass, WIDE_TAG = wide pass */
    unsigned int *counter;
    unsigned int *n_over;
    int32_t *over_list;
    unsigned long long *stat; /* [0] pops, [1] sectors */
    uint32_t *heads_wide;     /* HeadsWide32 only: per lane n_buckets heads + mask words */
    int heads_wide_stride;
    int arena_by_work;        /* arenas are indexed by work item (passes with fewer work items than lanes) instead of by lane */
    unsigned int *n_rec_full; /* counts the reads that ran out of record slab (null: not wanted) */
    int pop_batch; /* lanes of a warp that must wait for a memory pop before the warp takes them */
    int prep_rounds; /* pops / prunes a lane may go through per warp iteration before the warp moves on */
    /* Parking: once the work queue is dry, a warp in which at most susp_thresh lanes are still
     * searching writes their state out (SearchLane::save_state + bucket heads + open group, SUSP_STRIDE words per
     * lane) into a ring of tickets and then tries to take 32 parked searches back out of it: stragglers keep
     * running in full warps, sparse warps leave the SM.  What a launch leaves in the ring (fewer than 32 at the
     * time the last warps looked) is picked up by a following launch with resume_base set.  A launch's
     * stragglers then cost the slots of a few dense warps instead of one sparse warp each. */
    int susp_thresh;
    uint32_t *park_buf;        /* [PARK_CAP][SUSP_STRIDE] */
    uint32_t *park_flag;       /* [PARK_CAP]: ticket + 1 once the slot holds that ticket's search */
    unsigned int *park_ring;   /* [0] tail: tickets handed to parking lanes, [1] head: tickets taken back, [2] error flag */
    int resume;                /* 1: work item w of this launch is the parked search with ticket resume_base + w */
    unsigned int resume_base;
};
#define PARK_CAP (1u << 17)    /* ring slots; parked searches never outnumber the lanes of the launch (113 664) */
#define SUSP_GROUP_AT B2_SAVE_WORDS             /* open group words */
#define SUSP_HEADS_AT 48                        /* bucket heads, two 16-bit heads per word */
#define SUSP_STRIDE (SUSP_HEADS_AT + 80)        /* words per parked lane (n_buckets <= 160: fast_heads_ok) */
static_assert(B2_SAVE_WORDS + OG_WORDS <= SUSP_HEADS_AT, "parked-lane layout");

template <class Heads> struct HeadsFactory;
template <> struct HeadsFactory<HeadsStrided16> {
    static __device__ __forceinline__ HeadsStrided16 make(const SearchArgs &, size_t)
    {
        extern __shared__ uint16_t sm_heads[]; /* [n_buckets][blockDim.x] */
        HeadsStrided16 hd;
        hd.h = sm_heads + threadIdx.x;
        hd.stride = (int)blockDim.x;
        return hd;
    }
};
template <> struct HeadsFactory<HeadsStrided32> {
    static __device__ __forceinline__ HeadsStrided32 make(const SearchArgs &, size_t)
    {
        extern __shared__ uint32_t sm_heads32[]; /* [n_buckets][blockDim.x] */
        HeadsStrided32 hd;
        hd.h = sm_heads32 + threadIdx.x;
        hd.stride = (int)blockDim.x;
        return hd;
    }
};
template <> struct HeadsFactory<HeadsWide32> {
    static __device__ __forceinline__ HeadsWide32 make(const SearchArgs &A, size_t gl)
    {
        HeadsWide32 hd;
        hd.h = A.heads_wide + gl * (size_t)A.heads_wide_stride;
        return hd;
    }
};

template <class Heads> struct HeadsClear { /* all lanes of the warp empty the bucket heads of lane `src` */
    static __device__ __forceinline__ void run(const Heads &, int, int, int) {}
    static constexpr bool cooperative = false;
};
template <> struct HeadsClear<HeadsStrided16> {
    static __device__ __forceinline__ void run(const HeadsStrided16 &hd, int nb, int lane, int src)
    { /* hd.h is this lane's column; column of lane src = hd.h - lane + src */
        uint16_t *col = hd.h - lane + src;
        for (int b = lane; b < nb; b += 32) col[(size_t)b * hd.stride] = 0xffffu;
    }
    static constexpr bool cooperative = true;
};
template <> struct HeadsClear<HeadsStrided32> {
    static __device__ __forceinline__ void run(const HeadsStrided32 &hd, int nb, int lane, int src)
    {
        uint32_t *col = hd.h - lane + src;
        for (int b = lane; b < nb; b += 32) col[(size_t)b * hd.stride] = B2_NIL;
    }
    static constexpr bool cooperative = true;
};

/* takes the parked search with `ticket` out of the ring into this lane; false (and an error flag for the host) when
 * its slot was never published — the wait is bounded so that a bug cannot hang the GPU */
template <class Lane, class Heads>
__device__ __forceinline__ bool unpark(const SearchArgs &A, unsigned ticket, Lane &L, const Heads &heads, GroupStore gs, int &r, unsigned &w)
{
    const unsigned slot = ticket & (PARK_CAP - 1u);
    const volatile uint32_t *flag = A.park_flag + slot;
    unsigned spins = 0;
    while (*flag != ticket + 1u) { /* the parking lane reserved the ticket and is still writing */
        __nanosleep(200);
        if (++spins > (1u << 24)) { atomicExch(A.park_ring + 2, 1u); return false; }
    }
    __threadfence(); /* also drops this SM's L1 lines (CCTL.IVALL): the search may have run here before and been edited elsewhere since */
    const uint32_t *sv = A.park_buf + (size_t)slot * SUSP_STRIDE;
    L.load_state(sv);
    L.bk = heads;
    L.gs = gs;
    r = (int)sv[B2_SAVE_WORDS - 3];
    w = sv[B2_SAVE_WORDS - 2];
    for (int i = 0; i < OG_WORDS; ++i) gs.set(i, sv[SUSP_GROUP_AT + i]);
    for (int b = 0; b < A.env.P.n_buckets; ++b) L.bk.set(b, sv[SUSP_HEADS_AT + (b >> 1)] >> (16 * (b & 1)) & 0xffffu);
    return true;
}

/* BLK: lanes per block.  128 (four warps) is the default; 32 makes the warp the unit that leaves the SM when the
 * work queue has run dry, so that the next launch's blocks (another stream) move in while stragglers finish. */
template <class Heads, bool REUSE, int MINB, bool STATS, int BLK = 128, int QB = 32>
__global__ void __launch_bounds__(BLK, MINB * (128 / BLK)) k_search(const __grid_constant__ SearchArgs A)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const size_t gl = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    SearchLane<Heads, REUSE, STATS, QB> L;
    L.finished = true;
    const Heads heads = HeadsFactory<Heads>::make(A, gl);
    __shared__ uint32_t sm_group[OG_WORDS * BLK]; /* the lanes' open groups, one conflict-free column each */
    GroupStore gs;
    gs.p = sm_group + threadIdx.x;
    gs.stride = BLK;
    bool alive = true, active = false;
    int r = -1;
    unsigned w = 0;
    unsigned long long pops = 0, sectors = 0;

    while (__any_sync(FULL, alive)) {
        const bool need = alive && !active;
        const unsigned m = __ballot_sync(FULL, need);
        if (m) { /* warp-aggregated claim of the next reads */
            const int leader = __ffs((int)m) - 1;
            unsigned base = 0;
            if (lane == leader) base = atomicAdd(A.counter, (unsigned)__popc(m));
            base = __shfl_sync(FULL, base, leader);
            if (HeadsClear<Heads>::cooperative) { /* shared-memory heads of the claiming lanes, emptied by the whole warp */
                for (unsigned mm = m; mm; mm &= mm - 1u) HeadsClear<Heads>::run(heads, A.env.P.n_buckets, lane, __ffs((int)mm) - 1);
                __syncwarp();
            }
            if (need) {
                w = base + (unsigned)__popc(m & ((1u << lane) - 1u));
                if (w < (unsigned)A.n_work && A.resume) { /* a parked search goes on in this lane */
                    active = unpark(A, A.resume_base + w, L, heads, gs, r, w);
                } else if (w < (unsigned)A.n_work) {
                    r = A.work_list ? A.work_list[w] : (int)w;
                    if (!B2_CHECK(r >= 0 && (A.work_list || r < A.n_work), CHK_WORK, w, r, A.n_work)) r = 0;
                    const int len = A.lens[r];
                    L.begin(A.env, heads, gs, A.arena_by_work ? w : (uint32_t)gl, A.rows_by_work ? w : (uint32_t)r, A.recs_by_work ? w : (uint32_t)r,
                            len, A.md[len], A.n_amb[r], HeadsClear<Heads>::cooperative);
                    active = true;
                } else alive = false;
            }
        }
        /* one search step per lane, the warp re-converged around the lookup so that all of its
         * loads are issued together */
        const bool running = active && !L.finished;
        const bool ready = running && L.ready();
        const unsigned wait_mask = __ballot_sync(FULL, running && !ready);
        const unsigned ready_mask = __ballot_sync(FULL, ready);
        /* memory pops are taken in batches: when enough lanes wait for one, or nobody else can move */
        const bool allow_pop = __popc(wait_mask) >= A.pop_batch || ready_mask == 0;
        int mode = L.NONE;
        if (running && (ready || allow_pop)) mode = L.prepare(A.env, allow_pop, A.prep_rounds);
        __syncwarp();
        uint32_t nk4[4], nl4[4], ns = 0;
        if (mode != L.NONE) L.lookup(A.env, mode, nk4, nl4, ns);
        __syncwarp();
        if (mode != L.NONE) L.apply(A.env, mode, nk4, nl4, ns);
        __syncwarp();
        if (active) {
            if (L.finished) {
                if (STATS) {
                    pops += L.n_pops;
                    sectors += L.n_lookups;
                }
                if (L.status != LANE_OK) {
                    A.n_aln[r] = -L.status; /* counts as no records until a later pass succeeds; the host dies on what is left */
                    if (A.over_list) {
                        unsigned idx = atomicAdd(A.n_over, 1u);
                        A.over_list[idx] = r;
                    }
                    if (A.n_rec_full && L.status == LANE_REC_FULL) atomicAdd(A.n_rec_full, 1u);
                } else {
                    A.n_aln[r] = L.n_aln;
                    if (A.over_slot) A.over_slot[r] = (int32_t)w | A.slot_tag;
                }
                active = false;
            }
        }
        if (A.susp_thresh > 0) { /* the queue is dry and this warp has become sparse: park what is left, refill or leave */
            const unsigned dry = __ballot_sync(FULL, !alive), act = __ballot_sync(FULL, active);
            if (dry && act && __popc(act) <= A.susp_thresh) {
                const int leader = __ffs((int)act) - 1;
                unsigned base = 0;
                if (lane == leader) base = atomicAdd(A.park_ring + 0, (unsigned)__popc(act));
                base = __shfl_sync(FULL, base, leader);
                if (active) {
                    const unsigned ticket = base + (unsigned)__popc(act & ((1u << lane) - 1u));
                    uint32_t *sv = A.park_buf + (size_t)(ticket & (PARK_CAP - 1u)) * SUSP_STRIDE;
                    L.save_state(sv);
                    sv[B2_SAVE_WORDS - 3] = (uint32_t)r;
                    sv[B2_SAVE_WORDS - 2] = w;
                    for (int i = 0; i < OG_WORDS; ++i) sv[SUSP_GROUP_AT + i] = gs.get(i);
                    const int nb = A.env.P.n_buckets;
                    for (int b = 0; b < nb; b += 2)
                        sv[SUSP_HEADS_AT + (b >> 1)] = (L.bk.get(b) & 0xffffu) | (b + 1 < nb ? L.bk.get(b + 1) : 0xffffu) << 16;
                    __threadfence();
                    *(volatile uint32_t *)(A.park_flag + (ticket & (PARK_CAP - 1u))) = ticket + 1u; /* published */
                    active = false;
                }
                alive = false;
                __syncwarp();
                /* a full warp's worth of parked searches waiting?  take them (lane 0 moves the head) */
                unsigned got = 0xffffffffu;
                if (lane == 0) {
                    for (;;) {
                        const unsigned h = *(volatile unsigned int *)(A.park_ring + 1), t = *(volatile unsigned int *)(A.park_ring + 0);
                        if (t - h < 32u) break;
                        if (atomicCAS(A.park_ring + 1, h, h + 32u) == h) { got = h; break; }
                    }
                }
                got = __shfl_sync(FULL, got, 0);
                if (got != 0xffffffffu) { /* every lane of the warp continues one of them */
                    active = unpark(A, got + (unsigned)lane, L, heads, gs, r, w);
                    alive = active;
                }
            }
        }
    }
    if (STATS && pops) {
        atomicAdd(A.stat + 0, pops);
        atomicAdd(A.stat + 1, sectors);
    }
}

/* exclusive prefix sum of n_aln (int32) into int64 offsets: block totals, their scan, then local scans */
#define SCAN_ITEMS 2048
__global__ void __launch_bounds__(256) k_scan_totals(const int32_t *in, int n, int64_t *block_tot, unsigned int *n_bad)
{ /* n_bad (optional): counts negative entries — reads that no pass could hold (k_search leaves -status) */
    __shared__ int64_t sh[256];
    int64_t s = 0;
    const int base = blockIdx.x * SCAN_ITEMS;
    for (int i = threadIdx.x; i < SCAN_ITEMS; i += 256) {
        int idx = base + i;
        if (idx < n) {
            s += in[idx] > 0 ? in[idx] : 0;
            if (n_bad && in[idx] < 0) atomicAdd(n_bad, 1u);
        }
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) sh[threadIdx.x] += sh[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) block_tot[blockIdx.x] = sh[0];
}
__global__ void k_scan_blocks(int64_t *block_tot, int nb, int64_t *total)
{ /* few thousand values at most: one thread */
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int64_t run = 0;
        for (int i = 0; i < nb; ++i) {
            int64_t t = block_tot[i];
            block_tot[i] = run;
            run += t;
        }
        *total = run;
    }
}
__global__ void __launch_bounds__(256) k_scan_apply(const int32_t *in, int n, const int64_t *block_off, int64_t *out)
{
    __shared__ int64_t sh[256];
    const int per = SCAN_ITEMS / 256, base = blockIdx.x * SCAN_ITEMS + threadIdx.x * per;
    int64_t loc[SCAN_ITEMS / 256], s = 0;
    for (int i = 0; i < per; ++i) {
        int idx = base + i;
        int v = idx < n ? (in[idx] > 0 ? in[idx] : 0) : 0;
        loc[i] = s;
        s += v;
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t run = block_off[blockIdx.x];
        for (int i = 0; i < 256; ++i) {
            int64_t t = sh[i];
            sh[i] = run;
            run += t;
        }
    }
    __syncthreads();
    const int64_t off = sh[threadIdx.x];
    for (int i = 0; i < per; ++i) {
        int idx = base + i;
        if (idx < n) out[idx] = off + loc[i];
    }
}

__global__ void __launch_bounds__(256) k_compact(int n, const int32_t *n_aln, const int64_t *off, const Rec *recs,
                                                 int rec_cap, const Rec *recs_mid, int rec_cap_mid,
                                                 const Rec *recs_big, int rec_cap_big, const int32_t *over_slot,
                                                 Rec *out)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const int c = n_aln[r];
        if (c <= 0) continue;
        const int sl = over_slot[r];
        const Rec *src = sl < 0 ? recs + (size_t)r * rec_cap
                         : (sl & WIDE_TAG) ? recs_big + (size_t)(sl & ~WIDE_TAG) * rec_cap_big
                                           : recs_mid + (size_t)sl * rec_cap_mid;
        Rec *dst = out + off[r];
        for (int j = 0; j < c; ++j) dst[j] = src[j];
    }
}

/* The batch as the bytes bwa_aln_core writes for it (bwtaln.c:227-231): per read n_aln, then its records.
 * Read r's part starts at word r + 4 * off[r] of the stream; one thread per read (most reads have one or two records). */
__global__ void __launch_bounds__(256) k_sai_pack(int n, const int32_t *n_aln, const int64_t *off, const Rec *packed,
                                                  uint32_t *out)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const int c = n_aln[r] > 0 ? n_aln[r] : 0;
        uint32_t *dst = out + (size_t)r + 4 * (size_t)off[r];
        const uint32_t *src = reinterpret_cast<const uint32_t *>(packed + off[r]);
        dst[0] = (uint32_t)c;
        for (int j = 0; j < 4 * c; ++j) dst[1 + j] = src[j];
    }
}

/* random sector gather: the roofline denominator.  span = 1: independent random
 * 32-byte sectors; span = 2: random 64-byte aligned pairs of sectors (tells whether the memory
 * system moves 64 B per miss anyway). */
/* row N4: one thread per read merges the read's alignments from every stream (alngrp_core.cuh) */
#define B2_MAX_STREAMS 16
struct GrpArgs {
    int n_streams, s_mm;
    int64_t n_reads;
    const int32_t *n_aln[B2_MAX_STREAMS];
    const int64_t *rec_off[B2_MAX_STREAMS];
    const Rec *recs[B2_MAX_STREAMS];
    const int64_t *out_off;
    int32_t *out_n;
    Rec *out_rec;
    uint32_t *out_db;
};
__global__ void __launch_bounds__(128) k_alngrp_totals(const __grid_constant__ GrpArgs A, int32_t *tot)
{
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < A.n_reads; r += (int64_t)gridDim.x * blockDim.x) {
        int32_t t = 0;
        for (int s = 0; s < A.n_streams; ++s) t += A.n_aln[s][r];
        tot[r] = t;
    }
}
__global__ void __launch_bounds__(128) k_alngrp(const __grid_constant__ GrpArgs A)
{
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < A.n_reads; r += (int64_t)gridDim.x * blockDim.x)
        A.out_n[r] = alngrp_merge_one(A.n_streams, r, A.n_aln, A.rec_off, A.recs, A.s_mm, A.out_off[r], A.out_rec, A.out_db);
}

__global__ void __launch_bounds__(256) k_sector_gather(const OccBlk *blk0, uint64_t n0, const OccBlk *blk1, uint64_t n1,
                                                       uint64_t loads_per_thread, int span, unsigned long long *sink)
{ /* uniformly random sectors over BOTH device indexes (the footprint the search gathers from) */
    uint64_t s = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    uint32_t acc = 0;
    const uint64_t u0 = n0 / (uint64_t)span, units = u0 + n1 / (uint64_t)span;
    for (uint64_t i = 0; i < loads_per_thread; i += 8) {
        const OccBlk *ptr[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            const uint64_t u = (uint64_t)(((unsigned __int128)(s >> 11) * units) >> 53);
            ptr[j] = u < u0 ? blk0 + u * (uint64_t)span : blk1 + (u - u0) * (uint64_t)span;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            OccBlk a = ld_blk(ptr[j]);
            acc += a.cnt.x ^ a.bits.w;
            if (span == 2) {
                OccBlk b = ld_blk(ptr[j] + 1);
                acc += b.cnt.y ^ b.bits.z;
            }
        }
    }
    if (acc == 0x7fffffffu) atomicAdd(sink, 1ull);
}

/* -------------------------------------------------------------- context ---- */

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    void need(size_t bytes)
    {
        if (bytes <= cap) return;
        if (p) CK(cudaFree(p));
        p = nullptr;
        size_t want = bytes + bytes / 8 + 256;
        if (want < cap + cap / 2) want = cap + cap / 2; /* grow geometrically: freeing and allocating synchronise the device */
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            size_t fr = 0, tot = 0;
            cudaMemGetInfo(&fr, &tot);
            fprintf(stderr, "[b200aln] cannot allocate %.2f GB of device memory (%.2f GB free of %.2f GB): %s. Abort!\n",
                    want / 1e9, fr / 1e9, tot / 1e9, cudaGetErrorString(e));
            abort();
        }
        cap = want;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

struct HostBuf {
    void *p = nullptr;
    size_t cap = 0;
    void need(size_t bytes)
    {
        if (bytes <= cap) return;
        if (p) CK(cudaFreeHost(p));
        size_t want = bytes + bytes / 8 + 256;
        if (want < cap + cap / 2) want = cap + cap / 2;
        CK(cudaHostAlloc(&p, want, cudaHostAllocDefault));
        cap = want;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

struct b200aln_ctx {
    int device = 0;
    int n_sm = 0;
    bool owns_index = true; /* false for b200aln_clone()d contexts */
    FmView fm[2];
    OccBlk *d_idx[2] = {nullptr, nullptr};
    uint64_t n_blk[2] = {0, 0};
    cudaStream_t st = nullptr;    /* everything but the fast search pass; highest priority (see make_streams) */
    cudaStream_t st_lo = nullptr; /* the fast pass of k_search: lowest priority */
    cudaEvent_t ev[8];
    cudaEvent_t tm[2];
    cudaEvent_t ev_wait = nullptr; /* blocking-sync event: host waits sleep instead of spinning (wait_stream) */
    /* tuning */
    int search_blocks_per_sm = 6, width_blocks_per_sm = 5;
    uint32_t arena_cap = 2048, arena_cap_big = 0; /* 64-byte records per lane; big 0: max_entries + 64 */
    int rec_cap = 8, rec_cap_big = 1 << 13, big_lanes = 128; /* wide pass: 128 lanes x (max_entries+64) x 64 B = 16 GB */
    uint32_t arena_cap_mid = 12288; /* middle pass: 16-bit heads in shared memory, free-list arena */
    int rec_cap_mid = 512, mid_lanes = 148 * 192; /* x 12288 records x 64 B = 22 GB, allocated when a batch first needs it */
    int prefetch_fast = 0, prefetch_mid = 1; /* L2 prefetch of the next pop candidate, per pass */
    int order = 1;         /* fast pass takes the reads by work class, longest searches first (0: arrival order) */
    int search_block = 128; /* lanes per block of the fast pass (128 or 32) */
    int q16 = 1;            /* 16-bit width records in the fast pass when the options allow them (0: always 32-bit) */
    int susp = -16;         /* fast pass: a warp with at most |susp| lanes still searching once the queue is dry parks them.  > 0: always;
                             * < 0: only while at least susp_calls batches are in flight on this device index (parking trades a batch's
                             * latency for SM slots, which pays when other batches are there to use them); 0: never */
    int susp_calls = 4;
    std::atomic<int> active_calls{0}; /* owner: batch calls in flight on this index (all contexts sharing it) */
    int susp_min = 4096;    /* resume rounds park again only while more than this many searches are left */
    int lane_reads = 0;     /* fast pass: lanes = reads / lane_reads for batches too small to give the full grid that many (0: always the full grid) */
    int prep_rounds = 1;   /* pruned pops a lane may go through per warp iteration before the warp moves on */
    int reserve_reads = 0; /* size the per-batch buffers for at least this many reads */
    bool scratch_checked = false; /* the pool of b200aln_prealloc has been asked once (with the first batch) */
    int count = 0;         /* 1: fast pass with the pop / sector counters (b200aln_stats_t pops, occ_lookups) */
    int pop_batch = 1;     /* memory pops are taken when this many lanes of a warp wait for one */
    int lut_k = 14;        /* levels of the path-k-mer interval table (0 = off) */
    uint32_t *d_lut[2] = {nullptr, nullptr}; /* [0]: the table of both indexes */
    uint32_t *d_sa[2] = {nullptr, nullptr}; /* sampled suffix arrays (.sa / .rsa), row N2 */
    uint32_t sa_intv[2] = {0, 0};
    uint64_t n_sa[2] = {0, 0};
    int lut_pin_levels = 10; /* top levels of the table kept persisting in L2 (0 = no window) */
    size_t lut_pin_bytes = 0;
    int batch_max_len = 0; /* > 0: the reference batch this call is a shard of has this longest read */
    /* device buffers */
    DevBuf lens, offs, codes, md, Q, W, n_amb, ent, recs, n_aln, over_slot, over_list, misc, off64,
        blk_tot, packed, dkey, order_buf, Q_re, W_re, ent_big, recs_big, heads_wide, heads_wide_big, ent_mid, recs_mid, over_list2, over_list3, sa_in, sa_out, grp_in, grp_out, sai;
    HostBuf h_in, h_out, h_misc, h_nout, h_ring;
    b200aln_stats_t stats;
    /* chunk pipeline: a call with more than chunk_reads * 1.5 reads is cut into chunks that run on
     * `slots` sibling contexts (own stream, pinned staging and scratch; slot 0 is this context), so that the copies,
     * the width pass and the drain of one chunk's search overlap the search of the others */
    int chunk_reads = 1 << 22, slots = 3;
    int chunk_reads_device = 0; /* device-resident calls: 0 = one launch (nothing to overlap, and large launches are the efficient ones) */
    b200aln_ctx *parent = nullptr;   /* clones: the context that owns the index */
    int n_clones = 0;                /* owner: live clones (public and internal) */
    std::vector<b200aln_ctx *> slot; /* internal siblings 1 .. slots-1 */
    DevBuf asm_n_aln, asm_packed;    /* assembled results of a pipelined device-resident call */
    DevBuf park_buf, park_flag, park_ring; /* parked searches: ring of PARK_CAP slots, their published flags, {tail, head, error} (never reset: tickets do not repeat) */
};

static cudaEvent_t g_origin = nullptr; /* B200ALN_TIMELINE: common clock of the stage timeline (see timeline()) */

[[noreturn]] static void die(const char *func, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "[%s] ", func);
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, " Abort!\n");
    va_end(ap);
    abort();
}

#ifdef B2_CHECKED
extern "C" const char *b200aln_version(void) { return "b200aln 0.2 (sm_100a; device occ block 32 B / 64 bp; B2_CHECKED bounds-checked build)"; }
/* the first violation the device recorded (aln_core.cuh: B2_CHECK), reported like any fatal error */
static void report_device_checks(const char *where)
{
    unsigned int rec[4] = {0, 0, 0, 0};
    CK(cudaMemcpyFromSymbol(rec, b2::b2_check_rec, sizeof rec));
    if (rec[0]) {
        fprintf(stderr, "[%s] B2_CHECKED: violation %u (1 arena slot, 2 bucket, 3 empty pop, 4 width-record position, 5 record fill, "
                        "6 interval table, 7 occ block, 8 group member, 9 work item, 10 entry count) at row %u (read %u): %u vs %u. Abort!\n",
                where, rec[0], rec[1], rec[1] / 2, rec[2], rec[3]);
        abort();
    }
}
#else
extern "C" const char *b200aln_version(void) { return "b200aln 0.2 (sm_100a; device occ block 32 B / 64 bp)"; }
static inline void report_device_checks(const char *) {}
#endif

extern "C" void b200aln_opt_init(b200aln_opt_t *o)
{ /* gap_init_opt, bwtaln.c:21-37 */
    memset(o, 0, sizeof *o);
    o->s_mm = 3; o->s_gapo = 11; o->s_gape = 4;
    o->max_diff = -1; o->max_gapo = 1; o->max_gape = 6;
    o->indel_end_skip = 5; o->max_del_occ = 10; o->max_entries = 2000000;
    o->mode = 0x01 | 0x02; /* GAPE | COMPREAD */
    o->seed_len = 32; o->max_seed_diff = 2;
    o->fnr = 0.04f;
    o->n_threads = 1;
    o->max_top2 = 30;
    o->trim_qual = 0;
}

extern "C" int b200aln_cal_maxdiff(int len, double err, double thres) { return b2host::cal_maxdiff(len, err, thres); }

extern "C" int b200aln_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

/* creates the device's CUDA context (first-call cost) — the CLI driver does this while it reads the index files */
extern "C" void b200aln_warm_device(int device)
{
    if (cudaSetDevice(device) == cudaSuccess) (void)cudaFree(nullptr);
    (void)cudaGetLastError();
}

/* Two streams per context.  The long kernel of a batch (the fast search pass) runs on a low-priority stream, all
 * the short work (copies, width pass, ordering, re-run passes, scan, compact) on a high-priority one: when several
 * contexts share a GPU, the block scheduler then hands SM slots that a draining search kernel frees to the
 * pending short kernels of every batch first, instead of queueing them behind the whole grid of the next search
 * kernel (measured: a 0.1 ms compaction waited 35-94 ms that way, profiles/README.md). */
static void make_streams(b200aln_ctx *c)
{
    int least = 0, greatest = 0;
    CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    const char *e = getenv("B200ALN_PRIO");
    const bool on = !e || atoi(e) != 0;
    CK(cudaStreamCreateWithPriority(&c->st, cudaStreamNonBlocking, on ? greatest : least));
    CK(cudaStreamCreateWithPriority(&c->st_lo, cudaStreamNonBlocking, least));
    CK(cudaEventCreateWithFlags(&c->ev_wait, cudaEventBlockingSync | cudaEventDisableTiming));
}

/* The host side of a batch waits for its stream several times (counters back, records back).  With many contexts
 * on few cores (eight ranks x several batches in flight) spinning waiters starve the threads that feed the GPUs:
 * wait on an event created with cudaEventBlockingSync, which puts the thread to sleep. */
static void wait_stream(b200aln_ctx *c, cudaStream_t st)
{
    CK(cudaEventRecord(c->ev_wait, st));
    CK(cudaEventSynchronize(c->ev_wait));
}

static void upload_index(b200aln_ctx *c, int which, const b200aln_bwt_view_t *v)
{
    const uint64_t expect = ((uint64_t)v->seq_len + 15) / 16 + 4 * (((uint64_t)v->seq_len + 127) / 128 + 1);
    if (v->bwt_size != expect)
        die("b200aln_open", "BWT payload has %llu words, expected %llu for seq_len %u (bwtio.c:51-70).",
            (unsigned long long)v->bwt_size, (unsigned long long)expect, v->seq_len);
    uint32_t *d_raw = nullptr;
    CK(cudaMalloc(&d_raw, v->bwt_size * 4));
    CK(cudaMemcpyAsync(d_raw, v->bwt, v->bwt_size * 4, cudaMemcpyHostToDevice, c->st));
    RefBwt r;
    r.w = d_raw; r.n_words = v->bwt_size; r.seq_len = v->seq_len;
    for (int i = 0; i < 4; ++i) r.L2[i] = v->L2[i];
    const uint64_t nb = fm_num_blocks(v->seq_len);
    CK(cudaMalloc(&c->d_idx[which], nb * sizeof(OccBlk)));
    k_convert_index<<<c->n_sm * 8, 256, 0, c->st>>>(r, c->d_idx[which], nb);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->st));
    CK(cudaFree(d_raw));
    c->n_blk[which] = nb;
    c->fm[which].blk = c->d_idx[which];
    c->fm[which].primary = v->primary;
    c->fm[which].seq_len = v->seq_len;
    c->fm[which].lut = nullptr;
    c->fm[which].lut_k = 0;
}

static void build_luts(b200aln_ctx *c);

/* levels <= lut_pin_levels of the interval table (contiguous at its start, both indexes) persist in L2 */
static void apply_l2_window(b200aln_ctx *c)
{
    if (!c->lut_pin_bytes || !c->d_lut[0]) return;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, c->device));
    size_t bytes = c->lut_pin_bytes;
    if (prop.persistingL2CacheMaxSize <= 0 || prop.accessPolicyMaxWindowSize <= 0) return;
    if (bytes > (size_t)prop.accessPolicyMaxWindowSize) bytes = (size_t)prop.accessPolicyMaxWindowSize;
    size_t carve = bytes < (size_t)prop.persistingL2CacheMaxSize ? bytes : (size_t)prop.persistingL2CacheMaxSize;
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) != cudaSuccess) { (void)cudaGetLastError(); return; }
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof attr);
    attr.accessPolicyWindow.base_ptr = c->d_lut[0];
    attr.accessPolicyWindow.num_bytes = bytes;
    attr.accessPolicyWindow.hitRatio = carve >= bytes ? 1.0f : (float)carve / (float)bytes;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(c->st, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) (void)cudaGetLastError();
    if (c->st_lo && cudaStreamSetAttribute(c->st_lo, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) (void)cudaGetLastError();
    if (getenv("B200ALN_VERBOSE"))
        fprintf(stderr, "[b200aln] L2 window: %.1f MB of the interval table persisting (carve-out %.1f MB)\n", bytes / 1e6, carve / 1e6);
}

/* ---- scratch allocated ahead of the contexts (b200aln_prealloc) ------------------------------------------------
 * Allocating a context's per-batch buffers takes a few hundred milliseconds (the fast pass's arena alone is
 * 15 GB), and the first batch on a fresh context used to pay for it.  A driver that knows how many contexts it is
 * going to use can have the buffers allocated while it is still reading the index files: the sets wait in a
 * per-process pool, and a context of that device takes one with its first batch. */
struct ScratchSet {
    int device = 0;
    DevBuf ent, W, Q, recs, n_aln, over_slot, over_list, off64, n_amb, dkey, order_buf, lens, offs, codes, sai, packed;
    HostBuf h_out, h_nout;
};
static std::mutex g_scratch_mu;
static std::condition_variable g_scratch_cv;
static std::vector<ScratchSet *> g_scratch;
static int g_scratch_pending[64]; /* per device: sets being allocated right now */

extern "C" void b200aln_prealloc(int device, int n_contexts, int n_reads, int max_len)
{
    if (n_contexts <= 0 || n_reads <= 0 || cudaSetDevice(device) != cudaSuccess) { (void)cudaGetLastError(); return; }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    const b200aln_ctx defaults{};
    if (max_len < 1) max_len = 1;
    const size_t n = (size_t)n_reads, lanes = (size_t)prop.multiProcessorCount * defaults.search_blocks_per_sm * 128;
    const int strideQ = (max_len + 15) & ~15, strideW = round_up8(max_len + 1);
    for (int i = 0; i < n_contexts; ++i) {
        ScratchSet *s = new ScratchSet;
        s->device = device;
        {
            std::lock_guard<std::mutex> g(g_scratch_mu);
            ++g_scratch_pending[device & 63];
        }
        s->ent.need(lanes * defaults.arena_cap * sizeof(StackRec));
        s->W.need(n * 2 * strideW * 4 + 64);
        s->Q.need(n * 2 * strideQ * 2 + 64); /* the 16-bit records of the default options; grows for others */
        s->recs.need(n * defaults.rec_cap * 16);
        s->n_aln.need(n * 4);
        s->over_slot.need(n * 4);
        s->over_list.need(n * 4);
        s->off64.need(n * 8);
        s->n_amb.need(n * 4);
        s->dkey.need(n * 2);
        s->order_buf.need(n * 4);
        s->lens.need(n * 4);
        s->offs.need(n * 8);
        s->codes.need(n * (size_t)max_len + 16);
        s->sai.need(n * 28);
        s->packed.need(n * 24);
        s->h_out.need(n * 28);
        s->h_nout.need(n * 4);
        std::lock_guard<std::mutex> g(g_scratch_mu);
        g_scratch.push_back(s);
        --g_scratch_pending[device & 63];
        g_scratch_cv.notify_all();
    }
}

extern "C" void b200aln_prealloc_release(int device)
{ /* frees the sets of `device` that no context has taken (device < 0: of every device) */
    std::vector<ScratchSet *> mine;
    {
        std::unique_lock<std::mutex> g(g_scratch_mu);
        if (device >= 0) g_scratch_cv.wait(g, [&] { return g_scratch_pending[device & 63] == 0; });
        for (size_t i = 0; i < g_scratch.size();)
            if (device < 0 || g_scratch[i]->device == device) {
                mine.push_back(g_scratch[i]);
                g_scratch.erase(g_scratch.begin() + (long)i);
            } else ++i;
    }
    for (ScratchSet *s : mine) {
        if (cudaSetDevice(s->device) != cudaSuccess) { (void)cudaGetLastError(); continue; }
        DevBuf *d[] = {&s->ent, &s->W, &s->Q, &s->recs, &s->n_aln, &s->over_slot, &s->over_list, &s->off64, &s->n_amb,
                       &s->dkey, &s->order_buf, &s->lens, &s->offs, &s->codes, &s->sai, &s->packed};
        for (DevBuf *b : d) b->release();
        s->h_out.release();
        s->h_nout.release();
        delete s;
    }
}

/* With a context's first batch: takes a set allocated ahead for its device, waiting for one that is being allocated
 * right now (allocating a second one next to it would take as long and leave the first without an owner). */
static void adopt_scratch(b200aln_ctx *c)
{
    if (c->scratch_checked) return;
    c->scratch_checked = true;
    ScratchSet *s = nullptr;
    {
        std::unique_lock<std::mutex> g(g_scratch_mu);
        auto find = [&]() -> long {
            for (size_t i = 0; i < g_scratch.size(); ++i)
                if (g_scratch[i]->device == c->device) return (long)i;
            return -1;
        };
        g_scratch_cv.wait(g, [&] { return find() >= 0 || g_scratch_pending[c->device & 63] == 0; });
        const long i = find();
        if (i >= 0) {
            s = g_scratch[(size_t)i];
            g_scratch.erase(g_scratch.begin() + i);
        }
    }
    if (!s || c->ent.p) { /* (a context that already has buffers keeps them; the set goes back) */
        if (s) {
            std::lock_guard<std::mutex> g(g_scratch_mu);
            g_scratch.push_back(s);
            g_scratch_cv.notify_all();
        }
        return;
    }
    c->ent = s->ent; c->W = s->W; c->Q = s->Q; c->recs = s->recs; c->n_aln = s->n_aln; c->over_slot = s->over_slot;
    c->over_list = s->over_list; c->off64 = s->off64; c->n_amb = s->n_amb; c->dkey = s->dkey; c->order_buf = s->order_buf;
    c->lens = s->lens; c->offs = s->offs; c->codes = s->codes; c->sai = s->sai; c->packed = s->packed;
    c->h_out = s->h_out; c->h_nout = s->h_nout;
    delete s; /* (the buffers have no destructor: they live on in the context) */
}

extern "C" b200aln_ctx *b200aln_open(const b200aln_bwt_view_t *bwt, const b200aln_bwt_view_t *rbwt, int device)
{
    int n = b200aln_device_count();
    if (n <= 0) die("b200aln_open", "no CUDA device available; this engine has no CPU fallback.");
    if (device < 0 || device >= n) die("b200aln_open", "device %d out of range (%d visible).", device, n);
    if (bwt->seq_len != rbwt->seq_len) die("b200aln_open", ".bwt and .rbwt describe different lengths.");
    b200aln_ctx *c = new b200aln_ctx();
    c->device = device;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->n_sm = prop.multiProcessorCount;
    {   /* Random 32-byte occ sectors are the dominant traffic: ask the L2 to fetch 32 B per miss
         * instead of its 64-B default, which would double the DRAM traffic of every lookup. */
        const char *e = getenv("B200ALN_L2_FETCH");
        size_t gran = e ? (size_t)atoi(e) : 32;
        if (gran) {
            cudaError_t le = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
            if (le != cudaSuccess) { (void)cudaGetLastError(); fprintf(stderr, "[b200aln_open] note: L2 fetch granularity hint refused\n"); }
            size_t got = 0;
            if (cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity) == cudaSuccess && getenv("B200ALN_VERBOSE"))
                fprintf(stderr, "[b200aln_open] L2 fetch granularity: asked %zu, device reports %zu\n", gran, got);
        }
    }
    make_streams(c);
    for (int i = 0; i < 8; ++i) CK(cudaEventCreate(&c->ev[i]));
    for (int i = 0; i < 2; ++i) CK(cudaEventCreate(&c->tm[i]));
    if (!g_origin && getenv("B200ALN_TIMELINE")) {
        CK(cudaEventCreate(&g_origin));
        CK(cudaEventRecord(g_origin, c->st));
        CK(cudaEventSynchronize(g_origin));
    }
    upload_index(c, 0, bwt);
    upload_index(c, 1, rbwt);
    {
        const char *e = getenv("B200ALN_LUT_K");
        if (e) c->lut_k = atoi(e);
        e = getenv("B200ALN_LUT_PIN");
        if (e) c->lut_pin_levels = atoi(e);
    }
    build_luts(c);
    memset(&c->stats, 0, sizeof c->stats);
    return c;
}

/* (re)builds the interval table of both indexes for c->lut_k levels (clamped to the index size) and pins
 * its small, hot top levels in L2 with an access-policy window on the engine's stream */
static void build_luts(b200aln_ctx *c)
{
    int k = c->lut_k;
    if (k > 14) k = 14; /* path word: 5 bits depth + 2 * (k - 1) bits k-mer */
    while (k > 0 && ((uint64_t)1 << (2 * k)) > 4ull * ((uint64_t)c->fm[0].seq_len + 1)) --k; /* tiny indexes */
    if (c->d_lut[0]) { CK(cudaFree(c->d_lut[0])); c->d_lut[0] = nullptr; }
    for (int w = 0; w < 2; ++w) { c->fm[w].lut = nullptr; c->fm[w].lut_k = 0; c->fm[w].lut_w = w; }
    if (k <= 0) return;
    CK(cudaMalloc(&c->d_lut[0], lut_total_pairs(k) * 8 + 64));
    for (int w = 0; w < 2; ++w) {
        FmView f = c->fm[w];
        for (int level = 0; level < k; ++level) {
            const uint64_t n = (uint64_t)1 << (2 * level);
            const int blocks = (int)((n + 255) / 256 < (uint64_t)c->n_sm * 16 ? (n + 255) / 256 : (uint64_t)c->n_sm * 16);
            k_lut_build<<<blocks, 256, 0, c->st>>>(f, c->d_lut[0], level, n);
            CK(cudaGetLastError());
        }
    }
    for (int w = 0; w < 2; ++w) { c->fm[w].lut = c->d_lut[0]; c->fm[w].lut_k = k; }
    CK(cudaStreamSynchronize(c->st));
    c->lut_pin_bytes = 0;
    if (c->lut_pin_levels > 0) {
        const int pl = c->lut_pin_levels < k ? c->lut_pin_levels : k;
        c->lut_pin_bytes = lut_total_pairs(pl) * 8;
        apply_l2_window(c);
    }
}

extern "C" b200aln_ctx *b200aln_clone(b200aln_ctx *p)
{ /* a second in-flight batch on the same GPU: own stream, staging and scratch, shared index */
    b200aln_ctx *c = new b200aln_ctx();
    c->device = p->device;
    c->n_sm = p->n_sm;
    c->owns_index = false;
    c->parent = p->parent ? p->parent : p;
    ++c->parent->n_clones;
    c->chunk_reads = p->chunk_reads; c->chunk_reads_device = p->chunk_reads_device; c->slots = p->slots;
    CK(cudaSetDevice(c->device));
    for (int i = 0; i < 2; ++i) {
        c->fm[i] = p->fm[i];
        c->d_idx[i] = p->d_idx[i];
        c->n_blk[i] = p->n_blk[i];
        c->d_lut[i] = p->d_lut[i];
    }
    c->lut_k = p->lut_k;
    c->lut_pin_levels = p->lut_pin_levels; c->lut_pin_bytes = p->lut_pin_bytes;
    c->search_blocks_per_sm = p->search_blocks_per_sm; c->width_blocks_per_sm = p->width_blocks_per_sm;
    c->arena_cap = p->arena_cap; c->arena_cap_big = p->arena_cap_big;
    c->rec_cap = p->rec_cap; c->rec_cap_big = p->rec_cap_big; c->big_lanes = p->big_lanes;
    c->arena_cap_mid = p->arena_cap_mid; c->rec_cap_mid = p->rec_cap_mid; c->mid_lanes = p->mid_lanes;
    c->pop_batch = p->pop_batch; c->count = p->count; c->reserve_reads = p->reserve_reads; c->prep_rounds = p->prep_rounds;
    c->prefetch_fast = p->prefetch_fast; c->prefetch_mid = p->prefetch_mid; c->order = p->order; c->search_block = p->search_block; c->q16 = p->q16; c->susp = p->susp; c->susp_min = p->susp_min; c->susp_calls = p->susp_calls; c->lane_reads = p->lane_reads;
    make_streams(c);
    for (int i = 0; i < 8; ++i) CK(cudaEventCreate(&c->ev[i]));
    for (int i = 0; i < 2; ++i) CK(cudaEventCreate(&c->tm[i]));
    memset(&c->stats, 0, sizeof c->stats);
    apply_l2_window(c); /* same window on the clone's stream */
    return c;
}

static bool read_bwt_file(const char *fn, std::vector<uint32_t> &words, b200aln_bwt_view_t *v)
{ /* bwt_restore_bwt, bwtio.c:51-70 */
    FILE *fp = fopen(fn, "rb");
    if (!fp) return false;
    fseek(fp, 0, SEEK_END);
    long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    if (sz < 20) { fclose(fp); return false; }
    uint32_t hdr[5];
    if (fread(hdr, 4, 5, fp) != 5) { fclose(fp); return false; }
    size_t nw = ((size_t)sz - 20) >> 2;
    words.resize(nw);
    if (nw && fread(words.data(), 4, nw, fp) != nw) { fclose(fp); return false; }
    fclose(fp);
    v->primary = hdr[0];
    v->L2[0] = 0;
    for (int i = 0; i < 4; ++i) v->L2[i + 1] = hdr[1 + i];
    v->seq_len = v->L2[4];
    v->bwt_size = nw;
    v->bwt = words.data();
    return true;
}

extern "C" b200aln_ctx *b200aln_open_prefix(const char *prefix, int device)
{
    std::string p(prefix);
    std::vector<uint32_t> w0, w1;
    b200aln_bwt_view_t v0, v1;
    if (!read_bwt_file((p + ".bwt").c_str(), w0, &v0)) die("b200aln_open_prefix", "fail to open file '%s.bwt'.", prefix);
    if (!read_bwt_file((p + ".rbwt").c_str(), w1, &v1)) die("b200aln_open_prefix", "fail to open file '%s.rbwt'.", prefix);
    return b200aln_open(&v0, &v1, device);
}

extern "C" void b200aln_close(b200aln_ctx *c)
{
    if (!c) return;
    for (b200aln_ctx *s : c->slot) b200aln_close(s);
    c->slot.clear();
    if (c->parent) --c->parent->n_clones;
    else if (c->n_clones > 0)
        fprintf(stderr, "[b200aln_close] warning: %d clones of this context are still open; they must be closed first\n", c->n_clones);
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->st);
    DevBuf *bufs[] = {&c->lens, &c->offs, &c->codes, &c->md, &c->Q, &c->W, &c->n_amb, &c->ent,
                      &c->recs, &c->n_aln, &c->over_slot, &c->over_list, &c->misc, &c->off64, &c->blk_tot,
                      &c->packed, &c->dkey, &c->order_buf, &c->Q_re, &c->W_re, &c->ent_big, &c->recs_big, &c->heads_wide, &c->heads_wide_big, &c->ent_mid, &c->recs_mid,
                      &c->over_list2, &c->over_list3, &c->sa_in, &c->sa_out};
    for (DevBuf *b : bufs) b->release();
    c->grp_in.release(); c->grp_out.release(); c->sai.release();
    c->asm_n_aln.release(); c->asm_packed.release();
    c->park_buf.release(); c->park_flag.release(); c->park_ring.release();
    c->h_in.release(); c->h_out.release(); c->h_misc.release(); c->h_nout.release(); c->h_ring.release();
    if (c->owns_index) {
        for (int i = 0; i < 2; ++i) if (c->d_sa[i]) cudaFree(c->d_sa[i]);
        for (int i = 0; i < 2; ++i) if (c->d_idx[i]) cudaFree(c->d_idx[i]);
        for (int i = 0; i < 2; ++i) if (c->d_lut[i]) cudaFree(c->d_lut[i]);
    }
    for (int i = 0; i < 8; ++i) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 2; ++i) cudaEventDestroy(c->tm[i]);
    cudaStreamDestroy(c->st);
    cudaStreamDestroy(c->st_lo);
    cudaEventDestroy(c->ev_wait);
    delete c;
}

extern "C" void b200aln_set_int(b200aln_ctx *c, const char *key, int64_t v)
{
    if (!strcmp(key, "chunk_reads")) { c->chunk_reads = (int)v; return; }
    if (!strcmp(key, "chunk_reads_device")) { c->chunk_reads_device = (int)v; return; }
    if (!strcmp(key, "slots")) {
        if (v < 1 || v > 8) die("b200aln_set_int", "slots must be 1..8.");
        for (b200aln_ctx *s : c->slot) b200aln_close(s); /* re-created with the next pipelined call */
        c->slot.clear();
        c->slots = (int)v;
        return;
    }
    if (strcmp(key, "lut_k") && strcmp(key, "batch_max_len"))
        for (b200aln_ctx *s : c->slot) b200aln_set_int(s, key, v); /* the pipeline's siblings follow their context's knobs */
    if (!strcmp(key, "search_blocks_per_sm")) c->search_blocks_per_sm = (int)v;
    else if (!strcmp(key, "width_blocks_per_sm")) c->width_blocks_per_sm = (int)v;
    else if (!strcmp(key, "arena_cap")) c->arena_cap = (uint32_t)v;
    else if (!strcmp(key, "arena_cap_big")) c->arena_cap_big = (uint32_t)v;
    else if (!strcmp(key, "rec_cap")) c->rec_cap = (int)v;
    else if (!strcmp(key, "rec_cap_big")) c->rec_cap_big = (int)v;
    else if (!strcmp(key, "big_lanes")) c->big_lanes = (int)v;
    else if (!strcmp(key, "pop_batch")) c->pop_batch = (int)v;
    else if (!strcmp(key, "count")) c->count = (int)v;
    else if (!strcmp(key, "reserve_reads")) c->reserve_reads = (int)v;
    else if (!strcmp(key, "prep_rounds")) c->prep_rounds = (int)v;
    else if (!strcmp(key, "order")) c->order = (int)v;
    else if (!strcmp(key, "search_block")) c->search_block = (int)v;
    else if (!strcmp(key, "q16")) c->q16 = (int)v;
    else if (!strcmp(key, "susp")) c->susp = (int)v;
    else if (!strcmp(key, "susp_calls")) c->susp_calls = (int)v;
    else if (!strcmp(key, "susp_min")) c->susp_min = (int)v;
    else if (!strcmp(key, "lane_reads")) c->lane_reads = (int)v;
    else if (!strcmp(key, "prefetch_fast")) c->prefetch_fast = (int)v;
    else if (!strcmp(key, "prefetch_mid")) c->prefetch_mid = (int)v;
    else if (!strcmp(key, "arena_cap_mid")) c->arena_cap_mid = (uint32_t)v;
    else if (!strcmp(key, "rec_cap_mid")) c->rec_cap_mid = (int)v;
    else if (!strcmp(key, "mid_lanes")) c->mid_lanes = (int)v;
    else if (!strcmp(key, "batch_max_len")) c->batch_max_len = (int)v;
    else if (!strcmp(key, "lut_k")) {
        if (!c->owns_index) die("b200aln_set_int", "lut_k must be set on the context that owns the index.");
        for (b200aln_ctx *s : c->slot) b200aln_close(s); /* internal siblings hold pointers into the table */
        c->slot.clear();
        if (c->n_clones > 0) die("b200aln_set_int", "lut_k cannot change while %d clones share the interval table; close them first.", c->n_clones);
        c->lut_k = (int)v;
        build_luts(c);
    }
    else die("b200aln_set_int", "unknown key '%s'.", key);
}

extern "C" void b200aln_timer_start(b200aln_ctx *c)
{
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->tm[0], c->st));
}

extern "C" double b200aln_timer_stop(b200aln_ctx *c)
{
    float ms = 0;
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->tm[1], c->st));
    CK(cudaEventSynchronize(c->tm[1]));
    CK(cudaEventElapsedTime(&ms, c->tm[0], c->tm[1]));
    return (double)ms;
}

extern "C" void b200aln_last_stats(const b200aln_ctx *c, b200aln_stats_t *out) { *out = c->stats; }

/* device-side counters of one batch */
struct Misc {
    unsigned int counter, n_over, counter_big, counter_mid, n_over2, n_over3, n_rec_full, n_bad;
    unsigned int counter_res; /* the resume rounds' work counter */
    unsigned long long stat[2];
    long long total;
    unsigned int class_cnt[B2_N_CLASSES], class_fill[B2_N_CLASSES];
};

/* fast pass: 16-bit heads in shared memory when the score range and the arena allow it */
/* 16-bit heads in shared memory: slot << 1 | group must fit 16 bits and n_buckets columns the shared memory */
static bool fast_heads_ok(const Params &P, uint32_t arena_cap) { return P.n_buckets <= 160 && arena_cap <= 32767u; }

static void launch_search_mid(b200aln_ctx *c, SearchArgs &A, int blocks)
{ /* middle pass: 16-bit shared-memory heads, free-list arena (capacity = stack high-water, not total pushes) */
    const size_t smem = (size_t)A.env.P.n_buckets * 128 * sizeof(uint16_t);
    k_search<HeadsStrided16, true, 1, true><<<blocks, 128, smem, c->st>>>(A);
    CK(cudaGetLastError());
}

static void launch_search_fast(b200aln_ctx *c, SearchArgs &A, int blocks, bool q16, cudaStream_t st)
{
    if (fast_heads_ok(A.env.P, A.env.arena_cap)) {
        const size_t smem = (size_t)A.env.P.n_buckets * 128 * sizeof(uint16_t);
        const bool six = c->search_blocks_per_sm == 6;
        if (q16) { /* 16-bit width records (run_batch_device: q16_ok) */
            if (c->count) k_search<HeadsStrided16, false, 6, true, 128, 16><<<blocks, 128, smem, st>>>(A);
            else if (c->search_block == 32 && six) k_search<HeadsStrided16, false, 6, false, 32, 16><<<blocks * 4, 32, smem / 4, st>>>(A);
            else if (six) k_search<HeadsStrided16, false, 6, false, 128, 16><<<blocks, 128, smem, st>>>(A);
            else k_search<HeadsStrided16, false, 1, false, 128, 16><<<blocks, 128, smem, st>>>(A);
        } else if (c->count) k_search<HeadsStrided16, false, 6, true><<<blocks, 128, smem, st>>>(A); /* with pop / sector counters */
        else if (c->search_block == 32 && six) /* same lanes, one warp per block */
            k_search<HeadsStrided16, false, 6, false, 32><<<blocks * 4, 32, smem / 4, st>>>(A);
        else if (six) k_search<HeadsStrided16, false, 6, false><<<blocks, 128, smem, st>>>(A);
        else k_search<HeadsStrided16, false, 1, false><<<blocks, 128, smem, st>>>(A);
    } else {
        A.heads_wide_stride = A.env.P.n_buckets + (A.env.P.n_buckets + 31) / 32;
        c->heads_wide.need((size_t)blocks * 128 * A.heads_wide_stride * 4);
        A.heads_wide = c->heads_wide.as<uint32_t>();
        k_search<HeadsWide32, false, 1, true><<<blocks, 128, 0, st>>>(A);
    }
    CK(cudaGetLastError());
}

static void launch_search_big(b200aln_ctx *c, SearchArgs &A, int blocks, int threads)
{
    const size_t smem32 = (size_t)A.env.P.n_buckets * threads * sizeof(uint32_t);
    if ((size_t)A.env.P.n_buckets * 128 * sizeof(uint32_t) + OG_WORDS * 128 * sizeof(uint32_t) <= 48 * 1024) { /* 32-bit heads still fit in shared memory (next to the open groups) */
        k_search<HeadsStrided32, true, 1, true><<<blocks, threads, smem32, c->st>>>(A);
        CK(cudaGetLastError());
        return;
    }
    A.heads_wide_stride = A.env.P.n_buckets + (A.env.P.n_buckets + 31) / 32;
    c->heads_wide_big.need((size_t)blocks * threads * A.heads_wide_stride * 4);
    A.heads_wide = c->heads_wide_big.as<uint32_t>();
    k_search<HeadsWide32, true, 1, true><<<blocks, threads, 0, c->st>>>(A);
    CK(cudaGetLastError());
}

/* the device part of one batch; inputs already on the device */
static void run_batch_device(b200aln_ctx *c, int n_reads, int max_len, const int32_t *d_lens, const int64_t *d_offs,
                             const uint8_t *d_codes, const b200aln_opt_t *opt, const std::vector<int> &md,
                             const Params &P, int64_t *total_out)
{
    adopt_scratch(c);
    uint64_t launches = 0;
    /* 16-bit width records when every field fits (aln_core.cuh: QF<16>) and the fast pass runs the kernel built for them */
    int md_max = 0;
    for (int l = 0; l <= max_len && l < (int)md.size(); ++l) md_max = md[l] > md_max ? md[l] : md_max;
    const bool q16 = c->q16 != 0 && md_max < 7 && P.max_seed_diff < 3 && fast_heads_ok(P, c->arena_cap);
    const int strideQ32 = round_up8(max_len > 0 ? max_len : 1), strideW = round_up8(max_len + 1); /* 32-byte aligned rows */
    const int strideQ = q16 ? ((max_len > 0 ? max_len : 1) + 15) & ~15 : strideQ32; /* the fast pass's rows; re-runs use QF<32> */
    const int wblocks = c->n_sm * c->width_blocks_per_sm;
    /* The fast pass's grid: every resident slot for a large batch; for a small one only as many lanes as give each
     * about lane_reads reads (a lane with two reads is idle half the launch, waiting for the batch's longest), so
     * that several small batches in flight share the SMs side by side, each with the arenas of its own lanes only. */
    int sblocks = c->n_sm * c->search_blocks_per_sm;
    if (c->lane_reads > 0) {
        const int64_t want = ((int64_t)n_reads + (int64_t)c->lane_reads * 128 - 1) / ((int64_t)c->lane_reads * 128);
        if (want < sblocks) sblocks = (int)(want > c->n_sm ? want : c->n_sm);
    }
    const size_t lanes = (size_t)sblocks * 128;

    c->md.need(md.size() * 4);
    /* buffers are sized for at least reserve_reads reads, so that a driver whose launches vary in size does not
     * keep growing them (freeing and allocating synchronise the device) */
    const size_t n_alloc = (size_t)(n_reads > c->reserve_reads ? n_reads : c->reserve_reads);
    c->Q.need(n_alloc * 2 * strideQ * (q16 ? 2 : 4) + 64);
    c->W.need(n_alloc * 2 * strideW * 4 + 64);
    c->n_amb.need(n_alloc * 4);
    const bool ordered = c->order != 0 && n_reads > 1;
    if (ordered) {
        c->dkey.need(n_alloc * 2);
        c->order_buf.need(n_alloc * 4);
    }
    c->ent.need(lanes * c->arena_cap * sizeof(StackRec));
    c->recs.need(n_alloc * c->rec_cap * 16);
    c->n_aln.need(n_alloc * 4);
    c->over_slot.need(n_alloc * 4);
    c->over_list.need(n_alloc * 4);
    c->misc.need(sizeof(Misc));
    c->off64.need(n_alloc * 8);
    const int nscan = (n_reads + SCAN_ITEMS - 1) / SCAN_ITEMS;
    c->blk_tot.need((size_t)nscan * 8 + 8);
    c->h_misc.need(sizeof(Misc));

    if (c->susp != 0 && !c->park_ring.p && fast_heads_ok(P, c->arena_cap)) {
        /* The parking ring, with the first batch (not with the first parked one: allocating synchronises the device).
         * Cleared HERE, ahead of ev[2]: the fast pass runs on another stream and waits for that event only — cleared
         * after it (as this once was), the memsets could land in the middle of the first launch and wipe the
         * tickets of searches already parked, whose reads then came back without hits. */
        c->park_buf.need((size_t)PARK_CAP * SUSP_STRIDE * 4);
        c->park_flag.need((size_t)PARK_CAP * 4);
        c->park_ring.need(64);
        c->h_ring.need(64);
        CK(cudaMemsetAsync(c->park_flag.p, 0, (size_t)PARK_CAP * 4, c->st));
        CK(cudaMemsetAsync(c->park_ring.p, 0, 64, c->st));
    }

    CK(cudaMemcpyAsync(c->md.p, md.data(), md.size() * 4, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemsetAsync(c->misc.p, 0, sizeof(Misc), c->st));
    CK(cudaMemsetAsync(c->over_slot.p, 0xff, (size_t)n_reads * 4, c->st));
    Misc *dm = c->misc.as<Misc>();

    CK(cudaEventRecord(c->ev[1], c->st));
    WidthArgs WA;
    WA.fm[0] = c->fm[0]; WA.fm[1] = c->fm[1];
    WA.n_reads = n_reads; WA.work_list = nullptr; WA.lens = d_lens; WA.offs = d_offs; WA.codes = d_codes;
    WA.comp = (opt->mode & MODE_COMPREAD) ? 1 : 0; WA.seed_len = opt->seed_len;
    WA.strideQ = strideQ; WA.strideW = strideW;
    WA.Q = c->Q.as<uint32_t>(); WA.W = c->W.as<uint32_t>();
    WA.rows_by_work = 0;
    WA.n_amb = c->n_amb.as<int32_t>();
    WA.dkey = ordered ? c->dkey.as<uint8_t>() : nullptr;
    if (q16) k_width<16><<<wblocks, 128, 0, c->st>>>(WA);
    else k_width<32><<<wblocks, 128, 0, c->st>>>(WA);
    CK(cudaGetLastError());
    ++launches;
    if (ordered) { /* part of the width stage's time */
        OrderArgs OA;
        OA.n_reads = n_reads; OA.lens = d_lens; OA.md = c->md.as<int32_t>(); OA.dkey = c->dkey.as<uint8_t>();
        OA.cnt = dm->class_cnt; OA.fill = dm->class_fill; OA.order = c->order_buf.as<int32_t>();
        const int oblocks = (n_reads + 2047) / 2048 < c->n_sm * 8 ? (n_reads + 2047) / 2048 : c->n_sm * 8;
        k_class_count<<<oblocks, 256, 0, c->st>>>(OA);
        k_class_scatter<<<oblocks, 256, 0, c->st>>>(OA);
        CK(cudaGetLastError());
        launches += 2;
    }
    CK(cudaEventRecord(c->ev[2], c->st));

    SearchArgs SA;
    SA.env.fm[0] = c->fm[0]; SA.env.fm[1] = c->fm[1]; SA.env.P = P;
    SA.env.prefetch_next = c->prefetch_fast;
    SA.n_work = n_reads; SA.work_list = ordered ? c->order_buf.as<int32_t>() : nullptr;
    SA.lens = d_lens; SA.n_amb = c->n_amb.as<int32_t>(); SA.md = c->md.as<int32_t>();
    SA.env.Q = WA.Q; SA.env.W = WA.W; SA.env.strideQ = strideQ; SA.env.strideW = strideW;
    SA.rows_by_work = 0;
    SA.env.ent = c->ent.as<StackRec>(); SA.env.arena_cap = c->arena_cap;
    SA.env.recs = c->recs.as<Rec>(); SA.env.rec_cap = c->rec_cap; SA.recs_by_work = 0;
    SA.n_aln = c->n_aln.as<int32_t>(); SA.over_slot = nullptr; SA.slot_tag = 0;
    SA.counter = &dm->counter; SA.n_over = &dm->n_over; SA.over_list = c->over_list.as<int32_t>();
    SA.stat = dm->stat;
    SA.heads_wide = nullptr; SA.heads_wide_stride = 0;
    SA.arena_by_work = 0; SA.n_rec_full = nullptr;
    SA.pop_batch = c->pop_batch;
    SA.prep_rounds = c->prep_rounds;
    /* parking of sparse warps (SearchArgs): only the shared-memory-heads kernels, and not the counting instance */
    const b200aln_ctx *owner = c->parent ? c->parent : c;
    const int park_at = c->susp > 0 ? c->susp : (c->susp < 0 && owner->active_calls.load() >= c->susp_calls ? -c->susp : 0);
    const bool parking = park_at > 0 && !c->count && fast_heads_ok(P, c->arena_cap);
    SA.susp_thresh = 0; SA.park_buf = nullptr; SA.park_flag = nullptr; SA.park_ring = nullptr; SA.resume = 0; SA.resume_base = 0;
    if (parking) {
        if (lanes > PARK_CAP) die("b200aln_batch", "internal: %zu lanes exceed the parking ring.", lanes);
        SA.susp_thresh = park_at; SA.park_buf = c->park_buf.as<uint32_t>(); SA.park_flag = c->park_flag.as<uint32_t>();
        SA.park_ring = c->park_ring.as<unsigned int>();
    }
    CK(cudaStreamWaitEvent(c->st_lo, c->ev[2], 0)); /* the fast pass: on the low-priority stream, fenced on both sides */
    launch_search_fast(c, SA, sblocks, q16, c->st_lo);
    ++launches;
    CK(cudaEventRecord(c->ev[3], c->st_lo));
    CK(cudaStreamWaitEvent(c->st, c->ev[3], 0));
    if (parking) { /* what the launch left in the ring: dense again, on the high-priority stream, until nothing is left */
        c->h_ring.need(64);
        unsigned int *ring = c->h_ring.as<unsigned int>();
        unsigned last_parked = 0xffffffffu;
        for (int round = 0;; ++round) {
            CK(cudaMemcpyAsync(ring, c->park_ring.p, 12, cudaMemcpyDeviceToHost, c->st));
            wait_stream(c, c->st);
            const unsigned tail = ring[0], head = ring[1], n_parked = tail - head;
            if (ring[2]) die("b200aln_batch", "internal: a parked search was never published (ring tail %u head %u).", tail, head);
            if (!n_parked) break;
            if (n_parked > lanes || round > 64) die("b200aln_batch", "internal: %u parked searches of %zu lanes after %d rounds.", n_parked, lanes, round);
            SearchArgs SR = SA;
            SR.resume = 1; SR.resume_base = head; SR.n_work = (int)n_parked; SR.work_list = nullptr;
            SR.counter = &dm->counter_res;
            /* park again only while that can still pack warps: enough searches left, and fewer than last time */
            SR.susp_thresh = n_parked > (unsigned)(c->susp_min > 64 ? c->susp_min : 64) && n_parked < last_parked ? park_at : 0;
            last_parked = n_parked;
            CK(cudaMemsetAsync(&dm->counter_res, 0, 4, c->st));
            CK(cudaMemcpyAsync(c->park_ring.as<unsigned int>() + 1, &ring[0], 4, cudaMemcpyHostToDevice, c->st)); /* head = tail: the launch owns [head, tail) */
            int rblocks = (int)((n_parked + 127) / 128);
            if (rblocks > sblocks) rblocks = sblocks;
            launch_search_fast(c, SR, rblocks, q16, c->st);
            ++launches;
        }
    }

    /* Reads whose stack or record slab outgrew the fast pass are searched again from scratch:
     * middle pass = the same fast kernel with a larger arena / record slab (repeat-rich reads with
     * many hits land here), then the wide pass (arena of max_entries, 32-bit heads in memory).
     * An aborted attempt may have edited W/Q in place (gap_shadow), so the widths are rebuilt first. */
    CK(cudaMemcpyAsync(c->h_misc.p, c->misc.p, sizeof(Misc), cudaMemcpyDeviceToHost, c->st));
    wait_stream(c, c->st);
    unsigned n_over = c->h_misc.as<Misc>()->n_over;
    c->stats.overflow_reads = n_over;
    const bool verbose = getenv("B200ALN_VERBOSE") != nullptr;
    struct timespec tv0;
    clock_gettime(CLOCK_MONOTONIC, &tv0);
    auto since = [&]() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (t.tv_sec - tv0.tv_sec) * 1e3 + (t.tv_nsec - tv0.tv_nsec) * 1e-6; };
    if (verbose) {
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]);
        fprintf(stderr, "[b200aln] fast pass %.1f ms, %u of %d reads flagged\n", ms, n_over, n_reads);
    }
    const int32_t *wide_list = c->over_list.as<int32_t>();
    unsigned n_wide = n_over;
    if (n_over && fast_heads_ok(P, c->arena_cap_mid) &&
        (c->arena_cap_mid > c->arena_cap || c->rec_cap_mid > c->rec_cap)) {
        int lanes_mid = c->mid_lanes;
        if ((unsigned)lanes_mid > ((n_over + 127u) / 128u) * 128u) lanes_mid = (int)(((n_over + 127u) / 128u) * 128u);
        const int mblocks = (lanes_mid + 127) / 128;
        c->ent_mid.need((size_t)mblocks * 128 * c->arena_cap_mid * sizeof(StackRec));
        c->recs_mid.need((size_t)n_over * c->rec_cap_mid * 16);
        c->over_list2.need((size_t)n_over * 4);
        c->Q_re.need((size_t)n_over * 2 * strideQ32 * 4 + 64);
        c->W_re.need((size_t)n_over * 2 * strideW * 4 + 64);
        WidthArgs WM = WA;
        WM.dkey = nullptr; WM.rows_by_work = 1; WM.strideQ = strideQ32;
        WM.Q = c->Q_re.as<uint32_t>(); WM.W = c->W_re.as<uint32_t>();
        WM.n_reads = (int)n_over; WM.work_list = c->over_list.as<int32_t>();
        k_width<32><<<wblocks, 128, 0, c->st>>>(WM);
        CK(cudaGetLastError());
        ++launches;
        SearchArgs SM = SA;
        SM.susp_thresh = 0; SM.resume = 0; /* the re-run passes run every search to its end */
        SM.env.Q = WM.Q; SM.env.W = WM.W; SM.env.strideQ = strideQ32; SM.rows_by_work = 1;
        SM.n_work = (int)n_over; SM.work_list = c->over_list.as<int32_t>();
        SM.env.ent = c->ent_mid.as<StackRec>(); SM.env.arena_cap = c->arena_cap_mid;
        SM.env.recs = c->recs_mid.as<Rec>(); SM.env.rec_cap = c->rec_cap_mid; SM.recs_by_work = 1;
        SM.over_slot = c->over_slot.as<int32_t>(); SM.slot_tag = 0;
        SM.env.prefetch_next = c->prefetch_mid;
        SM.prep_rounds = 1 << 30; /* few lanes are busy here: a lane should not wait for the others to prune */
        SM.counter = &dm->counter_mid; SM.n_over = &dm->n_over2; SM.over_list = c->over_list2.as<int32_t>();
        launch_search_mid(c, SM, mblocks);
        ++launches;
        CK(cudaMemcpyAsync(c->h_misc.p, c->misc.p, sizeof(Misc), cudaMemcpyDeviceToHost, c->st));
        wait_stream(c, c->st);
        n_wide = c->h_misc.as<Misc>()->n_over2;
        wide_list = c->over_list2.as<int32_t>();
        if (verbose) fprintf(stderr, "[b200aln] middle pass %.1f ms (%d lanes), %u reads left for the wide pass\n", since(), lanes_mid, n_wide);
    }
    if (n_wide) {
        /* The reference's hit array grows without limit (bwtgap.c:187-191) and its stack holds max_entries entries:
         * the wide pass gives every read an arena of max_entries + 64 records (arenas only for the reads that need
         * one) and a record slab that is enlarged, and the pass repeated, until every read fits. */
        const uint32_t cap_big = c->arena_cap_big ? c->arena_cap_big : (uint32_t)opt->max_entries + 64u;
        c->over_list3.need((size_t)n_wide * 4);
        for (int attempt = 0;; ++attempt) {
            int big_lanes = c->big_lanes;
            if ((unsigned)big_lanes > ((n_wide + 31u) / 32u) * 32u) big_lanes = (int)(((n_wide + 31u) / 32u) * 32u);
            const int bthreads = big_lanes < 128 ? big_lanes : 128, bblocks = (big_lanes + bthreads - 1) / bthreads;
            const bool by_work = n_wide <= (unsigned)(bblocks * bthreads);
            c->ent_big.need((size_t)(by_work ? n_wide : (unsigned)(bblocks * bthreads)) * cap_big * sizeof(StackRec));
            c->recs_big.need((size_t)n_wide * c->rec_cap_big * 16);
            c->Q_re.need((size_t)n_wide * 2 * strideQ32 * 4 + 64);
            c->W_re.need((size_t)n_wide * 2 * strideW * 4 + 64);
            WidthArgs WB = WA;
            WB.dkey = nullptr; WB.rows_by_work = 1; WB.strideQ = strideQ32;
            WB.Q = c->Q_re.as<uint32_t>(); WB.W = c->W_re.as<uint32_t>();
            WB.n_reads = (int)n_wide; WB.work_list = wide_list;
            k_width<32><<<wblocks, 128, 0, c->st>>>(WB);
            CK(cudaGetLastError());
            ++launches;
            SearchArgs SB = SA;
            SB.susp_thresh = 0; SB.resume = 0;
            SB.env.Q = WB.Q; SB.env.W = WB.W; SB.env.strideQ = strideQ32; SB.rows_by_work = 1;
            SB.n_work = (int)n_wide; SB.work_list = wide_list;
            SB.env.ent = c->ent_big.as<StackRec>(); SB.env.arena_cap = cap_big; SB.arena_by_work = by_work ? 1 : 0;
            SB.env.recs = c->recs_big.as<Rec>(); SB.env.rec_cap = c->rec_cap_big; SB.recs_by_work = 1;
            SB.over_slot = c->over_slot.as<int32_t>(); SB.slot_tag = WIDE_TAG;
            SB.env.prefetch_next = c->prefetch_mid;
            SB.prep_rounds = 1 << 30;
            SB.counter = &dm->counter_big; SB.n_over = &dm->n_over3; SB.over_list = c->over_list3.as<int32_t>();
            SB.n_rec_full = &dm->n_rec_full;
            launch_search_big(c, SB, bblocks, bthreads);
            ++launches;
            CK(cudaMemcpyAsync(c->h_misc.p, c->misc.p, sizeof(Misc), cudaMemcpyDeviceToHost, c->st));
            wait_stream(c, c->st);
            const Misc hw = *c->h_misc.as<Misc>();
            if (verbose) fprintf(stderr, "[b200aln] wide pass done at %.1f ms (%d lanes, %d records per read), %u reads do not fit\n", since(), bblocks * bthreads, c->rec_cap_big, hw.n_over3);
            if (!hw.n_over3) break;
            if (hw.n_over3 > hw.n_rec_full)
                die("b200aln_batch", "%u reads exceeded the large per-read arena of %u records (knob arena_cap_big; the default holds max_entries).",
                    hw.n_over3 - hw.n_rec_full, cap_big);
            if (c->rec_cap_big > (1 << 26) || attempt > 16) die("b200aln_batch", "a read has more than %d hits.", c->rec_cap_big);
            c->rec_cap_big *= 4; /* everything the pass produced is redone with the larger slabs */
            CK(cudaMemsetAsync(&dm->counter_big, 0, 4, c->st));
            CK(cudaMemsetAsync(&dm->n_over3, 0, 8, c->st)); /* n_over3, n_rec_full */
        }
    }
    CK(cudaEventRecord(c->ev[4], c->st));

    k_scan_totals<<<nscan, 256, 0, c->st>>>(c->n_aln.as<int32_t>(), n_reads, c->blk_tot.as<int64_t>(), &dm->n_bad);
    k_scan_blocks<<<1, 32, 0, c->st>>>(c->blk_tot.as<int64_t>(), nscan, (int64_t *)&dm->total);
    k_scan_apply<<<nscan, 256, 0, c->st>>>(c->n_aln.as<int32_t>(), n_reads, c->blk_tot.as<int64_t>(),
                                           c->off64.as<int64_t>());
    CK(cudaGetLastError());
    launches += 3;
    CK(cudaMemcpyAsync(c->h_misc.p, c->misc.p, sizeof(Misc), cudaMemcpyDeviceToHost, c->st));
    wait_stream(c, c->st);
    const Misc hm = *c->h_misc.as<Misc>();
    report_device_checks("b200aln_batch");
    if (hm.n_bad) /* never a partial result (include/b200aln.h); the wide pass already grows what it can */
        die("b200aln_batch", "%u reads exceeded the per-read arena or record capacity of every pass (knobs arena_cap_big / rec_cap_big).", hm.n_bad);
    const int64_t total = hm.total;
    c->packed.need((size_t)(total > 0 ? total : 1) * 16);
    k_compact<<<c->n_sm * 4, 256, 0, c->st>>>(n_reads, c->n_aln.as<int32_t>(), c->off64.as<int64_t>(),
                                              c->recs.as<Rec>(), c->rec_cap, c->recs_mid.as<Rec>(), c->rec_cap_mid,
                                              c->recs_big.as<Rec>(), c->rec_cap_big, c->over_slot.as<int32_t>(),
                                              c->packed.as<Rec>());
    CK(cudaGetLastError());
    ++launches;
    CK(cudaEventRecord(c->ev[5], c->st));
    c->stats.kernel_launches = launches;
    c->stats.pops = hm.stat[0];
    c->stats.occ_lookups = hm.stat[1];
    *total_out = total;
}

/* B200ALN_TIMELINE=1: where each batch's stages fall on a common clock (an event recorded when the first context
 * was opened) — shows how the launches of contexts that share a GPU interleave */
static void timeline(b200aln_ctx *c)
{
    static const bool on = getenv("B200ALN_TIMELINE") != nullptr;
    if (!on || !g_origin) return;
    float t[6];
    for (int i = 1; i <= 5; ++i) CK(cudaEventElapsedTime(&t[i], g_origin, c->ev[i]));
    fprintf(stderr, "[timeline] ctx %p width %.2f-%.2f fast %.2f-%.2f passes-end %.2f compact-end %.2f\n", (void *)c, t[1], t[2], t[2], t[3], t[4], t[5]);
}

static void finish_stats(b200aln_ctx *c, bool with_copies)
{
    float ms = 0;
    CK(cudaEventSynchronize(c->ev[with_copies ? 6 : 5]));
    timeline(c);
    CK(cudaEventElapsedTime(&ms, c->ev[1], c->ev[2])); c->stats.ms_width = ms;
    CK(cudaEventElapsedTime(&ms, c->ev[2], c->ev[4])); c->stats.ms_search = ms;
    CK(cudaEventElapsedTime(&ms, c->ev[4], c->ev[5])); c->stats.ms_compact = ms;
    if (with_copies) {
        CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1])); c->stats.ms_h2d = ms;
        CK(cudaEventElapsedTime(&ms, c->ev[5], c->ev[6])); c->stats.ms_d2h = ms;
        CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[6])); c->stats.ms_total = ms;
    } else {
        c->stats.ms_h2d = c->stats.ms_d2h = 0;
        CK(cudaEventElapsedTime(&ms, c->ev[1], c->ev[5])); c->stats.ms_total = ms;
    }
}

static bool host_pinned(const void *p)
{ /* page-locked (cudaHostAlloc / cudaHostRegister) memory can be the end point of an asynchronous copy */
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

static void add_stats(b200aln_stats_t &acc, const b200aln_stats_t &s)
{
    acc.ms_h2d += s.ms_h2d; acc.ms_width += s.ms_width; acc.ms_search += s.ms_search; acc.ms_compact += s.ms_compact;
    acc.ms_d2h += s.ms_d2h; acc.kernel_launches += s.kernel_launches; acc.overflow_reads += s.overflow_reads;
    acc.pops += s.pops; acc.occ_lookups += s.occ_lookups;
}

/* counts a public batch call in on its device index for its duration (SearchArgs parking: `susp` < 0) */
struct CallInFlight {
    b200aln_ctx *o;
    explicit CallInFlight(b200aln_ctx *c) : o(c->parent ? c->parent : c) { o->active_calls.fetch_add(1); }
    ~CallInFlight() { o->active_calls.fetch_sub(1); }
};

/* the siblings a pipelined call runs its chunks on (created once, kept) */
static void ensure_slots(b200aln_ctx *c)
{
    while ((int)c->slot.size() < c->slots - 1) {
        b200aln_ctx *s = b200aln_clone(c);
        s->slots = 1; /* a sibling never splits further */
        c->slot.push_back(s);
    }
}

static int pipeline_chunks(const b200aln_ctx *c, int n_reads, int chunk_reads)
{
    if (c->slots < 2 || chunk_reads <= 0 || (int64_t)n_reads <= (int64_t)chunk_reads + chunk_reads / 2) return 1;
    return (int)(((int64_t)n_reads + chunk_reads - 1) / chunk_reads);
}

/* In-order hand-over of the chunks' results: a chunk learns where its records go once every earlier chunk has
 * reported its total.  Copies into the assembled buffer hold `grow` shared; enlarging the buffer holds it exclusive. */
struct ChunkOrder {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<int64_t> tot;
    std::shared_mutex grow;
    std::atomic<int> next{0};
    explicit ChunkOrder(int n) : tot((size_t)n, -1) {}
    int64_t report(int ci, int64_t t)
    { /* returns the record offset of chunk ci */
        std::unique_lock<std::mutex> lk(mu);
        tot[(size_t)ci] = t;
        cv.notify_all();
        int64_t off = 0;
        for (int j = 0; j < ci; ++j) {
            cv.wait(lk, [&] { return tot[(size_t)j] >= 0; });
            off += tot[(size_t)j];
        }
        return off;
    }
};

/* one chunk of host buffers on context s: H2D (through s's pinned staging when the caller's memory is pageable),
 * the device passes; leaves n_aln / packed records on the device; returns the number of records */
static int64_t run_chunk_host(b200aln_ctx *s, int n, const int32_t *lens, const int64_t *offs, const uint8_t *codes,
                              const b200aln_opt_t *opt, const Params &P, const std::vector<int> &md, bool pinned_in)
{
    adopt_scratch(s);
    int max_len = 0;
    int64_t lo = INT64_MAX, hi = 0;
    for (int r = 0; r < n; ++r) {
        if (lens[r] > max_len) max_len = lens[r];
        if (offs[r] < lo) lo = offs[r];
        if (offs[r] + lens[r] > hi) hi = offs[r] + lens[r];
    }
    if (lo > hi) lo = hi;
    const size_t nb = (size_t)(hi - lo);
    const size_t n_alloc = (size_t)(n > s->reserve_reads ? n : s->reserve_reads);
    s->lens.need(n_alloc * 4);
    s->offs.need(n_alloc * 8);
    s->codes.need((size_t)((double)nb * ((double)n_alloc / (double)n)) + 16);
    const void *h_lens = lens, *h_offs = offs, *h_codes = codes + lo;
    if (!pinned_in) { /* pageable caller memory: through this context's page-locked staging buffer */
        const size_t at_offs = ((size_t)n * 4 + 63) & ~(size_t)63, at_codes = (at_offs + (size_t)n * 8 + 63) & ~(size_t)63;
        s->h_in.need(at_codes + nb + 64);
        char *st = s->h_in.as<char>();
        memcpy(st, lens, (size_t)n * 4);
        memcpy(st + at_offs, offs, (size_t)n * 8);
        memcpy(st + at_codes, codes + lo, nb);
        h_lens = st; h_offs = st + at_offs; h_codes = st + at_codes;
    }
    CK(cudaEventRecord(s->ev[0], s->st));
    CK(cudaMemcpyAsync(s->lens.p, h_lens, (size_t)n * 4, cudaMemcpyHostToDevice, s->st));
    CK(cudaMemcpyAsync(s->offs.p, h_offs, (size_t)n * 8, cudaMemcpyHostToDevice, s->st));
    if (nb) CK(cudaMemcpyAsync(s->codes.p, h_codes, nb, cudaMemcpyHostToDevice, s->st));
    int64_t tot = 0;
    /* offs stay absolute: the kernels address codes_base + offs[r] */
    run_batch_device(s, n, max_len, s->lens.as<int32_t>(), s->offs.as<int64_t>(), s->codes.as<uint8_t>() - lo, opt, md, P, &tot);
    return tot;
}

/* D2H of a finished chunk: records to rec_dst (page-locked), counts to the caller's n_aln (staged when pageable) */
static void fetch_chunk_host(b200aln_ctx *s, int n, int64_t tot, b200aln_rec_t *rec_dst, int32_t *n_aln, bool pinned_out)
{
    int32_t *dst_n = n_aln;
    if (!pinned_out) {
        s->h_nout.need((size_t)n * 4);
        dst_n = s->h_nout.as<int32_t>();
    }
    CK(cudaMemcpyAsync(dst_n, s->n_aln.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s->st));
    if (tot) CK(cudaMemcpyAsync(rec_dst, s->packed.p, (size_t)tot * 16, cudaMemcpyDeviceToHost, s->st));
    CK(cudaEventRecord(s->ev[6], s->st));
    wait_stream(s, s->st);
    if (!pinned_out) memcpy(n_aln, dst_n, (size_t)n * 4);
    finish_stats(s, true);
}

static void grow_host_keep(HostBuf &b, size_t bytes)
{ /* like need(), but the contents survive */
    if (bytes <= b.cap) return;
    void *np = nullptr;
    size_t want = bytes + bytes / 2 + 256;
    CK(cudaHostAlloc(&np, want, cudaHostAllocDefault));
    if (b.p) { memcpy(np, b.p, b.cap); CK(cudaFreeHost(b.p)); }
    b.p = np;
    b.cap = want;
}
static void grow_dev_keep(DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return;
    void *np = nullptr;
    size_t want = bytes + bytes / 2 + 256;
    CK(cudaMalloc(&np, want));
    if (b.p) { CK(cudaMemcpy(np, b.p, b.cap, cudaMemcpyDeviceToDevice)); CK(cudaFree(b.p)); }
    b.p = np;
    b.cap = want;
}

extern "C" const b200aln_rec_t *b200aln_batch(b200aln_ctx *c, int n_reads, const int32_t *lens, const int64_t *offs,
                                              const uint8_t *codes, const b200aln_opt_t *opt, int32_t *n_aln,
                                              int64_t *total)
{
    CK(cudaSetDevice(c->device));
    CallInFlight in_flight(c);
    *total = 0;
    if (n_reads <= 0) return c->h_out.as<b200aln_rec_t>();
    int max_len = 0;
    for (int r = 0; r < n_reads; ++r) {
        if (lens[r] < 0) die("b200aln_batch", "read %d has negative length %d.", r, lens[r]);
        if (lens[r] > max_len) max_len = lens[r];
    }
    Params P;
    std::vector<int> md;
    if (c->batch_max_len > 0 && c->batch_max_len < max_len)
        die("b200aln_batch", "batch_max_len %d is smaller than a read of this shard (%d).", c->batch_max_len, max_len);
    /* a shard of a larger reference batch: the batch-level clamp (bwtaln.c:89-92) uses the whole batch's longest read */
    b2host::make_params(*opt, c->batch_max_len > 0 ? c->batch_max_len : max_len, lens, n_reads, P, md);
    if ((int)md.size() < max_len + 1) md.resize((size_t)max_len + 1, opt->max_diff);
    const bool pinned_in = host_pinned(lens) && host_pinned(offs) && host_pinned(codes), pinned_out = host_pinned(n_aln);

    const int n_chunks = pipeline_chunks(c, n_reads, c->chunk_reads);
    if (n_chunks == 1) {
        const int64_t tot = run_chunk_host(c, n_reads, lens, offs, codes, opt, P, md, pinned_in);
        c->h_out.need((size_t)(tot > 0 ? tot : 1) * 16);
        fetch_chunk_host(c, n_reads, tot, c->h_out.as<b200aln_rec_t>(), n_aln, pinned_out);
        *total = tot;
        return c->h_out.as<b200aln_rec_t>();
    }

    /* pipelined: chunks of the call on this context and its siblings, results assembled in read order */
    ensure_slots(c);
    const auto t0 = std::chrono::steady_clock::now();
    const int per = (n_reads + n_chunks - 1) / n_chunks, n_workers = c->slots < n_chunks ? c->slots : n_chunks;
    if (c->h_out.cap < (size_t)n_reads * 24) c->h_out.need((size_t)n_reads * 24); /* grown below when the hits outnumber that */
    ChunkOrder ord(n_chunks);
    b200aln_stats_t acc;
    memset(&acc, 0, sizeof acc);
    std::mutex acc_mu;
    auto worker = [&](int si) {
        b200aln_ctx *s = si == 0 ? c : c->slot[(size_t)si - 1];
        CK(cudaSetDevice(s->device));
        for (;;) {
            const int ci = ord.next.fetch_add(1);
            if (ci >= n_chunks) break;
            const int lo = ci * per, n = (lo + per < n_reads ? lo + per : n_reads) - lo;
            const int64_t tot = run_chunk_host(s, n, lens + lo, offs + lo, codes, opt, P, md, pinned_in);
            const int64_t off = ord.report(ci, tot);
            if ((size_t)(off + tot) * 16 > c->h_out.cap) {
                std::unique_lock<std::shared_mutex> g(ord.grow); /* no copy into the buffer is in flight now */
                grow_host_keep(c->h_out, (size_t)(off + tot) * 16);
            }
            {
                std::shared_lock<std::shared_mutex> g(ord.grow);
                fetch_chunk_host(s, n, tot, c->h_out.as<b200aln_rec_t>() + off, n_aln + lo, pinned_out);
            }
            std::lock_guard<std::mutex> lk(acc_mu);
            add_stats(acc, s->stats);
        }
    };
    std::vector<std::thread> th;
    for (int si = 1; si < n_workers; ++si) th.emplace_back(worker, si);
    worker(0);
    for (auto &t : th) t.join();
    int64_t tot_all = 0;
    for (int64_t t : ord.tot) tot_all += t;
    acc.ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    c->stats = acc;
    *total = tot_all;
    return c->h_out.as<b200aln_rec_t>();
}

extern "C" const void *b200aln_batch_sai(b200aln_ctx *c, int n_reads, const int32_t *lens, const int64_t *offs,
                                         const uint8_t *codes, const b200aln_opt_t *opt, int64_t *n_bytes)
{
    CK(cudaSetDevice(c->device));
    CallInFlight in_flight(c);
    *n_bytes = 0;
    if (n_reads <= 0) return c->h_out.p;
    int max_len = 0;
    for (int r = 0; r < n_reads; ++r) {
        if (lens[r] < 0) die("b200aln_batch_sai", "read %d has negative length %d.", r, lens[r]);
        if (lens[r] > max_len) max_len = lens[r];
    }
    Params P;
    std::vector<int> md;
    if (c->batch_max_len > 0 && c->batch_max_len < max_len)
        die("b200aln_batch_sai", "batch_max_len %d is smaller than a read of this shard (%d).", c->batch_max_len, max_len);
    b2host::make_params(*opt, c->batch_max_len > 0 ? c->batch_max_len : max_len, lens, n_reads, P, md);
    if ((int)md.size() < max_len + 1) md.resize((size_t)max_len + 1, opt->max_diff);
    const bool pinned_in = host_pinned(lens) && host_pinned(offs) && host_pinned(codes);
    const int64_t tot = run_chunk_host(c, n_reads, lens, offs, codes, opt, P, md, pinned_in);
    const size_t bytes = (size_t)n_reads * 4 + (size_t)tot * 16;
    const size_t reserve = (size_t)c->reserve_reads * 28; /* (a driver whose launches vary in size: sized once) */
    c->sai.need(bytes > reserve ? bytes : reserve);
    c->h_out.need(bytes > reserve ? bytes : reserve);
    k_sai_pack<<<c->n_sm * 4, 256, 0, c->st>>>(n_reads, c->n_aln.as<int32_t>(), c->off64.as<int64_t>(), c->packed.as<Rec>(),
                                               c->sai.as<uint32_t>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(c->h_out.p, c->sai.p, bytes, cudaMemcpyDeviceToHost, c->st));
    CK(cudaEventRecord(c->ev[6], c->st));
    wait_stream(c, c->st);
    finish_stats(c, true);
    c->stats.kernel_launches += 1;
    *n_bytes = (int64_t)bytes;
    return c->h_out.p;
}

/* page-locks caller memory (and releases it) so that b200aln_batch / b200aln_batch_sai copy straight from it */
extern "C" int b200aln_pin(void *p, size_t bytes)
{
    if (!p || !bytes) return 0;
    const cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return -1; }
    return 0;
}
extern "C" void b200aln_unpin(void *p)
{
    if (p && cudaHostUnregister(p) != cudaSuccess) (void)cudaGetLastError();
}

extern "C" void b200aln_batch_device(b200aln_ctx *c, int n_reads, int max_len, const int32_t *d_lens,
                                     const int64_t *d_offs, const uint8_t *d_codes, const b200aln_opt_t *opt,
                                     const int32_t **d_n_aln, const b200aln_rec_t **d_recs, int64_t *total)
{
    CK(cudaSetDevice(c->device));
    CallInFlight in_flight(c);
    Params P;
    std::vector<int> md;
    std::vector<int32_t> one(1, max_len);
    /* all lengths up to max_len are tabulated; reads longer than 1024 bp need the host-buffer entry point */
    if (max_len > 1024) die("b200aln_batch_device", "max_len > 1024 needs b200aln_batch (per-length max_diff table).");
    b2host::make_params(*opt, c->batch_max_len > 0 ? c->batch_max_len : max_len, one.data(), 1, P, md);
    if ((int)md.size() < max_len + 1) md.resize((size_t)max_len + 1, opt->max_diff);
    const int n_chunks = n_reads > 0 ? pipeline_chunks(c, n_reads, c->chunk_reads_device) : 1;
    if (n_chunks == 1) {
        int64_t tot = 0;
        run_batch_device(c, n_reads, max_len, d_lens, d_offs, d_codes, opt, md, P, &tot);
        wait_stream(c, c->st);
        finish_stats(c, false);
        *d_n_aln = c->n_aln.as<int32_t>();
        *d_recs = c->packed.as<b200aln_rec_t>();
        *total = tot;
        return;
    }
    ensure_slots(c);
    const auto t0 = std::chrono::steady_clock::now();
    const int per = (n_reads + n_chunks - 1) / n_chunks, n_workers = c->slots < n_chunks ? c->slots : n_chunks;
    c->asm_n_aln.need((size_t)n_reads * 4);
    if (c->asm_packed.cap < (size_t)n_reads * 24) c->asm_packed.need((size_t)n_reads * 24);
    ChunkOrder ord(n_chunks);
    b200aln_stats_t acc;
    memset(&acc, 0, sizeof acc);
    std::mutex acc_mu;
    auto worker = [&](int si) {
        b200aln_ctx *s = si == 0 ? c : c->slot[(size_t)si - 1];
        CK(cudaSetDevice(s->device));
        for (;;) {
            const int ci = ord.next.fetch_add(1);
            if (ci >= n_chunks) break;
            const int lo = ci * per, n = (lo + per < n_reads ? lo + per : n_reads) - lo;
            int64_t tot = 0;
            run_batch_device(s, n, max_len, d_lens + lo, d_offs + lo, d_codes, opt, md, P, &tot);
            const int64_t off = ord.report(ci, tot);
            if ((size_t)(off + tot) * 16 > c->asm_packed.cap) {
                std::unique_lock<std::shared_mutex> g(ord.grow);
                grow_dev_keep(c->asm_packed, (size_t)(off + tot) * 16);
            }
            {
                std::shared_lock<std::shared_mutex> g(ord.grow);
                CK(cudaMemcpyAsync(c->asm_n_aln.as<int32_t>() + lo, s->n_aln.p, (size_t)n * 4, cudaMemcpyDeviceToDevice, s->st));
                if (tot) CK(cudaMemcpyAsync(c->asm_packed.as<b200aln_rec_t>() + off, s->packed.p, (size_t)tot * 16, cudaMemcpyDeviceToDevice, s->st));
                wait_stream(s, s->st);
            }
            finish_stats(s, false);
            std::lock_guard<std::mutex> lk(acc_mu);
            add_stats(acc, s->stats);
        }
    };
    std::vector<std::thread> th;
    for (int si = 1; si < n_workers; ++si) th.emplace_back(worker, si);
    worker(0);
    for (auto &t : th) t.join();
    int64_t tot_all = 0;
    for (int64_t t : ord.tot) tot_all += t;
    acc.ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    c->stats = acc;
    *d_n_aln = c->asm_n_aln.as<int32_t>();
    *d_recs = c->asm_packed.as<b200aln_rec_t>();
    *total = tot_all;
}

/* ---- row N2: SA row -> position ---------------------------------------------------------------------- */

extern "C" void b200aln_sa_load(b200aln_ctx *c, int which, const b200aln_sa_view_t *v)
{ /* bwt_restore_sa (bwtio.c:29-49) onto the device */
    CK(cudaSetDevice(c->device));
    if (which < 0 || which > 1) die("b200aln_sa_load", "which must be 0 (.sa) or 1 (.rsa).");
    if (c->parent) die("b200aln_sa_load", "load the suffix arrays on the context that owns the index (clones see them).");
    if (v->primary != c->fm[which].primary) die("b200aln_sa_load", "SA-BWT inconsistency: primary is not the same.");
    if (v->seq_len != c->fm[which].seq_len) die("b200aln_sa_load", "SA-BWT inconsistency: seq_len is not the same.");
    if (v->sa_intv <= 0) die("b200aln_sa_load", "bad SA interval.");
    const uint64_t n_sa = ((uint64_t)v->seq_len + (uint64_t)v->sa_intv) / (uint64_t)v->sa_intv;
    if (v->n_sa != n_sa) die("b200aln_sa_load", "SA has %llu samples, expected %llu.", (unsigned long long)v->n_sa, (unsigned long long)n_sa);
    if (c->d_sa[which]) CK(cudaFree(c->d_sa[which]));
    CK(cudaMalloc(&c->d_sa[which], n_sa * 4));
    CK(cudaMemcpy(c->d_sa[which], v->sa, n_sa * 4, cudaMemcpyHostToDevice));
    c->sa_intv[which] = (uint32_t)v->sa_intv;
    c->n_sa[which] = n_sa;
}

extern "C" void b200aln_bwt_sa(b200aln_ctx *c, int which, int64_t n, const uint32_t *rows, uint32_t *pos)
{
    CK(cudaSetDevice(c->device));
    const b200aln_ctx *o = c->parent ? c->parent : c; /* suffix arrays live with the owner of the index */
    if (which < 0 || which > 1 || !o->d_sa[which]) die("b200aln_bwt_sa", "no suffix array loaded for index %d.", which);
    if (n <= 0) return;
    c->sa_in.need((size_t)n * 4);
    c->sa_out.need((size_t)n * 8);
    CK(cudaMemcpyAsync(c->sa_in.p, rows, (size_t)n * 4, cudaMemcpyHostToDevice, c->st));
    k_bwt_sa<<<c->n_sm * 8, 256, 0, c->st>>>(c->fm[which], o->d_sa[which], o->sa_intv[which], n, c->sa_in.as<uint32_t>(),
                                            c->sa_out.as<uint32_t>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(pos, c->sa_out.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->st));
    wait_stream(c, c->st);
}

extern "C" void b200aln_sa2seq(b200aln_ctx *c, int64_t n, const uint8_t *strand, const uint32_t *rows, const int32_t *lens,
                               uint64_t *pos)
{
    CK(cudaSetDevice(c->device));
    const b200aln_ctx *o = c->parent ? c->parent : c;
    if (!o->d_sa[0] || !o->d_sa[1]) die("b200aln_sa2seq", "both suffix arrays (.sa and .rsa) must be loaded.");
    if (n <= 0) return;
    c->sa_in.need((size_t)n * 9 + 64);
    c->sa_out.need((size_t)n * 8);
    uint32_t *d_rows = c->sa_in.as<uint32_t>();
    int32_t *d_lens = (int32_t *)(d_rows + n);
    uint8_t *d_strand = (uint8_t *)(d_lens + n);
    CK(cudaMemcpyAsync(d_rows, rows, (size_t)n * 4, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(d_lens, lens, (size_t)n * 4, cudaMemcpyHostToDevice, c->st));
    CK(cudaMemcpyAsync(d_strand, strand, (size_t)n, cudaMemcpyHostToDevice, c->st));
    k_sa2seq<<<c->n_sm * 8, 256, 0, c->st>>>(c->fm[0], o->d_sa[0], c->fm[1], o->d_sa[1], o->sa_intv[0], o->sa_intv[1], n,
                                            d_strand, d_rows, d_lens, c->sa_out.as<uint64_t>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(pos, c->sa_out.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->st));
    wait_stream(c, c->st);
}

static void scan_i32(b200aln_ctx *c, const int32_t *in, int n, int64_t *out, int64_t *total_dev)
{ /* exclusive prefix sum on the engine's stream (the three scan kernels of the batch path) */
    const int nscan = (n + SCAN_ITEMS - 1) / SCAN_ITEMS;
    c->blk_tot.need((size_t)nscan * 8 + 8);
    k_scan_totals<<<nscan, 256, 0, c->st>>>(in, n, c->blk_tot.as<int64_t>(), nullptr);
    k_scan_blocks<<<1, 32, 0, c->st>>>(c->blk_tot.as<int64_t>(), nscan, total_dev);
    k_scan_apply<<<nscan, 256, 0, c->st>>>(in, n, c->blk_tot.as<int64_t>(), out);
    CK(cudaGetLastError());
}

extern "C" int64_t b200aln_alngrp_merge(b200aln_ctx *c, int n_streams, int n_reads, const int32_t *const *n_aln,
                                        const b200aln_rec_t *const *recs, int s_mm, int64_t *out_off, int32_t *out_n,
                                        b200aln_rec_t *out_recs, uint32_t *out_dbidx)
{
    CK(cudaSetDevice(c->device));
    if (n_streams < 1 || n_streams > B2_MAX_STREAMS) die("b200aln_alngrp_merge", "1 to %d streams, got %d.", B2_MAX_STREAMS, n_streams);
    if (n_reads <= 0) return 0;
    std::vector<int64_t> tot(n_streams, 0);
    int64_t all = 0;
    for (int s = 0; s < n_streams; ++s) {
        for (int r = 0; r < n_reads; ++r) {
            if (n_aln[s][r] < 0) die("b200aln_alngrp_merge", "stream %d read %d: negative count.", s, r);
            tot[s] += n_aln[s][r];
        }
        all += tot[s];
    }
    /* device input: per stream [n_aln | offsets | records], then the per-read totals; output: offsets | n | records | db */
    const size_t nr = (size_t)n_reads;
    size_t in_bytes = 0;
    std::vector<size_t> at_n(n_streams), at_off(n_streams), at_rec(n_streams);
    auto take = [&](size_t bytes) { size_t at = in_bytes; in_bytes += (bytes + 63) & ~(size_t)63; return at; };
    for (int s = 0; s < n_streams; ++s) {
        at_n[s] = take(nr * 4);
        at_off[s] = take(nr * 8);
        at_rec[s] = take((size_t)tot[s] * 16 + 16);
    }
    const size_t at_tot = take(nr * 4), at_sum = take(64);
    c->grp_in.need(in_bytes);
    size_t out_bytes = 0;
    auto take_o = [&](size_t bytes) { size_t at = out_bytes; out_bytes += (bytes + 63) & ~(size_t)63; return at; };
    const size_t o_off = take_o(nr * 8), o_n = take_o(nr * 4), o_rec = take_o((size_t)all * 16 + 16), o_db = take_o((size_t)all * 4 + 16);
    c->grp_out.need(out_bytes);
    char *din = c->grp_in.as<char>(), *dout = c->grp_out.as<char>();
    GrpArgs A;
    A.n_streams = n_streams; A.s_mm = s_mm; A.n_reads = n_reads;
    for (int s = 0; s < n_streams; ++s) {
        CK(cudaMemcpyAsync(din + at_n[s], n_aln[s], nr * 4, cudaMemcpyHostToDevice, c->st));
        if (tot[s]) CK(cudaMemcpyAsync(din + at_rec[s], recs[s], (size_t)tot[s] * 16, cudaMemcpyHostToDevice, c->st));
        A.n_aln[s] = (const int32_t *)(din + at_n[s]);
        A.rec_off[s] = (const int64_t *)(din + at_off[s]);
        A.recs[s] = (const Rec *)(din + at_rec[s]);
        scan_i32(c, A.n_aln[s], n_reads, (int64_t *)(din + at_off[s]), (int64_t *)(din + at_sum));
    }
    A.out_off = (const int64_t *)(dout + o_off); A.out_n = (int32_t *)(dout + o_n);
    A.out_rec = (Rec *)(dout + o_rec); A.out_db = (uint32_t *)(dout + o_db);
    k_alngrp_totals<<<c->n_sm * 4, 128, 0, c->st>>>(A, (int32_t *)(din + at_tot));
    CK(cudaGetLastError());
    scan_i32(c, (const int32_t *)(din + at_tot), n_reads, (int64_t *)(dout + o_off), (int64_t *)(din + at_sum));
    k_alngrp<<<c->n_sm * 8, 128, 0, c->st>>>(A);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_off, dout + o_off, nr * 8, cudaMemcpyDeviceToHost, c->st));
    CK(cudaMemcpyAsync(out_n, dout + o_n, nr * 4, cudaMemcpyDeviceToHost, c->st));
    if (all) {
        CK(cudaMemcpyAsync(out_recs, dout + o_rec, (size_t)all * 16, cudaMemcpyDeviceToHost, c->st));
        CK(cudaMemcpyAsync(out_dbidx, dout + o_db, (size_t)all * 4, cudaMemcpyDeviceToHost, c->st));
    }
    wait_stream(c, c->st);
    return all;
}

extern "C" double b200aln_sector_roofline(b200aln_ctx *c, uint64_t n_loads, int repeats)
{
    /* repeats > 0: independent random 32-byte sectors; repeats < 0: random 64-byte pairs (|repeats| runs).
     * Returns GB/s of the bytes asked for (best run). */
    CK(cudaSetDevice(c->device));
    c->misc.need(sizeof(Misc));
    const int span = repeats < 0 ? 2 : 1;
    if (repeats < 0) repeats = -repeats;
    const int blocks = c->n_sm * 8, threads = 256;
    uint64_t per = (n_loads / ((uint64_t)blocks * threads) + 7) / 8 * 8;
    if (per < 8) per = 8;
    double best = 0;
    for (int it = 0; it < repeats + 1; ++it) {
        CK(cudaEventRecord(c->ev[0], c->st));
        k_sector_gather<<<blocks, threads, 0, c->st>>>(c->d_idx[0], c->n_blk[0], c->d_idx[1], c->n_blk[1], per, span,
                                                       (unsigned long long *)c->misc.p);
        CK(cudaGetLastError());
        CK(cudaEventRecord(c->ev[1], c->st));
        CK(cudaEventSynchronize(c->ev[1]));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
        double gbs = (double)per * blocks * threads * 32.0 * span / (ms * 1e-3) / 1e9;
        if (it > 0 && gbs > best) best = gbs;
    }
    return best;
}
