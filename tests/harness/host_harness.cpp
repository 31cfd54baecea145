// host_harness.cpp — TEST INFRASTRUCTURE: compiles the product's per-read state
// machines (ibwa_b200/csrc/aln_core.cuh, fm_layout.cuh) with g++ and runs them
// read by read on the CPU, so that tests can check the kernel LOGIC against the
// oracle in a container without a GPU.  Not part of libb200aln.so and never
// used by the product path.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifdef B2_DEBUG_COUNTS
#include <cstdint>
namespace b2 { uint64_t b2_dbg[16]; }
extern "C" uint64_t *hh_dbg() { return b2::b2_dbg; }
#endif
#include "../../include/b200aln.h"
#include "../../ibwa_b200/csrc/aln_core.cuh"
#include "../../ibwa_b200/csrc/alngrp_core.cuh"
#include "../../ibwa_b200/csrc/fm_layout.cuh"
#include "../../ibwa_b200/csrc/host_params.h"

using namespace b2;

static int g_rounds = 1 << 30; /* SearchLane::prepare's max_rounds (the fast kernel runs with 1) */
extern "C" void hh_set_rounds(int r) { g_rounds = r > 0 ? r : 1 << 30; }

static std::vector<OccBlk> convert(const b200aln_bwt_view_t *v)
{
    RefBwt r;
    r.w = v->bwt; r.n_words = v->bwt_size; r.seq_len = v->seq_len;
    for (int c = 0; c < 4; ++c) r.L2[c] = v->L2[c];
    uint64_t nb = fm_num_blocks(v->seq_len);
    std::vector<OccBlk> out(nb);
    for (uint64_t b = 0; b < nb; ++b) out[b] = fm_convert_block(r, b);
    return out;
}

template <class Heads> static Heads make_heads(std::vector<uint32_t> &store);
template <> HeadsStrided16 make_heads<HeadsStrided16>(std::vector<uint32_t> &store)
{
    HeadsStrided16 hd;
    hd.h = reinterpret_cast<uint16_t *>(store.data());
    hd.stride = 1;
    return hd;
}
template <> HeadsWide32 make_heads<HeadsWide32>(std::vector<uint32_t> &store)
{
    HeadsWide32 hd;
    hd.h = store.data();
    return hd;
}

static int g_susp = 0; /* > 0: park the lane every g_susp steps and resume it in a fresh SearchLane (save_state / load_state) */
extern "C" void hh_set_suspend_every(int n) { g_susp = n; }

/* runs `lane` to the end; with g_susp the search is interrupted every g_susp steps exactly as the kernel parks a
 * straggler: state words, bucket heads and open group copied out, the originals scribbled over, a new lane loaded */
template <class Lane, class Heads>
static void run_lane(Lane &lane, const SearchEnv &e, Heads heads, std::vector<uint32_t> &hstore, GroupStore gs, int n_buckets)
{
    if (g_susp <= 0) {
        while (!lane.finished) lane.step(e, g_rounds);
        return;
    }
    std::vector<uint32_t> words(B2_SAVE_WORDS), hcopy, gcopy(OG_WORDS);
    int steps = 0;
    while (!lane.finished) {
        lane.step(e, g_rounds);
        if (++steps % g_susp == 0 && !lane.finished) {
            lane.save_state(words.data());
            hcopy.assign(hstore.begin(), hstore.end());
            for (int w = 0; w < OG_WORDS; ++w) gcopy[w] = gs.get(w);
            for (auto &x : hstore) x = 0xA5A5A5A5u;
            for (int w = 0; w < OG_WORDS; ++w) gs.set(w, 0xDEADBEEFu);
            Lane fresh;
            memset((void *)&fresh, 0x5A, sizeof fresh);
            fresh.load_state(words.data());
            fresh.bk = heads;
            fresh.gs = gs;
            hstore.assign(hcopy.begin(), hcopy.end());
            for (int w = 0; w < OG_WORDS; ++w) gs.set(w, gcopy[w]);
            lane = fresh;
        }
    }
    (void)n_buckets;
}

static int g_q16 = 0; /* 1: the first pass uses the 16-bit width records when the options allow them (like the product) */
extern "C" void hh_set_q16(int v) { g_q16 = v; }
static int g_last_q16 = 0; /* whether the last hh_aln_batch call really ran on the 16-bit records */
extern "C" int hh_last_q16() { return g_last_q16; }

template <class Heads, bool REUSE, int QB>
static int64_t run(const SearchEnv &env, const std::vector<int> &md, int n_reads, const int32_t *lens,
                   const int64_t *offs, const uint8_t *codes, bool comp, int seed_len, uint32_t arena_cap, int rec_cap,
                   int32_t *n_aln, std::vector<Rec> &all, uint64_t *counters, uint32_t big_cap)
{
    int max_len = 0;
    for (int r = 0; r < n_reads; ++r) if (lens[r] > max_len) max_len = lens[r];
    typedef typename QF<QB>::T QT;
    const int strideQ = (max_len + 15) & ~15, strideW = round_up8(max_len + 1);
    std::vector<uint32_t> Q(2 * (size_t)strideQ + 8); /* room for either format */
    QT *Qt = reinterpret_cast<QT *>(Q.data());
    std::vector<uint32_t> W(2 * (size_t)strideW + 8);
    std::vector<StackRec> ent(arena_cap);
    std::vector<Rec> recs(rec_cap);
    int64_t n_status = 0;
    std::vector<uint32_t> hstore(2048 + 64);
    uint32_t gstore[OG_WORDS];
    GroupStore gs; gs.p = gstore; gs.stride = 1;
    std::vector<StackRec> ent2;
    std::vector<Rec> recs2;
    for (int r = 0; r < n_reads; ++r) {
        const uint8_t *fwd = codes + offs[r];
        int len = lens[r];
        const FmView *fm = env.fm;
        int n_amb = width_pass<QB>(fm[0], fwd, len, 0, comp, seed_len, W.data(), Qt).n_amb;
        width_pass<QB>(fm[1], fwd, len, 1, comp, seed_len, W.data() + strideW, Qt + strideQ);
        SearchLane<Heads, REUSE, true, QB> lane;
        SearchEnv e1 = env;
        e1.Q = Q.data(); e1.W = W.data(); e1.strideQ = strideQ; e1.strideW = strideW;
        e1.recs = recs.data(); e1.rec_cap = rec_cap; e1.ent = ent.data(); e1.arena_cap = arena_cap;
        lane.begin(e1, make_heads<Heads>(hstore), gs, 0, 0, 0, len, md[len], n_amb);
        run_lane(lane, e1, make_heads<Heads>(hstore), hstore, gs, env.P.n_buckets);
        if (lane.status != LANE_OK && big_cap) {
            /* the product's large pass: widths rebuilt (the aborted pass shadowed them), free-list arena */
            ++n_status;
            if (ent2.size() < big_cap) { ent2.resize(big_cap); recs2.resize(1 << 16); }
            n_amb = width_pass<32>(fm[0], fwd, len, 0, comp, seed_len, W.data(), Q.data()).n_amb;
            width_pass<32>(fm[1], fwd, len, 1, comp, seed_len, W.data() + strideW, Q.data() + strideQ);
            SearchLane<HeadsWide32, true> big;
            SearchEnv e2 = e1;
            e2.recs = recs2.data(); e2.rec_cap = 1 << 16; e2.ent = ent2.data(); e2.arena_cap = big_cap;
            big.begin(e2, make_heads<HeadsWide32>(hstore), gs, 0, 0, 0, len, md[len], n_amb);
            run_lane(big, e2, make_heads<HeadsWide32>(hstore), hstore, gs, env.P.n_buckets);
            if (big.status != LANE_OK) { n_aln[r] = -big.status; continue; }
            n_aln[r] = big.n_aln;
            all.insert(all.end(), recs2.begin(), recs2.begin() + big.n_aln);
            if (counters) { counters[0] += big.n_pops; counters[1] += big.n_lookups; }
            continue;
        }
        if (lane.status != LANE_OK) { ++n_status; n_aln[r] = -lane.status; continue; }
        n_aln[r] = lane.n_aln;
        all.insert(all.end(), recs.begin(), recs.begin() + lane.n_aln);
        if (counters) { counters[0] += lane.n_pops; counters[1] += lane.n_lookups; }
    }
    return n_status;
}

static std::vector<uint32_t> build_lut(FmView fm[2], int lut_k)
{
    std::vector<uint32_t> lut(2 * lut_total_pairs(lut_k) + 8);
    for (int w = 0; w < 2; ++w) {
        fm[w].lut = nullptr;
        fm[w].lut_k = 0;
        fm[w].lut_w = w;
        for (int level = 0; level < lut_k; ++level)
            for (uint64_t X = 0; X < ((uint64_t)1 << (2 * level)); ++X) lut_build_node(fm[w], lut.data(), level, X);
    }
    return lut;
}

static int g_lut_k = 0;
extern "C" void hh_set_lut_k(int k) { g_lut_k = k; }

extern "C" int64_t hh_aln_batch(const b200aln_bwt_view_t *bwt, const b200aln_bwt_view_t *rbwt, int n_reads,
                                const int32_t *lens, const int64_t *offs, const uint8_t *codes,
                                const b200aln_opt_t *opt, uint32_t arena_cap, int rec_cap, int reuse, uint32_t big_cap,
                                int batch_max_len, int32_t *n_aln,
                                Rec **records, int64_t *n_overflow, uint64_t *counters)
{
    std::vector<OccBlk> i0 = convert(bwt), i1 = convert(rbwt);
    SearchEnv env;
    env.prefetch_next = 0;
    FmView *fm = env.fm;
    fm[0].blk = i0.data(); fm[0].primary = bwt->primary; fm[0].seq_len = bwt->seq_len;
    fm[1].blk = i1.data(); fm[1].primary = rbwt->primary; fm[1].seq_len = rbwt->seq_len;
    std::vector<uint32_t> lut = build_lut(fm, g_lut_k);
    fm[0].lut = fm[1].lut = lut.data();
    fm[0].lut_k = fm[1].lut_k = g_lut_k;
    int max_len = 0;
    for (int r = 0; r < n_reads; ++r) if (lens[r] > max_len) max_len = lens[r];
    Params &P = env.P;
    std::vector<int> md;
    b2host::make_params(*opt, batch_max_len > 0 ? batch_max_len : max_len, lens, n_reads, P, md);
    std::vector<Rec> all;
    bool comp = opt->mode & MODE_COMPREAD;
    int64_t ov;
    int md_max = 0;
    for (int l = 0; l <= max_len && l < (int)md.size(); ++l) md_max = md[l] > md_max ? md[l] : md_max;
    const bool q16 = g_q16 && md_max < 7 && P.max_seed_diff < 3;
    g_last_q16 = q16;
#define HH_RUN(H, R, QB) run<H, R, QB>(env, md, n_reads, lens, offs, codes, comp, opt->seed_len, arena_cap, rec_cap, n_aln, all, counters, big_cap)
    if (P.n_buckets <= 128 && arena_cap <= 32767) {
        if (q16) ov = reuse ? HH_RUN(HeadsStrided16, true, 16) : HH_RUN(HeadsStrided16, false, 16);
        else ov = reuse ? HH_RUN(HeadsStrided16, true, 32) : HH_RUN(HeadsStrided16, false, 32);
    } else {
        if (q16) ov = reuse ? HH_RUN(HeadsWide32, true, 16) : HH_RUN(HeadsWide32, false, 16);
        else ov = reuse ? HH_RUN(HeadsWide32, true, 32) : HH_RUN(HeadsWide32, false, 32);
    }
#undef HH_RUN
    *n_overflow = ov;
    Rec *out = (Rec *)malloc(sizeof(Rec) * (all.size() + 1));
    memcpy(out, all.data(), sizeof(Rec) * all.size());
    *records = out;
    return (int64_t)all.size();
}

/* row N2: the device inv_psi / sa_of_row code on the CPU */
extern "C" void hh_bwt_sa(const b200aln_bwt_view_t *bwt, const uint32_t *sa, uint32_t sa_intv, int64_t n,
                          const uint32_t *rows, uint32_t *out)
{
    std::vector<OccBlk> idx = convert(bwt);
    FmView f;
    f.blk = idx.data(); f.primary = bwt->primary; f.seq_len = bwt->seq_len; f.lut = nullptr; f.lut_k = 0; f.lut_w = 0;
    for (int64_t i = 0; i < n; ++i) out[i] = sa_of_row(f, sa, sa_intv, rows[i]);
}

/* row N4: the device merge code (alngrp_core.cuh) on the CPU, read by read */
extern "C" int64_t hh_alngrp_merge(int n_streams, int n_reads, const int32_t *const *n_aln, const Rec *const *recs, int s_mm,
                                   int64_t *out_off, int32_t *out_n, Rec *out_recs, uint32_t *out_db)
{
    std::vector<std::vector<int64_t>> off(n_streams, std::vector<int64_t>(n_reads));
    std::vector<const int64_t *> offp(n_streams);
    std::vector<int32_t> tot(n_reads, 0);
    for (int s = 0; s < n_streams; ++s) {
        int64_t at = 0;
        for (int r = 0; r < n_reads; ++r) { off[s][r] = at; at += n_aln[s][r]; tot[r] += n_aln[s][r]; }
        offp[s] = off[s].data();
    }
    int64_t at = 0;
    for (int r = 0; r < n_reads; ++r) {
        out_off[r] = at;
        out_n[r] = alngrp_merge_one(n_streams, r, n_aln, offp.data(), recs, s_mm, at, out_recs, out_db);
        at += tot[r];
    }
    return at;
}

extern "C" void hh_free(void *p) { free(p); }
