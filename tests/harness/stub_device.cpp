// stub_device.cpp — TEST INFRASTRUCTURE: stands in for the CUDA half of the library (ibwa_b200/csrc/b200aln.cu)
// underneath the host driver (ibwa_b200/csrc/aln_host.cpp), so that the driver's own logic — parse units, launches
// cut at the batch-level clamp, worker threads, output order — can be tested in a container without a GPU.  A batch
// is computed by the CPU build of the product's state machines (host_harness.cpp: hh_aln_batch).  Linked only into
// tests/harness/libdriverstub.so; never part of libb200aln.so and never used by the product path.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "../../include/b200aln.h"
#include "../../ibwa_b200/csrc/host_params.h"

struct HhRec { uint32_t packed, k, l; int32_t score; };
extern "C" int64_t hh_aln_batch(const b200aln_bwt_view_t *bwt, const b200aln_bwt_view_t *rbwt, int n_reads,
                                const int32_t *lens, const int64_t *offs, const uint8_t *codes,
                                const b200aln_opt_t *opt, uint32_t arena_cap, int rec_cap, int reuse, uint32_t big_cap,
                                int batch_max_len, int32_t *n_aln, HhRec **records, int64_t *n_overflow,
                                uint64_t *counters);
extern "C" void hh_free(void *p);

struct StubIndex {
    std::vector<uint32_t> w[2];
    b200aln_bwt_view_t v[2];
};
struct b200aln_ctx {
    std::shared_ptr<StubIndex> idx;
    int batch_max_len = 0;
    int device = 0;
    std::vector<char> out;
    std::vector<int32_t> n_aln;
};

extern "C" const char *b200aln_version(void) { return "stub"; }
extern "C" void b200aln_opt_init(b200aln_opt_t *o)
{
    memset(o, 0, sizeof *o);
    o->s_mm = 3; o->s_gapo = 11; o->s_gape = 4;
    o->max_diff = -1; o->max_gapo = 1; o->max_gape = 6;
    o->indel_end_skip = 5; o->max_del_occ = 10; o->max_entries = 2000000;
    o->mode = 0x01 | 0x02;
    o->seed_len = 32; o->max_seed_diff = 2;
    o->fnr = 0.04f;
    o->n_threads = 1;
    o->max_top2 = 30;
}
extern "C" int b200aln_cal_maxdiff(int len, double err, double thres) { return b2host::cal_maxdiff(len, err, thres); }
extern "C" int b200aln_device_count(void)
{
    const char *e = getenv("B200ALN_STUB_DEVICES");
    return e ? atoi(e) : 1;
}
extern "C" void b200aln_warm_device(int) {}
extern "C" int b200aln_pin(void *, size_t) { return -1; }
extern "C" void b200aln_unpin(void *) {}
extern "C" void b200aln_prealloc(int, int, int, int) {}
extern "C" void b200aln_prealloc_release(int) {}

extern "C" b200aln_ctx *b200aln_open(const b200aln_bwt_view_t *bwt, const b200aln_bwt_view_t *rbwt, int device)
{
    b200aln_ctx *c = new b200aln_ctx;
    c->idx.reset(new StubIndex);
    const b200aln_bwt_view_t *src[2] = {bwt, rbwt};
    for (int j = 0; j < 2; ++j) { /* the driver frees its copy of the files right after the open */
        c->idx->w[j].assign(src[j]->bwt, src[j]->bwt + src[j]->bwt_size);
        c->idx->v[j] = *src[j];
        c->idx->v[j].bwt = c->idx->w[j].data();
    }
    c->device = device;
    return c;
}
extern "C" b200aln_ctx *b200aln_clone(b200aln_ctx *p)
{
    b200aln_ctx *c = new b200aln_ctx;
    c->idx = p->idx;
    c->device = p->device;
    return c;
}
extern "C" void b200aln_close(b200aln_ctx *c) { delete c; }
extern "C" void b200aln_last_stats(const b200aln_ctx *, b200aln_stats_t *out) { memset(out, 0, sizeof *out); }
extern "C" void b200aln_set_int(b200aln_ctx *c, const char *key, int64_t v)
{
    if (!strcmp(key, "batch_max_len")) c->batch_max_len = (int)v;
}

static int64_t run(b200aln_ctx *c, int n, const int32_t *lens, const int64_t *offs, const uint8_t *codes,
                   const b200aln_opt_t *opt, int32_t *n_aln, HhRec **rec)
{
    int64_t nov = 0;
    uint64_t counters[2] = {0, 0};
    const int64_t tot = hh_aln_batch(&c->idx->v[0], &c->idx->v[1], n, lens, offs, codes, opt, 1u << 15, 1 << 12, 1,
                                     1u << 21, c->batch_max_len, n_aln, rec, &nov, counters);
    for (int r = 0; r < n; ++r)
        if (n_aln[r] < 0) {
            fprintf(stderr, "[stub_device] read %d does not fit the harness arenas\n", r);
            abort();
        }
    return tot;
}

extern "C" const b200aln_rec_t *b200aln_batch(b200aln_ctx *c, int n, const int32_t *lens, const int64_t *offs,
                                              const uint8_t *codes, const b200aln_opt_t *opt, int32_t *n_aln,
                                              int64_t *total)
{
    HhRec *rec = nullptr;
    *total = n > 0 ? run(c, n, lens, offs, codes, opt, n_aln, &rec) : 0;
    c->out.resize((size_t)*total * 16 + 16);
    if (*total) memcpy(c->out.data(), rec, (size_t)*total * 16);
    if (rec) hh_free(rec);
    return reinterpret_cast<const b200aln_rec_t *>(c->out.data());
}

extern "C" const void *b200aln_batch_sai(b200aln_ctx *c, int n, const int32_t *lens, const int64_t *offs,
                                         const uint8_t *codes, const b200aln_opt_t *opt, int64_t *n_bytes)
{
    *n_bytes = 0;
    if (n <= 0) return c->out.data();
    HhRec *rec = nullptr;
    c->n_aln.assign((size_t)n, 0);
    const int64_t tot = run(c, n, lens, offs, codes, opt, c->n_aln.data(), &rec);
    c->out.resize((size_t)n * 4 + (size_t)tot * 16);
    char *w = c->out.data();
    const HhRec *p = rec;
    for (int r = 0; r < n; ++r) { /* bwtaln.c:227-231 */
        memcpy(w, &c->n_aln[(size_t)r], 4);
        w += 4;
        memcpy(w, p, (size_t)c->n_aln[(size_t)r] * 16);
        w += (size_t)c->n_aln[(size_t)r] * 16;
        p += c->n_aln[(size_t)r];
    }
    hh_free(rec);
    *n_bytes = (int64_t)c->out.size();
    return c->out.data();
}
