/* seam_cal_sa_reg_gap.c — TEST INFRASTRUCTURE, not reference code.
 *
 * The reference-side binding of INTEGRATION.md §3, built as a test: Makefile.ref links the UNMODIFIED
 * reference objects into oracle/_ref/ibwa_seam, with ONE symbol of a copy of bwtaln.o weakened
 * (objcopy --weaken-symbol=bwa_cal_sa_reg_gap) so that the definition below takes its place.  Everything
 * else of `ibwa aln` is the reference's own code: option parsing (bwtaln.c:243-328), bwa_read_seq
 * (bwaseqio.c:145-208), the batch loop and the fwrite loop of bwa_aln_core (bwtaln.c:193-232) and
 * bwa_free_read_seq (bwaseqio.c:210-222).  The batch operator (bwtaln.h:148) is the engine's
 * b200aln_cal_sa_reg_gap, loaded from libb200aln.so at run time.
 *
 * tests/test_seam.py runs `ibwa_seam aln ...` next to `ibwa aln ...` on the GPU box and compares bytes.
 */
#include <dlfcn.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "bwtaln.h"
#include "bwt.h"
#include "../../include/b200aln.h"

/* what a maintainer asserts next to the call */
_Static_assert(sizeof(gap_opt_t) == sizeof(b200aln_opt_t), "gap_opt_t layout");
_Static_assert(sizeof(bwa_seq_t) == 176, "bwa_seq_t layout assumed by b200aln_cal_sa_reg_gap");
_Static_assert(offsetof(bwa_seq_t, seq) == 8 && offsetof(bwa_seq_t, n_aln) == 48 && offsetof(bwa_seq_t, aln) == 56, "bwa_seq_t offsets");
_Static_assert(sizeof(bwt_aln1_t) == sizeof(b200aln_rec_t), "bwt_aln1_t layout");

static b200aln_ctx *(*p_open)(const b200aln_bwt_view_t *, const b200aln_bwt_view_t *, int);
static void (*p_close)(b200aln_ctx *);
static void (*p_cal)(b200aln_ctx *, int, void *, const b200aln_opt_t *);
static void (*p_layout)(b200aln_seq_layout_t *);
static b200aln_ctx *g_ctx;
static const bwt_t *g_bwt[2];

static void seam_close(void)
{
    if (g_ctx) p_close(g_ctx);
    g_ctx = 0;
}

static void seam_load(void)
{
    char path[4096];
    const char *e = getenv("B200ALN_LIB");
    if (e) snprintf(path, sizeof path, "%s", e);
    else { /* <repo>/oracle/_ref/ibwa_seam -> <repo>/ibwa_b200/libb200aln.so */
        ssize_t n = readlink("/proc/self/exe", path, sizeof path - 64);
        if (n <= 0) { fprintf(stderr, "[seam] cannot locate the executable. Abort!\n"); abort(); }
        path[n] = 0;
        char *s = strrchr(path, '/');
        strcpy(s ? s : path, "/../../ibwa_b200/libb200aln.so");
    }
    void *h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) { fprintf(stderr, "[seam] %s. Abort!\n", dlerror()); abort(); }
    p_open = (b200aln_ctx * (*)(const b200aln_bwt_view_t *, const b200aln_bwt_view_t *, int)) dlsym(h, "b200aln_open");
    p_close = (void (*)(b200aln_ctx *))dlsym(h, "b200aln_close");
    p_cal = (void (*)(b200aln_ctx *, int, void *, const b200aln_opt_t *))dlsym(h, "b200aln_cal_sa_reg_gap");
    p_layout = (void (*)(b200aln_seq_layout_t *))dlsym(h, "b200aln_seq_layout");
    if (!p_open || !p_close || !p_cal || !p_layout) { fprintf(stderr, "[seam] libb200aln.so lacks an entry point. Abort!\n"); abort(); }
    b200aln_seq_layout_t lay;
    p_layout(&lay); /* the library's idea of bwa_seq_t against the reference's own header */
    if (lay.size != sizeof(bwa_seq_t) || lay.off_seq != offsetof(bwa_seq_t, seq) || lay.off_rseq != offsetof(bwa_seq_t, rseq) ||
        lay.off_qual != offsetof(bwa_seq_t, qual) || lay.off_n_aln != offsetof(bwa_seq_t, n_aln) ||
        lay.off_aln != offsetof(bwa_seq_t, aln) || lay.off_sa != offsetof(bwa_seq_t, sa)) {
        fprintf(stderr, "[seam] bwa_seq_t layout mismatch between bwtaln.h and libb200aln.so. Abort!\n");
        abort();
    }
}

/* == bwtaln.h:148.  The reference calls it once per batch from the main thread with -t 1 (bwtaln.c:200-201)
 * and once per pthread otherwise (bwtaln.c:151-157): the engine takes the whole batch in thread 0's call. */
void bwa_cal_sa_reg_gap(int tid, bwt_t *const bwt[2], int n_seqs, bwa_seq_t *seqs, const gap_opt_t *opt)
{
    if (tid != 0) return;
    if (!g_ctx || g_bwt[0] != bwt[0] || g_bwt[1] != bwt[1]) {
        if (!p_open) seam_load();
        seam_close();
        b200aln_bwt_view_t v[2];
        int j;
        for (j = 0; j < 2; ++j) { /* bwt[0] = .bwt, bwt[1] = .rbwt as bwt_restore_bwt leaves them (bwtio.c:51-70) */
            v[j].primary = bwt[j]->primary;
            memcpy(v[j].L2, bwt[j]->L2, sizeof v[j].L2);
            v[j].seq_len = bwt[j]->seq_len;
            v[j].bwt_size = bwt[j]->bwt_size;
            v[j].bwt = bwt[j]->bwt;
        }
        const char *d = getenv("B200ALN_DEVICE");
        g_ctx = p_open(&v[0], &v[1], d ? atoi(d) : 0);
        g_bwt[0] = bwt[0];
        g_bwt[1] = bwt[1];
        atexit(seam_close);
    }
    p_cal(g_ctx, n_seqs, seqs, (const b200aln_opt_t *)opt);
}
