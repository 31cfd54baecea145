/*
 * fm_layout.cuh — re-layout of the reference's in-memory BWT (bwt.h:42-63,
 * producer bwtmisc.c:122-144: per 128 bases 4 cumulative-count words + 8 words
 * of 16 bases, 2 bits each, MSB first; last block truncated; one trailing
 * 4-word count block) into the device layout of aln_core.cuh (one 32-byte
 * block per 64 bases: counts with L2 pre-added + two bit planes).
 *
 * Compiled by nvcc (conversion kernel in b200aln.cu) and by g++ (tests/harness).
 */
#pragma once
#include "aln_core.cuh"

namespace b2 {

struct RefBwt { /* the reference's bwt_t, as a view */
    const uint32_t *w;
    uint64_t n_words;
    uint32_t seq_len;
    uint32_t L2[4]; /* L2[0..3] (L2[0] == 0) */
};

B2_HD uint64_t fm_num_blocks(uint32_t seq_len) { return (uint64_t)(seq_len >> 6) + 1u; }

/* spread the 16 bases of a reference word (MSB first) into plane bits [sh, sh+16) */
B2_HD void fm_spread16(uint32_t word, int sh, uint32_t &lo, uint32_t &hi)
{
    for (int j = 0; j < 16; ++j) {
        uint32_t s = word >> (30 - 2 * j) & 3u;
        lo |= (s & 1u) << (sh + j);
        hi |= (s >> 1) << (sh + j);
    }
}

/* device block `b` (bases [64b, 64b+64)) */
B2_HD OccBlk fm_convert_block(const RefBwt &r, uint64_t b)
{
    const uint64_t base0 = b * 64u;
    uint32_t cnt[4], wd[4] = {0, 0, 0, 0};
    const uint64_t payload_words = r.n_words - 4; /* words before the trailing count block */
    if (base0 >= r.seq_len) {
        for (int c = 0; c < 4; ++c) cnt[c] = r.w[r.n_words - 4 + c];
    } else {
        const uint64_t rb = base0 >> 7; /* reference block */
        const uint32_t *p = r.w + rb * 12u;
        for (int c = 0; c < 4; ++c) cnt[c] = p[c];
        const int half = (int)(b & 1u);
        if (half) { /* add the first 64 bases of the reference block */
            for (int j = 0; j < 4; ++j) {
                uint32_t word = p[4 + j];
                for (int t = 0; t < 16; ++t) cnt[word >> (30 - 2 * t) & 3u]++;
            }
        }
        for (int j = 0; j < 4; ++j) {
            uint64_t wi = rb * 12u + 4u + (uint64_t)(half * 4 + j);
            uint64_t first_base = rb * 128u + (uint64_t)(half * 4 + j) * 16u;
            if (wi < payload_words && first_base < r.seq_len) wd[j] = r.w[wi];
        }
        /* counts past seq_len inside the half-block were bumped by padding zeros
         * only if the padded words exist; they do not affect occ(q) for q <= seq_len
         * because the mask stops at q. */
    }
    U4 c4, pl;
    c4.x = r.L2[0] + cnt[0]; c4.y = r.L2[1] + cnt[1]; c4.z = r.L2[2] + cnt[2]; c4.w = r.L2[3] + cnt[3];
    uint32_t lo0 = 0, lo1 = 0, hi0 = 0, hi1 = 0;
    fm_spread16(wd[0], 0, lo0, hi0);
    fm_spread16(wd[1], 16, lo0, hi0);
    fm_spread16(wd[2], 0, lo1, hi1);
    fm_spread16(wd[3], 16, lo1, hi1);
    pl.x = lo0; pl.y = lo1; pl.z = hi0; pl.w = hi1;
    OccBlk o;
    o.cnt = c4;
    o.bits = pl;
    return o;
}

} // namespace b2
