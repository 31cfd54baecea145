/*
 * fast_inflate.h — a DEFLATE (RFC 1951) decoder for the read-ingest path (scope row N1).
 *
 * The reference opens every input through zlib (utils.c:56-66, kseq.h on gzread), and zlib's inflate is what a
 * gzip-compressed FASTQ is bound by (≈ 280 MB/s of output on one core).  This decoder produces the same bytes about
 * three times as fast: a 64-bit bit buffer refilled eight bytes at a time, one table lookup per symbol (11-bit
 * primary table for literals / lengths, 8-bit for distances, second-level tables for longer codes), matches copied
 * sixteen bytes at a time.  It is resumable at symbol boundaries, so a stream of any size is decoded chunk by chunk
 * into a caller-owned buffer that keeps the last 32 KB as history.
 *
 * It never decides on its own what a damaged stream means: any code-length set zlib would refuse, any distance
 * beyond the history, any truncation returns FI_ERROR, and the caller hands the stream to zlib from that point
 * (aln_host.cpp: InflateSource): what is delivered of a damaged or truncated stream, and where it ends, is zlib's
 * doing (a member whose CRC or length check fails loses its last chunk, as it does with gzread).
 */
#ifndef B200ALN_FAST_INFLATE_H
#define B200ALN_FAST_INFLATE_H

#include <stdint.h>
#include <string.h>

namespace fastinflate {

enum Status { FI_MORE_OUTPUT = 0, /* the output space is used up; call again with room */
              FI_DONE = 1,        /* the final block has ended */
              FI_ERROR = 2 };     /* not decodable by this decoder (damaged, truncated, or a code zlib refuses) */

enum { LITLEN_BITS = 11, OFF_BITS = 8, MAX_CODE_LEN = 15, N_LITLEN = 288, N_OFF = 32 };
/* table entry: bits 0-3 code length (bits to drop; for a pointer to a second-level table: the primary bits),
 * bits 4-7 kind, bits 8-12 extra bits (or second-level index bits), bits 16-31 value (literal, length base,
 * distance base, or start of the second-level table) */
enum { K_LITERAL = 1, K_LENGTH = 2, K_EOB = 3, K_SUB = 4, K_INVALID = 5, K_DIST = 6,
       K_LIT2 = 7, K_LIT3 = 8 /* two / three literals whose codes fit the primary index together */ };
/* literal entries (K_LITERAL, K_LIT2, K_LIT3) hold their bytes in bits 8-15, 16-23, 24-31 instead */

static inline uint32_t entry(unsigned len, unsigned kind, unsigned extra, unsigned value)
{
    return len | kind << 4 | extra << 8 | value << 16;
}

struct Decoder {
    /* input */
    const uint8_t *in = nullptr, *in_end = nullptr;
    uint64_t bitbuf = 0;
    int bitsleft = 0;
    /* block state */
    int mode = 0; /* 0: a block header comes next, 1: inside a stored block, 2: inside a Huffman block, 3: done */
    bool final_block = false;
    uint32_t stored_left = 0;
    uint32_t litlen[(1 << LITLEN_BITS) + N_LITLEN * 16];
    uint32_t dist[(1 << OFF_BITS) + N_OFF * 128];
    bool static_ready = false;
    uint32_t static_litlen[(1 << LITLEN_BITS) + 16];
    uint32_t static_dist[1 << OFF_BITS];
    bool use_static = false;

    void start(const uint8_t *p, size_t n)
    {
        in = p;
        in_end = p + n;
        bitbuf = 0;
        bitsleft = 0;
        mode = 0;
        final_block = false;
    }
    /* compressed bytes consumed so far (whole bytes still in the bit buffer are not counted) */
    const uint8_t *in_pos() const { return in - (bitsleft >> 3); }

    inline void refill()
    {
        if (in + 8 <= in_end) {
            uint64_t v;
            memcpy(&v, in, 8);
            bitbuf |= v << bitsleft;
            const int n = (63 - bitsleft) >> 3;
            in += n;
            bitsleft += n << 3;
        } else {
            while (bitsleft <= 56 && in < in_end) {
                bitbuf |= (uint64_t)*in++ << bitsleft;
                bitsleft += 8;
            }
        }
    }
    inline void drop(int n)
    {
        bitbuf >>= n;
        bitsleft -= n;
    }

    /* canonical Huffman table (RFC 1951 3.2.2) of `n` symbols with code lengths lens[]; false when the lengths are
     * not a code this decoder takes (zlib: over-subscribed, or incomplete other than a single one-bit distance code) */
    static bool build(const uint8_t *lens, int n, uint32_t *tab, int primary_bits, bool is_dist)
    {
        static const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        static const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
        static const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
        int count[MAX_CODE_LEN + 1] = {0};
        for (int s = 0; s < n; ++s) ++count[lens[s]];
        count[0] = 0;
        int max_len = 0, n_codes = 0;
        long left = 1;
        for (int l = 1; l <= MAX_CODE_LEN; ++l) {
            left = (left << 1) - count[l];
            if (left < 0) return false; /* over-subscribed */
            if (count[l]) max_len = l;
            n_codes += count[l];
        }
        if (n_codes == 0) { /* no code at all: every lookup is invalid (legal for distances when the block has no match) */
            if (!is_dist) return false;
            for (int i = 0; i < (1 << primary_bits); ++i) tab[i] = entry(1, K_INVALID, 0, 0);
            return true;
        }
        if (left > 0 && !(is_dist && n_codes == 1 && count[1] == 1)) return false; /* incomplete */
        int next_code[MAX_CODE_LEN + 2];
        next_code[1] = 0;
        for (int l = 1; l <= MAX_CODE_LEN; ++l) next_code[l + 1] = (next_code[l] + count[l]) << 1;
        const int n_primary = 1 << primary_bits;
        for (int i = 0; i < n_primary; ++i) tab[i] = entry(1, K_INVALID, 0, 0);
        const int sub_bits = max_len > primary_bits ? max_len - primary_bits : 0;
        int sub_next = n_primary;
        for (int s = 0; s < n; ++s) {
            const int l = lens[s];
            if (!l) continue;
            unsigned code = (unsigned)next_code[l]++, rev = 0;
            for (int b = 0; b < l; ++b) rev |= (code >> b & 1u) << (l - 1 - b); /* the stream holds codes MSB first */
            uint32_t e;
            if (is_dist) {
                if (s >= 30) e = entry((unsigned)l, K_INVALID, 0, 0);
                else e = entry(0, K_DIST, dist_extra[s], dist_base[s]);
            } else if (s < 256) e = (uint32_t)K_LITERAL << 4 | (uint32_t)s << 8;
            else if (s == 256) e = entry(0, K_EOB, 0, 0);
            else if (s < 286) e = entry(0, K_LENGTH, len_extra[s - 257], len_base[s - 257]);
            else e = entry(0, K_INVALID, 0, 0);
            if (l <= primary_bits) {
                e |= (unsigned)l;
                for (unsigned i = rev; i < (unsigned)n_primary; i += 1u << l) tab[i] = e;
            } else {
                const unsigned lo = rev & (unsigned)(n_primary - 1);
                if ((tab[lo] >> 4 & 15u) != K_SUB) {
                    tab[lo] = entry((unsigned)primary_bits, K_SUB, (unsigned)sub_bits, (unsigned)sub_next);
                    for (int i = 0; i < (1 << sub_bits); ++i) tab[sub_next + i] = entry(1, K_INVALID, 0, 0);
                    sub_next += 1 << sub_bits;
                }
                const unsigned base = tab[lo] >> 16, hi = rev >> primary_bits;
                e |= (unsigned)(l - primary_bits);
                for (unsigned i = hi; i < (1u << sub_bits); i += 1u << (l - primary_bits)) tab[base + i] = e;
            }
        }
        if (!is_dist) {
            /* Two literals per lookup where both codes fit the primary index: the decoder's speed is the length of
             * the chain lookup -> code length -> shift -> next lookup, and sequence data is short literal codes. */
            uint32_t one[1 << LITLEN_BITS];
            memcpy(one, tab, sizeof one);
            for (unsigned i = 0; i < (unsigned)n_primary; ++i) {
                const uint32_t a = one[i];
                if ((a >> 4 & 15u) != K_LITERAL) continue;
                const unsigned la = a & 15u;
                const uint32_t b = one[i >> la];
                const unsigned lb = b & 15u;
                if ((b >> 4 & 15u) != K_LITERAL || la + lb > (unsigned)primary_bits) continue;
                const uint32_t c = one[i >> (la + lb)];
                const unsigned lc = c & 15u;
                if ((c >> 4 & 15u) == K_LITERAL && la + lb + lc <= (unsigned)primary_bits)
                    tab[i] = (la + lb + lc) | (uint32_t)K_LIT3 << 4 | (a >> 8 & 255u) << 8 | (b >> 8 & 255u) << 16 | (c >> 8 & 255u) << 24;
                else
                    tab[i] = (la + lb) | (uint32_t)K_LIT2 << 4 | (a >> 8 & 255u) << 8 | (b >> 8 & 255u) << 16;
            }
        }
        return true;
    }

    bool read_dynamic_header()
    {
        static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        refill();
        if (bitsleft < 14) return false;
        const int hlit = (int)(bitbuf & 31u) + 257, hdist = (int)(bitbuf >> 5 & 31u) + 1, hclen = (int)(bitbuf >> 10 & 15u) + 4;
        drop(14);
        if (hlit > 286 || hdist > 30) return false; /* zlib refuses these too */
        uint8_t cl[19] = {0};
        for (int i = 0; i < hclen; ++i) {
            if (bitsleft < 3) refill();
            if (bitsleft < 3) return false;
            cl[order[i]] = (uint8_t)(bitbuf & 7u);
            drop(3);
        }
        uint32_t pre[1 << 7];
        {   /* the code-length code: at most 7 bits, complete (zlib refuses an incomplete one) */
            int count[8] = {0};
            for (int s = 0; s < 19; ++s) ++count[cl[s]];
            count[0] = 0;
            long left = 1;
            for (int l = 1; l <= 7; ++l) {
                left = (left << 1) - count[l];
                if (left < 0) return false;
            }
            if (left != 0) return false;
            int next_code[9];
            next_code[1] = 0;
            for (int l = 1; l <= 7; ++l) next_code[l + 1] = (next_code[l] + count[l]) << 1;
            for (int s = 0; s < 19; ++s) {
                const int l = cl[s];
                if (!l) continue;
                unsigned code = (unsigned)next_code[l]++, rev = 0;
                for (int b = 0; b < l; ++b) rev |= (code >> b & 1u) << (l - 1 - b);
                for (unsigned i = rev; i < 128u; i += 1u << l) pre[i] = (uint32_t)s << 8 | (uint32_t)l;
            }
        }
        uint8_t lens[286 + 30 + 140];
        int n = 0;
        const int total = hlit + hdist;
        while (n < total) {
            refill();
            if (bitsleft < 14) return false; /* 7 bits of code + 7 extra at most */
            const uint32_t e = pre[bitbuf & 127u];
            const int sym = (int)(e >> 8);
            drop((int)(e & 255u));
            if (sym < 16) lens[n++] = (uint8_t)sym;
            else {
                int rep, val = 0;
                if (sym == 16) {
                    if (n == 0) return false;
                    val = lens[n - 1];
                    rep = 3 + (int)(bitbuf & 3u);
                    drop(2);
                } else if (sym == 17) {
                    rep = 3 + (int)(bitbuf & 7u);
                    drop(3);
                } else {
                    rep = 11 + (int)(bitbuf & 127u);
                    drop(7);
                }
                if (n + rep > total) return false;
                memset(lens + n, val, (size_t)rep);
                n += rep;
            }
        }
        if (bitsleft < 0) return false;
        if (lens[256] == 0) return false; /* no end-of-block code (zlib: "invalid code -- missing end-of-block") */
        if (!build(lens, hlit, litlen, LITLEN_BITS, false)) return false;
        if (!build(lens + hlit, hdist, dist, OFF_BITS, true)) return false;
        return true;
    }

    void make_static()
    {
        uint8_t lens[288];
        for (int i = 0; i < 144; ++i) lens[i] = 8;
        for (int i = 144; i < 256; ++i) lens[i] = 9;
        for (int i = 256; i < 280; ++i) lens[i] = 7;
        for (int i = 280; i < 288; ++i) lens[i] = 8;
        build(lens, 288, static_litlen, LITLEN_BITS, false);
        uint8_t dl[32];
        for (int i = 0; i < 32; ++i) dl[i] = 5;
        /* 32 five-bit codes are a complete code; 30 and 31 never occur in a valid stream (K_INVALID) */
        build(dl, 32, static_dist, OFF_BITS, true);
        static_ready = true;
    }

    /*
     * Decodes into [out, out_end); the bytes [out_base, out) are history that matches may reach back into (at most
     * 32 KB are needed).  Returns FI_MORE_OUTPUT when fewer than 320 bytes of room are left (call again with room and
     * the history in place), FI_DONE at the end of the final block, FI_ERROR otherwise.  *out_pos is advanced.
     */
    Status run(uint8_t *out_base, uint8_t **out_pos, uint8_t *out_end)
    {
        uint8_t *out = *out_pos;
        for (;;) {
            if (mode == 3) { *out_pos = out; return FI_DONE; }
            if (mode == 0) {
                refill();
                if (bitsleft < 3) { *out_pos = out; return FI_ERROR; }
                final_block = bitbuf & 1u;
                const int type = (int)(bitbuf >> 1 & 3u);
                drop(3);
                if (type == 0) {
                    drop(bitsleft & 7); /* to the byte boundary */
                    refill();
                    if (bitsleft < 32) { *out_pos = out; return FI_ERROR; }
                    const uint32_t len = (uint32_t)(bitbuf & 0xffffu), nlen = (uint32_t)(bitbuf >> 16 & 0xffffu);
                    if ((len ^ nlen) != 0xffffu) { *out_pos = out; return FI_ERROR; }
                    drop(32);
                    stored_left = len;
                    mode = 1;
                } else if (type == 1) {
                    if (!static_ready) make_static();
                    use_static = true;
                    mode = 2;
                } else if (type == 2) {
                    if (!read_dynamic_header()) { *out_pos = out; return FI_ERROR; }
                    use_static = false;
                    mode = 2;
                } else { *out_pos = out; return FI_ERROR; }
            }
            if (mode == 1) {
                /* stored bytes: first what the bit buffer holds (whole bytes), then straight from the input */
                while (stored_left && bitsleft >= 8) {
                    if (out >= out_end) { *out_pos = out; return FI_MORE_OUTPUT; }
                    *out++ = (uint8_t)bitbuf;
                    drop(8);
                    --stored_left;
                }
                if (stored_left) {
                    bitbuf = 0; /* (fewer than 8 bits were left: they belong to bytes `in` has not passed) */
                    in -= bitsleft >> 3;
                    bitsleft = 0;
                    size_t k = stored_left;
                    if ((size_t)(in_end - in) < k) { *out_pos = out; return FI_ERROR; }
                    if ((size_t)(out_end - out) < k) k = (size_t)(out_end - out);
                    memcpy(out, in, k);
                    out += k;
                    in += k;
                    stored_left -= (uint32_t)k;
                    if (stored_left) { *out_pos = out; return FI_MORE_OUTPUT; }
                }
                mode = final_block ? 3 : 0;
                continue;
            }
            /* mode 2: symbols of a Huffman block.  The bit buffer and the input pointer live in locals here: the
             * byte stores to `out` may alias anything, and with the state in the object the compiler would have to
             * reload it after every store. */
            const uint32_t *lt = use_static ? static_litlen : litlen, *dt = use_static ? static_dist : dist;
            const uint8_t *ip = in;
            const uint8_t *const ie = in_end;
            uint64_t bb = bitbuf;
            int bl = bitsleft;
#define FI_REFILL()                                                         \
    do {                                                                    \
        if (ip + 8 <= ie) {                                                 \
            uint64_t v_;                                                    \
            memcpy(&v_, ip, 8);                                             \
            bb |= v_ << bl;                                                 \
            const int n_ = (63 - bl) >> 3;                                  \
            ip += n_;                                                       \
            bl += n_ << 3;                                                  \
        } else {                                                            \
            while (bl <= 56 && ip < ie) {                                   \
                bb |= (uint64_t)*ip++ << bl;                                \
                bl += 8;                                                    \
            }                                                               \
        }                                                                   \
    } while (0)
#define FI_DROP(n) do { const int d_ = (int)(n); bb >>= d_; bl -= d_; } while (0)
#define FI_LEAVE(status) do { in = ip; bitbuf = bb; bitsleft = bl; *out_pos = out; return (status); } while (0)
            for (;;) {
                if (out_end - out < 320) FI_LEAVE(FI_MORE_OUTPUT);
                FI_REFILL();
                uint32_t e = lt[bb & ((1u << LITLEN_BITS) - 1u)];
                if ((e >> 4 & 15u) == K_SUB) {
                    const unsigned sb = e >> 8 & 31u;
                    e = lt[(e >> 16) + (unsigned)(bb >> LITLEN_BITS & ((1u << sb) - 1u))];
                    FI_DROP(LITLEN_BITS);
                }
                FI_DROP(e & 15u);
                const unsigned kind = e >> 4 & 15u;
                if (__builtin_expect(kind == K_LENGTH, 1)) {
                    const unsigned lx = e >> 8 & 31u;
                    const unsigned length = (e >> 16) + (unsigned)(bb & ((1u << lx) - 1u));
                    FI_DROP(lx);
                    if (__builtin_expect(bl < 28, 0)) FI_REFILL(); /* 15 bits of distance code + 13 extra */
                    uint32_t d = dt[bb & ((1u << OFF_BITS) - 1u)];
                    if (__builtin_expect((d >> 4 & 15u) == K_SUB, 0)) {
                        const unsigned sb = d >> 8 & 31u;
                        d = dt[(d >> 16) + (unsigned)(bb >> OFF_BITS & ((1u << sb) - 1u))];
                        FI_DROP(OFF_BITS);
                    }
                    FI_DROP(d & 15u);
                    if (__builtin_expect((d >> 4 & 15u) != K_DIST, 0)) FI_LEAVE(FI_ERROR);
                    const unsigned dx = d >> 8 & 31u;
                    const unsigned distance = (d >> 16) + (unsigned)(bb & ((1u << dx) - 1u));
                    FI_DROP(dx);
                    if (__builtin_expect(bl < 0 || distance > (size_t)(out - out_base), 0)) FI_LEAVE(FI_ERROR);
                    const uint8_t *src = out - distance;
                    uint8_t *dst = out;
                    out += length;
                    if (__builtin_expect(distance >= 16, 1)) { /* sixteen bytes at a time; writes up to 15 bytes past the match (room is kept) */
                        memcpy(dst, src, 16);
                        if (__builtin_expect(length > 16, 0)) {
                            do {
                                dst += 16;
                                src += 16;
                                memcpy(dst, src, 16);
                            } while (dst + 16 < out);
                        }
                    } else if (distance == 1) {
                        memset(dst, *src, length);
                    } else {
                        do *dst++ = *src++; while (dst < out);
                    }
                    continue;
                }
                if (kind == K_LITERAL || kind >= K_LIT2) {
                    /* one to three literals: all three bytes are stored (room is kept), the pointer moves by the count */
                    out[0] = (uint8_t)(e >> 8);
                    out[1] = (uint8_t)(e >> 16);
                    out[2] = (uint8_t)(e >> 24);
                    out += kind == K_LITERAL ? 1 : (int)kind - (K_LIT2 - 2);
                    /* more literals from the same refill: a primary-table entry takes at most LITLEN_BITS bits, and 40
                     * bits are kept for whatever comes next */
                    while (bl >= 40) {
                        const uint32_t e2 = lt[bb & ((1u << LITLEN_BITS) - 1u)];
                        const unsigned k2 = e2 >> 4 & 15u;
                        if (k2 != K_LITERAL && k2 < K_LIT2) break;
                        FI_DROP(e2 & 15u);
                        out[0] = (uint8_t)(e2 >> 8);
                        out[1] = (uint8_t)(e2 >> 16);
                        out[2] = (uint8_t)(e2 >> 24);
                        out += k2 == K_LITERAL ? 1 : (int)k2 - (K_LIT2 - 2);
                    }
                    if (bl < 0) FI_LEAVE(FI_ERROR);
                    continue;
                }
                if (kind == K_EOB) {
                    if (bl < 0) FI_LEAVE(FI_ERROR);
                    break;
                }
                FI_LEAVE(FI_ERROR);
            }
            in = ip;
            bitbuf = bb;
            bitsleft = bl;
#undef FI_REFILL
#undef FI_DROP
#undef FI_LEAVE
            mode = final_block ? 3 : 0;
        }
    }
};

/* gzip member header (RFC 1952) at p: its length, or 0 when it is not one / not whole */
static inline size_t gzip_header_len(const uint8_t *p, size_t n)
{
    if (n < 10 || p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || (p[3] & 0xe0)) return 0;
    const unsigned flg = p[3];
    size_t at = 10;
    if (flg & 4) {
        if (at + 2 > n) return 0;
        at += 2 + ((size_t)p[at] | (size_t)p[at + 1] << 8);
    }
    if (flg & 8) {
        while (at < n && p[at]) ++at;
        ++at;
    }
    if (flg & 16) {
        while (at < n && p[at]) ++at;
        ++at;
    }
    if (flg & 2) at += 2;
    return at <= n ? at : 0;
}

} // namespace fastinflate

#endif
