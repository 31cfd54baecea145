// inflate_harness.cpp — TEST INFRASTRUCTURE: ibwa_b200/csrc/fast_inflate.h behind a C entry point, so that the tests
// can run the decoder on arbitrary deflate streams and compare it with zlib.  Not part of libb200aln.so.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../ibwa_b200/csrc/fast_inflate.h"

// Decodes the raw deflate stream in[0..n) in output chunks of `chunk` bytes (as the reader does: a chunk is decoded
// into a buffer that carries the last 32 KB before it).  Returns the status of the last call (1 = done, 2 = error),
// *out_len = bytes produced, *consumed = compressed bytes used.
extern "C" int fi_inflate(const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t chunk, size_t *out_len,
                          size_t *consumed)
{
    fastinflate::Decoder *d = new fastinflate::Decoder;
    d->start(in, n);
    const size_t H = 32768;
    std::vector<uint8_t> buf(H + chunk + 64);
    size_t total = 0, hist = 0;
    int st;
    for (;;) {
        uint8_t *base = buf.data() + H, *p = base;
        st = d->run(base - hist, &p, base + chunk);
        const size_t got = (size_t)(p - base);
        if (total + got > cap) { st = 3; break; }
        memcpy(out + total, base, got);
        total += got;
        if (st != fastinflate::FI_MORE_OUTPUT) break;
        /* history for the next chunk: the last 32 KB of everything produced so far */
        const size_t keep = total < H ? total : H;
        memcpy(base - keep, out + total - keep, keep);
        hist = keep;
    }
    *out_len = total;
    *consumed = (size_t)(d->in_pos() - in);
    delete d;
    return st;
}
