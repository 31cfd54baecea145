/*
 * aln_host.cpp — host side of the drop-in boundary: read input, the batch loop,
 * the .sai writer, the `aln` command line, and the bwa_seq_t seam.
 *
 * Reference interfaces replaced (file:line under the reference tree):
 *   bwa_aln            bwtaln.c:243-328   -> b200aln_aln_main
 *   bwa_aln_core       bwtaln.c:173-241   -> b200aln_aln_core
 *   bwa_read_seq       bwaseqio.c:145-208 -> SeqReader::next_batch (FASTA/FASTQ, gz)
 *   kseq_read          kseq.h:150-194     -> SeqReader::read_record
 *   bwa_trim_read      bwaseqio.c:74-87   -> trim_len
 *   bwa_cal_sa_reg_gap bwtaln.c:80-140    -> b200aln_cal_sa_reg_gap (on bwa_seq_t)
 * Compute happens in b200aln.cu; nothing here searches the index.
 */
#include <ctype.h>
#include <stdint.h>
#include <stdio.h>
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <future>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200aln.h"
#include "host_params.h"
#include "fast_inflate.h"

struct b200aln_reader;
extern "C" void b200aln_warm_device(int device); /* b200aln.cu: creates the CUDA context */
extern "C" int b200aln_pin(void *p, size_t bytes);
extern "C" void b200aln_unpin(void *p);
extern "C" void b200aln_prealloc(int device, int n_contexts, int n_reads, int max_len);
extern "C" void b200aln_prealloc_release(int device);

namespace {

enum {
    BWA_MODE_GAPE = 0x01, BWA_MODE_COMPREAD = 0x02, BWA_MODE_LOGGAP = 0x04, BWA_MODE_NONSTOP = 0x10,
    BWA_MODE_BAM = 0x20, BWA_MODE_BAM_SE = 0x40, BWA_MODE_BAM_READ1 = 0x80, BWA_MODE_BAM_READ2 = 0x100,
    BWA_MODE_IL13 = 0x200
};
const int BWA_MIN_RDLEN = 35;
const double BWA_AVG_ERR = 0.02;

/* nt4 code of a character (bntseq.c:39-56): ACGT/acgt -> 0-3, '-' -> 5, other -> 4 */
struct Nt4Table {
    uint8_t t[256];
    Nt4Table()
    {
        memset(t, 4, sizeof t);
        t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3;
        t['-'] = 5;
    }
};
const Nt4Table g_nt4;

/* conversion table of the fast path: nt4 code in the low bits, bit 7 set for a character the fast path must
 * not accept in a sequence line (not graphic, or one of '>', '+', '@' which end the sequence in kseq.h:167) */
struct SeqClassTable {
    uint8_t t[256];
    SeqClassTable()
    {
        for (int c = 0; c < 256; ++c) {
            const bool bad = c <= 32 || c >= 127 || c == '>' || c == '+' || c == '@';
            t[c] = (uint8_t)(g_nt4.t[c] | (bad ? 0x80 : 0));
        }
    }
};
const SeqClassTable g_seqclass;

/* Sequence line -> nt4 codes and the character checks of the fast path, on L characters: false when the
 * sequence holds a character the fast path must not accept (SeqClassTable) or the quality string one outside
 * 33..127 (kseq.h:185).  The table version, and an AVX2 version of the same function picked at run time. */
static bool convert_scalar(const unsigned char *sq, const unsigned char *ql, int L, uint8_t *codes)
{
    unsigned acc = 0, badq = 0;
    for (int i = 0; i < L; ++i) {
        const uint8_t v = g_seqclass.t[sq[i]];
        acc |= v;
        codes[i] = (uint8_t)(v & 7);
    }
    for (int i = 0; i < L; ++i) badq |= (unsigned)((unsigned char)(ql[i] - 33) > 94);
    return !((acc & 0x80u) | badq);
}
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2"))) static inline __m256i nt4_avx2(__m256i v, __m256i &bad)
{
    /* a letter of ACGT in either case has a low nibble of 1, 3, 7, 4: the nibble selects the code and the lower-case
     * letter the character has to be; everything else is 4, '-' is 5 (bntseq.c:39-56) */
    const __m256i lut_code = _mm256_setr_epi8(4, 0, 4, 1, 3, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 0, 4, 1, 3, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4);
    const __m256i lut_chr = _mm256_setr_epi8(0, 'a', 0, 'c', 't', 0, 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0, 0, 'a', 0, 'c', 't', 0, 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i low = _mm256_and_si256(v, _mm256_set1_epi8(0x0f));
    const __m256i is_nt = _mm256_cmpeq_epi8(_mm256_or_si256(v, _mm256_set1_epi8(0x20)), _mm256_shuffle_epi8(lut_chr, low));
    __m256i code = _mm256_blendv_epi8(_mm256_set1_epi8(4), _mm256_shuffle_epi8(lut_code, low), is_nt);
    code = _mm256_blendv_epi8(code, _mm256_set1_epi8(5), _mm256_cmpeq_epi8(v, _mm256_set1_epi8('-')));
    /* not graphic (<= 32, >= 127: as signed bytes everything below 33, and 127), or one of '>', '+', '@' */
    __m256i b = _mm256_cmpgt_epi8(_mm256_set1_epi8(33), v);
    b = _mm256_or_si256(b, _mm256_cmpeq_epi8(v, _mm256_set1_epi8(127)));
    b = _mm256_or_si256(b, _mm256_cmpeq_epi8(v, _mm256_set1_epi8('>')));
    b = _mm256_or_si256(b, _mm256_cmpeq_epi8(v, _mm256_set1_epi8('+')));
    b = _mm256_or_si256(b, _mm256_cmpeq_epi8(v, _mm256_set1_epi8('@')));
    bad = _mm256_or_si256(bad, b);
    return code;
}
__attribute__((target("avx2"))) static bool convert_avx2(const unsigned char *sq, const unsigned char *ql, int L, uint8_t *codes)
{
    if (L < 32) return convert_scalar(sq, ql, L, codes);
    __m256i bad = _mm256_setzero_si256();
    int i = 0;
    for (; i + 32 <= L; i += 32) {
        _mm256_storeu_si256((__m256i *)(codes + i), nt4_avx2(_mm256_loadu_si256((const __m256i *)(sq + i)), bad));
        bad = _mm256_or_si256(bad, _mm256_cmpgt_epi8(_mm256_set1_epi8(33), _mm256_loadu_si256((const __m256i *)(ql + i))));
    }
    if (i < L) { /* the last 32 characters once more, overlapping what is done */
        i = L - 32;
        _mm256_storeu_si256((__m256i *)(codes + i), nt4_avx2(_mm256_loadu_si256((const __m256i *)(sq + i)), bad));
        bad = _mm256_or_si256(bad, _mm256_cmpgt_epi8(_mm256_set1_epi8(33), _mm256_loadu_si256((const __m256i *)(ql + i))));
    }
    return _mm256_testz_si256(bad, bad) != 0;
}
static bool (*const g_convert)(const unsigned char *, const unsigned char *, int, uint8_t *) = [] {
    __builtin_cpu_init();
    return (getenv("B200ALN_NO_SIMD") == nullptr && __builtin_cpu_supports("avx2")) ? convert_avx2 : convert_scalar;
}();
#else
static bool (*const g_convert)(const unsigned char *, const unsigned char *, int, uint8_t *) = convert_scalar;
#endif

/* persistent worker threads for the record conversion (one pool per process) */
class ParsePool {
  public:
    static ParsePool &get()
    {
        static ParsePool p;
        return p;
    }
    unsigned size() const { return n_; }
    /* runs fn(0..size-1), fn(0) on the caller */
    void run(const std::function<void(unsigned)> &fn)
    {
        if (n_ == 1) { fn(0); return; }
        std::lock_guard<std::mutex> one_at_a_time(run_m_); /* several readers may share the pool */
        {
            std::lock_guard<std::mutex> g(m_);
            fn_ = &fn;
            pending_ = n_ - 1;
            ++epoch_;
        }
        cv_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }

  private:
    ParsePool()
    {
        const char *e = getenv("B200ALN_PARSE_THREADS");
        unsigned t = e ? (unsigned)atoi(e) : std::thread::hardware_concurrency();
        n_ = t < 1 ? 1u : (t > 32 ? 32u : t);
        for (unsigned i = 1; i < n_; ++i) th_.emplace_back([this, i] { loop(i); });
    }
    ~ParsePool()
    {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
            ++epoch_;
        }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    void loop(unsigned id)
    {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(unsigned)> *fn;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return epoch_ != seen; });
                seen = epoch_;
                if (stop_) return;
                fn = fn_;
            }
            (*fn)(id);
            std::lock_guard<std::mutex> g(m_);
            if (--pending_ == 0) done_.notify_one();
        }
    }
    unsigned n_ = 1;
    std::vector<std::thread> th_;
    std::mutex m_, run_m_;
    std::condition_variable cv_, done_;
    const std::function<void(unsigned)> *fn_ = nullptr;
    unsigned pending_ = 0;
    uint64_t epoch_ = 0;
    bool stop_ = false;
};

/*
 * Sequential source of decompressed bytes with gzread()'s contract (utils.c:56-66 opens every input
 * through zlib, "-" = stdin; plain data passes through unchanged).
 *   BGZF files (BAM, bgzip: every gzip member carries its size in a 'BC' extra field) are inflated block-parallel
 *   on the parse pool, a window of blocks at a time;
 *   every other stream goes through zlib's gzread on a helper thread that stays one buffer ahead of the parser.
 * The bytes delivered are the ones gzread would deliver; only who inflates them, and when, differs.
 */
class InflateSource {
  public:
    explicit InflateSource(const char *fn)
    {
        const bool is_stdin = strcmp(fn, "-") == 0;
        if (!is_stdin && !getenv("B200ALN_NO_BGZF")) {
            fd_ = open(fn, O_RDONLY);
            unsigned char h[18];
            if (fd_ >= 0 && pread(fd_, h, 18, 0) == 18 && h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4) &&
                h[12] == 'B' && h[13] == 'C' && h[14] == 2 && h[15] == 0)
                bgzf_ = true;
            else if (fd_ >= 0) { close(fd_); fd_ = -1; }
        }
        if (bgzf_) th_ = std::thread([this] { producer_bgzf(); });
        if (!bgzf_) {
            fn_ = fn;
            /* a gzip file on disk is mapped and decoded by fast_inflate.h; everything else (stdin, pipes, plain data)
             * and anything that decoder does not take goes through zlib */
            if (!is_stdin && !getenv("B200ALN_NO_FAST_INFLATE")) {
                const int fd = open(fn, O_RDONLY);
                struct stat st;
                unsigned char h[2] = {0, 0};
                if (fd >= 0 && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 18 && pread(fd, h, 2, 0) == 2 &&
                    h[0] == 0x1f && h[1] == 0x8b) {
                    void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
                    if (m != MAP_FAILED) {
                        madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
                        zmap_ = (const unsigned char *)m;
                        zmap_len_ = (size_t)st.st_size;
                    }
                }
                if (fd >= 0) close(fd);
            }
            if (!zmap_) open_zlib(is_stdin);
            for (auto &b : ring_) b.data.resize(kHist + kChunk);
            th_ = std::thread([this] { if (zmap_) producer_fast(); else producer(); });
        }
    }
    ~InflateSource()
    {
        if (th_.joinable()) {
            {
                std::lock_guard<std::mutex> g(m_);
                stop_ = true;
            }
            cv_.notify_all();
            th_.join();
        }
        if (gz_) gzclose(gz_);
        if (fd_ >= 0) close(fd_);
        if (zmap_) munmap(const_cast<unsigned char *>(zmap_), zmap_len_);
    }
    /* like gzread: the number of bytes delivered, less than n only at the end of the stream */
    int read(void *dst, unsigned n)
    {
        unsigned char *d = (unsigned char *)dst;
        unsigned got = 0;
        while (got < n) {
            {
                Chunk &c = ring_[rd_idx_ % kRing];
                {
                    std::unique_lock<std::mutex> g(m_);
                    cv_.wait(g, [&] { return c.full; });
                }
                if (c.len == 0) break; /* end of stream (stays "full": further reads return 0 too) */
                const size_t k = std::min<size_t>(n - got, c.len - c.pos);
                memcpy(d + got, c.data.data() + c.pos, k);
                c.pos += k;
                got += (unsigned)k;
                if (c.pos == c.len) {
                    {
                        std::lock_guard<std::mutex> g(m_);
                        c.full = false;
                    }
                    cv_.notify_all();
                    ++rd_idx_;
                }
            }
        }
        return (int)got;
    }

  private:
    /* ---- zlib stream, one helper thread ahead ---- */
    static const size_t kChunk = (size_t)16 << 20;
    static const size_t kHist = 32768; /* fast decoder: the 32 KB before a chunk's data, for matches that reach back */
    static const int kRing = 3;
    struct Chunk { std::vector<unsigned char> data; size_t len = 0, pos = 0; bool full = false; };
    void open_zlib(bool is_stdin)
    {
        gz_ = is_stdin ? gzdopen(fileno(stdin), "r") : gzopen(fn_.c_str(), "r");
        if (!gz_) b2host::fatal("b200aln_aln_core", (std::string("fail to open file '") + fn_ + "'.").c_str());
        gzbuffer(gz_, 1 << 20);
    }
    /* waits for ring slot i to be free; false when the reader is gone */
    bool claim(uint64_t i)
    {
        Chunk &c = ring_[i % kRing];
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return !c.full || stop_; });
        return !stop_;
    }
    void publish(uint64_t i, size_t pos, size_t len)
    {
        Chunk &c = ring_[i % kRing];
        c.pos = pos;
        c.len = len;
        {
            std::lock_guard<std::mutex> g(m_);
            c.full = true;
        }
        cv_.notify_all();
    }
    /* The mapped gzip file through fast_inflate.h, member by member (RFC 1952), chunk by chunk; the CRC and the
     * length of every member are checked like gzread does.  Whatever this path does not take for granted — a member
     * the decoder refuses, a check that fails, bytes after a member that are not another member — is left to zlib:
     * the stream is opened with gzopen, wound forward to the byte this path has delivered last, and goes on there,
     * so the bytes and the end of the stream are what gzread gives. */
    void producer_fast()
    {
        std::unique_ptr<fastinflate::Decoder> dec(new fastinflate::Decoder);
        size_t at = 0;          /* in the compressed file */
        uint64_t delivered = 0; /* plain bytes published so far */
        uint64_t i = 0;
        bool handed_over = false;
        while (at < zmap_len_ && !handed_over) {
            const size_t hl = fastinflate::gzip_header_len(zmap_ + at, zmap_len_ - at);
            if (!hl) { handed_over = true; break; }
            dec->start(zmap_ + at + hl, zmap_len_ - at - hl);
            /* the member's CRC: every chunk's on a thread of its own while the next chunk is decoded, put together with
             * crc32_combine; at most two are outstanding, so a ring slot is never reused under a running one */
            uLong crc = crc32(0L, Z_NULL, 0);
            std::deque<std::pair<std::future<uLong>, size_t>> crcs;
            auto settle = [&](size_t keep) {
                while (crcs.size() > keep) {
                    crc = crc32_combine(crc, crcs.front().first.get(), (z_off_t)crcs.front().second);
                    crcs.pop_front();
                }
            };
            uint64_t member_bytes = 0;
            size_t hist = 0; /* bytes of this member's output in front of the current chunk's data */
            const unsigned char *prev_tail = nullptr;
            for (;;) {
                settle(1);
                if (!claim(i)) { settle(0); return; }
                Chunk &c = ring_[i % kRing];
                unsigned char *base = c.data.data() + kHist;
                if (hist) memcpy(base - hist, prev_tail - hist, hist);
                uint8_t *out = base;
                const fastinflate::Status st = dec->run(base - hist, &out, base + kChunk);
                if (st == fastinflate::FI_ERROR) { settle(0); handed_over = true; break; }
                const size_t got = (size_t)(out - base);
                if (got)
                    crcs.emplace_back(std::async(std::launch::async, [base, got] { return crc32(crc32(0L, Z_NULL, 0), base, (uInt)got); }), got);
                member_bytes += got;
                if (st == fastinflate::FI_DONE) {
                    settle(0);
                    const unsigned char *t = dec->in_pos();
                    if (t + 8 > zmap_ + zmap_len_) { handed_over = true; break; }
                    const uint32_t want_crc = t[0] | t[1] << 8 | t[2] << 16 | (uint32_t)t[3] << 24;
                    const uint32_t want_len = t[4] | t[5] << 8 | t[6] << 16 | (uint32_t)t[7] << 24;
                    if (want_crc != (uint32_t)crc || want_len != (uint32_t)member_bytes) { handed_over = true; break; }
                    at = (size_t)(t + 8 - zmap_);
                }
                if (got) {
                    delivered += got;
                    publish(i, kHist, kHist + got);
                    ++i;
                }
                if (st == fastinflate::FI_DONE) break;
                hist = got < kHist ? std::min(kHist, hist + got) : kHist;
                if (got < kHist && hist > got) { /* (a short chunk: part of the history is the history of before) */
                    settle(0);
                    handed_over = true; /* does not happen with 16 MB chunks; leave it to zlib rather than stitch */
                    break;
                }
                prev_tail = base + got;
            }
        }
        if (handed_over) {
            open_zlib(false);
            if (delivered && gzseek(gz_, (z_off_t)delivered, SEEK_SET) < 0) { /* zlib cannot get there either: end of stream */
                if (claim(i)) publish(i, 0, 0);
                return;
            }
            producer(i);
            return;
        }
        if (claim(i)) publish(i, 0, 0); /* end of stream */
    }
    void producer(uint64_t first = 0)
    {
        for (uint64_t i = first;; ++i) {
            Chunk &c = ring_[i % kRing];
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return !c.full || stop_; });
                if (stop_) return;
            }
            const int k = gzread(gz_, c.data.data(), (unsigned)kChunk);
            c.len = k > 0 ? (size_t)k : 0;
            c.pos = 0;
            {
                std::lock_guard<std::mutex> g(m_);
                c.full = true;
            }
            cv_.notify_all();
            if (c.len == 0) return; /* end of stream or error: the zero-length chunk tells the reader */
        }
    }
    /* ---- BGZF, block-parallel, one window ahead of the parser ---- */
    struct Block { size_t in_off, in_len, out_off, out_len; uint32_t crc; };
    void producer_bgzf()
    {
        for (uint64_t i = 0;; ++i) {
            if (!claim(i)) return;
            size_t n = 0;
            while (!bgzf_eof_ && n == 0) n = fill_bgzf(ring_[i % kRing].data); /* (a window of empty blocks — the EOF marker: look further) */
            publish(i, 0, n);
            if (n == 0) return;
        }
    }
    /* the next window of whole blocks, inflated into `out` by threads of this source's own (the parse pool is busy
     * with the window before); returns the plain bytes, 0 with bgzf_eof_ at the end of the file */
    size_t fill_bgzf(std::vector<unsigned char> &out)
    {
        static const size_t kWindow = [] { /* compressed bytes per round (B200ALN_BGZF_WINDOW: tests shrink it) */
            const char *e = getenv("B200ALN_BGZF_WINDOW");
            const long v = e ? atol(e) : 0;
            return v >= 1024 ? (size_t)v : (size_t)8 << 20;
        }();
        cbuf_.resize(kWindow + 65536 + 64);
        const ssize_t have = pread(fd_, cbuf_.data(), cbuf_.size(), (off_t)file_pos_);
        if (have <= 0) { bgzf_eof_ = true; return 0; }
        std::vector<Block> blocks;
        size_t at = 0, out_total = 0;
        while (at + 18 <= (size_t)have && at < kWindow) {
            const unsigned char *h = cbuf_.data() + at;
            if (!(h[0] == 0x1f && h[1] == 0x8b && h[2] == 8 && (h[3] & 4)))
                b2host::fatal("b200aln_reader", "corrupt BGZF stream (bad member header).");
            const unsigned xlen = h[10] | h[11] << 8;
            if (at + 12 + xlen > (size_t)have) break;
            int bsize = -1;
            for (unsigned x = 0; x + 4 <= xlen;) { /* the BC subfield holds (block size - 1) */
                const unsigned char *f = h + 12 + x;
                const unsigned slen = f[2] | f[3] << 8;
                if (f[0] == 'B' && f[1] == 'C' && slen == 2 && x + 6 <= xlen) bsize = (f[4] | f[5] << 8) + 1;
                x += 4 + slen;
            }
            if (bsize < 0 || (size_t)bsize < 12 + xlen + 8) b2host::fatal("b200aln_reader", "corrupt BGZF stream (no block size).");
            if (at + (size_t)bsize > (size_t)have) break; /* the block continues beyond what was read */
            const unsigned char *tail = h + bsize - 8;
            Block b;
            b.in_off = at + 12 + xlen;
            b.in_len = (size_t)bsize - 12 - xlen - 8;
            b.crc = tail[0] | tail[1] << 8 | tail[2] << 16 | (uint32_t)tail[3] << 24;
            b.out_len = tail[4] | tail[5] << 8 | tail[6] << 16 | (uint32_t)tail[7] << 24;
            b.out_off = out_total;
            out_total += b.out_len;
            blocks.push_back(b);
            at += (size_t)bsize;
        }
        if (blocks.empty()) {
            if ((size_t)have >= 18) b2host::fatal("b200aln_reader", "truncated BGZF block.");
            bgzf_eof_ = true;
            return 0;
        }
        file_pos_ += at;
        if (out.size() < out_total) out.resize(out_total);
        const unsigned hw = std::thread::hardware_concurrency();
        const unsigned nw = std::max(1u, std::min<unsigned>({8u, hw ? hw / 2 : 1u, (unsigned)blocks.size()}));
        std::vector<int> bad(nw, 0);
        const bool fast = getenv("B200ALN_NO_FAST_INFLATE") == nullptr;
        auto work = [&](unsigned t) {
            z_stream zs;
            std::unique_ptr<fastinflate::Decoder> dec(fast ? new fastinflate::Decoder : nullptr);
            std::vector<unsigned char> scratch(fast ? 65536 + 1024 : 0); /* (the decoder keeps 320 bytes of room) */
            for (size_t i = t; i < blocks.size(); i += nw) {
                const Block &b = blocks[i];
                if (b.out_len == 0) continue;
                if (fast && b.out_len <= 65536) { /* fast_inflate.h first; what it refuses goes to zlib below */
                    dec->start(cbuf_.data() + b.in_off, b.in_len + 8); /* (the 8-byte trailer is slack for the bit reader) */
                    uint8_t *p = scratch.data();
                    if (dec->run(scratch.data(), &p, scratch.data() + scratch.size()) == fastinflate::FI_DONE &&
                        (size_t)(p - scratch.data()) == b.out_len &&
                        crc32(crc32(0L, Z_NULL, 0), scratch.data(), (uInt)b.out_len) == b.crc) {
                        memcpy(out.data() + b.out_off, scratch.data(), b.out_len);
                        continue;
                    }
                }
                memset(&zs, 0, sizeof zs);
                if (inflateInit2(&zs, -15) != Z_OK) { bad[t] = 1; return; }
                zs.next_in = cbuf_.data() + b.in_off;
                zs.avail_in = (uInt)b.in_len;
                zs.next_out = out.data() + b.out_off;
                zs.avail_out = (uInt)b.out_len;
                const int rc = inflate(&zs, Z_FINISH);
                inflateEnd(&zs);
                if (rc != Z_STREAM_END || zs.avail_out != 0 ||
                    crc32(crc32(0L, Z_NULL, 0), out.data() + b.out_off, (uInt)b.out_len) != b.crc) {
                    bad[t] = 1;
                    return;
                }
            }
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nw; ++t) th.emplace_back(work, t);
        work(0);
        for (auto &x : th) x.join();
        for (int x : bad) if (x) b2host::fatal("b200aln_reader", "corrupt BGZF block (inflate or CRC error).");
        return out_total;
    }

    bool bgzf_ = false, bgzf_eof_ = false;
    int fd_ = -1;
    uint64_t file_pos_ = 0;
    std::vector<unsigned char> cbuf_;
    gzFile gz_ = nullptr;
    std::string fn_;
    const unsigned char *zmap_ = nullptr; /* fast path: the compressed file */
    size_t zmap_len_ = 0;
    std::thread th_;
    std::mutex m_;
    std::condition_variable cv_;
    Chunk ring_[kRing];
    uint64_t rd_idx_ = 0;
    bool stop_ = false;
};

/* buffered reader + record parser with the reference parser's observable
 * behaviour (kseq.h:60-71,150-194): multi-line records, name = first
 * whitespace-delimited token, only isgraph() characters enter the sequence,
 * reading stops at the first truncated record. */
class SeqReader {
  public:
    explicit SeqReader(const char *fn)
    {
        /* A plain (not gzip) regular file is mapped: the whole file is "the buffer", nothing is copied.
         * Everything else — gzip, stdin, pipes — goes through zlib like the reference (utils.c:56-66). */
        if (strcmp(fn, "-") != 0 && !getenv("B200ALN_NO_MMAP")) {
            const int fd = open(fn, O_RDONLY);
            struct stat st;
            if (fd >= 0 && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
                unsigned char magic[2] = {0, 0};
                if (pread(fd, magic, 2, 0) == 2 && !(magic[0] == 0x1f && magic[1] == 0x8b)) {
                    void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
                    if (m != MAP_FAILED) {
                        madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
                        map_ = (unsigned char *)m;
                        map_len_ = (size_t)st.st_size;
                        data_ = map_;
                        end_ = (int64_t)map_len_;
                        is_eof_ = true;
                    }
                }
            }
            if (fd >= 0) close(fd);
        }
        if (!map_) {
            f_.reset(new InflateSource(fn));
            buf_.resize((size_t)160 << 20);
            data_ = buf_.data();
        }
    }
    ~SeqReader()
    {
        if (map_) munmap(map_, map_len_);
    }

    /* returns sequence length, -1 at end of file, -2 on a truncated quality string */
    int read_record()
    {
        int fast = read_record_fast();
        if (fast != -3) return fast;
        return read_record_exact();
    }

    /* Fast path for the ordinary 4-line FASTQ record lying whole in the buffer.  It is taken only when
     * the exact state machine below would provably produce the same record: header '@' is the next
     * byte, the sequence line holds only graphic characters other than '>', '+', '@', the separator
     * line starts with '+', the quality line has exactly as many characters (all in 33..127) and is
     * followed by a newline.  Anything else returns -3 and the exact parser runs from the same state. */
    int read_record_fast()
    {
        if (last_char_ != 0) return -3;
        refill_keep_tail(1 << 16);
        const unsigned char *b = data_;
        const int64_t e = end_;
        int64_t p = begin_;
        if (p >= e || b[p] != '@') return -3;
        const unsigned char *nl1 = (const unsigned char *)memchr(b + p + 1, '\n', (size_t)(e - p - 1));
        if (!nl1) return -3;
        int64_t name_end = p + 1;
        while (name_end < (int64_t)(nl1 - b) && !isspace(b[name_end])) ++name_end;
        if (name_end == p + 1) return -3; /* empty name: let the exact parser decide */
        const int64_t s0 = (int64_t)(nl1 - b) + 1;
        const unsigned char *nl2 = (const unsigned char *)memchr(b + s0, '\n', (size_t)(e - s0));
        if (!nl2) return -3;
        if ((int64_t)(nl2 - b) - s0 > 0x3fffffff) return -3;
        const int L = (int)((int64_t)(nl2 - b) - s0);
        if (L <= 0) return -3;
        for (int i = 0; i < L; ++i) {
            const unsigned char ch = b[s0 + i];
            if (ch <= 32 || ch >= 127 || ch == '>' || ch == '+' || ch == '@') return -3;
        }
        const int64_t q_plus = s0 + L + 1;
        if (q_plus >= e || b[q_plus] != '+') return -3;
        const unsigned char *nl3 = (const unsigned char *)memchr(b + q_plus, '\n', (size_t)(e - q_plus));
        if (!nl3) return -3;
        const int64_t q0 = (int64_t)(nl3 - b) + 1;
        if (q0 + L >= e) return -3; /* need the byte after the quality string too */
        for (int i = 0; i < L; ++i) {
            const unsigned char ch = b[q0 + i];
            if (ch < 33 || ch > 127) return -3;
        }
        if (b[q0 + L] != '\n') return -3; /* the exact parser swallows exactly one more byte here */
        name_.assign((const char *)b + p + 1, (size_t)(name_end - p - 1));
        seq_.assign((const char *)b + s0, (size_t)L);
        qual_.assign((const char *)b + q0, (size_t)L);
        begin_ = q0 + L + 1;
        return L;
    }

    int read_record_exact()
    {
        int c;
        if (last_char_ == 0) {
            while ((c = getc()) != -1 && c != '>' && c != '@') {}
            if (c == -1) return -1;
            last_char_ = c;
        }
        seq_.clear();
        qual_.clear();
        if (get_until(0, name_, &c) < 0) return -1;
        if (c != '\n') {
            std::string comment;
            get_until('\n', comment, nullptr);
        }
        while ((c = getc()) != -1 && c != '>' && c != '+' && c != '@')
            if (isgraph(c)) seq_.push_back((char)c);
        if (c == '>' || c == '@') last_char_ = c;
        if (c != '+') return (int)seq_.size();
        while ((c = getc()) != -1 && c != '\n') {}
        if (c == -1) return -2;
        while ((c = getc()) != -1 && qual_.size() < seq_.size())
            if (c >= 33 && c <= 127) qual_.push_back((char)c);
        last_char_ = 0;
        if (seq_.size() != qual_.size()) return -2;
        return (int)seq_.size();
    }

    /* One ordinary 4-line record found by the structural scan: where it starts in the buffer, and its sequence and
     * quality strings relative to that. */
    struct Extent {
        int64_t start;
        int32_t seq, qual, len; /* offsets from start; the record ends at start + qual + len + 1 */
        int64_t next() const { return start + qual + len + 1; }
    };

    /* One ordinary 4-line record at p, located by its newlines only (convert_extent() validates the characters).
     * `guess` = sequence length of the previous record (most files have one length): the sequence line is
     * taken to end at seq + guess when a newline sits there.  Guessing costs nothing when wrong in the safe
     * direction: a newline INSIDE the guessed region is a character <= 32, which convert_extent() refuses, and
     * the record then goes through the exact parser; a longer line fails the test. */
    bool walk_one(int64_t p, int &guess, Extent &x) const
    {
        const unsigned char *b = data_;
        const int64_t e = end_;
        if (p >= e || b[p] != '@') return false;
        const unsigned char *nl1 = (const unsigned char *)memchr(b + p + 1, '\n', (size_t)(e - p - 1));
        if (!nl1 || nl1 == b + p + 1 || isspace(b[p + 1])) return false;
        const int64_t s0 = (int64_t)(nl1 - b) + 1;
        if (s0 - p > 0x3fffffff) return false;
        int L;
        if (guess > 0 && s0 + guess < e && b[s0 + guess] == '\n') L = guess;
        else {
            const unsigned char *nl2 = (const unsigned char *)memchr(b + s0, '\n', (size_t)(e - s0));
            if (!nl2 || (int64_t)(nl2 - b) - s0 > 0x1fffffff) return false;
            L = (int)((int64_t)(nl2 - b) - s0);
        }
        const int64_t q_plus = s0 + L + 1;
        if (L <= 0 || q_plus >= e || b[q_plus] != '+') return false;
        int64_t q0;
        if (q_plus + 1 < e && b[q_plus + 1] == '\n') q0 = q_plus + 2;
        else {
            const unsigned char *nl3 = (const unsigned char *)memchr(b + q_plus, '\n', (size_t)(e - q_plus));
            if (!nl3) return false;
            q0 = (int64_t)(nl3 - b) + 1;
        }
        if (q0 - p > 0x7fffffff || q0 + L >= e || b[q0 + L] != '\n') return false; /* (a newline inside the quality string is refused by convert_extent) */
        x.start = p; x.seq = (int32_t)(s0 - p); x.len = L; x.qual = (int32_t)(q0 - p);
        guess = L;
        return true;
    }
    /* consecutive ordinary records from p while they start before `limit`, at most max_records; returns their bases */
    int64_t walk(int64_t p, int64_t limit, size_t max_records, std::vector<Extent> &out) const
    {
        int guess = -1;
        int64_t bases = 0;
        Extent x;
        while (out.size() < max_records && p < limit && walk_one(p, guess, x)) {
            out.push_back(x);
            bases += x.len;
            p = x.next();
        }
        return bases;
    }

    /* What one worker of the structural scan found in its slice of the byte range (kept between calls: no
     * allocation per batch).  The first `n` records of `ext` are the piece's share of the scan's result. */
    struct Piece {
        std::vector<Extent> ext;
        size_t n = 0;
        int64_t bases = 0; /* sum of len over the first n records */
    };
    const std::vector<Piece> &pieces() const { return pieces_; }
    unsigned n_pieces() const { return n_pieces_; }

    /* Structural scan of up to max_records consecutive ordinary records from the cursor; returns how many it found
     * and leaves them in pieces()[0 .. n_pieces()), in order.  Does NOT move the cursor.  Stops at the first record
     * that is not of the plain 4-line shape or not wholly in the buffer.
     * Large requests are scanned in parallel: the byte range is cut into one slice per worker, every worker
     * but the first looks for a place in its slice where two ordinary records follow a newline and walks on
     * from there, and a piece is accepted only when it begins exactly where the piece before it ended — so the
     * result is a chain of consecutive records from the cursor, like the serial walk's (a piece that does not fit is
     * dropped together with everything after it and scanned again by the next call). */
    int scan_fast(int max_records)
    {
        n_pieces_ = 0;
        if (last_char_ != 0) return 0;
        refill_keep_tail(1 << 26);
        ParsePool &pool = ParsePool::get();
        const unsigned T = pool.size();
        if (pieces_.size() < T) pieces_.resize(T);
        const int64_t avail = end_ - begin_;
        int64_t span = (int64_t)((double)max_records * avg_record_bytes_ * 1.02) + 4096;
        if (span > avail) span = avail;
        static const int64_t min_span = [] { /* B200ALN_PAR_SCAN_MIN: tests lower it to push small inputs through the parallel scan */
            const char *e = getenv("B200ALN_PAR_SCAN_MIN");
            return e ? (int64_t)atol(e) : (int64_t)1 << 20;
        }();
        int64_t total = 0;
        if (!(T < 2 || max_records < 8192 || span < min_span || span < (int64_t)T * 64 || getenv("B200ALN_SERIAL_SCAN"))) {
            const int64_t lo = begin_;
            pool.run([&](unsigned t) {
                Piece &pc = pieces_[t];
                pc.ext.clear();
                pc.n = 0;
                pc.bases = 0;
                const int64_t from = lo + span * (int64_t)t / (int64_t)T, to = lo + span * (int64_t)(t + 1) / (int64_t)T;
                int64_t p = from;
                if (t > 0) { /* the first position at or after `from` where two ordinary records follow a newline */
                    const unsigned char *b = data_;
                    p = -1;
                    for (int64_t q = from; q < to;) {
                        const unsigned char *nl = (const unsigned char *)memchr(b + q - 1, '\n', (size_t)(to - q + 1));
                        if (!nl) break;
                        const int64_t c = (int64_t)(nl - b) + 1;
                        int g = -1;
                        Extent x1, x2;
                        if (c < to && walk_one(c, g, x1) && (x1.next() >= end_ || walk_one(x1.next(), g, x2))) { p = c; break; }
                        q = c + 1;
                    }
                    if (p < 0) return;
                }
                pc.bases = walk(p, to, (size_t)max_records, pc.ext);
            });
            int64_t expect = begin_;
            for (unsigned t = 0; t < T && total < max_records; ++t) {
                Piece &pc = pieces_[t];
                if (pc.ext.empty() || pc.ext[0].start != expect) break;
                pc.n = pc.ext.size();
                if ((int64_t)pc.n > max_records - total) { /* the request is full inside this piece */
                    pc.n = (size_t)(max_records - total);
                    pc.bases = 0;
                    for (size_t i = 0; i < pc.n; ++i) pc.bases += pc.ext[i].len;
                }
                total += (int64_t)pc.n;
                expect = pc.ext[pc.n - 1].next();
                ++n_pieces_;
            }
        }
        if (total == 0) { /* a small request, or the first record is odd */
            Piece &pc = pieces_[0];
            pc.ext.clear();
            pc.bases = walk(begin_, end_, (size_t)max_records, pc.ext);
            pc.n = pc.ext.size();
            n_pieces_ = 1;
            total = (int64_t)pc.n;
        }
        if (total >= 64) {
            const Piece &last = pieces_[n_pieces_ - 1];
            avg_record_bytes_ = (double)(last.ext[last.n - 1].next() - begin_) / (double)total;
        }
        return (int)total;
    }
    /* character checks of the fast path (see read_record_fast) + conversion; false = let the exact parser decide */
    bool convert_extent(const Extent &x, uint8_t *codes, bool is_64, int trim_qual, int *len_out) const
    {
        const unsigned char *b = data_;
        const unsigned char *sq = b + x.start + x.seq, *ql = b + x.start + x.qual;
        if (!g_convert(sq, ql, x.len, codes)) return false;
        int len = x.len;
        if (trim_qual >= 1) { /* bwa_trim_read on the (offset-corrected) qualities */
            int sc = 0, best = 0, best_l = x.len - 1;
            for (int l = x.len - 1; l >= BWA_MIN_RDLEN - 1; --l) {
                const int q = (int)(unsigned char)(char)(ql[l] - (is_64 ? 31 : 0));
                sc += trim_qual - (q - 33);
                if (sc < 0) break;
                if (sc > best) { best = sc; best_l = l; }
            }
            len = best_l + 1;
        }
        *len_out = len;
        return true;
    }
    void set_cursor(int64_t pos) { begin_ = pos; }

    const std::string &seq() const { return seq_; }
    std::string &seq_mut() { return seq_; }
    std::string &qual_mut() { return qual_; }
    const std::string &name() const { return name_; }

  private:
    /* keep at least one large record's worth of bytes contiguous for the fast path */
    void refill_keep_tail(int want_bytes)
    {
        if (is_eof_ || end_ - begin_ >= want_bytes) return;
        const int64_t tail = end_ - begin_;
        if (tail > 0 && begin_ > 0) memmove(buf_.data(), buf_.data() + begin_, (size_t)tail);
        begin_ = 0;
        end_ = tail;
        const int want = (int)((int64_t)buf_.size() - end_);
        const int got = f_->read(buf_.data() + end_, (unsigned)want);
        if (got < want) is_eof_ = true;
        if (got > 0) end_ += got;
    }
    int getc()
    {
        if (is_eof_ && begin_ >= end_) return -1;
        if (begin_ >= end_) {
            begin_ = 0;
            end_ = f_->read(buf_.data(), (unsigned)buf_.size());
            if (end_ < (int64_t)buf_.size()) is_eof_ = true;
            if (end_ <= 0) { end_ = 0; return -1; }
        }
        return (int)data_[begin_++];
    }
    int get_until(int delim, std::string &str, int *dret)
    {
        if (dret) *dret = 0;
        str.clear();
        if (begin_ >= end_ && is_eof_) return -1;
        for (;;) {
            if (begin_ >= end_) {
                if (is_eof_) break;
                begin_ = 0;
                end_ = f_->read(buf_.data(), (unsigned)buf_.size());
                if (end_ < (int64_t)buf_.size()) is_eof_ = true;
                if (end_ <= 0) { end_ = 0; break; }
            }
            int64_t i = begin_;
            if (delim) while (i < end_ && data_[i] != delim) ++i;
            else while (i < end_ && !isspace(data_[i])) ++i;
            str.append((const char *)data_ + begin_, (size_t)(i - begin_));
            begin_ = i + 1;
            if (i < end_) {
                if (dret) *dret = data_[i];
                break;
            }
        }
        return (int)str.size();
    }

    std::unique_ptr<InflateSource> f_; /* zlib mode: where the bytes come from */
    std::vector<unsigned char> buf_;   /* zlib mode: the stream buffer */
    unsigned char *map_ = nullptr;     /* mapped mode: the file */
    size_t map_len_ = 0;
    const unsigned char *data_ = nullptr; /* buf_.data() or map_ */
    int64_t begin_ = 0, end_ = 0;
    double avg_record_bytes_ = 300.0; /* of the records scanned last: sizes the parallel scan's byte range */
    std::vector<Piece> pieces_;        /* scan_fast's result, one piece per worker */
    unsigned n_pieces_ = 0;
    bool is_eof_ = false;
    int last_char_ = 0;
    std::string name_, seq_, qual_;
};

/* bwa_trim_read (bwaseqio.c:74-87): length kept */
int trim_len(int trim_qual, int len, const char *qual)
{
    int s = 0, best = 0, best_l = len - 1;
    if (trim_qual < 1 || qual == nullptr) return len;
    for (int l = len - 1; l >= BWA_MIN_RDLEN - 1; --l) {
        s += trim_qual - ((unsigned char)qual[l] - 33);
        if (s < 0) break;
        if (s > best) { best = s; best_l = l; }
    }
    return best_l + 1;
}

/* The blocks of batch arrays that the driver has page-locked (ParseUnit::repin): a vector that grows gives its old
 * block back while it is still locked, so the allocator unlocks it first (cudaHostUnregister before free). */
struct PinRegistry {
    std::mutex mu;
    std::vector<const void *> blocks;
    static PinRegistry &get()
    {
        static PinRegistry r;
        return r;
    }
    void add(const void *p)
    {
        std::lock_guard<std::mutex> g(mu);
        blocks.push_back(p);
    }
    bool take(const void *p)
    { /* true: p was registered (and no longer is) */
        std::lock_guard<std::mutex> g(mu);
        for (size_t i = 0; i < blocks.size(); ++i)
            if (blocks[i] == p) {
                blocks[i] = blocks.back();
                blocks.pop_back();
                return true;
            }
        return false;
    }
};

/* std::allocator that leaves trivially constructible elements uninitialised: resize() of a code buffer that
 * is about to be overwritten by the converters does not have to zero it first */
template <class T> struct NoInitAlloc : std::allocator<T> {
    template <class U> struct rebind { typedef NoInitAlloc<U> other; };
    NoInitAlloc() = default;
    template <class U> NoInitAlloc(const NoInitAlloc<U> &) {}
    void deallocate(T *p, size_t n)
    {
        if (p && PinRegistry::get().take(p)) b200aln_unpin(p);
        std::allocator<T>::deallocate(p, n);
    }
    template <class U> void construct(U *p) noexcept { ::new ((void *)p) U; }
    template <class U, class... A> void construct(U *p, A &&...a) { ::new ((void *)p) U(std::forward<A>(a)...); }
};

struct PackedBatch {
    std::vector<int32_t, NoInitAlloc<int32_t>> lens;
    std::vector<int64_t, NoInitAlloc<int64_t>> offs;
    std::vector<uint8_t, NoInitAlloc<uint8_t>> codes;
    long n_trimmed = 0, n_tot = 0;
    void clear() { lens.clear(); offs.clear(); codes.clear(); n_trimmed = n_tot = 0; }
};

/* appends the record the reader currently holds (exact or single fast path), bwaseqio.c:158-192 */
static void append_current(SeqReader &rd, bool is_64, int l_bc, int trim_qual, PackedBatch &b)
{
    std::string &s = rd.seq_mut(), &q = rd.qual_mut();
    if (is_64) for (char &ch : q) ch = (char)(ch - 31);
    if ((int)s.size() <= l_bc) return;
    if (l_bc) {
        s.erase(0, (size_t)l_bc);
        if (!q.empty()) q.erase(0, (size_t)l_bc);
    }
    const int full = (int)s.size();
    int len = full;
    b.n_tot += full;
    if (!q.empty() && trim_qual >= 1) {
        len = trim_len(trim_qual, full, q.data());
        b.n_trimmed += full - len;
    }
    b.offs.push_back((int64_t)b.codes.size());
    b.lens.push_back(len);
    const size_t at = b.codes.size();
    b.codes.resize(at + (size_t)full);
    for (int i = 0; i < full; ++i) b.codes[at + (size_t)i] = g_nt4.t[(unsigned char)s[i]];
}

/*
 * bwa_read_seq (bwaseqio.c:145-208) into the packed form; returns reads stored.
 * Runs of ordinary 4-line FASTQ records are located by a serial structural scan (newlines only) and
 * validated + converted by worker threads; any record that is not provably parsed identically by the
 * reference's state machine (kseq.h:150-194) goes through the exact parser, one record at a time.
 */
int next_batch(SeqReader &rd, int n_needed, int mode, int trim_qual, PackedBatch &b, bool append = false)
{
    const bool is_64 = mode & BWA_MODE_IL13;
    const int l_bc = (mode >> 24) & 0xff;
    if (!append) b.clear();
    b.n_trimmed = b.n_tot = 0;
    const size_t batch_base = b.lens.size(); /* append: the batch is what this call adds */
    if (l_bc > 15) {
        fprintf(stderr, "[bwa_read_seq] the maximum barcode length is 15.\n");
        return 0;
    }
    ParsePool &pool = ParsePool::get();
    bool eof = false;
    while (!eof && (int)(b.lens.size() - batch_base) < n_needed) {
        const int n = l_bc == 0 ? rd.scan_fast(n_needed - (int)(b.lens.size() - batch_base)) : 0;
        if (n >= 64) {
            /* every piece of the scan is converted by the worker that found it, straight into its place */
            const std::vector<SeqReader::Piece> &pc = rd.pieces();
            const unsigned np = rd.n_pieces();
            const size_t base_reads = b.lens.size(), codes_base = b.codes.size();
            size_t first_read[33], first_code[33];
            first_read[0] = base_reads;
            first_code[0] = codes_base;
            for (unsigned t = 0; t < np; ++t) {
                first_read[t + 1] = first_read[t] + pc[t].n;
                first_code[t + 1] = first_code[t] + (size_t)pc[t].bases;
            }
            b.offs.resize(first_read[np]);
            b.lens.resize(first_read[np]);
            b.codes.resize(first_code[np]);
            long bad_at[32], trimmed[32]; /* per piece: the first record the fast path refuses (-1: none), bases trimmed */
            auto work = [&](unsigned t) {
                if (t >= np) return;
                bad_at[t] = -1;
                long tr = 0;
                size_t at = first_code[t];
                for (size_t i = 0; i < pc[t].n; ++i) {
                    const SeqReader::Extent &x = pc[t].ext[i];
                    int len;
                    if (!rd.convert_extent(x, b.codes.data() + at, is_64, trim_qual, &len)) {
                        bad_at[t] = (long)i;
                        break;
                    }
                    b.offs[first_read[t] + i] = (int64_t)at;
                    b.lens[first_read[t] + i] = len;
                    tr += x.len - len;
                    at += (size_t)x.len;
                }
                trimmed[t] = tr;
            };
            if (np == 1) work(0);
            else pool.run(work);
            /* keep the records before the first refused one; that one goes to the exact parser below */
            bool all = true;
            for (unsigned t = 0; t < np && all; ++t) {
                if (bad_at[t] < 0) {
                    b.n_tot += pc[t].bases;
                    b.n_trimmed += trimmed[t];
                    continue;
                }
                all = false;
                size_t kept_codes = first_code[t];
                for (long i = 0; i < bad_at[t]; ++i) {
                    kept_codes += (size_t)pc[t].ext[(size_t)i].len;
                    b.n_tot += pc[t].ext[(size_t)i].len;
                    b.n_trimmed += pc[t].ext[(size_t)i].len - b.lens[first_read[t] + (size_t)i];
                }
                b.offs.resize(first_read[t] + (size_t)bad_at[t]);
                b.lens.resize(first_read[t] + (size_t)bad_at[t]);
                b.codes.resize(kept_codes);
                rd.set_cursor(pc[t].ext[(size_t)bad_at[t]].start);
            }
            if (all) {
                const SeqReader::Piece &last = pc[np - 1];
                rd.set_cursor(last.ext[last.n - 1].next());
                continue;
            }
        }
        /* one record through the reference-exact path (also refills the buffer / detects the end) */
        const int l = rd.read_record();
        if (l < 0) eof = true;
        else append_current(rd, is_64, l_bc, trim_qual, b);
    }
    if (b.lens.size() > batch_base && trim_qual >= 1)
        fprintf(stderr, "[bwa_read_seq] %.1f%% bases are trimmed.\n", 100.0f * b.n_trimmed / b.n_tot);
    return (int)(b.lens.size() - batch_base);
}

/* BAM input (bwa_bam_open / bwa_read_bam, bwaseqio.c:21-31,89-141; record layout bamlite.c:76-104).
 * BGZF blocks are gzip members, so zlib's gzread delivers the plain BAM stream like bamlite does. */
class BamReader {
  public:
    BamReader(const char *fn, int which) : which_(which ? which : 7)
    {
        f_.reset(new InflateSource(fn));
        char magic[4];
        int32_t l_text = 0, n_ref = 0;
        if (f_->read(magic, 4) != 4 || memcmp(magic, "BAM\001", 4) != 0)
            b2host::fatal("bam_header_read", "invalid BAM binary header (this is not a BAM file).");
        rd(&l_text, 4);
        skip(l_text);
        rd(&n_ref, 4);
        for (int i = 0; i < n_ref; ++i) {
            int32_t l_name = 0, l_ref;
            rd(&l_name, 4);
            skip(l_name);
            rd(&l_ref, 4);
        }
    }

    /* bwa_read_bam: up to n_needed reads into the packed form */
    int next_batch(int n_needed, int trim_qual, PackedBatch &b, bool append = false)
    {
        static const uint8_t nt16_nt4[16] = {4, 0, 1, 4, 2, 4, 4, 4, 3, 4, 4, 4, 4, 4, 4, 4}; /* bwaseqio.c:11 */
        if (!append) b.clear();
        b.n_trimmed = b.n_tot = 0;
        const size_t base = b.lens.size();
        for (;;) {
            int32_t block_len;
            if (f_->read(&block_len, 4) != 4) break;
            uint32_t x[8];
            if (f_->read(x, 32) != 32) break;
            const int data_len = block_len - 32;
            data_.resize((size_t)(data_len > 0 ? data_len : 0));
            if (data_len > 0 && f_->read(data_.data(), (unsigned)data_len) != data_len) break;
            const uint32_t flag = x[3] >> 16, n_cigar = x[3] & 0xffff, l_qname = x[2] & 0xff;
            const int l = (int)x[4];
            bool go = false;
            if ((which_ & 1) && (flag & 64)) go = true;
            if ((which_ & 2) && (flag & 128)) go = true;
            if ((which_ & 4) && !(flag & 64) && !(flag & 128)) go = true;
            if (!go) continue;
            const uint8_t *sq = data_.data() + n_cigar * 4 + l_qname, *ql = sq + ((l + 1) >> 1);
            seq_.resize((size_t)l);
            qual_.resize((size_t)l);
            for (int i = 0; i < l; ++i) {
                seq_[i] = nt16_nt4[sq[i / 2] >> 4 * (1 - i % 2) & 0xf];
                qual_[i] = (uint8_t)(ql[i] + 33 < 126 ? ql[i] + 33 : 126);
            }
            if (flag & 16) { /* reverse strand: back to sequencing orientation (bwaseqio.c:123-126) */
                for (int i = 0; i < l / 2; ++i) {
                    uint8_t t1 = seq_[l - 1 - i], t2 = seq_[i];
                    seq_[i] = t1 < 4 ? (uint8_t)(3 - t1) : t1;
                    seq_[l - 1 - i] = t2 < 4 ? (uint8_t)(3 - t2) : t2;
                    std::swap(qual_[i], qual_[l - 1 - i]);
                }
                if (l & 1) seq_[l / 2] = seq_[l / 2] < 4 ? (uint8_t)(3 - seq_[l / 2]) : seq_[l / 2];
            }
            int len = l;
            b.n_tot += l;
            if (trim_qual >= 1) {
                len = trim_len(trim_qual, l, (const char *)qual_.data());
                b.n_trimmed += l - len;
            }
            b.offs.push_back((int64_t)b.codes.size());
            b.lens.push_back(len);
            b.codes.insert(b.codes.end(), seq_.begin(), seq_.end());
            if ((int)(b.lens.size() - base) == n_needed) break;
        }
        if (b.lens.size() > base && trim_qual >= 1)
            fprintf(stderr, "[bwa_read_seq] %.1f%% bases are trimmed.\n", 100.0f * b.n_trimmed / b.n_tot);
        return (int)(b.lens.size() - base);
    }

  private:
    void rd(void *p, int n) { if (f_->read(p, (unsigned)n) != n) b2host::fatal("bam_header_read", "truncated BAM header."); }
    void skip(int n)
    {
        char tmp[4096];
        while (n > 0) {
            int k = n < 4096 ? n : 4096;
            rd(tmp, k);
            n -= k;
        }
    }
    std::unique_ptr<InflateSource> f_;
    int which_;
    std::vector<uint8_t> data_, seq_, qual_;
};

/* mirror of the reference's bwa_seq_t (bwtaln.h:72-104) for the batch seam */
struct RefAln1 { uint32_t packed, k, l; int32_t score; };
struct RefSeq {
    char *name;
    uint8_t *seq, *rseq, *qual;
    uint32_t len : 20, strand : 1, type : 2, dummy : 1, extra_flag : 8;
    uint32_t n_mm : 8, n_gapo : 8, n_gape : 8, mapQ : 8;
    int score;
    int clip_len;
    int n_aln;
    RefAln1 *aln;
    int n_multi;
    void *multi;
    uint32_t sa;
    uint64_t pos;
    uint64_t remapped_pos;
    uint32_t dbidx;
    uint32_t remapped_dbidx;
    int32_t remapped_seqid;
    int remap_identical;
    uint64_t c1 : 28, c2 : 28, seQ : 8;
    int n_cigar;
    uint32_t *cigar;
    int tid;
    char bc[16];
    uint32_t full_len : 20, nm : 12;
    char *md;
};
static_assert(sizeof(RefSeq) == 176, "bwa_seq_t is 176 bytes on LP64 (bwtaln.h:72-104)");

} // namespace

/* bwa_open_reads (bwtaln.c:159-171): FASTA/FASTQ or, with BWA_MODE_BAM, BAM filtered by -0/-1/-2 */
struct b200aln_reader {
    std::unique_ptr<SeqReader> fq;
    std::unique_ptr<BamReader> bam;
    PackedBatch batch;
    b200aln_reader(const char *fn, int mode)
    {
        if (mode & BWA_MODE_BAM) {
            int which = 0;
            if (mode & BWA_MODE_BAM_SE) which |= 4;
            if (mode & BWA_MODE_BAM_READ1) which |= 1;
            if (mode & BWA_MODE_BAM_READ2) which |= 2;
            bam.reset(new BamReader(fn, which));
        } else fq.reset(new SeqReader(fn));
    }
    /* the next n_needed reads into b, or (append) behind what b holds; returns how many were read */
    int next(int n_needed, int mode, int trim_qual, PackedBatch &b, bool append = false)
    {
        if (bam) {
            if ((mode >> 24 & 0xff) > 15) {
                fprintf(stderr, "[bwa_read_seq] the maximum barcode length is 15.\n");
                return 0;
            }
            return bam->next_batch(n_needed, trim_qual, b, append);
        }
        return next_batch(*fq, n_needed, mode, trim_qual, b, append);
    }
};

extern "C" b200aln_reader *b200aln_reader_open(const char *fn, int mode) { return new b200aln_reader(fn, mode); }

extern "C" int b200aln_reader_next(b200aln_reader *r, int n_needed, int mode, int trim_qual, const int32_t **lens,
                                   const int64_t **offs, const uint8_t **codes, int64_t *codes_bytes)
{
    const int n = r->next(n_needed, mode, trim_qual, r->batch);
    *lens = r->batch.lens.data();
    *offs = r->batch.offs.data();
    *codes = r->batch.codes.data();
    *codes_bytes = (int64_t)r->batch.codes.size();
    return n;
}

extern "C" void b200aln_reader_close(b200aln_reader *r) { delete r; }

extern "C" void b200aln_seq_layout(b200aln_seq_layout_t *o)
{
    o->size = sizeof(RefSeq);
    o->off_name = offsetof(RefSeq, name);
    o->off_seq = offsetof(RefSeq, seq);
    o->off_rseq = offsetof(RefSeq, rseq);
    o->off_qual = offsetof(RefSeq, qual);
    o->off_lenword = offsetof(RefSeq, qual) + sizeof(void *);
    o->off_n_aln = offsetof(RefSeq, n_aln);
    o->off_aln = offsetof(RefSeq, aln);
    o->off_sa = offsetof(RefSeq, sa);
    o->off_c1c2 = offsetof(RefSeq, remap_identical) + sizeof(int);
}

extern "C" void b200aln_cal_sa_reg_gap(b200aln_ctx *ctx, int n_seqs, void *seqs_, const b200aln_opt_t *opt)
{
    RefSeq *seqs = (RefSeq *)seqs_;
    std::vector<int32_t> lens((size_t)n_seqs), n_aln((size_t)n_seqs);
    std::vector<int64_t> offs((size_t)n_seqs);
    std::vector<uint8_t> codes;
    for (int r = 0; r < n_seqs; ++r) {
        RefSeq *p = seqs + r;
        const int len = (int)p->len;
        lens[r] = len;
        offs[r] = (int64_t)codes.size();
        /* p->seq is the read reversed (bwaseqio.c:190); undo it */
        for (int j = 0; j < len; ++j) codes.push_back(p->seq[len - 1 - j]);
    }
    int64_t total = 0;
    const b200aln_rec_t *rec = b200aln_batch(ctx, n_seqs, lens.data(), offs.data(), codes.data(), opt, n_aln.data(), &total);
    int64_t at = 0;
    for (int r = 0; r < n_seqs; ++r) {
        RefSeq *p = seqs + r;
        p->sa = 0; p->type = 0; p->c1 = p->c2 = 0; /* bwtaln.c:114 */
        p->n_aln = n_aln[r];
        int cap = 4;
        while (cap < n_aln[r]) cap <<= 1; /* the reference's growth policy (bwtgap.c:113,187-191) */
        p->aln = (RefAln1 *)calloc((size_t)cap, sizeof(RefAln1));
        if (n_aln[r]) memcpy(p->aln, rec + at, sizeof(RefAln1) * (size_t)n_aln[r]);
        at += n_aln[r];
        free(p->name); free(p->seq); free(p->rseq); free(p->qual); /* bwtaln.c:134-135 */
        p->name = nullptr; p->seq = p->rseq = p->qual = nullptr;
    }
}

namespace {

/* A parse unit of the driver: up to B200ALN_MERGE consecutive reference batches (0x40000 reads each, bwtaln.c:193)
 * parsed into one set of arrays, which are page-locked so that the batch call copies straight from them.  The arrays
 * are recycled through a pool: no allocation and no page faults per batch. */
struct ParseUnit {
    PackedBatch b;
    const void *pin_p[3] = {nullptr, nullptr, nullptr};
    size_t pin_bytes[3] = {0, 0, 0};
    std::atomic<int> launches_left{0}; /* launches cut from this unit that are not finished yet */
    void repin(bool pin)
    { /* page-lock what the vectors own now (their whole capacity); a no-op while they have not moved */
        const void *p[3] = {b.lens.data(), b.offs.data(), b.codes.data()};
        const size_t n[3] = {b.lens.capacity() * 4, b.offs.capacity() * 8, b.codes.capacity()};
        for (int i = 0; i < 3; ++i) {
            if (p[i] == pin_p[i] && n[i] == pin_bytes[i]) continue;
            /* (a block the vector has given back was unlocked by the allocator: PinRegistry) */
            if (pin_p[i] && PinRegistry::get().take(pin_p[i])) b200aln_unpin(const_cast<void *>(pin_p[i]));
            pin_p[i] = nullptr;
            pin_bytes[i] = 0;
            if (pin && p[i] && n[i] && b200aln_pin(const_cast<void *>(p[i]), n[i]) == 0) {
                pin_p[i] = p[i];
                pin_bytes[i] = n[i];
                PinRegistry::get().add(p[i]);
            }
        }
    }
    /* (the vectors' allocator unlocks the blocks when they are freed) */
};

/* One launch: reads [lo, hi) of a unit — whole reference batches that agree on the batch-level max_gapo clamp. */
struct Launch {
    ParseUnit *unit = nullptr;
    int lo = 0, hi = 0;
    int64_t seq = 0; /* position in the output */
};

} // namespace

extern "C" int64_t b200aln_aln_core(const char *prefix, const char *fn_fa, const b200aln_opt_t *opt, int out_fd,
                                    int device)
{
    const bool trace = getenv("B200ALN_TRACE") != nullptr; /* wall-clock timeline of the driver on stderr */
    struct timespec tr0;
    clock_gettime(CLOCK_MONOTONIC, &tr0);
    auto stamp = [&](const char *what, long long n) {
        if (!trace) return;
        struct timespec t;
        clock_gettime(CLOCK_MONOTONIC, &t);
        fprintf(stderr, "[trace] %8.3f s  %s %lld\n", (t.tv_sec - tr0.tv_sec) + 1e-9 * (t.tv_nsec - tr0.tv_nsec), what, n);
    };
    b200aln_reader rd(fn_fa, opt->mode);
    /* devices: one, or all visible; per device n_slots contexts that share the device index (b200aln_clone), one
     * launch in flight on each */
    std::vector<int> devs;
    if (device >= 0) devs.push_back(device);
    else for (int d = 0; d < b200aln_device_count(); ++d) devs.push_back(d);
    if (devs.empty()) b2host::fatal("b200aln_aln_core", "no CUDA device available; this engine has no CPU fallback.");
    int n_slots = 4; /* launches in flight per GPU (B200ALN_INFLIGHT): the copies, the width pass and the drain of one
                      * launch's search overlap the search of the others; from four the engine parks stragglers */
    {
        const char *e = getenv("B200ALN_INFLIGHT");
        if (e) n_slots = atoi(e);
        if (n_slots < 1) n_slots = 1;
        if (n_slots > 8) n_slots = 8;
    }
    /* reads per reference batch: 0x40000 (bwtaln.c:193).  B200ALN_BATCH_READS is a hook for tests of this driver, which
     * cannot afford a quarter of a million reads per batch on the CPU; the output then equals the reference's only
     * for inputs whose batches all agree on the clamp. */
    const int batch_reads = getenv("B200ALN_BATCH_READS") ? std::max(1, atoi(getenv("B200ALN_BATCH_READS"))) : 0x40000;
    int merge = 8;
    {
        const char *e = getenv("B200ALN_MERGE");
        if (e) merge = atoi(e);
        if (merge < 1) merge = 1;
        if (merge > 64) merge = 64;
    }
    const bool pin = getenv("B200ALN_NO_PIN") == nullptr;
    const int n_workers = (int)devs.size() * n_slots;
    std::mutex mu;
    std::condition_variable cv_work, cv_unit, cv_turn;
    std::deque<Launch> queue;           /* launches waiting for a worker, in output order */
    std::vector<ParseUnit *> free_units; /* the pool */
    std::vector<std::unique_ptr<ParseUnit>> units;
    bool no_more = false;
    int64_t next_write = 0, written = 0;
    /* the pool's arrays are allocated and page-locked by a helper thread, one unit after the other, while the
     * first units are already at work */
    for (int i = 0; i < n_workers + 2; ++i) units.emplace_back(new ParseUnit);
    std::atomic<bool> parsed_all{false};
    std::thread pool_maker([&]() {
        for (auto &u : units) {
            if (parsed_all.load()) break; /* the input has ended: what is not prepared yet is not needed */
            u->b.lens.reserve((size_t)merge * batch_reads);
            u->b.offs.reserve((size_t)merge * batch_reads);
            u->b.codes.reserve((size_t)merge * batch_reads * 104);
            u->repin(pin);
            stamp("unit ready (arrays reserved and page-locked), MB", (long long)((u->pin_bytes[0] + u->pin_bytes[1] + u->pin_bytes[2]) >> 20));
            std::lock_guard<std::mutex> lk(mu);
            free_units.push_back(u.get());
            cv_unit.notify_one();
        }
    });
    std::vector<std::vector<b200aln_ctx *>> slot_ctx((size_t)n_slots);
    /* what the scratch allocated ahead of the contexts is sized by: the longest read of the first batch, known to
     * the setup thread once this thread has parsed it (-1: not yet, 0: there are no reads) */
    int first_max_len = -1;
    std::vector<std::thread> prealloc;
    auto load_index = [&]() {
        /* bwt_restore_bwt x2 once (bwtio.c:51-70) — both files at the same time, while the CUDA contexts come up
         * and the contexts' device buffers are allocated on other threads — then one upload per GPU in parallel */
        struct Raw { uint32_t *w = nullptr; size_t nw = 0; uint32_t hdr[5]; };
        Raw raw[2];
        b200aln_bwt_view_t v[2];
        const char *ext[2] = {".bwt", ".rbwt"};
        auto load = [&](int j) {
            const std::string fn = std::string(prefix) + ext[j];
            const int fd = open(fn.c_str(), O_RDONLY);
            if (fd < 0) b2host::fatal("b200aln_aln_core", (std::string("fail to open file '") + fn + "'.").c_str());
            struct stat st;
            if (fstat(fd, &st) != 0 || st.st_size < 20 || pread(fd, raw[j].hdr, 20, 0) != 20)
                b2host::fatal("b200aln_aln_core", "truncated BWT file.");
            const size_t nw = ((size_t)st.st_size - 20) >> 2;
            raw[j].nw = nw;
            raw[j].w = (uint32_t *)malloc(nw * 4 + 16);
            if (!raw[j].w) b2host::fatal("b200aln_aln_core", "out of memory reading the BWT.");
            size_t got = 0;
            while (got < nw * 4) {
                const ssize_t k = pread(fd, (char *)raw[j].w + got, nw * 4 - got, (off_t)(20 + got));
                if (k <= 0) b2host::fatal("b200aln_aln_core", "truncated BWT file.");
                got += (size_t)k;
            }
            close(fd);
        };
        std::thread warm([&]() { for (int d : devs) b200aln_warm_device(d); });
        std::thread t1([&]() { load(1); });
        load(0);
        t1.join();
        stamp("index files read", 2);
        warm.join();
        for (int j = 0; j < 2; ++j) {
            v[j].primary = raw[j].hdr[0];
            v[j].L2[0] = 0;
            for (int i = 0; i < 4; ++i) v[j].L2[i + 1] = raw[j].hdr[1 + i];
            v[j].seq_len = v[j].L2[4];
            v[j].bwt_size = raw[j].nw;
            v[j].bwt = raw[j].w;
        }
        for (auto &sc : slot_ctx) sc.resize(devs.size());
        std::vector<std::thread> th;
        for (size_t i = 0; i < devs.size(); ++i)
            th.emplace_back([&, i]() {
                slot_ctx[0][i] = b200aln_open(&v[0], &v[1], devs[i]);
                for (int sl = 1; sl < n_slots; ++sl) slot_ctx[(size_t)sl][i] = b200aln_clone(slot_ctx[0][i]);
            });
        for (auto &t : th) t.join();
        free(raw[0].w);
        free(raw[1].w);
        for (auto &sc : slot_ctx)
            for (b200aln_ctx *c : sc) {
                b200aln_set_int(c, "reserve_reads", (int64_t)merge * batch_reads + 1);
                if (const char *e = getenv("B200ALN_SET")) { /* engine knobs for experiments: "key=value,key=value" (b200aln_set_int) */
                    std::string kv(e);
                    for (size_t at = 0; at < kv.size();) {
                        size_t end = kv.find(',', at);
                        if (end == std::string::npos) end = kv.size();
                        const std::string one = kv.substr(at, end - at);
                        const size_t eq = one.find('=');
                        if (eq != std::string::npos) b200aln_set_int(c, one.substr(0, eq).c_str(), atoll(one.c_str() + eq + 1));
                        at = end + 1;
                    }
                }
            }
        stamp("index resident on devices", (long long)devs.size());
    };
    FILE *out = fdopen(dup(out_fd), "wb");
    if (!out) b2host::fatal("b200aln_aln_core", "cannot open the output descriptor.");
    fwrite(opt, sizeof(b200aln_opt_t), 1, out); /* bwtaln.c:192 */

    /* The reference works in batches of 0x40000 reads (bwtaln.c:193), and one thing is decided per batch: the
     * max_gapo clamp from the batch's longest read (bwtaln.c:89-92).  0x40000 reads are only two per lane of one
     * GPU, so consecutive reference batches that agree on the clamp go out as ONE launch (up to B200ALN_MERGE of
     * them).  Three stages run side by side:
     *   this thread parses units of up to B200ALN_MERGE batches into recycled page-locked arrays (the first units
     *   are smaller, so that the GPUs start at once and short inputs are not held back) and cuts them into launches
     *   — from the start, while a setup thread is still reading the index files and opening the contexts;
     *   one worker thread per (GPU, slot) takes the next launch, runs the operator — which returns the launch as
     *   the bytes of the .sai stream, formatted on the device — and
     *   writes them when every earlier launch has been written. */
    auto clamp_key = [&](int max_len) { /* what make_params derives from the batch's longest read */
        int gapo = opt->max_gapo;
        if (opt->fnr > 0.0f) {
            const int d = b2host::cal_maxdiff(max_len, BWA_AVG_ERR, opt->fnr);
            if (d < gapo) gapo = d;
        }
        return gapo;
    };

    auto worker = [&](int w) {
        b200aln_ctx *ctx = slot_ctx[(size_t)(w / (int)devs.size())][(size_t)(w % (int)devs.size())];
        for (;;) {
            Launch L;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return !queue.empty() || no_more; });
                if (queue.empty()) return;
                L = queue.front();
                queue.pop_front();
            }
            struct timespec t0, t1, t2;
            clock_gettime(CLOCK_MONOTONIC, &t0);
            stamp("worker takes launch", (long long)L.seq);
            const PackedBatch &b = L.unit->b;
            int max_len = 0;
            for (int r = L.lo; r < L.hi; ++r) if (b.lens[(size_t)r] > max_len) max_len = b.lens[(size_t)r];
            b200aln_set_int(ctx, "batch_max_len", max_len);
            int64_t n_bytes = 0;
            const void *sai = b200aln_batch_sai(ctx, L.hi - L.lo, b.lens.data() + L.lo, b.offs.data() + L.lo, b.codes.data(),
                                                opt, &n_bytes);
            clock_gettime(CLOCK_MONOTONIC, &t1);
            if (trace) {
                b200aln_stats_t st;
                b200aln_last_stats(ctx, &st);
                fprintf(stderr, "[trace]            launch %lld on worker %d: %d reads, call %.1f ms (h2d %.1f width %.1f search %.1f compact %.1f d2h %.1f)\n",
                        (long long)L.seq, w, L.hi - L.lo, 1e3 * ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec)),
                        st.ms_h2d, st.ms_width, st.ms_search, st.ms_compact, st.ms_d2h);
            }
            std::unique_lock<std::mutex> lk(mu);
            if (--L.unit->launches_left == 0) { /* the unit's arrays are free again */
                free_units.push_back(L.unit);
                cv_unit.notify_one();
            }
            cv_turn.wait(lk, [&] { return next_write == L.seq; });
            lk.unlock(); /* (only the launch whose turn it is gets here) */
            stamp("launch finished, reads", (long long)(L.hi - L.lo));
            if (n_bytes && fwrite(sai, 1, (size_t)n_bytes, out) != (size_t)n_bytes)
                b2host::fatal("b200aln_aln_core", "short write on the .sai output.");
            clock_gettime(CLOCK_MONOTONIC, &t2);
            written += L.hi - L.lo;
            fprintf(stderr, "[bwa_aln_core] calculate SA coordinate... %.2f sec\n", (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec));
            fprintf(stderr, "[bwa_aln_core] write to the disk... %.2f sec\n", (t2.tv_sec - t1.tv_sec) + 1e-9 * (t2.tv_nsec - t1.tv_nsec));
            fprintf(stderr, "[bwa_aln_core] %lld sequences have been processed.\n", (long long)written);
            lk.lock();
            ++next_write;
            cv_turn.notify_all();
        }
    };
    std::vector<std::thread> workers;
    std::thread setup([&]() {
        load_index();
        for (int w = 0; w < n_workers; ++w) workers.emplace_back(worker, w);
    });

    int64_t tot_seqs = 0, seq = 0;
    int unit_batches = 1;
    for (bool eof = false; !eof;) {
        ParseUnit *u;
        {
            std::unique_lock<std::mutex> lk(mu);
            if (free_units.empty()) stamp("parser waits for a free unit, reads so far", (long long)tot_seqs);
            cv_unit.wait(lk, [&] { return !free_units.empty(); });
            u = free_units.front(); /* (oldest first: the ones the helper has prepared come before recycled ones) */
            free_units.erase(free_units.begin());
        }
        PackedBatch &b = u->b;
        b.clear();
        /* consecutive reference batches; a launch is cut where the batch-level clamp changes */
        std::vector<Launch> cut;
        int cur_lo = 0, cur_key = 0, cur_max = 0;
        for (int k = 0; k < unit_batches; ++k) {
            const int at = (int)b.lens.size();
            const int n = rd.next(batch_reads, opt->mode, opt->trim_qual, b, true);
            if (n == 0) { eof = true; break; }
            tot_seqs += n;
            stamp("parsed reads", (long long)tot_seqs);
            int max_len = 0;
            for (int r = at; r < at + n; ++r) if (b.lens[(size_t)r] > max_len) max_len = b.lens[(size_t)r];
            const int key = clamp_key(max_len);
            if (at > cur_lo) {
                const int joint = max_len > cur_max ? max_len : cur_max;
                if (key != cur_key || clamp_key(joint) != cur_key) {
                    Launch L;
                    L.unit = u; L.lo = cur_lo; L.hi = at;
                    cut.push_back(L);
                    cur_lo = at;
                }
            }
            if (at == cur_lo) { cur_key = key; cur_max = max_len; }
            else if (max_len > cur_max) cur_max = max_len;
            if (n < batch_reads) { eof = true; break; }
        }
        if (first_max_len < 0) {
            /* the contexts' device buffers, sized by the first batch and allocated ahead on one thread per device while
             * the setup thread is still reading the index files (a context takes its set with its first launch) */
            first_max_len = b.lens.empty() ? 0 : cur_max;
            /* (an input that ends inside its first batch is one launch on one context: nothing to allocate ahead) */
            if (first_max_len > 0 && !eof && !getenv("B200ALN_NO_PREALLOC"))
                for (int d : devs)
                    prealloc.emplace_back([&, d]() {
                        b200aln_prealloc(d, n_slots, merge * batch_reads + 1, first_max_len);
                        stamp("device buffers of the contexts allocated ahead, device", (long long)d);
                    });
        }
        if ((int)b.lens.size() > cur_lo) {
            Launch L;
            L.unit = u; L.lo = cur_lo; L.hi = (int)b.lens.size();
            cut.push_back(L);
        }
        std::unique_lock<std::mutex> lk(mu);
        if (cut.empty()) free_units.push_back(u);
        else {
            lk.unlock();
            u->repin(pin);
            lk.lock();
            u->launches_left = (int)cut.size();
            for (Launch &L : cut) {
                L.seq = seq++;
                stamp("launch of reads", (long long)(L.hi - L.lo));
                queue.push_back(L);
            }
            cv_work.notify_all();
        }
        if (unit_batches < merge) unit_batches = unit_batches * 2 < merge ? unit_batches * 2 : merge;
    }
    parsed_all.store(true);
    {
        std::lock_guard<std::mutex> lk(mu);
        no_more = true;
    }
    setup.join();
    cv_work.notify_all();
    for (auto &t : workers) t.join();
    for (auto &t : prealloc) t.join();
    if (!getenv("B200ALN_FAST_EXIT")) for (int d : devs) b200aln_prealloc_release(d); /* sets no launch came to use */
    pool_maker.join();
    fclose(out);
    stamp("output closed, reads", (long long)tot_seqs);
    if (getenv("B200ALN_FAST_EXIT")) for (auto &u : units) (void)u.release(); /* (unlocking 0.2 GB per unit takes 0.1 s) */
    units.clear();
    if (!getenv("B200ALN_FAST_EXIT")) /* (the command line: the process ends here, the driver's teardown frees the device) */
        for (size_t i = 0; i < devs.size(); ++i)
            for (int sl = n_slots - 1; sl >= 0; --sl) b200aln_close(slot_ctx[(size_t)sl][i]);
    stamp("contexts closed", (long long)devs.size());
    return tot_seqs;
}

extern "C" int b200aln_aln_main(int argc, char *argv[])
{ /* bwa_aln, bwtaln.c:243-328 */
    int c, opte = -1;
    b200aln_opt_t o;
    b200aln_opt_init(&o);
    optind = 1;
    while ((c = getopt(argc, argv, "n:o:e:i:d:l:k:cLR:m:t:NM:O:E:q:f:b012IB:")) >= 0) {
        switch (c) {
        case 'n':
            if (strstr(optarg, ".")) { o.fnr = (float)atof(optarg); o.max_diff = -1; }
            else { o.max_diff = atoi(optarg); o.fnr = -1.0f; }
            break;
        case 'o': o.max_gapo = atoi(optarg); break;
        case 'e': opte = atoi(optarg); break;
        case 'M': o.s_mm = atoi(optarg); break;
        case 'O': o.s_gapo = atoi(optarg); break;
        case 'E': o.s_gape = atoi(optarg); break;
        case 'd': o.max_del_occ = atoi(optarg); break;
        case 'i': o.indel_end_skip = atoi(optarg); break;
        case 'l': o.seed_len = atoi(optarg); break;
        case 'k': o.max_seed_diff = atoi(optarg); break;
        case 'm': o.max_entries = atoi(optarg); break;
        case 't': o.n_threads = atoi(optarg); break;
        case 'L': o.mode |= BWA_MODE_LOGGAP; break;
        case 'R': o.max_top2 = atoi(optarg); break;
        case 'q': o.trim_qual = atoi(optarg); break;
        case 'c': o.mode &= ~BWA_MODE_COMPREAD; break;
        case 'N': o.mode |= BWA_MODE_NONSTOP; o.max_top2 = 0x7fffffff; break;
        case 'f':
            if (freopen(optarg, "wb", stdout) == nullptr) {
                fprintf(stderr, "[bwa_aln] fail to open file '%s': ", optarg);
                perror(nullptr);
                fprintf(stderr, "Abort!\n");
                abort();
            }
            break;
        case 'b': o.mode |= BWA_MODE_BAM; break;
        case '0': o.mode |= BWA_MODE_BAM_SE; break;
        case '1': o.mode |= BWA_MODE_BAM_READ1; break;
        case '2': o.mode |= BWA_MODE_BAM_READ2; break;
        case 'I': o.mode |= BWA_MODE_IL13; break;
        case 'B': o.mode |= atoi(optarg) << 24; break;
        default: return 1;
        }
    }
    if (opte > 0) {
        o.max_gape = opte;
        o.mode &= ~BWA_MODE_GAPE;
    }
    if (optind + 2 > argc) {
        fprintf(stderr, "\nUsage:   b200aln aln [options] <prefix> <in.fq>\n\n");
        fprintf(stderr, "Options: -n NUM    max #diff (int) or missing prob under %.2f err rate (float) [%.2f]\n", BWA_AVG_ERR, o.fnr);
        fprintf(stderr, "         -o INT    maximum number or fraction of gap opens [%d]\n", o.max_gapo);
        fprintf(stderr, "         -e INT    maximum number of gap extensions, -1 for disabling long gaps [-1]\n");
        fprintf(stderr, "         -i INT    do not put an indel within INT bp towards the ends [%d]\n", o.indel_end_skip);
        fprintf(stderr, "         -d INT    maximum occurrences for extending a long deletion [%d]\n", o.max_del_occ);
        fprintf(stderr, "         -l INT    seed length [%d]\n", o.seed_len);
        fprintf(stderr, "         -k INT    maximum differences in the seed [%d]\n", o.max_seed_diff);
        fprintf(stderr, "         -m INT    maximum entries in the queue [%d]\n", o.max_entries);
        fprintf(stderr, "         -t INT    number of threads (recorded in the header; the GPU engine ignores it) [%d]\n", o.n_threads);
        fprintf(stderr, "         -M INT    mismatch penalty [%d]\n", o.s_mm);
        fprintf(stderr, "         -O INT    gap open penalty [%d]\n", o.s_gapo);
        fprintf(stderr, "         -E INT    gap extension penalty [%d]\n", o.s_gape);
        fprintf(stderr, "         -R INT    stop searching when there are >INT equally best hits [%d]\n", o.max_top2);
        fprintf(stderr, "         -q INT    quality threshold for read trimming down to %dbp [%d]\n", BWA_MIN_RDLEN, o.trim_qual);
        fprintf(stderr, "         -f FILE   file to write output to instead of stdout\n");
        fprintf(stderr, "         -B INT    length of barcode\n");
        fprintf(stderr, "         -c        input sequences are in the color space\n");
        fprintf(stderr, "         -L        log-scaled gap penalty for long deletions\n");
        fprintf(stderr, "         -N        non-iterative mode: search for all n-difference hits (slooow)\n");
        fprintf(stderr, "         -I        the input is in the Illumina 1.3+ FASTQ-like format\n");
        fprintf(stderr, "         -b        the input read file is in the BAM format\n");
        fprintf(stderr, "         -0        use single-end reads only (effective with -b)\n");
        fprintf(stderr, "         -1        use the 1st read in a pair (effective with -b)\n");
        fprintf(stderr, "         -2        use the 2nd read in a pair (effective with -b)\n\n");
        return 1;
    }
    if (o.fnr > 0.0f) { /* bwtaln.c:317-324 */
        for (int i = 17, k = 0; i <= 250; ++i) {
            int l = b200aln_cal_maxdiff(i, BWA_AVG_ERR, o.fnr);
            if (l != k) fprintf(stderr, "[bwa_aln] %dbp reads: max_diff = %d\n", i, l);
            k = l;
        }
    }
    fflush(stdout);
    {
        const char *e = getenv("B200ALN_DEVICE"); /* default: every visible GPU */
        b200aln_aln_core(argv[optind], argv[optind + 1], &o, fileno(stdout), e ? atoi(e) : -1);
    }
    return 0;
}
