// Host-side readers of the two big text inputs:
//   * the Beagle genotype-likelihood file (replaces reader_cy.pyx:16-77), and
//   * the allele-depth matrix of the z-score modes (the reference's np.loadtxt, WGSassign.py:320, :399).
//
// Beagle format (reader_cy.pyx:31-68): gzip text, fields separated by runs of tab/space; header
// `marker allele1 allele2` then every sample name three times; each data row: site id, two
// allele codes (ignored), then three GLs per individual of which the third is dropped.
// Output contract: float32 [M, 2N] whose values equal (float)atof(token) bit for bit, plus
// the sample and site name lists.
//
// Pipeline (GzLines): inflating one gzip stream is inherently serial - a dedicated thread runs it
// and hands decompressed blocks, cut at line boundaries, through a bounded queue, so inflating
// block b+1.. overlaps the parsing of block b.  A BGZF file (bgzip; what ANGSD writes) is a chain of
// independent members whose sizes stand in their headers: those are inflated by several threads at once.  Parsing is not serial: a persistent pool of
// threads converts the lines of a block straight into the CALLER's row-major output (pinned
// memory in the CLI - no intermediate copy), and rows outside the caller's keep range (another
// rank's sites) are counted and named but never converted.  Plain "digits.digits" tokens (what
// ANGSD writes) are converted exactly: an integer below 2^53 divided by an exact power of ten is
// one correctly rounded double operation, i.e. the same double strtod returns; anything else
// goes to strtod.
#include "../../include/wgsassign_b200.h"

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

std::string g_reader_error;

inline bool is_delim(char c) { return c == '\t' || c == ' ' || c == '\n' || c == '\r'; }

const double kPow10[] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};

// (float)atof(tok) for tok = [p, q)
inline float parse_float(const char* p, const char* q)
{
    const char* s = p;
    bool neg = false;
    if (s < q && (*s == '-' || *s == '+')) { neg = *s == '-'; ++s; }
    unsigned long long v = 0;
    int digits = 0, frac = 0;
    bool seen_dot = false, simple = s < q;
    for (const char* c = s; c < q; ++c) {
        if (*c >= '0' && *c <= '9') {
            v = v * 10 + (unsigned)(*c - '0');
            if (v) ++digits;
            if (seen_dot) ++frac;
        } else if (*c == '.' && !seen_dot) {
            seen_dot = true;
        } else { simple = false; break; }
    }
    if (simple && digits <= 15 && frac <= 15) {
        double d = (double)v / kPow10[frac];
        return (float)(neg ? -d : d);
    }
    std::string tmp(p, q);
    return (float)atof(tmp.c_str());
}

bool blank(const char* p, const char* e)
{
    for (; p < e; ++p) if (!is_delim(*p)) return false;
    return true;
}

// ---- persistent worker pool: run(n, fn) calls fn(i) for i in [0, n) on all threads, chunks handed out dynamically ----
class Pool {
public:
    explicit Pool(int threads) : n_(std::max(1, threads)) {
        for (int t = 1; t < n_; ++t) th_.emplace_back([this] { loop(); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    void run(long n, long grain, const std::function<void(long, long)>& fn) {
        if (n <= 0) return;
        if (n_ == 1 || n <= grain) { fn(0, n); return; }
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &fn; total_ = n; grain_ = grain; next_.store(0); pending_ = n_ - 1; ++gen_;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }
    int size() const { return n_; }
private:
    void work() {
        for (;;) {
            const long a = next_.fetch_add(grain_);
            if (a >= total_) break;
            (*fn_)(a, std::min(total_, a + grain_));
        }
    }
    void loop() {
        unsigned long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            work();
            { std::lock_guard<std::mutex> lk(m_); if (--pending_ == 0) done_.notify_one(); }
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(long, long)>* fn_ = nullptr;
    long total_ = 0, grain_ = 1;
    std::atomic<long> next_{0};
    int pending_ = 0;
    unsigned long gen_ = 0;
    bool stop_ = false;
};

// ---- gzip text -> blocks of whole lines, inflated by a background thread ----
// A text block: raw storage that is never value-initialised (zero-filling 32 MB per block cost as much as parsing it)
// and that goes back to the producer for reuse when the consumer takes the next block (no fresh pages to fault in).
struct Block {
    std::unique_ptr<char[]> buf;
    size_t cap = 0, size = 0, end = 0;
    char* data() { return buf.get(); }
    const char* data() const { return buf.get(); }
    void reserve(size_t n) {                              // keeps the first `size` bytes
        if (n <= cap) return;
        std::unique_ptr<char[]> nb(new char[n]);
        if (size) memcpy(nb.get(), buf.get(), size);
        buf = std::move(nb);
        cap = n;
    }
    void assign(const std::string& t) { size = 0; reserve(t.size()); if (!t.empty()) memcpy(buf.get(), t.data(), t.size()); size = t.size(); }
};

// BGZF (bgzip; what ANGSD itself writes): a gzip file made of independent members of at most 64 KB of text, each with
// its compressed size in a "BC" extra subfield of the header and its text size in the trailer - the members of a batch
// can be located without inflating anything and inflated in parallel straight into their places of one text buffer.
// Returns the member's total size, 0 if [p, p + avail) does not hold the whole header yet, -1 if it is not BGZF.
inline long bgzf_member_size(const unsigned char* p, size_t avail)
{
    if (avail < 12) return 0;
    if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return -1;
    const size_t xlen = (size_t)p[10] | ((size_t)p[11] << 8);
    if (avail < 12 + xlen) return 0;
    for (size_t o = 12; o + 4 <= 12 + xlen;) {
        const size_t slen = (size_t)p[o + 2] | ((size_t)p[o + 3] << 8);
        if (p[o] == 'B' && p[o + 1] == 'C' && slen == 2 && o + 6 <= 12 + xlen)
            return (long)((size_t)p[o + 4] | ((size_t)p[o + 5] << 8)) + 1;
        o += 4 + slen;
    }
    return -1;
}

class GzLines {
public:
    // inflate_threads: workers of the BGZF path (a plain gzip stream is inflated by the one background thread).
    // part / parts (BGZF only): this reader yields the LINES THAT START in the part-th of `parts` equal byte ranges of
    // the compressed file - more exactly, with t0 / t1 the text positions where the members that start inside the range
    // begin / end, the lines whose first character lies in (t0, t1] (part 0: [0, t1]).  Every line of the file belongs to
    // exactly one part, the parts' lines concatenated in order are the file, and no part inflates more than its own
    // members plus the end of its last line.
    bool open(const char* path, std::string* err, int inflate_threads = 1, int part = 0, int parts = 1) {
        FILE* f = fopen(path, "rb");
        if (!f) { *err = std::string("cannot open ") + path; return false; }
        unsigned char head[64];
        const size_t got = fread(head, 1, sizeof head, f);
        fseek(f, 0, SEEK_END); compressed_size_ = ftell(f);
        const bool bgzf = bgzf_member_size(head, got) > 0;
        if (parts > 1 && !bgzf) {
            fclose(f);
            *err = std::string("byte-range parts need a BGZF (bgzip) file: ") + path;
            return false;
        }
        if (bgzf) {
            raw_ = f;
            nth_ = std::max(1, std::min(inflate_threads, 32));
            part_ = part; parts_ = std::max(parts, 1);
            if (parts_ > 1) {
                const long size = compressed_size_;
                const long lo = (long)((__int128)size * part / parts_);
                hi_ = (long)((__int128)size * (part + 1) / parts_);
                start_ = part == 0 ? 0 : find_member_start(f, lo, size);
                if (start_ < 0 || start_ >= hi_) { start_ = hi_; }            // no member starts inside the range: no lines
                compressed_size_ = hi_ - start_;
            }
            fseek(f, start_, SEEK_SET);
        } else {
            fclose(f);
            gz_ = gzopen(path, "rb");
            if (!gz_) { *err = std::string("cannot open ") + path; return false; }
            gzbuffer(gz_, 1 << 20);
        }
        th_ = std::thread([this] { if (raw_) bgzf_loop(); else inflate_loop(); });
        return true;
    }
    ~GzLines() {
        { std::lock_guard<std::mutex> lk(m_); abort_ = true; }
        cv_space_.notify_all();
        if (th_.joinable()) th_.join();
        if (gz_) gzclose(gz_);
        if (raw_) fclose(raw_);
    }
    bool is_bgzf() const { return raw_ != nullptr; }
    // next block of whole lines; false at end of file or on error (failed())
    bool next(Block* out) {
        std::unique_lock<std::mutex> lk(m_);
        if (out->cap) { out->size = out->end = 0; spare_.push_back(std::move(*out)); *out = Block(); }   // the consumed block's storage
        cv_data_.wait(lk, [this] { return !q_.empty() || eof_ || failed_; });
        if (q_.empty()) return false;
        *out = std::move(q_.front());
        q_.pop_front();
        lk.unlock();
        cv_space_.notify_one();
        return true;
    }
    bool failed() const { return failed_; }
    long compressed_size() const { return compressed_size_; }
    long compressed_pos() const { return compressed_pos_.load(); }
    long uncompressed_bytes() const { return uncompressed_.load(); }
    double inflate_seconds() const { return inflate_s_.load(); }
private:
    // cut a text block at its last line end (unless it is the file's last), keep the rest for the next one and queue it;
    // false when the consumer went away
    bool push_text(Block& b, std::string& carry, bool last) {
        size_t end = b.size;
        if (!last) {
            while (end > 0 && b.data()[end - 1] != '\n') --end;
            if (end == 0) { carry.assign(b.data(), b.size); recycle(b); return true; }   // one line longer than the block
        }
        carry.assign(b.data() + end, b.size - end);
        b.end = end;
        std::unique_lock<std::mutex> lk(m_);
        cv_space_.wait(lk, [this] { return q_.size() < 3 || abort_; });
        if (abort_) return false;
        q_.push_back(std::move(b));
        lk.unlock();
        cv_data_.notify_one();
        return true;
    }
    void fail_now() { std::lock_guard<std::mutex> lk(m_); failed_ = true; cv_data_.notify_all(); }
    Block fresh() {
        std::lock_guard<std::mutex> lk(m_);
        if (spare_.empty()) return Block();
        Block b = std::move(spare_.back());
        spare_.pop_back();
        return b;
    }
    void recycle(Block& b) { std::lock_guard<std::mutex> lk(m_); b.size = b.end = 0; spare_.push_back(std::move(b)); b = Block(); }

    struct Member { size_t in_off; unsigned csize, isize; size_t out_off; };
    // raw inflate of one member's deflate payload into dst; checks the text size and the CRC-32 of the trailer
    static bool inflate_member(const unsigned char* p, unsigned csize, unsigned isize, char* dst) {
        const size_t xlen = (size_t)p[10] | ((size_t)p[11] << 8), hdr = 12 + xlen;
        if ((size_t)csize < hdr + 8) return false;
        char nothing = 0;
        if (!dst) dst = &nothing;                          // an empty member into an empty block: zlib refuses a null next_out
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, -15) != Z_OK) return false;
        zs.next_in = const_cast<unsigned char*>(p + hdr);
        zs.avail_in = (unsigned)(csize - hdr - 8);
        zs.next_out = reinterpret_cast<unsigned char*>(dst);
        zs.avail_out = isize;
        const int rc = inflate(&zs, Z_FINISH);
        const bool ok = rc == Z_STREAM_END && zs.total_out == isize;
        inflateEnd(&zs);
        if (!ok) return false;
        const unsigned char* t = p + csize - 8;
        const unsigned long want = (unsigned long)t[0] | ((unsigned long)t[1] << 8) | ((unsigned long)t[2] << 16) | ((unsigned long)t[3] << 24);
        return crc32(crc32(0L, Z_NULL, 0), reinterpret_cast<const unsigned char*>(dst), isize) == want;
    }
    // first member start at or after `from`: a header whose own size leads to another valid header (or to the end of the
    // file), twice in a row - four magic bytes plus the BC subfield plus two chained sizes do not occur by accident
    static long find_member_start(FILE* f, long from, long size) {
        if (from >= size) return -1;
        std::vector<unsigned char> w((size_t)std::min<long>(size - from, 320 * 1024));
        fseek(f, from, SEEK_SET);
        const size_t got = fread(w.data(), 1, w.size(), f);
        for (size_t o = 0; o + 18 <= got && o <= 65536 + 18; ++o) {
            if (w[o] != 0x1f || w[o + 1] != 0x8b || w[o + 2] != 8 || !(w[o + 3] & 4)) continue;
            size_t q = o;
            int chained = 0;
            bool ok = true;
            while (chained < 3) {
                const long sz = bgzf_member_size(w.data() + q, got - q);
                if (sz <= 0) { ok = (long)(from + q) == size && chained > 0; break; }   // clean end of file after >= 1 member
                q += (size_t)sz;
                ++chained;
                if ((long)(from + q) == size) break;
                if (q + 18 > got) break;                                          // window exhausted: accept what chained
            }
            if (ok && chained > 0) return from + (long)o;
        }
        return -1;
    }
    // text of the members from the current file position up to and including the first line end (or the end of file)
    bool read_tail_line(std::string* out) {
        std::vector<unsigned char> m(1 << 17);
        std::vector<char> text(1 << 16);
        for (;;) {
            const size_t got = fread(m.data(), 1, 18, raw_);
            if (got == 0) return true;                                            // end of file
            if (got < 18) return false;
            const size_t xlen = (size_t)m[10] | ((size_t)m[11] << 8);
            if (12 + xlen > m.size()) return false;
            if (12 + xlen > 18 && fread(m.data() + 18, 1, 12 + xlen - 18, raw_) != 12 + xlen - 18) return false;
            const long sz = bgzf_member_size(m.data(), 12 + xlen);
            if (sz <= 0 || (size_t)sz > m.size() || (size_t)sz < 12 + xlen + 8) return false;
            const size_t had = std::max<size_t>(18, 12 + xlen);
            if (fread(m.data() + had, 1, (size_t)sz - had, raw_) != (size_t)sz - had) return false;
            const unsigned char* t = m.data() + sz - 4;
            const unsigned isize = (unsigned)t[0] | ((unsigned)t[1] << 8) | ((unsigned)t[2] << 16) | ((unsigned)t[3] << 24);
            if (isize > text.size()) text.resize(isize);
            if (!inflate_member(m.data(), (unsigned)sz, isize, text.data())) return false;
            const char* nl = (const char*)memchr(text.data(), '\n', isize);
            out->append(text.data(), nl ? (size_t)(nl - text.data()) + 1 : isize);
            if (nl) return true;
        }
    }
    void bgzf_loop() {
        const size_t RAW = (size_t)8 << 20, TEXT = (size_t)32 << 20;
        std::vector<unsigned char> raw(RAW + (1 << 17));
        size_t have = 0, pos = 0;
        long raw0 = start_;                                         // file offset of raw[0]
        long consumed = 0;
        std::string carry;
        bool file_done = false;
        const bool ranged = parts_ > 1;
        bool skipping = ranged && part_ > 0;                        // a part other than the first drops its text through the first line end
        bool own_done = false;                                      // the next member starts outside this part's byte range
        if (ranged && start_ >= hi_) {                              // no member of its own: no lines
            { std::lock_guard<std::mutex> lk(m_); eof_ = true; }
            cv_data_.notify_all();
            return;
        }
        for (;;) {
            if (pos > 0) { memmove(raw.data(), raw.data() + pos, have - pos); have -= pos; raw0 += (long)pos; pos = 0; }
            if (!file_done && have < RAW) {
                const size_t got = fread(raw.data() + have, 1, RAW - have, raw_);
                if (got == 0) file_done = true;
                have += got;
            }
            std::vector<Member> mem;
            size_t text = 0;
            while (pos < have && text < TEXT) {
                if (ranged && raw0 + (long)pos >= hi_) { own_done = true; break; }
                const long sz = bgzf_member_size(raw.data() + pos, have - pos);
                if (sz < 0 || (sz == 0 && file_done)) { fail_now(); return; }      // not BGZF after all, or a truncated member
                if (sz == 0 || pos + (size_t)sz > have) {
                    if (file_done) { fail_now(); return; }
                    break;                                                         // the rest of this member is still in the file
                }
                const unsigned char* t = raw.data() + pos + sz - 4;
                const unsigned isize = (unsigned)t[0] | ((unsigned)t[1] << 8) | ((unsigned)t[2] << 16) | ((unsigned)t[3] << 24);
                mem.push_back({pos, (unsigned)sz, isize, text});
                pos += (size_t)sz;
                text += isize;
            }
            const bool at_eof = file_done && pos == have;
            const bool last = at_eof || own_done;
            if (mem.empty() && !last) continue;                                    // (a member is at most 64 KB: the next read completes it)
            Block b = fresh();
            b.assign(carry);
            const size_t off = b.size;
            b.reserve(off + text);
            b.size = off + text;
            const auto t0 = std::chrono::steady_clock::now();
            std::atomic<size_t> nexti{0};
            std::atomic<bool> bad{false};
            auto work = [&] {
                for (size_t i; (i = nexti.fetch_add(1)) < mem.size();)
                    if (!inflate_member(raw.data() + mem[i].in_off, mem[i].csize, mem[i].isize, b.data() + off + mem[i].out_off)) bad = true;
            };
            std::vector<std::thread> ws;
            for (int w = 1; w < nth_ && (size_t)w < mem.size(); ++w) ws.emplace_back(work);
            work();
            for (auto& w : ws) w.join();
            inflate_s_.store(inflate_s_.load() + std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            if (bad) { fail_now(); return; }
            uncompressed_.fetch_add((long)text);
            for (const Member& m : mem) consumed += m.csize;
            compressed_pos_.store(consumed);
            if (skipping) {                                                        // (carry is empty while skipping)
                const char* nl = (const char*)memchr(b.data(), '\n', b.size);
                if (nl) {
                    const size_t drop = (size_t)(nl - b.data()) + 1;
                    memmove(b.data(), b.data() + drop, b.size - drop);
                    b.size -= drop;
                    skipping = false;
                } else {
                    b.size = 0;
                }
            }
            if (last && own_done && !skipping) {
                // the line in progress at the end of this part's text - or, if that text ends with a line end, the line
                // that starts exactly there - is finished from the members of the next part
                std::string tail;
                fseek(raw_, raw0 + (long)pos, SEEK_SET);
                if (!read_tail_line(&tail)) { fail_now(); return; }
                b.reserve(b.size + tail.size());
                memcpy(b.data() + b.size, tail.data(), tail.size());
                b.size += tail.size();
            }
            if (last && skipping) b.size = 0;                                      // no line starts inside this part
            if (!push_text(b, carry, last)) return;
            if (last) break;
        }
        { std::lock_guard<std::mutex> lk(m_); eof_ = true; }
        cv_data_.notify_all();
    }
    void inflate_loop() {
        const size_t CH = (size_t)32 << 20;
        std::string carry;
        bool eof = false;
        while (!eof) {
            Block b = fresh();
            b.assign(carry);
            const size_t off = b.size;
            b.reserve(off + CH);
            const auto t0 = std::chrono::steady_clock::now();
            const int got = gzread(gz_, b.data() + off, (unsigned)CH);
            inflate_s_.store(inflate_s_.load() + std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            if (got < 0) { fail_now(); return; }
            b.size = off + (size_t)got;
            uncompressed_.fetch_add(got);
            compressed_pos_.store((long)gzoffset(gz_));
            eof = (size_t)got < CH;
            if (!push_text(b, carry, eof)) return;
        }
        { std::lock_guard<std::mutex> lk(m_); eof_ = true; }
        cv_data_.notify_all();
    }
    gzFile gz_ = nullptr;
    FILE* raw_ = nullptr;                                 // BGZF: the compressed file itself
    int nth_ = 1;
    int part_ = 0, parts_ = 1;                            // BGZF byte-range parts
    long start_ = 0, hi_ = 0;                             // first own member / end of the byte range
    std::thread th_;
    std::mutex m_;
    std::condition_variable cv_data_, cv_space_;
    std::deque<Block> q_;
    std::vector<Block> spare_;                            // storage of consumed blocks, reused by the producer
    bool eof_ = false, failed_ = false, abort_ = false;
    long compressed_size_ = 0;
    std::atomic<long> compressed_pos_{0}, uncompressed_{0};
    std::atomic<double> inflate_s_{0.0};
};

// "D.DDDDDD" (what ANGSD's %f writes) at p, already known to be followed by a delimiter or the line end: the seven
// digits as one integer, or -1 if the eight characters are anything else.  (double)v / 1e6 is the one correctly rounded
// operation that parse_float performs for such a token, so the value is identical.
inline long fixed6(const char* p)
{
    const unsigned d0 = (unsigned)(p[0] - '0'), d2 = (unsigned)(p[2] - '0'), d3 = (unsigned)(p[3] - '0'), d4 = (unsigned)(p[4] - '0'),
                   d5 = (unsigned)(p[5] - '0'), d6 = (unsigned)(p[6] - '0'), d7 = (unsigned)(p[7] - '0');
    if ((p[1] != '.') | (d0 > 9) | (d2 > 9) | (d3 > 9) | (d4 > 9) | (d5 > 9) | (d6 > 9) | (d7 > 9)) return -1;
    return (long)(d0 * 1000000u + d2 * 100000u + d3 * 10000u + d4 * 1000u + d5 * 100u + d6 * 10u + d7);
}

// parse one Beagle data line [p, e) into out[0 .. 2n); returns false on a short line
bool parse_line(const char* p, const char* e, int n_ind, float* out)
{
    auto skip_delims = [&] { while (p < e && is_delim(*p)) ++p; };
    auto skip_token = [&] { while (p < e && !is_delim(*p)) ++p; };
    for (int k = 0; k < 3; ++k) {                         // site id (kept by the caller), allele1, allele2
        skip_delims();
        if (p >= e) return false;
        skip_token();
    }
    for (int i = 0; i < n_ind; ++i) {
#pragma GCC unroll 3
        for (int g = 0; g < 3; ++g) {                     // the third GL is dropped (reader_cy.pyx:62-63)
            if (p < e && is_delim(*p)) { ++p; if (p < e && is_delim(*p)) skip_delims(); }
            if (p >= e) return false;
            const long room = e - p;
            if (room >= 8 && (room == 8 || is_delim(p[8]))) {     // an eight-character token: the fixed-point fast path
                const long v = fixed6(p);
                if (v >= 0) {
                    if (g < 2) out[2 * i + g] = (float)((double)v / 1e6);
                    p += 8;
                    continue;
                }
            }
            const char* a = p;
            skip_token();
            if (g < 2) out[2 * i + g] = parse_float(a, p);
        }
    }
    return true;
}

// one allele-depth row: 2n non-negative integers -> saturating uint8 (255 = "255 or more") or int32
template <class T>
bool parse_ad_line(const char* p, const char* e, int ncol, T* out, bool* negative)
{
    for (int c = 0; c < ncol; ++c) {
        while (p < e && is_delim(*p)) ++p;
        if (p >= e) return false;
        {   // one- and two-digit counts (nearly every token of a low-coverage file) without the general loop
            const unsigned d0 = (unsigned)(p[0] - '0');
            if (d0 <= 9) {
                if (p + 1 == e || is_delim(p[1])) { out[c] = (T)d0; p += 1; continue; }
                const unsigned d1 = (unsigned)(p[1] - '0');
                if (d1 <= 9 && (p + 2 == e || is_delim(p[2]))) { out[c] = (T)(d0 * 10 + d1); p += 2; continue; }
            }
        }
        bool neg = false;
        if (*p == '-') { neg = true; ++p; } else if (*p == '+') ++p;
        long v = 0;
        bool any = false;
        while (p < e && *p >= '0' && *p <= '9') { v = std::min(v * 10 + (*p - '0'), 2000000000L); ++p; any = true; }
        if (p < e && *p == '.') { ++p; while (p < e && *p >= '0' && *p <= '9') ++p; }   // "3.0" as np.loadtxt(dtype=int32) would refuse; accept the integer part
        if (!any || (p < e && !is_delim(*p))) return false;
        if (neg && v != 0) *negative = true;
        if (sizeof(T) == 1) out[c] = (T)std::min(v, 255L);
        else out[c] = (T)(neg ? -v : v);
    }
    while (p < e && is_delim(*p)) ++p;
    return p >= e;                                         // no extra columns
}

struct LineRef { const char *p, *e; };

// ---- streaming state shared by the two readers ----
struct Stream {
    GzLines gz;
    Pool* pool = nullptr;
    int n_ind = 0, ncol = 0;                              // Beagle: individuals; AD: columns (2N)
    bool is_ad = false, header_done = false;
    std::vector<std::string> samples;
    Block cur;                                            // block being consumed
    std::vector<LineRef> lines;
    size_t line_pos = 0;
    long rows_seen = 0;                                   // data rows handed out or skipped so far
    long keep_lo = 0, keep_hi = -1;                       // rows to convert (keep_hi < 0: all)
    std::vector<std::string> sites;                       // Beagle: names of ALL rows seen (cheap; the CLI prints and intersects them)
    bool keep_names = true;
    double parse_s = 0.0;
    ~Stream() { delete pool; }
};

bool index_block(Stream* S)
{
    S->lines.clear();
    S->line_pos = 0;
    const char* p = S->cur.data();
    const char* e = p + S->cur.end;
    if (!S->is_ad && !S->header_done) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
        const char* he = nl ? nl : e;
        int tok = 0;
        const char* q = p;
        while (q < he) {
            while (q < he && is_delim(*q)) ++q;
            if (q >= he) break;
            const char* a = q;
            while (q < he && !is_delim(*q)) ++q;
            ++tok;
            if (tok > 3 && (tok - 3) % 3 == 1) S->samples.emplace_back(a, q);   // every 3rd GL column names a sample
        }
        if (tok < 6) { g_reader_error = "Beagle header has fewer than 6 columns"; return false; }
        S->n_ind = (tok - 3) / 3;
        S->header_done = true;
        p = nl ? nl + 1 : e;
    }
    while (p < e) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
        const char* le = nl ? nl : e;
        if (!blank(p, le)) {
            if (S->is_ad && !S->header_done) {                // width of the matrix from its first row
                int tok = 0;
                const char* q = p;
                while (q < le) { while (q < le && is_delim(*q)) ++q; if (q >= le) break; while (q < le && !is_delim(*q)) ++q; ++tok; }
                S->ncol = tok;
                S->header_done = true;
            }
            S->lines.push_back(LineRef{p, le});
        }
        p = nl ? nl + 1 : e;
    }
    return true;
}

// make sure a block with unread lines is current; false at end of input
bool ensure_lines(Stream* S)
{
    while (S->line_pos >= S->lines.size()) {
        if (!S->gz.next(&S->cur)) {
            if (S->gz.failed()) g_reader_error = "gzread failed (corrupt gzip stream?)";
            return false;
        }
        if (!index_block(S)) return false;
    }
    return true;
}

Stream* open_stream(const char* path, int threads, bool is_ad, int part = 0, int parts = 1)
{
    Stream* S = new Stream();
    S->is_ad = is_ad;
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    if (parts > 1 && part > 0 && !is_ad) {
        // the header line belongs to part 0: read it from the start of the file (one member is inflated)
        Stream* H = open_stream(path, 1, false);
        if (!H) { delete S; return nullptr; }
        S->samples = H->samples;
        S->n_ind = H->n_ind;
        S->header_done = true;
        delete H;
    }
    // BGZF input: three quarters of the threads inflate - with the fixed-point token path a core parses the text about
    // eight times faster than it inflates it (measured on the 16-core GPU host: 200 MB of Beagle text inflated in
    // 0.31 core-seconds x 4 threads, parsed in 0.01 s of wall time by 16)
    if (!S->gz.open(path, &g_reader_error, std::max(2, threads * 3 / 4), part, parts)) { delete S; return nullptr; }
    S->pool = new Pool(std::min(threads, 64));
    g_reader_error.clear();
    if (!ensure_lines(S) && (!S->header_done || !g_reader_error.empty())) {
        if (g_reader_error.empty()) g_reader_error = is_ad ? "empty allele-depth file" : "empty Beagle file";
        delete S;
        return nullptr;
    }
    return S;
}

// convert up to max_rows of the rows inside the keep range into out (row-major, `width` elements per row);
// rows outside the range are consumed without conversion.  Returns rows written, 0 at end of input, -1 on error.
template <class T, class ParseFn>
long stream_next(Stream* S, T* out, long max_rows, int width, ParseFn parse)
{
    long written = 0;
    while (written < max_rows) {
        if (!ensure_lines(S)) { if (!g_reader_error.empty()) return -1; break; }
        const long avail = (long)(S->lines.size() - S->line_pos);
        // rows of this block before the keep range: skip (names only)
        long skip = 0;
        if (S->rows_seen < S->keep_lo) skip = std::min(avail, S->keep_lo - S->rows_seen);
        const bool past = S->keep_hi >= 0 && S->rows_seen + skip >= S->keep_hi;
        long take = past ? 0 : avail - skip;
        if (S->keep_hi >= 0) take = std::min(take, std::max(0L, S->keep_hi - (S->rows_seen + skip)));
        take = std::min(take, max_rows - written);
        const long consume = past ? avail : skip + take;
        if (!S->is_ad && S->keep_names) {
            const size_t base = S->sites.size();
            S->sites.resize(base + (size_t)consume);
            S->pool->run(consume, 4096, [&](long a, long b) {
                for (long r = a; r < b; ++r) {
                    const LineRef& ln = S->lines[S->line_pos + (size_t)r];
                    const char* p = ln.p;
                    while (p < ln.e && is_delim(*p)) ++p;
                    const char* q = p;
                    while (q < ln.e && !is_delim(*q)) ++q;
                    S->sites[base + (size_t)r].assign(p, q);
                }
            });
        }
        if (take > 0) {
            std::atomic<int> bad{0};
            const auto t0 = std::chrono::steady_clock::now();
            const size_t first = S->line_pos + (size_t)skip;
            S->pool->run(take, 16, [&](long a, long b) {
                for (long r = a; r < b; ++r) {
                    const LineRef& ln = S->lines[first + (size_t)r];
                    if (!parse(ln.p, ln.e, out + (size_t)(written + r) * width)) bad.store(1);
                }
            });
            S->parse_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (bad.load()) {
                g_reader_error = std::string(S->is_ad ? "allele-depth row" : "Beagle row") + " with a wrong number of columns or a malformed field near row " +
                                 std::to_string(S->rows_seen + skip);
                return -1;
            }
        }
        S->line_pos += (size_t)consume;
        S->rows_seen += consume;
        written += take;
        if (past && !S->keep_names) {                      // nothing more to keep and nobody wants the names: stop reading
            break;
        }
    }
    return written;
}

}  // namespace

extern "C" {

const char* wgs_beagle_last_error(void) { return g_reader_error.c_str(); }

// ---- streaming interface ----
int32_t wgs_beagle_stream_open(const char* path, int32_t threads, wgs_beagle_stream** out)
{
    *out = nullptr;
    Stream* S = open_stream(path, threads, false);
    if (!S) return 1;
    S->ncol = 2 * S->n_ind;
    *out = (wgs_beagle_stream*)S;
    return 0;
}
// The part-th of `parts` byte ranges of a BGZF Beagle file (see GzLines::open): the rows that start inside it, in file
// order; the parts of a file are disjoint, cover it and can be read by different processes at the same time, each
// inflating only its own share.  Row numbers (wgs_beagle_stream_keep, rows_seen, site) count from the part's first row.
// Fails for a plain gzip file (one deflate stream cannot be entered in the middle).
int32_t wgs_beagle_stream_open_part(const char* path, int32_t threads, int32_t part, int32_t parts, wgs_beagle_stream** out)
{
    *out = nullptr;
    if (parts < 1 || part < 0 || part >= parts) { g_reader_error = "part must lie in [0, parts)"; return 1; }
    Stream* S = open_stream(path, threads, false, part, parts);
    if (!S) return 1;
    S->ncol = 2 * S->n_ind;
    *out = (wgs_beagle_stream*)S;
    return 0;
}
int32_t wgs_beagle_stream_inds(const wgs_beagle_stream* s) { return ((const Stream*)s)->n_ind; }
const char* wgs_beagle_stream_sample(const wgs_beagle_stream* s, int32_t i) { return ((const Stream*)s)->samples[(size_t)i].c_str(); }
int32_t wgs_beagle_stream_keep(wgs_beagle_stream* s, int64_t row_lo, int64_t row_hi)
{
    Stream* S = (Stream*)s;
    S->keep_lo = row_lo; S->keep_hi = row_hi;
    return 0;
}
int32_t wgs_beagle_stream_names(wgs_beagle_stream* s, int32_t keep) { ((Stream*)s)->keep_names = keep != 0; return 0; }
int64_t wgs_beagle_stream_next(wgs_beagle_stream* s, float* out, int64_t max_rows)
{
    Stream* S = (Stream*)s;
    const int n = S->n_ind;
    return stream_next<float>(S, out, max_rows, 2 * n, [n](const char* p, const char* e, float* o) { return parse_line(p, e, n, o); });
}
int64_t wgs_beagle_stream_rows_seen(const wgs_beagle_stream* s) { return ((const Stream*)s)->rows_seen; }
// all site names seen so far, each followed by '\n', into out[0 .. cap); returns the bytes needed (call with out = NULL
// first).  One call instead of one per site: 20 M names through a per-name FFI call take longer than parsing them.
int64_t wgs_beagle_stream_sites_joined(const wgs_beagle_stream* s, char* out, int64_t cap)
{
    const Stream* S = (const Stream*)s;
    int64_t need = 0;
    for (const std::string& t : S->sites) need += (int64_t)t.size() + 1;
    if (!out || cap < need) return need;
    char* w = out;
    for (const std::string& t : S->sites) { memcpy(w, t.data(), t.size()); w += t.size(); *w++ = '\n'; }
    return need;
}
const char* wgs_beagle_stream_site(const wgs_beagle_stream* s, int64_t row) { return ((const Stream*)s)->sites[(size_t)row].c_str(); }
// estimate of the total number of data rows from what has been read so far (exact once the stream is exhausted):
// compressed size x the compression ratio and bytes per row seen so far
int64_t wgs_beagle_stream_estimate_rows(const wgs_beagle_stream* s)
{
    const Stream* S = (const Stream*)s;
    const long cpos = S->gz.compressed_pos(), csize = S->gz.compressed_size(), ubytes = S->gz.uncompressed_bytes();
    // rows indexed so far: those handed out plus the rest of the current block; the queue holds more bytes than rows we counted, so use the block sizes
    const long rows_indexed = S->rows_seen + (long)(S->lines.size() - S->line_pos);
    if (cpos <= 0 || ubytes <= 0 || rows_indexed <= 0 || csize <= 0) return -1;
    // bytes per row from the CURRENT block only (its size is known exactly)
    const double bpr = S->lines.empty() ? 0.0 : (double)S->cur.end / (double)S->lines.size();
    if (bpr <= 0.0) return -1;
    const double total_u = (double)ubytes * ((double)csize / (double)cpos);
    return (int64_t)(total_u / bpr) + 1;
}
// 1 when the stream is BGZF (members inflated in parallel), 0 for a plain gzip stream; either kind of stream handle
int32_t wgs_stream_is_bgzf(const void* s) { return s && ((const Stream*)s)->gz.is_bgzf() ? 1 : 0; }

int32_t wgs_beagle_stream_stats(const wgs_beagle_stream* s, double* inflate_s, double* parse_s, int64_t* compressed_bytes, int64_t* uncompressed_bytes)
{
    const Stream* S = (const Stream*)s;
    *inflate_s = S->gz.inflate_seconds(); *parse_s = S->parse_s;
    *compressed_bytes = S->gz.compressed_size(); *uncompressed_bytes = S->gz.uncompressed_bytes();
    return 0;
}
void wgs_beagle_stream_close(wgs_beagle_stream* s) { delete (Stream*)s; }

// ---- allele depths: whitespace-separated integers, optionally gzipped (zlib reads plain text transparently) ----
int32_t wgs_ad_stream_open(const char* path, int32_t threads, wgs_ad_stream** out)
{
    *out = nullptr;
    Stream* S = open_stream(path, threads, true);
    if (!S) return 1;
    if (S->ncol <= 0 || (S->ncol & 1)) { g_reader_error = "allele-depth rows must hold two counts per individual"; delete S; return 1; }
    S->n_ind = S->ncol / 2;
    *out = (wgs_ad_stream*)S;
    return 0;
}
int32_t wgs_ad_stream_inds(const wgs_ad_stream* s) { return ((const Stream*)s)->n_ind; }
int32_t wgs_ad_stream_keep(wgs_ad_stream* s, int64_t row_lo, int64_t row_hi) { return wgs_beagle_stream_keep((wgs_beagle_stream*)s, row_lo, row_hi); }
int64_t wgs_ad_stream_next_u8(wgs_ad_stream* s, uint8_t* out, int64_t max_rows)
{
    Stream* S = (Stream*)s;
    const int nc = S->ncol;
    bool negative = false;
    const long r = stream_next<uint8_t>(S, out, max_rows, nc, [nc, &negative](const char* p, const char* e, uint8_t* o) { return parse_ad_line<uint8_t>(p, e, nc, o, &negative); });
    if (r >= 0 && negative) { g_reader_error = "negative allele depth"; return -1; }
    return r;
}
int64_t wgs_ad_stream_next_i32(wgs_ad_stream* s, int32_t* out, int64_t max_rows)
{
    Stream* S = (Stream*)s;
    const int nc = S->ncol;
    bool negative = false;
    return stream_next<int32_t>(S, out, max_rows, nc, [nc, &negative](const char* p, const char* e, int32_t* o) { return parse_ad_line<int32_t>(p, e, nc, o, &negative); });
}
int64_t wgs_ad_stream_rows_seen(const wgs_ad_stream* s) { return ((const Stream*)s)->rows_seen; }
int64_t wgs_ad_stream_estimate_rows(const wgs_ad_stream* s) { return wgs_beagle_stream_estimate_rows((const wgs_beagle_stream*)s); }
int32_t wgs_ad_stream_stats(const wgs_ad_stream* s, double* inflate_s, double* parse_s, int64_t* compressed_bytes, int64_t* uncompressed_bytes)
{
    return wgs_beagle_stream_stats((const wgs_beagle_stream*)s, inflate_s, parse_s, compressed_bytes, uncompressed_bytes);
}
void wgs_ad_stream_close(wgs_ad_stream* s) { delete (Stream*)s; }

// ---- whole-file interface (kept for callers that want the reference's one-call behaviour) ----
struct Beagle {
    std::vector<std::string> samples, sites;
    std::vector<std::vector<float>> blocks;               // row-major pieces, concatenated on copy-out
    int n_ind = 0;
    long rows = 0;
};

int32_t wgs_beagle_open(const char* path, int32_t threads, wgs_beagle** out)
{
    *out = nullptr;
    wgs_beagle_stream* st = nullptr;
    if (wgs_beagle_stream_open(path, threads, &st)) return 1;
    Stream* S = (Stream*)st;
    Beagle* B = new Beagle();
    B->n_ind = S->n_ind;
    const long chunk = std::max<long>(1, ((long)64 << 20) / std::max(1, 8 * S->n_ind));
    for (;;) {
        std::vector<float> blk((size_t)chunk * 2 * S->n_ind);
        const long r = wgs_beagle_stream_next(st, blk.data(), chunk);
        if (r < 0) { wgs_beagle_stream_close(st); delete B; return 1; }
        if (r == 0) break;
        blk.resize((size_t)r * 2 * S->n_ind);
        B->blocks.push_back(std::move(blk));
        B->rows += r;
    }
    B->samples = S->samples;
    B->sites = std::move(S->sites);
    wgs_beagle_stream_close(st);
    *out = (wgs_beagle*)B;
    return 0;
}

int64_t wgs_beagle_sites(const wgs_beagle* b) { return ((const Beagle*)b)->rows; }
int32_t wgs_beagle_inds(const wgs_beagle* b) { return ((const Beagle*)b)->n_ind; }
const char* wgs_beagle_sample(const wgs_beagle* b, int32_t i) { return ((const Beagle*)b)->samples[(size_t)i].c_str(); }
const char* wgs_beagle_site(const wgs_beagle* b, int64_t s) { return ((const Beagle*)b)->sites[(size_t)s].c_str(); }

int32_t wgs_beagle_copy(const wgs_beagle* b, float* L_out)
{
    const Beagle* B = (const Beagle*)b;
    size_t off = 0;
    for (size_t k = 0; k < B->blocks.size(); ++k) {
        memcpy(L_out + off, B->blocks[k].data(), B->blocks[k].size() * sizeof(float));
        off += B->blocks[k].size();
    }
    return 0;
}

void wgs_beagle_close(wgs_beagle* b) { delete (Beagle*)b; }

}  // extern "C"
