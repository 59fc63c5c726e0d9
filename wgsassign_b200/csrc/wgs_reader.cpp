// Host-side Beagle genotype-likelihood reader (replaces reader_cy.pyx:16-77).
//
// Format (reader_cy.pyx:31-68): gzip text, fields separated by runs of tab/space; header
// `marker allele1 allele2` then every sample name three times; each data row: site id, two
// allele codes (ignored), then three GLs per individual of which the third is dropped.
// Output contract: float32 [M, 2N] whose values equal (float)atof(token) bit for bit, plus
// the sample and site name lists.
//
// inflate is inherently serial (one gzip stream); parsing is not: each decompressed block
// is cut at line boundaries and its lines are parsed by a pool of threads straight into
// the row-major output.  Plain "digits.digits" tokens (what ANGSD writes) are converted
// exactly: an integer below 2^53 divided by an exact power of ten is one correctly rounded
// double operation, i.e. the same double strtod returns; anything else goes to strtod.
#include "../../include/wgsassign_b200.h"

#include <zlib.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Beagle {
    std::vector<std::string> samples, sites;
    std::vector<std::vector<float>> blocks;   // row-major pieces, concatenated on copy-out
    std::vector<long> block_rows;
    int n_ind = 0;
    long rows = 0;
    std::string err;
};

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

// parse one data line [p, e) into out[0 .. 2n); returns false on a short line
bool parse_line(const char* p, const char* e, int n_ind, float* out, std::string* site)
{
    auto next = [&](const char*& a, const char*& b) -> bool {
        while (p < e && is_delim(*p)) ++p;
        if (p >= e) return false;
        a = p;
        while (p < e && !is_delim(*p)) ++p;
        b = p;
        return true;
    };
    const char *a, *b;
    if (!next(a, b)) return false;
    site->assign(a, b);
    if (!next(a, b) || !next(a, b)) return false;     // allele1, allele2
    for (int i = 0; i < n_ind; ++i) {
        if (!next(a, b)) return false;
        out[2 * i] = parse_float(a, b);
        if (!next(a, b)) return false;
        out[2 * i + 1] = parse_float(a, b);
        if (!next(a, b)) return false;                // third GL: dropped (reader_cy.pyx:62-63)
    }
    return true;
}

bool blank(const char* p, const char* e)
{
    for (; p < e; ++p) if (!is_delim(*p)) return false;
    return true;
}

}  // namespace

extern "C" {

typedef struct wgs_beagle wgs_beagle;

const char* wgs_beagle_last_error(void) { return g_reader_error.c_str(); }

int32_t wgs_beagle_open(const char* path, int32_t threads, wgs_beagle** out)
{
    *out = nullptr;
    gzFile gz = gzopen(path, "rb");
    if (!gz) { g_reader_error = std::string("cannot open ") + path; return 1; }
    gzbuffer(gz, 1 << 20);
    Beagle* B = new Beagle();
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    threads = std::min(threads, 64);

    const size_t CH = (size_t)64 << 20;
    std::vector<char> buf;
    std::string carry;
    bool header_done = false, eof = false;
    while (!eof) {
        buf.assign(carry.begin(), carry.end());
        size_t off = buf.size();
        buf.resize(off + CH);
        int got = gzread(gz, buf.data() + off, (unsigned)CH);
        if (got < 0) { g_reader_error = "gzread failed (corrupt gzip stream?)"; gzclose(gz); delete B; return 1; }
        buf.resize(off + (size_t)got);
        eof = (size_t)got < CH;
        // cut at the last newline unless this is the final block
        size_t end = buf.size();
        if (!eof) {
            while (end > 0 && buf[end - 1] != '\n') --end;
            if (end == 0) { carry.assign(buf.begin(), buf.end()); continue; }   // one line longer than the block
        }
        carry.assign(buf.begin() + end, buf.end());
        const char* p = buf.data();
        const char* e = buf.data() + end;
        if (!header_done) {
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
                if (tok > 3 && (tok - 3) % 3 == 1) B->samples.emplace_back(a, q);   // every 3rd GL column names a sample
            }
            if (tok < 6) { g_reader_error = "Beagle header has fewer than 6 columns"; gzclose(gz); delete B; return 1; }
            B->n_ind = (tok - 3) / 3;
            header_done = true;
            p = nl ? nl + 1 : e;
        }
        // index the lines of this block
        std::vector<std::pair<const char*, const char*>> lines;
        while (p < e) {
            const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
            const char* le = nl ? nl : e;
            if (!blank(p, le)) lines.emplace_back(p, le);
            p = nl ? nl + 1 : e;
        }
        if (lines.empty()) continue;
        const long nl_ = (long)lines.size();
        const int n = B->n_ind;
        B->blocks.emplace_back((size_t)nl_ * 2 * n);
        B->block_rows.push_back(nl_);
        float* dst = B->blocks.back().data();
        std::vector<std::string> names((size_t)nl_);
        std::vector<int> bad(threads, 0);
        auto work = [&](int t) {
            for (long r = t; r < nl_; r += threads)
                if (!parse_line(lines[r].first, lines[r].second, n, dst + (size_t)r * 2 * n, &names[r])) bad[t] = 1;
        };
        if (threads == 1 || nl_ < 64) { for (int t = 0; t < threads; ++t) work(t); }
        else {
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
            for (auto& th : pool) th.join();
        }
        for (int t = 0; t < threads; ++t) if (bad[t]) {
            g_reader_error = "Beagle row with fewer columns than the header near site " + std::to_string(B->rows);
            gzclose(gz); delete B; return 1;
        }
        for (auto& s : names) B->sites.push_back(std::move(s));
        B->rows += nl_;
    }
    gzclose(gz);
    if (!header_done) { g_reader_error = "empty Beagle file"; delete B; return 1; }
    *out = (wgs_beagle*)B;
    return 0;
}

int64_t wgs_beagle_sites(const wgs_beagle* b) { return ((const Beagle*)b)->rows; }
int32_t wgs_beagle_inds(const wgs_beagle* b) { return ((const Beagle*)b)->n_ind; }
const char* wgs_beagle_sample(const wgs_beagle* b, int32_t i) { return ((const Beagle*)b)->samples[i].c_str(); }
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
