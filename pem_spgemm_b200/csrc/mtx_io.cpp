// Matrix Market coordinate reader / writer (host side of the C ABI).
//
// Replaces the reference's use of the un-vendored fast_matrix_market v1.7.6 in
// read_matrix_market (/root/reference/spgemm.cu:43-110): header + triplets, `pattern` values
// become 1.0, `complex` keeps the real part (:99-107), symmetric / skew-symmetric / hermitian
// files are expanded to general form.  Unlike fast_matrix_market's default, a diagonal entry of
// a symmetric file is emitted ONCE (no extra zero-valued duplicate): duplicates are undefined
// input for the tile conversion.
//
// The body is parsed by all host threads: the file is mapped (no read into a staging buffer: the page
// faults of the mapping are taken by the parsing threads, in parallel), cut at line boundaries into one
// chunk per thread, the chunks' lines are counted as newline bytes, and every chunk is parsed in ONE pass
// (digit loops for the coordinates, std::from_chars for the value) straight into its place of the output.
//
// The writers (Matrix Market files, and the one-number-per-line files of the reference's COO dump,
// spgemm.cu:1545-1560) format with std::to_chars on all host threads: every thread fills its own buffer with
// a slice of the lines, the byte offsets follow from the slice lengths, and the threads write their slices
// with pwrite at those offsets.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "fixed17.h"
#include "pemspgemm.h"

namespace {

void set_err(char* err, size_t n, const std::string& msg)
{
    if (err && n) snprintf(err, n, "%s", msg.c_str());
}

inline const char* skip_ws(const char* p, const char* e)
{
    while (p < e && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
    return p;
}

inline const char* parse_i64(const char* p, const char* e, long long& v, bool& ok)
{
    p = skip_ws(p, e);
    if (p < e && *p == '+') ++p;
    auto r = std::from_chars(p, e, v);
    ok = ok && r.ec == std::errc();
    return r.ptr;
}

inline const char* parse_f64(const char* p, const char* e, double& v, bool& ok)
{
    p = skip_ws(p, e);
    if (p < e && *p == '+') ++p;
    auto r = std::from_chars(p, e, v);
    if (r.ec != std::errc()) {  // inf/nan spellings or exotic formats: fall back to strtod
        char* endp = nullptr;
        std::string tmp(p, std::min<size_t>((size_t)(e - p), 64));
        v = strtod(tmp.c_str(), &endp);
        if (endp == tmp.c_str()) { ok = false; return p; }
        return p + (endp - tmp.c_str());
    }
    return r.ptr;
}

struct Chunk {
    const char* b;
    const char* e;
    size_t lines = 0;
    size_t out = 0;
    bool ok = true;
};

inline bool blank_or_comment(const char* p, const char* e)
{
    p = skip_ws(p, e);
    return p >= e || *p == '%' || *p == '\n';
}

// a read-only mapping of a whole file
struct Mapping {
    const char* p = nullptr;
    size_t n = 0;
    ~Mapping() { if (p && n) munmap(const_cast<char*>(p), n); }
};

bool write_all(int fd, const char* b, size_t n, long long at)
{
    while (n) {
        const ssize_t w = pwrite(fd, b, n, (off_t)at);
        if (w <= 0) return false;
        b += w; n -= (size_t)w; at += w;
    }
    return true;
}

// n lines, formatted by fmt(i, p) -> end of line i (at most max_len bytes, newline included), appended to fd at
// byte `at`.  Rounds of nt slices: every thread formats its slice of the round, learns its byte offset from the
// lengths of the slices before it, writes the slice with pwrite and goes on to the next round without waiting for
// the others' writes (writes to one file serialise in the kernel; the formatting of the next round runs under them).
// Returns the end offset, or -1 on a write error.
template <class Fmt>
long long write_lines(int fd, long long at, int64_t n, size_t max_len, Fmt fmt)
{
    unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    if (n < (1 << 16)) nt = 1;
    const int64_t per = std::max<int64_t>(1, std::min<int64_t>((n + nt - 1) / nt, 1 << 18));   // lines per slice
    const int64_t rounds = (n + per * nt - 1) / (per * nt);
    std::vector<size_t> len((size_t)2 * nt);              // slice lengths of the even / odd rounds
    std::atomic<long long> formatted{0};                  // slices formatted so far, all rounds
    std::atomic<bool> ok{true};
    std::vector<long long> end_at(nt, at);
    auto work = [&](unsigned t) {
        std::vector<char> b;
        long long base = at;                              // where the current round starts in the file
        for (int64_t r = 0; r < rounds; ++r) {
            const int64_t lo = std::min<int64_t>(n, (r * nt + t) * per), hi = std::min<int64_t>(n, lo + per);
            if (b.size() < max_len) b.resize(std::max<size_t>(max_len, (size_t)(hi - lo) * 24));
            size_t pos = 0;
            for (int64_t i = lo; i < hi; ++i) {
                if (b.size() - pos < max_len) b.resize(b.size() * 2);
                pos = (size_t)(fmt(i, b.data() + pos) - b.data());
            }
            size_t* L = len.data() + (size_t)(r & 1) * nt;
            L[t] = pos;
            formatted.fetch_add(1, std::memory_order_release);
            while (formatted.load(std::memory_order_acquire) < (r + 1) * (long long)nt) std::this_thread::yield();
            long long mine = base;
            for (unsigned u = 0; u < nt; ++u) {
                if (u == t) mine = base;
                base += (long long)L[u];
            }
            if (pos && !write_all(fd, b.data(), pos, mine)) ok = false;
        }
        end_at[t] = base;
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    return ok ? end_at[0] : -1;
}

// open for a fresh file or for appending; *at = the offset the new lines start at
int open_out(const char* path, int append, long long* at)
{
    const int fd = open(path, O_WRONLY | O_CREAT | (append ? 0 : O_TRUNC), 0644);
    if (fd < 0) return -1;
    *at = 0;
    if (append) {
        const off_t e = lseek(fd, 0, SEEK_END);
        if (e < 0) { close(fd); return -1; }
        *at = (long long)e;
    }
    return fd;
}

}  // namespace

extern "C" {

void pem_free_host(void* p) { free(p); }

int pem_mtx_read(const char* path, int32_t* rows, int32_t* cols, int64_t* nnz,
                 int32_t** I, int32_t** J, double** V, int* is_symmetric, char* err, size_t err_len)
{
    if (!path || !rows || !cols || !nnz || !I || !J || !V) return PEM_ERR_ARG;
    *I = *J = nullptr; *V = nullptr;
    Mapping map;
    {
        const int fd = open(path, O_RDONLY);
        if (fd < 0) { set_err(err, err_len, std::string("cannot open ") + path); return PEM_ERR_IO; }
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { close(fd); set_err(err, err_len, std::string("cannot read ") + path); return PEM_ERR_IO; }
        if (st.st_size > 0) {
            void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) { close(fd); set_err(err, err_len, std::string("cannot map ") + path); return PEM_ERR_IO; }
            map.p = (const char*)m; map.n = (size_t)st.st_size;
            madvise(m, map.n, MADV_WILLNEED);
        }
        close(fd);
    }
    // every scan below is bounded by `end` (memchr, from_chars), so the mapping needs no terminator
    const char* p = map.p;
    const char* end = map.p + map.n;
    if (!map.n) { set_err(err, err_len, "missing %%MatrixMarket banner"); return PEM_ERR_IO; }

    // banner
    const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
    if (!eol) eol = end;
    std::string banner(p, eol);
    std::transform(banner.begin(), banner.end(), banner.begin(), [](unsigned char c) { return (char)tolower(c); });
    if (banner.rfind("%%matrixmarket", 0) != 0) { set_err(err, err_len, "missing %%MatrixMarket banner"); return PEM_ERR_IO; }
    if (banner.find("coordinate") == std::string::npos) { set_err(err, err_len, "only coordinate format is supported"); return PEM_ERR_IO; }
    const bool pattern = banner.find("pattern") != std::string::npos;
    const bool complex_ = banner.find("complex") != std::string::npos;
    const bool skew = banner.find("skew-symmetric") != std::string::npos;
    const bool symm = !skew && (banner.find("symmetric") != std::string::npos || banner.find("hermitian") != std::string::npos);
    if (is_symmetric) *is_symmetric = symm ? 1 : 0;
    p = eol < end ? eol + 1 : end;
    // comments, then the size line
    long long R = 0, Cc = 0, N = 0;
    for (;;) {
        if (p >= end) { set_err(err, err_len, "missing size line"); return PEM_ERR_IO; }
        eol = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        if (!blank_or_comment(p, eol)) {
            bool ok = true;
            const char* q = parse_i64(p, eol, R, ok);
            q = parse_i64(q, eol, Cc, ok);
            q = parse_i64(q, eol, N, ok);
            if (!ok || R < 0 || Cc < 0 || N < 0 || R > 0x7fffffffLL || Cc > 0x7fffffffLL) {
                set_err(err, err_len, "bad size line");
                return PEM_ERR_IO;
            }
            p = eol < end ? eol + 1 : end;
            break;
        }
        p = eol < end ? eol + 1 : end;
    }

    // cut the body into chunks at line boundaries
    unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    size_t body = (size_t)(end - p);
    if (body < (1u << 20)) nt = 1;
    std::vector<Chunk> ch(nt);
    {
        const char* b = p;
        for (unsigned t = 0; t < nt; ++t) {
            const char* e = (t + 1 == nt) ? end : p + body * (t + 1) / nt;
            if (e < b) e = b;
            if (t + 1 != nt) {
                const char* nl = (const char*)memchr(e, '\n', (size_t)(end - e));
                e = nl ? nl + 1 : end;
            }
            ch[t].b = b; ch[t].e = e;
            b = e;
        }
    }
    // Lines per chunk, counted as newline bytes (a vectorised compare-and-add, several GB/s per thread): an upper
    // bound of the chunk's entries that is exact unless the body holds blank or comment lines, so every chunk can be
    // parsed straight into its place of the output; the rare gaps are closed afterwards.
    auto count_lines = [&](unsigned t) {
        const char* b = ch[t].b;
        const char* const e = ch[t].e;
        size_t n = 0;
        if (b < e && e[-1] != '\n') n = 1;                   // the file's last line has no terminator
        while (b < e) {
            const size_t m = std::min<size_t>((size_t)(e - b), 4096);
            unsigned k = 0;
            for (size_t x = 0; x < m; ++x) k += (b[x] == '\n');
            n += k;
            b += m;
        }
        ch[t].lines = n;
    };
    auto on_threads = [&](auto&& fn) {
        if (nt == 1) { fn(0u); return; }
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(fn, t);
        for (auto& x : th) x.join();
    };
    on_threads(count_lines);
    size_t room = 0;
    for (auto& c : ch) { c.out = room; room += c.lines; }
    if ((long long)room < N) { set_err(err, err_len, "entry count differs from the size line"); return PEM_ERR_IO; }
    const size_t cap = (symm || skew) ? 2 * room : room;
    int32_t* oi = (int32_t*)malloc(std::max<size_t>(cap, 1) * 4);
    int32_t* oj = (int32_t*)malloc(std::max<size_t>(cap, 1) * 4);
    double* ov = (double*)malloc(std::max<size_t>(cap, 1) * 8);
    if (!oi || !oj || !ov) { free(oi); free(oj); free(ov); set_err(err, err_len, "out of host memory"); return PEM_ERR_IO; }
    // One pass per chunk, no line search up front: coordinates by a digit loop, the value by from_chars (which stops
    // at the line end by itself), then on to the newline.  An entry uses up at least one line, so a chunk never
    // writes past its room.
    std::vector<size_t> found(nt, 0);
    auto parse_chunk = [&](unsigned t) {
        Chunk& c = ch[t];
        size_t o = c.out;
        const char* q = c.b;
        const char* const e = c.e;
        auto coord = [&](long long lim, int32_t& out) {    // [ws] [+] digits, 1 <= value <= lim
            q = skip_ws(q, e);
            if (q < e && *q == '+') ++q;
            unsigned long long x = 0;
            const char* d0 = q;
            while (q < e && (unsigned)(*q - '0') < 10u && q - d0 < 18) x = x * 10 + (unsigned)(*q++ - '0');
            if (q == d0 || (q < e && (unsigned)(*q - '0') < 10u) || x < 1 || x > (unsigned long long)lim) { c.ok = false; out = 0; return; }
            out = (int32_t)(x - 1);
        };
        while (q < e) {
            q = skip_ws(q, e);
            if (q < e && *q != '\n' && *q != '%') {
                double v = 1.0;
                coord(R, oi[o]);
                coord(Cc, oj[o]);
                if (!pattern && c.ok) {                      // complex: the real part comes first
                    q = skip_ws(q, e);
                    if (q < e && *q == '+') ++q;
                    auto r = std::from_chars(q, e, v);
                    if (r.ec == std::errc()) {
                        q = r.ptr;
                    } else {                                 // inf / nan spellings, exotic formats: strtod on this line only
                        const char* nl = (const char*)memchr(q, '\n', (size_t)(e - q));
                        bool ok = true;
                        q = parse_f64(q, nl ? nl : e, v, ok);
                        if (!ok) c.ok = false;
                    }
                }
                ov[o] = v;
                ++o;
            }
            const char* nl = (const char*)memchr(q, '\n', (size_t)(e - q));     // rest of the line (usually nothing)
            q = nl ? nl + 1 : e;
        }
        found[t] = o - c.out;
    };
    on_threads(parse_chunk);
    (void)complex_;
    for (auto& c : ch)
        if (!c.ok) { free(oi); free(oj); free(ov); set_err(err, err_len, "malformed or out-of-range entry"); return PEM_ERR_IO; }
    size_t total = 0;
    for (unsigned t = 0; t < nt; ++t) {                     // close the gaps blank / comment lines left (usually none)
        if (found[t] && total != ch[t].out) {
            memmove(oi + total, oi + ch[t].out, found[t] * 4);
            memmove(oj + total, oj + ch[t].out, found[t] * 4);
            memmove(ov + total, ov + ch[t].out, found[t] * 8);
        }
        total += found[t];
    }
    if ((long long)total != N) { free(oi); free(oj); free(ov); set_err(err, err_len, "entry count differs from the size line"); return PEM_ERR_IO; }
    size_t n = total;
    if (symm || skew) {
        for (size_t e = 0; e < total; ++e)
            if (oi[e] != oj[e]) { oi[n] = oj[e]; oj[n] = oi[e]; ov[n] = skew ? -ov[e] : ov[e]; ++n; }
    }
    *rows = (int32_t)R; *cols = (int32_t)Cc; *nnz = (int64_t)n;
    *I = oi; *J = oj; *V = ov;
    return PEM_OK;
}

int pem_mtx_write(const char* path, int32_t rows, int32_t cols, int64_t nnz,
                  const int32_t* I, const int32_t* J, const double* V)
{
    if (!path || nnz < 0 || (nnz && (!I || !J || !V))) return PEM_ERR_ARG;
    long long at = 0;
    const int fd = open_out(path, 0, &at);
    if (fd < 0) return PEM_ERR_IO;
    char head[128];
    const int hn = snprintf(head, sizeof head, "%%%%MatrixMarket matrix coordinate real general\n%d %d %lld\n", rows, cols, (long long)nnz);
    bool ok = write_all(fd, head, (size_t)hn, 0);
    if (ok)
        ok = write_lines(fd, hn, nnz, 96, [&](int64_t e, char* q) {
                 q = std::to_chars(q, q + 16, I[e] + 1).ptr; *q++ = ' ';
                 q = std::to_chars(q, q + 16, J[e] + 1).ptr; *q++ = ' ';
                 q = std::to_chars(q, q + 40, V[e]).ptr; *q++ = '\n';   // shortest round-trip representation
                 return q;
             }) >= 0;
    ok = (close(fd) == 0) && ok;
    return ok ? PEM_OK : PEM_ERR_IO;
}

// One number per line: the files of the reference's COO dump (spgemm.cu:1545-1560).  Integers as they are;
// doubles in fixed notation with max_digits10 = 17 decimals, the digits `std::fixed << std::setprecision(17)`
// prints there.  append != 0 continues an existing file (results dumped panel by panel).
int pem_write_lines_i32(const char* path, const int32_t* x, int64_t n, int append)
{
    if (!path || n < 0 || (n && !x)) return PEM_ERR_ARG;
    long long at = 0;
    const int fd = open_out(path, append, &at);
    if (fd < 0) return PEM_ERR_IO;
    bool ok = write_lines(fd, at, n, 16, [&](int64_t i, char* q) {
                  q = std::to_chars(q, q + 15, x[i]).ptr; *q++ = '\n';
                  return q;
              }) >= 0;
    ok = (close(fd) == 0) && ok;
    return ok ? PEM_OK : PEM_ERR_IO;
}

int pem_write_lines_f64(const char* path, const double* x, int64_t n, int append)
{
    if (!path || n < 0 || (n && !x)) return PEM_ERR_ARG;
    long long at = 0;
    const int fd = open_out(path, append, &at);
    if (fd < 0) return PEM_ERR_IO;
    bool ok = write_lines(fd, at, n, 400, [&](int64_t i, char* q) {      // 1.8e308 in fixed notation: 309 + 1 + 17 digits
                  q = pem_fmt::fixed17(x[i], q);      // printf("%.17f") digits from integer arithmetic (fixed17.h)
                  *q++ = '\n';
                  return q;
              }) >= 0;
    ok = (close(fd) == 0) && ok;
    return ok ? PEM_OK : PEM_ERR_IO;
}

}  // extern "C"
