// Matrix Market coordinate reader / writer (host side of the C ABI).
//
// Replaces the reference's use of the un-vendored fast_matrix_market v1.7.6 in
// read_matrix_market (/root/reference/spgemm.cu:43-110): header + triplets, `pattern` values
// become 1.0, `complex` keeps the real part (:99-107), symmetric / skew-symmetric / hermitian
// files are expanded to general form.  Unlike fast_matrix_market's default, a diagonal entry of
// a symmetric file is emitted ONCE (no extra zero-valued duplicate): duplicates are undefined
// input for the tile conversion.
//
// The body is parsed by all host threads: the file is read in one piece, cut at line boundaries
// into one chunk per thread, lines are counted, then parsed with std::from_chars straight into
// the output arrays at each chunk's offset.
#include <algorithm>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

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

}  // namespace

extern "C" {

void pem_free_host(void* p) { free(p); }

int pem_mtx_read(const char* path, int32_t* rows, int32_t* cols, int64_t* nnz,
                 int32_t** I, int32_t** J, double** V, int* is_symmetric, char* err, size_t err_len)
{
    if (!path || !rows || !cols || !nnz || !I || !J || !V) return PEM_ERR_ARG;
    *I = *J = nullptr; *V = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) { set_err(err, err_len, std::string("cannot open ") + path); return PEM_ERR_IO; }
    fseek(f, 0, SEEK_END);
    long long fsz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)fsz + 1);
    size_t got = fread(buf.data(), 1, (size_t)fsz, f);
    fclose(f);
    if ((long long)got != fsz) { set_err(err, err_len, "short read"); return PEM_ERR_IO; }
    buf[(size_t)fsz] = '\n';
    const char* p = buf.data();
    const char* end = buf.data() + fsz;

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
    auto for_each_line = [](Chunk& c, auto&& fn) {
        const char* q = c.b;
        while (q < c.e) {
            const char* nl = (const char*)memchr(q, '\n', (size_t)(c.e - q));
            if (!nl) nl = c.e;
            if (!blank_or_comment(q, nl)) fn(q, nl);
            q = nl + 1;
        }
    };
    {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t)
            th.emplace_back([&, t] { for_each_line(ch[t], [&](const char*, const char*) { ++ch[t].lines; }); });
        for (auto& x : th) x.join();
    }
    size_t total = 0;
    for (auto& c : ch) { c.out = total; total += c.lines; }
    if ((long long)total != N) { set_err(err, err_len, "entry count differs from the size line"); return PEM_ERR_IO; }
    const size_t cap = (symm || skew) ? 2 * total : total;
    int32_t* oi = (int32_t*)malloc(std::max<size_t>(cap, 1) * 4);
    int32_t* oj = (int32_t*)malloc(std::max<size_t>(cap, 1) * 4);
    double* ov = (double*)malloc(std::max<size_t>(cap, 1) * 8);
    if (!oi || !oj || !ov) { free(oi); free(oj); free(ov); set_err(err, err_len, "out of host memory"); return PEM_ERR_IO; }
    {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t)
            th.emplace_back([&, t] {
                size_t o = ch[t].out;
                for_each_line(ch[t], [&](const char* q, const char* nl) {
                    long long i = 0, j = 0;
                    double v = 1.0;
                    bool ok = true;
                    q = parse_i64(q, nl, i, ok);
                    q = parse_i64(q, nl, j, ok);
                    if (!pattern) q = parse_f64(q, nl, v, ok);  // complex: the real part comes first
                    if (!ok || i < 1 || j < 1 || i > R || j > Cc) { ch[t].ok = false; i = j = 1; }
                    oi[o] = (int32_t)(i - 1); oj[o] = (int32_t)(j - 1); ov[o] = v;
                    ++o;
                });
            });
        for (auto& x : th) x.join();
    }
    (void)complex_;
    for (auto& c : ch)
        if (!c.ok) { free(oi); free(oj); free(ov); set_err(err, err_len, "malformed or out-of-range entry"); return PEM_ERR_IO; }
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
    FILE* f = fopen(path, "wb");
    if (!f) return PEM_ERR_IO;
    fprintf(f, "%%%%MatrixMarket matrix coordinate real general\n%d %d %lld\n", rows, cols, (long long)nnz);
    std::vector<char> buf(1 << 22);
    size_t pos = 0;
    for (int64_t e = 0; e < nnz; ++e) {
        if (pos + 96 > buf.size()) { fwrite(buf.data(), 1, pos, f); pos = 0; }
        char* q = buf.data() + pos;
        q = std::to_chars(q, q + 16, I[e] + 1).ptr; *q++ = ' ';
        q = std::to_chars(q, q + 16, J[e] + 1).ptr; *q++ = ' ';
        q = std::to_chars(q, q + 40, V[e]).ptr; *q++ = '\n';   // shortest round-trip representation
        pos = (size_t)(q - buf.data());
    }
    fwrite(buf.data(), 1, pos, f);
    int rc = ferror(f) ? PEM_ERR_IO : PEM_OK;
    fclose(f);
    return rc;
}

}  // extern "C"
