// TEST INFRASTRUCTURE — minimal stand-in for fast_matrix_market v1.7.6 (not vendored by the
// reference and absent offline): just the three names /root/reference/spgemm.cu:60-83 uses.
// Coordinate files only; symmetric / skew / hermitian inputs are expanded to general form with
// the diagonal emitted once.  Written from the call sites and the Matrix Market format spec.
#pragma once
#include <algorithm>
#include <cctype>
#include <complex>
#include <cstdlib>
#include <istream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace fast_matrix_market {

enum class symmetry_type { general, symmetric, skew_symmetric, hermitian };
enum class field_type { real, double_, complex, integer, pattern };

struct matrix_market_header {
    symmetry_type symmetry = symmetry_type::general;
    field_type field = field_type::real;
    long long nrows = 0, ncols = 0, nnz = 0;
};

inline void read_header(std::istream& in, matrix_market_header& h)
{
    std::string line;
    if (!std::getline(in, line)) throw std::runtime_error("empty matrix market file");
    std::transform(line.begin(), line.end(), line.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    if (line.rfind("%%matrixmarket", 0) != 0) throw std::runtime_error("missing banner");
    if (line.find("coordinate") == std::string::npos) throw std::runtime_error("only coordinate files");
    if (line.find("complex") != std::string::npos) h.field = field_type::complex;
    else if (line.find("pattern") != std::string::npos) h.field = field_type::pattern;
    else if (line.find("integer") != std::string::npos) h.field = field_type::integer;
    else h.field = field_type::real;
    if (line.find("skew-symmetric") != std::string::npos) h.symmetry = symmetry_type::skew_symmetric;
    else if (line.find("symmetric") != std::string::npos) h.symmetry = symmetry_type::symmetric;
    else if (line.find("hermitian") != std::string::npos) h.symmetry = symmetry_type::hermitian;
    else h.symmetry = symmetry_type::general;
    while (std::getline(in, line)) {
        size_t p = line.find_first_not_of(" \t\r");
        if (p == std::string::npos || line[p] == '%') continue;
        std::istringstream ss(line);
        ss >> h.nrows >> h.ncols >> h.nnz;
        break;
    }
}

namespace detail {
inline void set_value(double& dst, double re, double) { dst = re; }
inline void set_value(std::complex<double>& dst, double re, double im) { dst = {re, im}; }
inline double negate(double v) { return -v; }
inline std::complex<double> negate(std::complex<double> v) { return -v; }
}  // namespace detail

template <typename IT, typename VT>
void read_matrix_market_triplet(std::istream& in, IT& nrows, IT& ncols, std::vector<IT>& rows,
                                std::vector<IT>& cols, std::vector<VT>& vals)
{
    matrix_market_header h;
    read_header(in, h);
    nrows = (IT)h.nrows;
    ncols = (IT)h.ncols;
    // slurp the body and parse it with strtol/strtod
    std::string body((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    const bool expand = h.symmetry != symmetry_type::general;
    rows.reserve((size_t)h.nnz * (expand ? 2 : 1));
    cols.reserve(rows.capacity());
    vals.reserve(rows.capacity());
    const char* p = body.c_str();
    const char* end = p + body.size();
    while (p < end) {
        while (p < end && std::isspace((unsigned char)*p)) ++p;
        if (p >= end) break;
        if (*p == '%') { while (p < end && *p != '\n') ++p; continue; }
        char* q = nullptr;
        long long i = std::strtoll(p, &q, 10); p = q;
        long long j = std::strtoll(p, &q, 10); p = q;
        double re = 1.0, im = 0.0;
        if (h.field != field_type::pattern) { re = std::strtod(p, &q); p = q; }
        if (h.field == field_type::complex) { im = std::strtod(p, &q); p = q; }
        VT v;
        detail::set_value(v, re, im);
        rows.push_back((IT)(i - 1)); cols.push_back((IT)(j - 1)); vals.push_back(v);
        if (expand && i != j) {
            rows.push_back((IT)(j - 1)); cols.push_back((IT)(i - 1));
            vals.push_back(h.symmetry == symmetry_type::skew_symmetric ? detail::negate(v) : v);
        }
    }
}

}  // namespace fast_matrix_market
