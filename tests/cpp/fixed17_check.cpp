// Checks pem_fmt::fixed17 (pem_spgemm_b200/csrc/fixed17.h: the digits of the reference's COO dump,
// /root/reference/spgemm.cu:1529, 1556-1558) against std::to_chars(fixed, 17) = printf("%.17f"):
// random bit patterns, every binade around the fast path's limits, ties at the 17th decimal, carries.
// usage: fixed17_check [millions of random probes]; prints "checked N, mismatches M", exit status 1 on a mismatch.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "fixed17.h"

static long long bad = 0, total = 0;

static void check(double x)
{
    char a[512], b[512];
    char* ea = pem_fmt::fixed17(x, a);
    char* eb = std::to_chars(b, b + 399, x, std::chars_format::fixed, 17).ptr;
    ++total;
    if (ea - a != eb - b || std::memcmp(a, b, (size_t)(ea - a))) {
        if (bad++ < 10) {
            *ea = 0; *eb = 0;
            std::printf("MISMATCH %a: %s vs %s\n", x, a, b);
        }
    }
}

int main(int argc, char** argv)
{
    const long long millions = argc > 1 ? std::atoll(argv[1]) : 2;
    std::mt19937_64 g(12345);
    for (long long i = 0; i < millions * 1000000; ++i) {            // all bit patterns (nan / inf included)
        uint64_t u = g();
        double x;
        std::memcpy(&x, &u, 8);
        check(x);
    }
    std::uniform_real_distribution<double> U(-1, 1);
    for (long long i = 0; i < millions * 500000; ++i) check(U(g));
    for (int e = -140; e <= 70; ++e)                                  // every binade from below 2^-128 to above 2^52
        for (int i = 0; i < 2000; ++i) check(std::ldexp(U(g), e));
    for (long long j = 0; j < 300000; ++j) {                          // odd multiples of 2^-18 are exact ties at the 17th decimal
        const double x = std::ldexp((double)j, -18);
        check(x); check(-x); check(std::nextafter(x, 1e9)); check(std::nextafter(x, -1e9));
        check(123456.0 + std::ldexp((double)(2 * j + 1), -18));
    }
    for (int k = 1; k <= 30; ++k)
        for (long long j = 1; j < 20000; j += 2) check(std::ldexp((double)j, -k));
    for (double b : {1.0, 10.0, 100.0, 4503599627370496.0, 9007199254740992.0, 0.1, 1e-5, 1e-17, 5e-18, 1e-18}) {   // carries, the 2^52 limit
        double x = b;
        for (int i = 0; i < 500; ++i) { check(x); check(-x); x = std::nextafter(x, 0.0); }
        x = b;
        for (int i = 0; i < 500; ++i) { check(x); x = std::nextafter(x, 1e300); }
    }
    for (double x : {0.0, -0.0, 5e-324, -5e-324, 2.2250738585072014e-308, 1.7976931348623157e308,
                     (double)INFINITY, -(double)INFINITY})
        check(x);
    std::printf("checked %lld, mismatches %lld\n", total, bad);
    return bad != 0;
}
