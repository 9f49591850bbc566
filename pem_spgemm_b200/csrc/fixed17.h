// Fixed notation with 17 decimals, the digits `std::fixed << std::setprecision(17)` prints in the reference's
// COO dump (/root/reference/spgemm.cu:1529, 1556-1558), i.e. printf("%.17f"): the exact binary value rounded
// half-to-even at the 17th decimal.
//
// std::to_chars(fixed, 17) does this for every double at ~120 ns per value; a result holds 10^8..10^9 values.
// For |x| < 2^53 the digits come from integer arithmetic instead: x = m * 2^-k (m < 2^53), the integer part is
// m >> k, and the 17 decimals are round_half_even(f * 10^17 / 2^k) for the k-bit fraction f, which fits a
// 128-bit product (2^53 * 10^17 < 2^110).  Everything else (huge values, inf, nan) goes through std::to_chars.
#pragma once
#include <charconv>
#include <cstdint>
#include <cstring>
#include <limits>

namespace pem_fmt {

inline char* put_2digits(char* p, unsigned v)
{
    static const char lut[] =
        "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960616263646566676869"
        "707172737475767778798081828384858687888990919293949596979899";
    std::memcpy(p, lut + 2 * v, 2);
    return p + 2;
}

// exactly 8 digits of v < 10^8
inline char* put_8digits(char* p, uint32_t v)
{
    const uint32_t hi = v / 10000, lo = v % 10000;
    p = put_2digits(p, hi / 100);
    p = put_2digits(p, hi % 100);
    p = put_2digits(p, lo / 100);
    return put_2digits(p, lo % 100);
}

// writes x as printf("%.17f") would and returns the end; needs room for 400 characters
inline char* fixed17(double x, char* p)
{
    uint64_t bits;
    std::memcpy(&bits, &x, 8);
    const unsigned ex = (unsigned)(bits >> 52) & 0x7FFu;
    if (ex >= 1075u) {      // |x| >= 2^52 (an integer), inf or nan: the general routine
        return std::to_chars(p, p + 399, x, std::chars_format::fixed, std::numeric_limits<double>::max_digits10).ptr;
    }
    if (bits >> 63) *p++ = '-';
    const uint64_t frac = bits & ((uint64_t(1) << 52) - 1);
    const uint64_t m = ex ? (frac | (uint64_t(1) << 52)) : frac;      // x = m * 2^-k
    const unsigned k = ex ? 1075u - ex : 1074u;                        // 1 .. 1074
    uint64_t ip = 0, dec = 0;                                          // integer part, the 17 decimals as an integer
    if (k < 128) {
        const uint64_t f = k < 64 ? (m & ((uint64_t(1) << k) - 1)) : m;
        ip = k < 64 ? (m >> k) : 0;
        const unsigned __int128 prod = (unsigned __int128)f * 100000000000000000ull;          // f * 10^17 < 2^110
        const unsigned __int128 rem = prod & ((((unsigned __int128)1) << k) - 1), half = ((unsigned __int128)1) << (k - 1);
        dec = (uint64_t)(prod >> k);
        if (rem > half || (rem == half && (dec & 1))) ++dec;
        if (dec == 100000000000000000ull) { dec = 0; ++ip; }
    }                                                                  // k >= 128: |x| < 2^-75, all 17 decimals are 0
    p = std::to_chars(p, p + 20, ip).ptr;
    *p++ = '.';
    const uint64_t top = dec / 100000000ull;                           // 9 digits
    const uint32_t low = (uint32_t)(dec % 100000000ull);               // 8 digits
    *p++ = (char)('0' + top / 100000000ull);
    p = put_8digits(p, (uint32_t)(top % 100000000ull));
    return put_8digits(p, low);
}

}  // namespace pem_fmt
