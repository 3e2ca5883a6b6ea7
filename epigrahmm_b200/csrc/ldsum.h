// ldsum.h -- R's cumsum() on doubles, bit for bit, without an 80-bit floating-point unit.
//
// R accumulates in `long double` (x87 extended: 64-bit significand) and stores every partial sum rounded
// to double (src/main/cum.c: LDOUBLE sum; sum += x[i]; s[i] = (double) sum).  fdrControl
// (R/base-utils.R:21) compares those stored sums, divided by 1..n, with the FDR level, so the called
// windows depend on the two roundings.  Here the running sum of NON-NEGATIVE doubles is kept as an
// integer significand m in [2^63, 2^64) and an exponent q (value = m * 2^q); an addition aligns the
// smaller operand in 128 bits, adds, and rounds to nearest-even at 64 bits exactly as the x87 does; the
// stored value rounds the 64-bit significand to nearest-even at 53 bits.  Compiled for the host too
// (tests/test_ldsum.py checks it against the compiler's own long double).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LDSUM_FN __host__ __device__ inline
#else
#define LDSUM_FN inline
#endif

struct LdAcc {
    uint64_t m;  // 0 = the sum is zero; else the top bit is set
    int q;
};

LDSUM_FN int ld_clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return __builtin_clzll(x);
#endif
}

LDSUM_FN uint64_t ld_bits(double x) {
    union { double d; uint64_t u; } c;
    c.d = x;
    return c.u;
}
LDSUM_FN double ld_from_bits(uint64_t u) {
    union { double d; uint64_t u; } c;
    c.u = u;
    return c.d;
}

// a finite double x >= 0 as (m, q), normalised
LDSUM_FN LdAcc ld_from_double(double x) {
    const uint64_t b = ld_bits(x);
    const int e = (int)((b >> 52) & 0x7ff);
    uint64_t f = b & 0x000fffffffffffffull;
    LdAcc r;
    if (e == 0) {
        if (f == 0) { r.m = 0; r.q = 0; return r; }
        const int z = ld_clz64(f);
        r.m = f << z;
        r.q = -1074 - z;
        return r;
    }
    f |= 0x0010000000000000ull;
    r.m = f << 11;
    r.q = e - 1075 - 11;
    return r;
}

// S + x rounded to nearest-even at 64 significant bits (both >= 0)
LDSUM_FN LdAcc ld_add(LdAcc a, LdAcc b) {
    if (b.m == 0) return a;
    if (a.m == 0) return b;
    if (a.q < b.q) { const LdAcc t = a; a = b; b = t; }
    const int d = a.q - b.q;
    // both shifted right by one so that the sum fits 128 bits: value = X * 2^(a.q - 63)
    const unsigned __int128 A = (unsigned __int128)a.m << 63;
    unsigned __int128 Bv = 0;
    bool sticky = false;
    if (d < 127) {
        const unsigned __int128 b0 = (unsigned __int128)b.m << 63;
        Bv = b0 >> d;
        sticky = (Bv << d) != b0;
    } else {
        sticky = true;
    }
    const unsigned __int128 s = A + Bv;
    LdAcc r;
    uint64_t rem_hi;   // the discarded bits, left-aligned in 64 bits
    bool rem_lo;       // ... and whether anything non-zero lies below them
    if (s >> 127) {    // carried into the next binade
        r.m = (uint64_t)(s >> 64);
        rem_hi = (uint64_t)s;
        rem_lo = sticky;
        r.q = a.q + 1;
    } else {
        r.m = (uint64_t)(s >> 63);
        rem_hi = (uint64_t)s << 1;
        rem_lo = sticky;
        r.q = a.q;
    }
    const uint64_t half = 0x8000000000000000ull;
    const bool up = (rem_hi > half) || (rem_hi == half && (rem_lo || (r.m & 1ull)));
    if (up) {
        r.m += 1;
        if (r.m == 0) { r.m = half; r.q += 1; }
    }
    return r;
}

// (double) of the accumulator: nearest-even at 53 bits (sums of probabilities are far from the double range's ends)
LDSUM_FN double ld_to_double(LdAcc a) {
    if (a.m == 0) return 0.0;
    uint64_t f = a.m >> 11;
    const uint64_t rem = a.m & 0x7ffull;
    int e = a.q + 11 + 1075;  // biased exponent of f * 2^(q+11) with f in [2^52, 2^53)
    if (rem > 0x400ull || (rem == 0x400ull && (f & 1ull))) {
        f += 1;
        if (f == 0x0020000000000000ull) { f >>= 1; e += 1; }
    }
    if (e <= 0 || e >= 0x7ff) return (e <= 0) ? 0.0 : ld_from_bits(0x7ff0000000000000ull);
    return ld_from_bits(((uint64_t)e << 52) | (f & 0x000fffffffffffffull));
}
