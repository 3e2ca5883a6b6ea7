// mtpoly.h -- GF(2) polynomial arithmetic for jumping ahead in R's Mersenne-Twister (MT19937)
// stream.  Host-only, header-only C++ (also compiled by the CPU test tests/test_mtpoly.py).
//
// The raw word sequence w_n (n >= 1) of MT19937 is a linear recurring sequence over GF(2) in every
// bit position, with the generator's degree-19937 characteristic polynomial phi.  If
// g_J(x) = x^J mod phi = sum_i c_i x^i, then  w[m + J] = XOR_{i : c_i = 1} w[m + i]  for all m >= 1,
// so a 624-word window J words ahead is a GF(2) correlation of 19937 + 623 consecutive words with
// the coefficient vector c.  phi is obtained once by Berlekamp-Massey from 2*19937 output bits and
// x^(J*2^k) mod phi by repeated squaring (Haramoto, Matsumoto, Nishimura, Panneton, L'Ecuyer 2008).
#pragma once
#include <stdint.h>
#include <string.h>
#include <vector>

namespace mtpoly {

constexpr int DEG = 19937;
constexpr int NW = (DEG + 63) / 64;  // 312 words hold coefficients 0..19936 (+ room to bit 19967)
constexpr int MT_N = 624, MT_M = 397;

typedef std::vector<uint64_t> Poly;  // little-endian bit order: bit i of the poly = coefficient of x^i

inline bool get(const Poly &p, int i) { return (p[i >> 6] >> (i & 63)) & 1ull; }
inline void flip(Poly &p, int i) { p[i >> 6] ^= 1ull << (i & 63); }

// one step of the MT19937 recurrence on a 624-word ring: returns the next raw word
struct RawMT {
    uint32_t mt[MT_N];
    int pos = 0;  // index of the oldest word
    uint32_t next() {
        const uint32_t a = mt[pos], b = mt[(pos + 1) % MT_N], c = mt[(pos + MT_M) % MT_N];
        const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
        const uint32_t v = c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        mt[pos] = v;
        pos = (pos + 1) % MT_N;
        return v;
    }
};

// characteristic polynomial phi (degree 19937) as NW+1 words, by Berlekamp-Massey
inline Poly min_poly() {
    RawMT g;
    uint32_t s = 19650218u;
    for (int i = 0; i < MT_N; i++) { s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i; g.mt[i] = s; }
    const int NB = 2 * DEG + 64;
    std::vector<uint8_t> seq(NB);
    for (int i = 0; i < NB; i++) seq[i] = g.next() & 1u;
    const int W = NW + 2;
    Poly C(W, 0), B(W, 0), H(W, 0), T;
    C[0] = 1;
    B[0] = 1;
    int L = 0, m = 1;
    for (int n = 0; n < NB; n++) {
        // history H: bit i = seq[n - i]
        for (int w = W - 1; w > 0; w--) H[w] = (H[w] << 1) | (H[w - 1] >> 63);
        H[0] = (H[0] << 1) | seq[n];
        uint64_t acc = 0;
        const int lw = (L >> 6) + 1;
        for (int w = 0; w < lw && w < W; w++) acc ^= C[w] & H[w];
        const int d = __builtin_parityll(acc);
        if (!d) { m++; continue; }
        if (2 * L <= n) T = C;
        // C ^= B << m
        const int ws = m >> 6, bs = m & 63;
        for (int w = W - 1; w >= ws; w--) {
            uint64_t v = B[w - ws] << bs;
            if (bs && w - ws - 1 >= 0) v |= B[w - ws - 1] >> (64 - bs);
            C[w] ^= v;
        }
        if (2 * L <= n) { L = n + 1 - L; B = T; m = 1; } else { m++; }
    }
    // connection polynomial C(x) = sum c_i x^i with sum_i c_i s[n-i] = 0; phi(x) = x^L C(1/x)
    Poly phi(NW + 1, 0);
    if (L != DEG) return Poly();  // caller checks
    for (int i = 0; i <= L; i++)
        if (get(C, i)) flip(phi, L - i);
    return phi;
}

// r = a^2 mod phi   (a: NW words, degree < DEG)
inline Poly sqr_mod(const Poly &a, const Poly &phi) {
    std::vector<uint64_t> t(2 * NW + 2, 0);
    for (int w = 0; w < NW; w++) {  // spread bits: coefficient i -> 2i
        uint64_t x = a[w];
        uint64_t lo = x & 0xffffffffull, hi = x >> 32;
        auto spread = [](uint64_t v) {
            v = (v | (v << 16)) & 0x0000ffff0000ffffull;
            v = (v | (v << 8)) & 0x00ff00ff00ff00ffull;
            v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0full;
            v = (v | (v << 2)) & 0x3333333333333333ull;
            v = (v | (v << 1)) & 0x5555555555555555ull;
            return v;
        };
        t[2 * w] = spread(lo);
        t[2 * w + 1] = spread(hi);
    }
    // reduce from the top bit down to DEG
    for (int d = 2 * DEG - 2; d >= DEG; d--) {
        if (!((t[d >> 6] >> (d & 63)) & 1ull)) continue;
        const int sh = d - DEG, ws = sh >> 6, bs = sh & 63;
        for (int w = 0; w <= NW; w++) {
            const uint64_t v = phi[w];
            t[w + ws] ^= v << bs;
            if (bs) t[w + ws + 1] ^= v >> (64 - bs);
        }
    }
    Poly r(NW, 0);
    for (int w = 0; w < NW; w++) r[w] = t[w];
    r[NW - 1] &= (DEG & 63) ? ((1ull << (DEG & 63)) - 1) : ~0ull;
    return r;
}

// jump polynomials x^(2^log2J * 2^k) mod phi for k = 0 .. levels-1
inline std::vector<Poly> jump_polys(int log2J, int levels) {
    std::vector<Poly> out;
    Poly phi = min_poly();
    if (phi.empty()) return out;
    Poly p(NW, 0);
    p[0] = 2;  // x
    for (int i = 0; i < log2J; i++) p = sqr_mod(p, phi);
    out.push_back(p);
    for (int k = 1; k < levels; k++) {
        p = sqr_mod(p, phi);
        out.push_back(p);
    }
    return out;
}

// apply a jump polynomial to a word sequence: out[p] = XOR_{i : c_i} w[p + i], p = 0..623
// (w holds DEG + 623 consecutive raw words)
inline void apply(const Poly &c, const uint32_t *w, uint32_t *out) {
    for (int p = 0; p < MT_N; p++) out[p] = 0;
    for (int i = 0; i < DEG; i++)
        if (get(c, i))
            for (int p = 0; p < MT_N; p++) out[p] ^= w[p + i];
}

}  // namespace mtpoly
