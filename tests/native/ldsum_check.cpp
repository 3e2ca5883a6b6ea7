// ldsum_check.cpp -- epigrahmm_b200/csrc/ldsum.h against the compiler's own 80-bit long double
// (x86: the arithmetic R's cumsum uses).  Prints the number of mismatching partial sums.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <random>
#include <vector>
#include <algorithm>
#include "../../epigrahmm_b200/csrc/ldsum.h"

static long check(const std::vector<double> &x) {
    long double s = 0.0L;
    LdAcc a = {0, 0};
    long bad = 0;
    for (size_t i = 0; i < x.size(); i++) {
        s += x[i];
        a = ld_add(a, ld_from_double(x[i]));
        const double want = (double)s, got = ld_to_double(a);
        if (std::memcmp(&want, &got, sizeof(double)) != 0) bad++;
    }
    return bad;
}

int main() {
    static_assert(sizeof(long double) >= 10, "needs the x87 extended type");
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    long bad = 0, n = 0;
    for (int rep = 0; rep < 40; rep++) {
        std::vector<double> x(50000);
        for (double &v : x) {
            switch (rep % 8) {
                case 0: v = U(rng); break;                                         // generic
                case 1: v = 1.0 - std::exp(-30.0 * U(rng)); break;                 // 1 - posterior
                case 2: v = std::ldexp(std::floor(U(rng) * 4096.0), -12 - (int)(rng() % 50)); break;  // few bits: ties
                case 3: v = (rng() % 3 == 0) ? 0.0 : std::ldexp(1.0, -(int)(rng() % 64)); break;      // powers of two, zeros
                case 4: v = std::ldexp(U(rng), -(int)(rng() % 1000)); break;       // tiny, incl. towards subnormal sums
                case 5: v = 1.0 - std::ldexp(1.0, -(int)(rng() % 53) - 1); break;  // just below one
                case 6: v = std::ldexp((double)(rng() >> 11), -53 - (int)(rng() % 12)); break;        // full significands
                default: v = U(rng) < 0.5 ? 0.5 : std::ldexp(1.0, -11) * (double)(rng() % 4096); break;
            }
        }
        if (rep % 8 != 4) std::sort(x.begin(), x.end());   // fdrControl adds in ascending order
        bad += check(x);
        n += (long)x.size();
    }
    std::printf("%ld %ld\n", bad, n);
    return bad ? 1 : 0;
}
