/*
 * ehmm_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Single-threaded CPU restatement of the epigraHMM EM hot path, written from
 * the reference's behaviour (file:line citations are relative to
 * /root/reference).  It is the checker for the CUDA library in
 * epigrahmm_b200/csrc; nothing in the product path may call it.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * load this library.
 *
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction, so every + and * is
 * rounded exactly as in the reference's scalar C++ code).
 *
 * PARITY STATUS.  The reference cannot be built in this image (every
 * src/ file needs Rcpp, RcppArmadillo and the HDF5 C++ API; R itself is
 * absent), and its test-suite holds no numeric golden vectors for this path
 * (SURVEY.md section 4).  The native functions are restated operation for
 * operation from src/ (pinned by reading, cross-checked by a numpy twin in
 * tests/); pieces that live in third-party code not under /root/reference
 * are restated from their published algorithms and are "parity unpinned":
 *   [EXT-R]    R nmath (dnbinom_mu, dbinom_raw, stirlerr, bd0, rbinom,
 *              Mersenne-Twister unif_rand, set.seed scrambling) -- R 4.3.x
 *              algorithms; the reference's DESCRIPTION pins no R version.
 *   [EXT-R]    digamma (R uses AMOS dpsifn; here: recurrence + asymptotic
 *              series, checked against scipy/mpmath to 1e-14 in tests).
 *   [EXT-ARMA] Armadillo accu() two-accumulator order, zero-filled mat(),
 *              index_max first-maximum -- Armadillo 10.5+.  The order restated
 *              here is Armadillo's single-threaded one.  src/Makevars:15-16
 *              passes the OpenMP flags, and with ARMA_USE_OPENMP
 *              accu(exp(...)) over a large matrix (src/maxStepProb.cpp:60,63)
 *              takes Armadillo's multi-threaded path, whose summation order
 *              depends on the thread count: "bit-faithful" for maxStepProb and
 *              computeQFunction can therefore only mean "to the rounding of a
 *              differently ordered sum" (~1e-16 * sqrt(M) relative), which is
 *              what the 1e-9 gates allow for.
 *   [EXT-R]    cumsum() (fdrControl): long double accumulation, doubles stored.
 * dnbinom is additionally pinned against the closed forms used by the
 * reference's only emission KAT (tests/testthat/test-base.R:6-42) and
 * against scipy.stats.nbinom in tests/.
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ */
/* [EXT-R] nmath: stirlerr, bd0, dbinom_raw, dnbinom_mu                      */
/* ------------------------------------------------------------------------ */

#define M_LN_SQRT_2PI_ 0.918938533204672741780329736406 /* log(sqrt(2*pi)) */
#define M_LN_2PI_ 1.837877066409345483560659472811      /* log(2*pi) */

/* stirlerr(n) = log(n!) - log( sqrt(2*pi*n)*(n/e)^n ); table for n = 0, 0.5, ... 15
 * (values recomputed with mpmath at 40 digits). */
static const double sferr_halves[31] = {
    0.0,
    0.1534264097200273452913839393, 0.08106146679532725821967026359,
    0.05481412105191765389613870235, 0.04134069595540929409382208141,
    0.03316287351993628748511050974, 0.02767792568499833914878929275,
    0.02374616365629749597133027909, 0.02079067210376509311152277177,
    0.01848845053267318523077935748, 0.01664469118982119216319486537,
    0.01513497322191737887351383688, 0.01387612882307074799874572702,
    0.01281046524292022692425065528, 0.01189670994589177009505572412,
    0.0111045597582069173266307552, 0.01041126526197209649747856713,
    0.009799416126158803298390373402, 0.009255462182712732917728636633,
    0.008768700134139385462955047269, 0.00833056343336287125646931866,
    0.007934114564314020547249562491, 0.007573675487951840794972024212,
    0.007244554301320383179546196602, 0.006942840107209529865664152663,
    0.006665247032707682442356180895, 0.006408994188004207068439631083,
    0.006171712263039457647534604798, 0.005951370112758847735624416046,
    0.005746216513010115682026102477, 0.00555473355196280137103868996};

static double stirlerr(double n) {
    const double S0 = 0.083333333333333333333;        /* 1/12 */
    const double S1 = 0.00277777777777777777778;      /* 1/360 */
    const double S2 = 0.00079365079365079365079365;   /* 1/1260 */
    const double S3 = 0.000595238095238095238095238;  /* 1/1680 */
    const double S4 = 0.0008417508417508417508417508; /* 1/1188 */
    double nn;
    if (n <= 15.0) {
        nn = n + n;
        if (nn == (int)nn) return sferr_halves[(int)nn];
        return lgamma(n + 1.) - (n + 0.5) * log(n) + n - M_LN_SQRT_2PI_;
    }
    nn = n * n;
    if (n > 500) return (S0 - S1 / nn) / n;
    if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / n;
    if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

/* bd0(x, np) = x log(x/np) + np - x, evaluated without cancellation */
static double bd0(double x, double np) {
    if (!isfinite(x) || !isfinite(np) || np == 0.0) return NAN;
    if (fabs(x - np) < 0.1 * (x + np)) {
        double v = (x - np) / (x + np);
        double s = (x - np) * v;
        if (fabs(s) < DBL_MIN) return s;
        double ej = 2 * x * v;
        v = v * v;
        for (int j = 1; j < 1000; j++) {
            ej *= v;
            double s1 = s + ej / ((j << 1) + 1);
            if (s1 == s) return s1;
            s = s1;
        }
    }
    return x * log(x / np) + np - x;
}

static double dbinom_raw(double x, double n, double p, double q, int give_log) {
    double lf, lc;
    if (p == 0) return (x == 0) ? (give_log ? 0. : 1.) : (give_log ? -INFINITY : 0.);
    if (q == 0) return (x == n) ? (give_log ? 0. : 1.) : (give_log ? -INFINITY : 0.);
    if (x == 0) {
        if (n == 0) return give_log ? 0. : 1.;
        lc = (p < 0.1) ? -bd0(n, n * q) - n * p : n * log(q);
        return give_log ? lc : exp(lc);
    }
    if (x == n) {
        lc = (q < 0.1) ? -bd0(n, n * p) - n * q : n * log(p);
        return give_log ? lc : exp(lc);
    }
    if (x < 0 || x > n) return give_log ? -INFINITY : 0.;
    lc = stirlerr(n) - stirlerr(x) - stirlerr(n - x) - bd0(x, n * p) - bd0(n - x, n * q);
    lf = M_LN_2PI_ + log(x) + log1p(-x / n);
    return give_log ? (lc - 0.5 * lf) : exp(lc - 0.5 * lf);
}

/* stats::dnbinom(x, size=, mu=, log=) for integer x >= 0, finite size > 0.
 * Call sites: R/base-loglikelihood.R:16,53,67,115,121; R/base-optim.R:111,190,197,204;
 * R/base-EM.R:27. */
EO_API double eo_dnbinom_mu(double x, double size, double mu, int give_log) {
    if (isnan(x) || isnan(size) || isnan(mu)) return x + size + mu;
    if (mu < 0 || size < 0) return NAN;
    if (x < 0 || !isfinite(x)) return give_log ? -INFINITY : 0.;
    if (x == 0 && size == 0) return give_log ? 0. : 1.;
    x = nearbyint(x);
    if (x == 0) {
        double v = size * (size < mu ? log(size / (size + mu)) : log1p(-mu / (size + mu)));
        return give_log ? v : exp(v);
    }
    if (x < 1e-10 * size) {
        double p = (size < mu ? log(size / (1 + size / mu)) : log(mu / (1 + mu / size)));
        double v = x * p - mu - lgamma(x + 1) + log1p(x * (x - 1) / (2 * size));
        return give_log ? v : exp(v);
    } else {
        double p = size / (size + x);
        double ans = dbinom_raw(size, x + size, size / (size + mu), mu / (size + mu), give_log);
        return give_log ? log(p) + ans : p * ans;
    }
}

/* [EXT-R] digamma for x > 0 (call sites R/base-optim.R:127-128,227,239,251) */
EO_API double eo_digamma(double x) {
    double r = 0.0;
    while (x < 12.0) { r -= 1.0 / x; x += 1.0; }
    double xi = 1.0 / x, x2 = xi * xi;
    /* B2k/(2k): 1/12, -1/120, 1/252, -1/240, 1/132, -691/32760, 1/12 */
    double s = x2 * (1.0 / 12 - x2 * (1.0 / 120 - x2 * (1.0 / 252 - x2 * (1.0 / 240 -
               x2 * (1.0 / 132 - x2 * (691.0 / 32760 - x2 * (1.0 / 12)))))));
    return r + (log(x) - 0.5 * xi - s);
}

/* ------------------------------------------------------------------------ */
/* [EXT-R] RNG: Mersenne-Twister + set.seed scrambling + rbinom(1, p)        */
/* ------------------------------------------------------------------------ */

typedef struct {
    uint32_t mti;      /* R: dummy[0] */
    uint32_t mt[624];  /* R: dummy[1..624] */
} eo_rng;

EO_API void eo_set_seed(eo_rng *g, uint32_t seed) {
    for (int j = 0; j < 50; j++) seed = 69069u * seed + 1u;
    for (int j = 0; j < 625; j++) {
        seed = 69069u * seed + 1u;
        if (j == 0) g->mti = seed; else g->mt[j - 1] = seed;
    }
    g->mti = 624; /* FixupSeeds: dummy[0] = N */
}

static uint32_t mt_next(eo_rng *g) {
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
    uint32_t y;
    if (g->mti >= 624) {
        uint32_t *mt = g->mt;
        int kk;
        for (kk = 0; kk < 624 - 397; kk++) {
            y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
            mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
        }
        for (; kk < 623; kk++) {
            y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
            mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
        }
        y = (mt[623] & UPPER) | (mt[0] & LOWER);
        mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
        g->mti = 0;
    }
    y = g->mt[g->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

EO_API double eo_unif_rand(eo_rng *g) {
    const double i2_32m1 = 2.328306437080797e-10; /* 1/(2^32 - 1) */
    double x = (double)mt_next(g) * 2.3283064365386963e-10;
    if (x <= 0.0) return 0.5 * i2_32m1;
    if ((1.0 - x) <= 0.0) return 1.0 - 0.5 * i2_32m1;
    return x;
}

EO_API void eo_unif_fill(eo_rng *g, double *u, int64_t n) {
    for (int64_t i = 0; i < n; i++) u[i] = eo_unif_rand(g);
}

/* Uniform source shared by reweight: either a live generator or a
 * pre-drawn buffer (so that the GPU and the oracle can consume one stream). */
typedef struct {
    eo_rng *rng;
    const double *buf;
    int64_t pos, cap;
} eo_usrc;

static double usrc_next(eo_usrc *s) {
    if (s->buf) return (s->pos < s->cap) ? s->buf[s->pos++] : (s->pos++, NAN);
    s->pos++;
    return eo_unif_rand(s->rng);
}

/* R::rbinom(1, pp): inversion branch (n*min(p,1-p) < 30), src/rbinomVectorized.cpp:17 */
static double rbinom1(eo_usrc *s, double pp) {
    if (!isfinite(pp) || pp < 0. || pp > 1.) return NAN;
    if (pp == 0.) return 0.;
    if (pp == 1.) return 1.;
    const int n = 1;
    double p = fmin(pp, 1. - pp), q = 1. - p;
    double r = p / q, g = r * (n + 1), qn = q; /* R_pow_di(q, 1) */
    int ix;
    for (;;) {
        ix = 0;
        double f = qn, u = usrc_next(s);
        if (isnan(u)) return NAN;
        for (;;) {
            if (u < f) goto finis;
            if (ix > 110) break;
            u -= f;
            ix++;
            f *= (g / ix - r);
        }
    }
finis:
    if (pp > 0.5) ix = n - ix;
    return (double)ix;
}

/* reweight(x,p): src/reweight.cpp:16-22 (+ rbinomVectorized.cpp:15-19).
 * Returns the number of uniforms consumed. */
static int64_t reweight_src(double *x, int64_t n, double p, eo_usrc *s) {
    int64_t start = s->pos;
    for (int64_t i = 0; i < n; i++)
        if (x[i] < p) x[i] = rbinom1(s, x[i] / p) * p;
    return s->pos - start;
}

EO_API int64_t eo_reweight(double *x, int64_t n, double p, eo_rng *g) {
    eo_usrc s = {g, NULL, 0, 0};
    return reweight_src(x, n, p, &s);
}

EO_API int64_t eo_reweight_buf(double *x, int64_t n, double p, const double *u, int64_t ucap) {
    eo_usrc s = {NULL, u, 0, ucap};
    return reweight_src(x, n, p, &s);
}

/* ------------------------------------------------------------------------ */
/* [EXT-ARMA] accu(): two running sums over a linear range                   */
/* ------------------------------------------------------------------------ */
#define ACCU2(n, EXPR, out)                                   \
    do {                                                      \
        double v1_ = 0, v2_ = 0;                              \
        int64_t i_, j_;                                       \
        for (i_ = 0, j_ = 1; j_ < (n); i_ += 2, j_ += 2) {    \
            { int64_t I = i_; v1_ += (EXPR); }                \
            { int64_t I = j_; v2_ += (EXPR); }                \
        }                                                     \
        if (i_ < (n)) { int64_t I = i_; v1_ += (EXPR); }      \
        (out) = v1_ + v2_;                                    \
    } while (0)

static double lse_row(const double *a, int K) {
    double mx = a[0];
    for (int i = 1; i < K; i++) if (a[i] > mx) mx = a[i];
    double s;
    ACCU2(K, exp(a[I] - mx), s);
    return mx + log(s);
}

/* ------------------------------------------------------------------------ */
/* expStep: src/expStep.cpp:45-123.  All matrices column-major M x K.        */
/* gamma is column-major K x K (R matrix), gamma[i + k*K] = gamma(i,k).      */
/* ------------------------------------------------------------------------ */
EO_API void eo_expstep(int64_t M, int K, const double *pi, const double *gamma, const double *logf,
                       double *logFP, double *logBP, double *logProb1, double *logProb2) {
    double a[16], b[16], lg[256];
    for (int i = 0; i < K * K; i++) lg[i] = log(gamma[i]);
    for (int k = 0; k < K; k++) {
        logFP[0 + k * M] = log(pi[k]) + logf[0 + k * M]; /* l.67 */
        logBP[(M - 1) + k * M] = 0.0;                    /* l.68 */
    }
    for (int64_t j = 1; j < M; j++) { /* l.69-83 */
        for (int k = 0; k < K; k++) {
            for (int i = 0; i < K; i++) a[i] = (logFP[(j - 1) + i * M] + lg[i + k * K]) + logf[j + k * M];
            logFP[j + k * M] = lse_row(a, K);
            for (int i = 0; i < K; i++) b[i] = (logBP[(M - j) + i * M] + lg[k + i * K]) + logf[(M - j) + i * M];
            logBP[(M - j - 1) + k * M] = lse_row(b, K);
        }
    }
    double d[16];
    for (int k = 0; k < K; k++) d[k] = logFP[(M - 1) + k * M];
    int KK = K * K;
    for (int64_t j = 0; j < M; j++) { /* l.102-109 */
        double ll = lse_row(d, K);    /* recomputed per use in the reference; same value */
        for (int k = 0; k < K; k++) logProb1[j + k * M] = (logFP[j + k * M] + logBP[j + k * M]) - ll;
        for (int c = 0; c < KK; c++) {
            int i = c / K, l = c % K; /* K_each, K_rep */
            if (j >= 1)
                logProb2[j + c * M] = (((logFP[(j - 1) + i * M] + lg[i + l * K]) + logf[j + l * M]) + logBP[j + l * M]) - ll;
            else
                logProb2[j + c * M] = 0.0; /* zero-filled mat(M,K*K), never assigned for j == 0 */
        }
    }
    for (int64_t j = 0; j < M; j++) { /* l.112-113: sum(exp(.),1) adds columns left to right */
        double s = exp(logProb1[j]);
        for (int k = 1; k < K; k++) s += exp(logProb1[j + k * M]);
        s = log(s);
        for (int k = 0; k < K; k++) logProb1[j + k * M] -= s;
        s = exp(logProb2[j]);
        for (int c = 1; c < KK; c++) s += exp(logProb2[j + c * M]);
        s = log(s);
        for (int c = 0; c < KK; c++) logProb2[j + c * M] -= s;
    }
}

/* NOT a restatement of the reference: the yardstick of SURVEY.md H4.  The same forward-backward with every
 * vector normalised at every bin and long double (x87, 64-bit significand) arithmetic, so that its posteriors
 * carry ~1e-19 relative error instead of the ulp(|logLik|) ~ 1e-8 that the reference's unscaled log-space
 * recursion accumulates at 15 M bins.  Used only to say which of two differing results is the
 * accurate one.  P1: M x K posteriors; returns the log-likelihood. */
EO_API double eo_expstep_truth(int64_t M, int K, const double *pi, const double *gamma, const double *logf, double *P1) {
    long double *A = (long double *)malloc(sizeof(long double) * (size_t)M * K);
    long double ll = 0.0L, a[16], b[16], e[16], t[16];
    if (!A) return NAN;
    for (int64_t j = 0; j < M; j++) {
        long double mx = logf[j];
        for (int k = 1; k < K; k++) if (logf[j + k * M] > mx) mx = logf[j + k * M];
        for (int k = 0; k < K; k++) e[k] = expl((long double)logf[j + k * M] - mx);
        if (j == 0) {
            for (int k = 0; k < K; k++) a[k] = (long double)pi[k] * e[k];
        } else {
            for (int k = 0; k < K; k++) {
                long double s = 0.0L;
                for (int i = 0; i < K; i++) s += A[(j - 1) * K + i] * (long double)gamma[i + k * K];
                a[k] = s * e[k];
            }
        }
        long double c = 0.0L;
        for (int k = 0; k < K; k++) c += a[k];
        for (int k = 0; k < K; k++) A[j * K + k] = a[k] / c;
        ll += logl(c) + mx;
    }
    for (int k = 0; k < K; k++) b[k] = 1.0L;
    for (int64_t j = M - 1; j >= 0; j--) {
        long double z = 0.0L;
        for (int k = 0; k < K; k++) z += A[j * K + k] * b[k];
        for (int k = 0; k < K; k++) P1[j + k * M] = (double)(A[j * K + k] * b[k] / z);
        if (j == 0) break;
        long double mx = logf[j];
        for (int k = 1; k < K; k++) if (logf[j + k * M] > mx) mx = logf[j + k * M];
        for (int k = 0; k < K; k++) e[k] = expl((long double)logf[j + k * M] - mx) * b[k];
        long double c = 0.0L;
        for (int i = 0; i < K; i++) {
            long double s = 0.0L;
            for (int k = 0; k < K; k++) s += (long double)gamma[i + k * K] * e[k];
            t[i] = s;
            c += s;
        }
        for (int i = 0; i < K; i++) b[i] = t[i] / c;
    }
    free(A);
    return (double)ll;
}

/* maxStepProb: src/maxStepProb.cpp:37-81.  gamma_out column-major K x K. */
EO_API void eo_maxstepprob(int64_t M, int K, const double *logProb1, const double *logProb2,
                           double *pi_out, double *gamma_out) {
    int64_t R = M - 1;
    for (int k1 = 0; k1 < K; k1++) {
        /* denom = accu(exp(submat(1, idx1, M-1, idx1+K-1))): column loop, two accumulators shared */
        double v1 = 0, v2 = 0;
        for (int c = 0; c < K; c++) {
            const double *col = logProb2 + (int64_t)(k1 * K + c) * M + 1;
            int64_t i, j;
            for (i = 0, j = 1; j < R; i += 2, j += 2) { v1 += exp(col[i]); v2 += exp(col[j]); }
            if (i < R) v1 += exp(col[i]);
        }
        double denom = v1 + v2;
        for (int k2 = 0; k2 < K; k2++) {
            const double *col = logProb2 + (int64_t)(k1 * K + k2) * M + 1;
            double num;
            ACCU2(R, exp(col[I]), num);
            gamma_out[k1 + k2 * K] = num / denom;
        }
    }
    double s;
    for (int k = 0; k < K; k++) {
        pi_out[k] = exp(logProb1[0 + k * M]);
        if (pi_out[k] < 0.000001) pi_out[k] = 0.000001;
    }
    ACCU2(K, pi_out[I], s);
    for (int k = 0; k < K; k++) pi_out[k] /= s;
    for (int i = 0; i < K * K; i++) if (gamma_out[i] < 0.000001) gamma_out[i] = 0.000001;
    for (int k1 = 0; k1 < K; k1++) {
        ACCU2(K, fabs(gamma_out[k1 + I * K]), s);
        for (int k2 = 0; k2 < K; k2++) gamma_out[k1 + k2 * K] /= s;
    }
}

/* computeQFunction: src/computeQFunction.cpp:12-33 */
EO_API double eo_qfunction(int64_t M, int K, const double *logf, const double *logProb1,
                           const double *logProb2, const double *pi, const double *gamma) {
    double q0 = 0;
    for (int k = 0; k < K; k++) q0 += exp(logProb1[0 + k * M]) * log(pi[k]);
    double q1;
    int64_t n1 = M * K;
    ACCU2(n1, exp(logProb1[I]) * logf[I], q1);
    /* Prob2.rows(1,M-1) * diagmat(log(vectorise(gamma,1))) -> (M-1) x K^2 matrix, then accu */
    int KK = K * K;
    int64_t R = M - 1, n2 = R * KK;
    double lgv[256];
    for (int i = 0; i < K; i++) for (int l = 0; l < K; l++) lgv[i * K + l] = log(gamma[i + l * K]);
    double q2;
    ACCU2(n2, exp(logProb2[(I % R) + 1 + (I / R) * M]) * lgv[I / R], q2);
    return q0 + q1 + q2;
}

/* computeBIC: src/computeBIC.cpp:12-29 */
EO_API double eo_bic(int64_t M, int K, const double *logFP, int numPar, int numSamples) {
    double C = logFP[M - 1];
    for (int k = 1; k < K; k++) if (logFP[(M - 1) + k * M] > C) C = logFP[(M - 1) + k * M];
    double s;
    ACCU2(K, exp(logFP[(M - 1) + I * M] - C), s);
    /* numSamples*M is an int product in the reference (int M = n_rows) */
    return -2 * (C + log(s)) + numPar * log((double)((int64_t)numSamples * M));
}

/* innerMaxStepProb: src/innerMaxStepProb.cpp:15-38 (logProb1 has K = 3 columns) */
EO_API void eo_inner_maxstepprob(int64_t M, int B, const double *eta, const double *logProb1, double *delta) {
    const double *c1 = logProb1 + M;
    double den;
    ACCU2(M, exp(c1[I]), den);
    for (int b = 0; b < B; b++) {
        const double *e = eta + (int64_t)b * M;
        double num;
        ACCU2(M, exp(c1[I]) * e[I], num);
        delta[b] = num / den;
    }
}

/* aggregate(x,f): src/aggregate.cpp:20-45.  Dense form: bucket[g] = sum of x[i] with f[i]==g
 * (sequential order), g in [0,G].  The reference's output table is the non-zero
 * buckets in ascending g; eo_aggregate_compact extracts it. */
EO_API void eo_aggregate_dense(const double *x, int64_t n, const int32_t *f, int64_t G, double *bucket) {
    memset(bucket, 0, sizeof(double) * (size_t)(G + 1));
    for (int64_t i = 0; i < n; i++) bucket[f[i]] += x[i];
}

EO_API int64_t eo_aggregate_compact(const double *bucket, int64_t G, double *sums, double *ids) {
    int64_t s = 0;
    for (int64_t g = 0; g <= G; g++)
        if (bucket[g] != 0) { if (sums) { sums[s] = bucket[g]; ids[s] = (double)g; } s++; }
    return s;
}

/* consensusRejectionControlled: src/consensusRejectionControlled.cpp:14-31.
 * f has M*N entries but only f[0..M-1] is read (aggregate loops x.size() = M).
 * buckets: K x (G+1) row-major by state.  ubuf != NULL -> consume that buffer. */
EO_API int64_t eo_consensus_rc(int64_t M, int K, const double *logProb1, const int32_t *f, int64_t G,
                               double p, eo_rng *g, const double *ubuf, int64_t ucap, double *buckets) {
    double *w = (double *)malloc(sizeof(double) * (size_t)M);
    eo_usrc s = {g, ubuf, 0, ucap};
    for (int k = 0; k < K; k++) {
        for (int64_t j = 0; j < M; j++) w[j] = exp(logProb1[j + k * M]);
        reweight_src(w, M, p, &s);
        eo_aggregate_dense(w, M, f, G, buckets + (int64_t)k * (G + 1));
    }
    free(w);
    return s.pos;
}

/* differentialRejectionControlled: src/differentialRejectionControlled.cpp:14-48.
 * Order: background, components 0..B-1, enrichment; each weight vector is
 * thinned once and then replicated N times (repmat) before aggregation over
 * the M*N group ids.  buckets: (B+2) x (G+1). */
EO_API int64_t eo_differential_rc(int64_t M, int B, int N, const double *logProb1, const double *eta,
                                  const int32_t *f, int64_t G, double p, eo_rng *g, const double *ubuf,
                                  int64_t ucap, double *buckets) {
    double *w = (double *)malloc(sizeof(double) * (size_t)M);
    eo_usrc s = {g, ubuf, 0, ucap};
    for (int c = 0; c < B + 2; c++) {
        if (c == 0) for (int64_t j = 0; j < M; j++) w[j] = exp(logProb1[j]);
        else if (c == B + 1) for (int64_t j = 0; j < M; j++) w[j] = exp(logProb1[j + 2 * M]);
        else for (int64_t j = 0; j < M; j++) w[j] = exp(logProb1[j + M]) * eta[j + (int64_t)(c - 1) * M];
        reweight_src(w, M, p, &s);
        double *bk = buckets + (int64_t)c * (G + 1);
        memset(bk, 0, sizeof(double) * (size_t)(G + 1));
        for (int i = 0; i < N; i++)
            for (int64_t j = 0; j < M; j++) bk[f[j + (int64_t)i * M]] += w[j];
    }
    free(w);
    return s.pos;
}

/* computeViterbiSequence: src/computeViterbiSequence.cpp:12-54.  NOTE the
 * reference feeds logFP (forward log-probabilities) as the per-bin score. */
EO_API void eo_viterbi(int64_t M, int K, const double *logFP, const double *pi, const double *gamma, double *S) {
    double *V = (double *)malloc(sizeof(double) * (size_t)(M * K));
    double lg[256], mc[16];
    for (int i = 0; i < K * K; i++) lg[i] = log(gamma[i]);
    for (int k = 0; k < K; k++) V[0 + k * M] = log(pi[k]) + logFP[0 + k * M];
    for (int64_t j = 0; j < M - 1; j++) {
        for (int k = 0; k < K; k++) {
            double mx = V[j] + lg[0 + k * K];
            for (int i = 1; i < K; i++) { double c = V[j + i * M] + lg[i + k * K]; if (c > mx) mx = c; }
            mc[k] = mx;
        }
        for (int k = 0; k < K; k++) V[(j + 1) + k * M] = mc[k] + logFP[(j + 1) + k * M];
    }
    int best = 0;
    for (int k = 1; k < K; k++) if (V[(M - 1) + k * M] > V[(M - 1) + best * M]) best = k;
    S[M - 1] = best;
    for (int64_t j = M - 2; j >= 0; j--) {
        int s1 = (int)S[j + 1];
        int bi = 0;
        double bv = V[j] + lg[0 + s1 * K];
        for (int i = 1; i < K; i++) { double c = V[j + i * M] + lg[i + s1 * K]; if (c > bv) { bv = c; bi = i; } }
        S[j] = bi;
    }
    free(V);
}

/* ------------------------------------------------------------------------ */
/* Emission log-likelihoods (R/base-loglikelihood.R, R/base-EM.R)            */
/* counts y: int32 M x N column-major (the reference holds as.numeric()).    */
/* rowSums() accumulates in long double (R's LDOUBLE).                       */
/* ------------------------------------------------------------------------ */

/* loglikConsensusHMM, NB branch: R/base-loglikelihood.R:104-122.
 * P = 2 (beta0,size) or 3 (beta0,beta_ctrl,size); ctrl = log1p(controls) or NULL.
 *
 * QUIRK (reproduced, not fixed): the control term is written
 *     ifelse(useControl, theta[2]*controls, 0)         (l.113, 119; ZINB l.126-129, 137)
 * and R's ifelse() returns a value shaped like its TEST, which is the length-1 logical
 * useControl -- so only the FIRST element of theta[2]*controls survives and is recycled over all
 * M*N rows: every cell gets theta[2]*controls[1] (bin 1 of sample 1), not its own control.  The
 * M-step (glmNB's x matrix, R/base-EM.R:131-136) does use the per-cell controls, so the reference
 * is inconsistent with itself here and a faithful restatement has to be, too. */
EO_API void eo_loglik_consensus(int64_t M, int N, const int32_t *y, const double *off, const double *ctrl,
                                const double *th1, const double *th2, int P, double minZero, double *LL) {
    const double *th[2] = {th1, th2};
    for (int k = 0; k < 2; k++) {
        double size = th[k][P - 1];
        for (int64_t j = 0; j < M; j++) {
            long double acc = 0;
            for (int i = 0; i < N; i++) {
                int64_t c = j + (int64_t)i * M;
                double mu = exp(th[k][0] + (ctrl ? th[k][1] * ctrl[0] : 0) + off[c]);  /* ctrl[0]: see QUIRK */
                acc += log(eo_dnbinom_mu((double)y[c], size, mu, 0) + minZero);
            }
            LL[j + k * M] = (double)acc;
        }
    }
}

/* loglikConsensusHMM, ZINB branch: R/base-loglikelihood.R:123-141.  The background state is
 * zero-inflated: th1 = (lambda0[, lambda_ctrl], beta0[, beta_ctrl], size), Q = 3 or 5 entries;
 * the enrichment state is the NB of the other branch, th2 = (beta0[, beta_ctrl], size).
 * Note that the offsets enter the zero-inflation predictor as well (l.125). */
EO_API void eo_loglik_consensus_zinb(int64_t M, int N, const int32_t *y, const double *off, const double *ctrl,
                                     const double *th1, const double *th2, int Q, double minZero, double *LL) {
    const int useC = (Q == 5);
    for (int64_t j = 0; j < M; j++) {
        long double acc = 0;
        for (int i = 0; i < N; i++) {
            int64_t c = j + (int64_t)i * M;
            double xbeta = (useC ? th1[0] + th1[1] * ctrl[0] : th1[0]) + off[c];  /* ctrl[0]: the ifelse quirk above */
            double mu = exp((useC ? th1[2] + th1[3] * ctrl[0] : th1[1]) + off[c]);
            double size = useC ? th1[4] : th1[2];
            double ll00 = log(eo_dnbinom_mu(0.0, size, mu, 0) + minZero);
            double ll01 = log(eo_dnbinom_mu((double)y[c], size, mu, 0) + minZero);
            double zip = 1 / (1 + exp(-xbeta));
            double yvec0 = (y[c] == 0) ? 1.0 : 0.0;
            acc += yvec0 * log(zip + exp(log(1 - zip) + ll00)) + (1 - yvec0) * (log(1 - zip) + ll01);
        }
        LL[j] = (double)acc;
    }
    int P = useC ? 3 : 2;
    double size2 = th2[P - 1];
    for (int64_t j = 0; j < M; j++) {
        long double acc = 0;
        for (int i = 0; i < N; i++) {
            int64_t c = j + (int64_t)i * M;
            double mu = exp((useC ? th2[0] + th2[1] * ctrl[0] : th2[0]) + off[c]);
            acc += log(eo_dnbinom_mu((double)y[c], size2, mu, 0) + minZero);
        }
        LL[j + M] = (double)acc;
    }
}

/* emInitialization emission: R/base-EM.R:24-30 (N = 1, dnbinom(log=TRUE), psi_k=(beta,size)) */
EO_API void eo_loglik_init(int64_t M, const int32_t *y, const double *off, const double *psi1,
                           const double *psi2, double *LL) {
    const double *ps[2] = {psi1, psi2};
    for (int k = 0; k < 2; k++)
        for (int64_t j = 0; j < M; j++)
            LL[j + k * M] = eo_dnbinom_mu((double)y[j], ps[k][1], exp(ps[k][0] + off[j]), 1);
}

/* loglikDifferential, NB: R/base-loglikelihood.R:11-22.
 * design[b*N + i] in {0,1}: Dsg.Mix_b for sample i.  ll: M x B. */
EO_API void eo_loglik_differential(int64_t M, int N, int B, const int32_t *y, const double *off,
                                   const uint8_t *design, const double *delta, const double *psi, double *ll) {
    for (int b = 0; b < B; b++) {
        double ld = log(delta[b]);
        for (int64_t j = 0; j < M; j++) {
            long double acc = 0;
            for (int i = 0; i < N; i++) {
                int64_t c = j + (int64_t)i * M;
                double D = design[b * N + i];
                acc += eo_dnbinom_mu((double)y[c], exp(psi[1] + psi[3] * D), exp(psi[0] + psi[2] * D + off[c]), 1);
            }
            ll[j + (int64_t)b * M] = ld + (double)acc;
        }
    }
}

/* loglikDifferentialHMM, NB: R/base-loglikelihood.R:45-71,96.  psi = (b_bg, l_bg, db, dl). LL: M x 3 */
EO_API void eo_loglik_differential_hmm(int64_t M, int N, int B, const int32_t *y, const double *off,
                                       const uint8_t *design, const double *delta, const double *psi,
                                       double minZero, double *LL) {
    double *ll = (double *)malloc(sizeof(double) * (size_t)(M * B));
    eo_loglik_differential(M, N, B, y, off, design, delta, psi, ll);
    for (int64_t j = 0; j < M; j++) {
        long double a0 = 0, a2 = 0;
        for (int i = 0; i < N; i++) {
            int64_t c = j + (int64_t)i * M;
            a0 += eo_dnbinom_mu((double)y[c], exp(psi[1]), exp(psi[0] + off[c]), 1);
            a2 += eo_dnbinom_mu((double)y[c], exp(psi[1] + psi[3]), exp(psi[0] + psi[2] + off[c]), 1);
        }
        double s = exp(ll[j]); /* Reduce(`+`, lapply(ll, exp)) : left to right */
        for (int b = 1; b < B; b++) s = s + exp(ll[j + (int64_t)b * M]);
        double v[3] = {(double)a0, log(s), (double)a2};
        for (int k = 0; k < 3; k++) LL[j + k * M] = isinf(v[k]) ? log(minZero) : v[k];
    }
    free(ll);
}

/* innerExpStep: R/base-EM.R:302-314.  eta: M x B */
EO_API void eo_inner_estep(int64_t M, int N, int B, const int32_t *y, const double *off, const uint8_t *design,
                           const double *delta, const double *psi, double minZero, double *eta) {
    eo_loglik_differential(M, N, B, y, off, design, delta, psi, eta);
    double lmz = log(minZero);
    for (int64_t j = 0; j < M; j++) {
        double s = 0;
        for (int b = 0; b < B; b++) {
            double v = eta[j + (int64_t)b * M];
            if (exp(v) == 0) v = lmz;
            v = exp(v);
            eta[j + (int64_t)b * M] = v;
            s = (b == 0) ? v : s + v;
        }
        for (int b = 0; b < B; b++) eta[j + (int64_t)b * M] /= s;
    }
}

/* ---- dist = 'zinb', differential model: psi = (lambda, b_bg, l_bg, db, dl) ---------------- */
/* dzinb: R/base-loglikelihood.R:152-158 */
static double eo_dzinb(double x, double size, double mu, double zip, int lg, double minZero) {
    double v = zip * (x == 0 ? 1.0 : 0.0) + (1 - zip) * (eo_dnbinom_mu(x, size, mu, 0) + minZero);
    return lg ? log(v) : v;
}

/* loglikDifferential, ZINB: R/base-loglikelihood.R:23-36 (the masks multiply, as in R) */
EO_API void eo_loglik_differential_zinb(int64_t M, int N, int B, const int32_t *y, const double *off,
                                        const uint8_t *design, const double *delta, const double *psi,
                                        double minZero, double *ll) {
    double zip = exp(psi[0]) / (1 + exp(psi[0]));
    for (int b = 0; b < B; b++) {
        double ld = log(delta[b]);
        for (int64_t j = 0; j < M; j++) {
            long double acc = 0;
            for (int i = 0; i < N; i++) {
                int64_t c = j + (int64_t)i * M;
                double D = design[b * N + i];
                double nb = eo_dnbinom_mu((double)y[c], exp(psi[2] + psi[4]), exp(psi[1] + psi[3] + off[c]), 1);
                double zi = eo_dzinb((double)y[c], exp(psi[2]), exp(psi[1] + off[c]), zip, 1, minZero);
                acc += (D == 1) * nb + (D == 0) * zi;
            }
            ll[j + (int64_t)b * M] = ld + (double)acc;
        }
    }
}

/* loglikDifferentialHMM, ZINB: R/base-loglikelihood.R:72-96 */
EO_API void eo_loglik_differential_hmm_zinb(int64_t M, int N, int B, const int32_t *y, const double *off,
                                            const uint8_t *design, const double *delta, const double *psi,
                                            double minZero, double *LL) {
    double *ll = (double *)malloc(sizeof(double) * (size_t)(M * B));
    eo_loglik_differential_zinb(M, N, B, y, off, design, delta, psi, minZero, ll);
    double zip = exp(psi[0]) / (1 + exp(psi[0]));
    for (int64_t j = 0; j < M; j++) {
        long double a0 = 0, a2 = 0;
        for (int i = 0; i < N; i++) {
            int64_t c = j + (int64_t)i * M;
            a0 += eo_dzinb((double)y[c], exp(psi[2]), exp(psi[1] + off[c]), zip, 1, minZero);
            a2 += eo_dnbinom_mu((double)y[c], exp(psi[2] + psi[4]), exp(psi[1] + psi[3] + off[c]), 1);
        }
        double s = exp(ll[j]);
        for (int b = 1; b < B; b++) s = s + exp(ll[j + (int64_t)b * M]);
        double v[3] = {(double)a0, log(s), (double)a2};
        for (int k = 0; k < 3; k++) LL[j + k * M] = isinf(v[k]) ? log(minZero) : v[k];
    }
    free(ll);
}

/* innerExpStep with dist = 'zinb': R/base-EM.R:302-314 */
EO_API void eo_inner_estep_zinb(int64_t M, int N, int B, const int32_t *y, const double *off, const uint8_t *design,
                                const double *delta, const double *psi, double minZero, double *eta) {
    eo_loglik_differential_zinb(M, N, B, y, off, design, delta, psi, minZero, eta);
    double lmz = log(minZero);
    for (int64_t j = 0; j < M; j++) {
        double s = 0;
        for (int b = 0; b < B; b++) {
            double v = eta[j + (int64_t)b * M];
            if (exp(v) == 0) v = lmz;
            v = exp(v);
            eta[j + (int64_t)b * M] = v;
            s = (b == 0) ? v : s + v;
        }
        for (int b = 0; b < B; b++) eta[j + (int64_t)b * M] /= s;
    }
}

/* glmMixZINB / derivMixZINB: R/base-optim.R:261-362.  Same long table as eo_glm_mix_nb;
 * par = (lambda, b_bg, l_bg, b_enr, l_enr) (all five free: optimDifferential passes b_bg + db). */
EO_API double eo_glm_mix_zinb(int64_t n, int B, int maxid, const double *par, const int32_t *id,
                              const double *w, const double *y, const double *off, const uint8_t *dsg,
                              double minZero, double *grad) {
    long double f = 0, g5[5] = {0, 0, 0, 0, 0};
    (void)B;
    double zip = exp(par[0]) / (1 + exp(par[0]));
    double phz = exp(par[2]), phn = exp(par[4]);
    for (int64_t r = 0; r < n; r++) {
        double D; /* 1: NB (enriched) row, 0: ZINB (background) row */
        if (id[r] == 1) D = 0;
        else if (id[r] == maxid) D = 1;
        else D = dsg[r + (int64_t)(id[r] - 2) * n];
        double muz = exp(par[1] + off[r]), mun = exp(par[3] + off[r]);
        double lnb = log(eo_dnbinom_mu(y[r], phn, mun, 0) + minZero);
        double lzi = eo_dzinb(y[r], phz, muz, zip, 1, minZero);
        if (id[r] == 1) f += -w[r] * lzi;
        else if (id[r] == maxid) f += -w[r] * lnb;
        else f += -w[r] * ((D == 1) * lnb + (D == 0) * lzi);
        if (grad) {
            double y0 = (y[r] == 0) ? 1.0 : 0.0, yn = 1 - y0;
            if (id[r] != maxid) { /* ZINB part (masked by Dsg.Mix == 0 on mixture rows) */
                double mk = (id[r] == 1) ? 1.0 : (D == 0);
                double dz = eo_dzinb(0.0, phz, muz, zip, 0, minZero);
                double g1 = -w[r] * mk * (y0 * (((1 - eo_dnbinom_mu(0.0, phz, muz, 0)) / dz) * zip * (1 - zip)) + yn * (-zip));
                double g2 = -w[r] * mk * (y0 * ((-(1 - zip) * pow(phz / (muz + phz), phz + 1) / dz) * muz) +
                                          yn * ((y[r] / muz - (y[r] + phz) / (muz + phz)) * muz));
                double g3 = -w[r] * mk * (y0 * ((1 - zip) * pow(phz / (muz + phz), phz) * (log(phz / (muz + phz)) + muz / (muz + phz)) * (1 / dz) * phz) +
                                          yn * ((eo_digamma(y[r] + phz) - eo_digamma(phz) + 1 + log(phz) - (y[r] + phz) / (muz + phz) - log(muz + phz)) * phz));
                g5[0] += g1; g5[1] += g2; g5[2] += g3;
            }
            if (id[r] != 1) { /* NB part (masked by Dsg.Mix == 1 on mixture rows) */
                double mk = (id[r] == maxid) ? 1.0 : (D == 1);
                double g4 = -w[r] * mk * (y[r] / mun - ((y[r] + phn) / (mun + phn))) * mun;
                double g5v = -w[r] * mk * (eo_digamma(y[r] + phn) - eo_digamma(phn) + 1 + log(phn) - (y[r] + phn) / (mun + phn) - log(mun + phn)) * phn;
                g5[3] += g4; g5[4] += g5v;
            }
        }
    }
    if (grad) for (int i = 0; i < 5; i++) grad[i] = (double)g5[i];
    return (double)f;
}

/* ------------------------------------------------------------------------ */
/* NB-GLM objective / gradient (R/base-optim.R).  sum()/colSums() use long   */
/* double accumulation as R does.                                            */
/* ------------------------------------------------------------------------ */

/* glmNB / derivNB: R/base-optim.R:109-130.  par = (beta[0..P-1], phi); x: G x P col-major */
EO_API double eo_glm_nb(int64_t G, int P, const double *par, const double *y, const double *x,
                        const double *offset, const double *w, double *grad) {
    double phi = par[P];
    long double f = 0, gphi = 0, gb[8] = {0};
    for (int64_t g = 0; g < G; g++) {
        double eta = 0;
        for (int c = 0; c < P; c++) eta += x[g + (int64_t)c * G] * par[c]; /* x %*% par */
        double mu = exp(eta + offset[g]);
        double l = eo_dnbinom_mu(y[g], 1 / phi, mu, 1);
        if (isinf(l)) l = log(1e-300);
        f += w[g] * l;
        if (grad) {
            double r = (y[g] - mu) / (1 + phi * mu);
            for (int c = 0; c < P; c++) gb[c] += w[g] * r * x[g + (int64_t)c * G];
            gphi += w[g] * (log(1 + phi * mu) + phi * (y[g] - mu) / (1 + phi * mu) -
                            eo_digamma(y[g] + 1 / phi) + eo_digamma(1 / phi)) / (phi * phi);
        }
    }
    if (grad) {
        for (int c = 0; c < P; c++) grad[c] = -(double)gb[c];
        grad[P] = -(double)gphi;
    }
    return -(double)f;
}

/* glmZINB / derivZINB: R/base-optim.R:136-181.  x: G x P column-major; par = (beta[P], phi,
 * lambda[P]); returns the value, grad (2P + 1 entries) when non-NULL. */
EO_API double eo_glm_zinb(int64_t G, int P, const double *par, const double *y, const double *x,
                          const double *offset, const double *w, double minZero, double *grad) {
    double phi = par[P];
    long double f = 0, gphi = 0, gb[8] = {0}, gl[8] = {0};
    for (int64_t g = 0; g < G; g++) {
        double eb = 0, el = 0;
        for (int c = 0; c < P; c++) {
            eb += x[g + (int64_t)c * G] * par[c];
            el += x[g + (int64_t)c * G] * par[P + 1 + c];
        }
        double mu = exp(eb + offset[g]);
        double z = 1 / (1 + exp(-(el + offset[g])));
        double y0 = (y[g] == 0) ? 1.0 : 0.0;
        double d0 = log(eo_dnbinom_mu(0.0, 1 / phi, mu, 0) + minZero);
        double d1 = log(eo_dnbinom_mu(y[g], 1 / phi, mu, 0) + minZero);
        double l0 = log(z + exp(log(1 - z) + d0));
        double l1 = log(1 - z) + d1;
        f += w[g] * (y0 * l0 + (1 - y0) * l1);
        if (grad) {
            double P0 = z + exp(log(1 - z) + d0);
            double P1 = exp(log(1 - z) + d1);
            double aux = 1 + phi * mu;
            double db1 = y0 * (w[g] / P0) * (1 - z) * (-mu * pow(1 + phi * mu, -(1 / phi + 1)));
            double db2 = (1 - y0) * (w[g] / P1) * (1 - z) * exp(d1) * (y[g] - mu) / (1 + phi * mu);
            double dp1 = y0 * (w[g] / P0) * (1 - z) * pow(aux, -1 / phi) * (log(aux) / (phi * phi) - mu / (phi * aux));
            double dp2 = (1 - y0) * (w[g] / P1) * (1 - z) * exp(d1) *
                         (log(aux) / (phi * phi) - mu * (y[g] + 1 / phi) / aux - eo_digamma(y[g] + 1 / phi) / (phi * phi) +
                          eo_digamma(1 / phi) / (phi * phi) + y[g] / phi);
            double dl1 = y0 * (w[g] / P0) * (1 - exp(d0)) * z * (1 - z);
            double dl2 = (1 - y0) * (w[g] / P1) * (-exp(d1)) * z * (1 - z);
            for (int c = 0; c < P; c++) {
                gb[c] += (db1 + db2) * x[g + (int64_t)c * G];
                gl[c] += (dl1 + dl2) * x[g + (int64_t)c * G];
            }
            gphi += dp1 + dp2;
        }
    }
    if (grad) {
        for (int c = 0; c < P; c++) { grad[c] = -(double)gb[c]; grad[P + 1 + c] = -(double)gl[c]; }
        grad[P] = -(double)gphi;
    }
    return -(double)f;
}

/* glmMixNB / derivMixNB: R/base-optim.R:187-255.  Long table of n rows with
 * id in 1..maxid (1 = background, maxid = enrichment, 2..maxid-1 = mixture
 * component id-1 whose design column is dsg[, id-2]).  par = (b, db, l, dl).
 * dsg: n x B column-major 0/1. */
EO_API double eo_glm_mix_nb(int64_t n, int B, int maxid, const double *par, const int32_t *id,
                            const double *w, const double *y, const double *off, const uint8_t *dsg,
                            double minZero, double *grad) {
    long double f = 0, g4[4] = {0, 0, 0, 0};
    for (int64_t r = 0; r < n; r++) {
        double D1, D2; /* multipliers of (db, dl) in mean / dispersion */
        if (id[r] == 1) D1 = 0;
        else if (id[r] == maxid) D1 = 1;
        else D1 = dsg[r + (int64_t)(id[r] - 2) * n];
        D2 = D1;
        double mu = exp(par[0] + par[1] * D1 + off[r]);
        double ph = exp(par[2] + par[3] * D2);
        f += -w[r] * log(eo_dnbinom_mu(y[r], ph, mu, 0) + minZero);
        if (grad) {
            double a = -w[r] * (y[r] / mu - ((y[r] + ph) / (mu + ph))) * mu;
            double c = -w[r] * (eo_digamma(y[r] + ph) - eo_digamma(ph) + 1 + log(ph) - (y[r] + ph) / (mu + ph) - log(mu + ph)) * ph;
            g4[0] += a; g4[1] += a * D1; g4[2] += c; g4[3] += c * D2;
        }
    }
    if (grad) for (int i = 0; i < 4; i++) grad[i] = (double)g4[i];
    (void)B;
    return (double)f;
}

EO_API int eo_version(void) { return 1; }

/* ------------------------------------------------------------------------ */
/* fdrControl: R/base-utils.R:11-25.  notpp = 1 - prob; data.table's order()  */
/* is a stable sort (ties keep window order); [EXT-R] cumsum() accumulates in */
/* long double and stores each partial sum as a double (src/main/cum.c);      */
/* FDR = cumsum / seq_len(.N) < fdr, put back in window order.                */
/* Returns -1 when a probability lies outside [0, 1] (the reference stops).   */
/* ------------------------------------------------------------------------ */
typedef struct { double v; int64_t i; } eo_kv;
static int eo_kv_cmp(const void *a, const void *b) {
    const eo_kv *x = (const eo_kv *)a, *y = (const eo_kv *)b;
    if (x->v < y->v) return -1;
    if (x->v > y->v) return 1;
    return (x->i > y->i) - (x->i < y->i);   /* ties: ascending window = a stable sort */
}
EO_API int64_t eo_fdr_control(int64_t n, const double *prob, double fdr, unsigned char *out) {
    eo_kv *kv = (eo_kv *)malloc(sizeof(eo_kv) * (size_t)(n > 0 ? n : 1));
    for (int64_t i = 0; i < n; i++) {
        if (!(prob[i] >= 0.0 && prob[i] <= 1.0)) { free(kv); return -1; }
        kv[i].v = 1.0 - prob[i];
        kv[i].i = i;
    }
    qsort(kv, (size_t)n, sizeof(eo_kv), eo_kv_cmp);
    long double sum = 0.0L;
    int64_t called = 0;
    for (int64_t r = 0; r < n; r++) {
        sum += kv[r].v;
        const double s = (double)sum;
        const unsigned char c = (s / (double)(r + 1)) < fdr;
        out[kv[r].i] = c;
        called += c;
    }
    free(kv);
    return called;
}
