// rcweights.cuh -- the posterior weights the rejection control thins, and the test that decides which of
// them consume a uniform.  Shared by every kernel that counts or applies flags (emission.cu, fb.cu,
// rc.cu): a count taken in one kernel and the ranks handed out in another must come from the same bits,
// so the expressions live here once and nothing in them can contract into an FMA.
#pragma once
#include <stdint.h>

namespace ehmm {

constexpr int RC_STRIPS = 8;
constexpr int RC_WT = 32 * RC_STRIPS;  // bins per warp tile: the unit of the flagged counts

inline int64_t rc_num_wtiles(int64_t M) { return (M + RC_WT - 1) / RC_WT; }

// x < p draws rbinom(1, x / p) (src/reweight.cpp:18-19); R's rbinom returns 0 WITHOUT drawing when the
// probability x / p is 0.  For a cut 0 < p <= 1 (a probability: the entry points reject anything else) and
// a weight x >= 0, x / p >= x, so the quotient is 0 exactly when x is: no division needed.
__device__ __forceinline__ bool rc_flagged(double x, double p) { return x < p && x != 0.0; }

// x / p for the thinning probability, with the cut's reciprocal rp = RN(1 / p) computed once on the host:
// q = RN(x * rp), r = x - q * p (exact: one FMA), q' = RN(q + r * rp).  Markstein's correction step: q' is the
// correctly rounded quotient (the significand of a cut 0.9^iter or probCut is never all ones, the one
// excluded case), i.e. the very double the reference's `x / p` holds, at 3 instructions instead of ~25.
__device__ __forceinline__ double rc_div_cut(double x, double p, double rp) {
    const double q = __dmul_rn(x, rp);
    const double r = __fma_rn(-q, p, x);
    return __fma_rn(r, rp, q);
}

// posterior from where the session keeps it: P1 itself (lin, written by the lean E-step) or logProb1
__device__ __forceinline__ double rc_post(double stored, int lin) { return lin ? stored : exp(stored); }

// weight of mixture component b in the differential model: P1[j,1] * eta[j,b]
// (src/differentialRejectionControlled.cpp:38-41)
__device__ __forceinline__ double rc_mix_weight(double p1_diff, double eta_b) { return __dmul_rn(p1_diff, eta_b); }

}  // namespace ehmm
