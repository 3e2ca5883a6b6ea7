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

// x < p draws rbinom(1, x / p); R's rbinom returns 0 without drawing when the probability is 0
__device__ __forceinline__ bool rc_flagged(double x, double p) { return x < p && x / p != 0.0; }

// weight of mixture component b in the differential model: P1[j,1] * eta[j,b]
// (src/differentialRejectionControlled.cpp:38-41)
__device__ __forceinline__ double rc_mix_weight(double p1_diff, double eta_b) { return __dmul_rn(p1_diff, eta_b); }

}  // namespace ehmm
