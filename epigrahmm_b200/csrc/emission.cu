// emission.cu -- negative-binomial emission log-likelihoods streamed from the bins x samples
// count matrix (replaces R/base-loglikelihood.R:11-143 and R/base-EM.R:24-30,302-314, which call
// stats::dnbinom once per (bin, sample, component) from R).
//
// NB log-density in the (mu, size) parameterisation, with t = log(mu/size) = eta - log(size):
//   log f(x) = [lgamma(x+size) - lgamma(size) - lgamma(x+1)] - size*softplus(t) - x*softplus(-t)
// The bracket depends only on the integer count and the state's size, so it is tabulated once per
// call (lgtab); the rest costs one exp and one log1p per (bin, sample, density).
#include "common.cuh"
#include "rcweights.cuh"
#include <algorithm>

namespace ehmm {

namespace {

constexpr int EM_NT = 256;
constexpr int TAB_MAX = 8192;  // counts >= TAB_MAX fall back to lgamma in the kernel
constexpr int MAX_DENS = 4;    // distinct (size) values per launch

struct EmisParams {
    double b0[MAX_DENS];     // intercept
    double size[MAX_DENS];   // NB size
    double lsize[MAX_DENS];  // log(size)
    double lgs[MAX_DENS];    // lgamma(size)
    double minZero, lminZero, thr;
    double z0;               // zero-inflation predictor of density 0 (ZINB background): z0 + offset
    double zip;              // differential model with dist = 'zinb': constant zero-inflation probability of density 0
    int zi_diff;             // ... and the flag that switches it on
    int tabn;
};

struct MixParams {
    uint64_t mask[EHMM_MAX_B + 2];  // bit i set <=> Dsg.Mix_b of sample i is 1
    double ldelta[EHMM_MAX_B + 2];  // log(delta_b)
    int B;
};

// tab[d*tabn + x] = lgamma(x+size_d) - lgamma(size_d) - lgamma(x+1)
__global__ void lgtab_kernel(EmisParams P, int nd, double *__restrict__ tab) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nd * P.tabn) return;
    int d = i / P.tabn, x = i % P.tabn;
    tab[i] = (lgamma((double)x + P.size[d]) - P.lgs[d]) - lgamma((double)x + 1.0);
}

// counts beyond the per-call table (rare): kept out of line, lgamma is long and would otherwise be
// inlined at every unrolled use of nb_logpmf
__device__ __noinline__ double lg_tail(double x, double size, double lgs) {
    return (lgamma(x + size) - lgs) - lgamma(x + 1.0);
}

__device__ __forceinline__ double nb_logpmf(int x, double eta, int d, const EmisParams &P,
                                            const double *__restrict__ tab) {
    const double t = eta - P.lsize[d];
    const double w = log1p(exp(-fabs(t)));
    const double sp = (t > 0.0) ? t + w : w;   // softplus(t)  = log(1 + mu/size)
    const double spm = (t > 0.0) ? w : w - t;  // softplus(-t) = log(1 + size/mu)
    const double c = (x < P.tabn) ? __ldg(tab + d * P.tabn + x) : lg_tail((double)x, P.size[d], P.lgs[d]);
    return (c - P.size[d] * sp) - (double)x * spm;
}

// log(dnbinom(log=FALSE) + minZero), R/base-loglikelihood.R:115,121
__device__ __forceinline__ double plus_min_zero(double l, const EmisParams &P) {
    return (l > P.thr) ? l : log(exp(l) + P.minZero);
}

// zero-inflated background cell (R/base-loglikelihood.R:124-133): l01 = log(dnbinom(x) + minZero);
// for x = 0 the reference's ll00 is the same number
__device__ __forceinline__ double zinb_cell(int x, double l01, double xbeta) {
    const double zip = 1.0 / (1.0 + exp(-xbeta));
    const double l1mz = log(1.0 - zip);
    return (x == 0) ? log(zip + exp(l1mz + l01)) : (l1mz + l01);
}

// dzinb(log = TRUE), R/base-loglikelihood.R:152-158, from the NB log-density l of the same cell
__device__ __forceinline__ double dzinb_log(int x, double l, const EmisParams &P) {
    return log((x == 0 ? P.zip : 0.0) + (1.0 - P.zip) * (exp(l) + P.minZero));
}

// loglikConsensusHMM: LL[j,k] = sum_i log(NB(y_ji; mu=exp(b_k + o_ji), size_k) + minZero), where
// b_k = th_k0 [+ th_k1*ctrl_11]: with a controls assay the reference adds theta[2]*controls[1], the
// control of the table's FIRST cell, to every cell -- its ifelse(useControl, theta[2]*controls, 0) has
// a length-1 test and so returns one element (R/base-loglikelihood.R:113,119,126-129,137).  That
// scalar is folded into b_k on the host (emis_consensus), so no kernel reads the controls matrix.
// ZI: the background state is zero-inflated (dist = 'zinb')
template <bool PLUS_MZ, bool ZI = false>
__global__ void __launch_bounds__(EM_NT)
    emis_consensus_kernel(int64_t M, int N, const int32_t *__restrict__ y, const double *__restrict__ off,
                          const __grid_constant__ EmisParams P, const double *__restrict__ tab,
                          double *__restrict__ LL) {
    for (int64_t j = (int64_t)blockIdx.x * EM_NT + threadIdx.x; j < M; j += (int64_t)gridDim.x * EM_NT) {
        double a0 = 0.0, a1 = 0.0;
        for (int i = 0; i < N; i++) {
            const int64_t c = j + (int64_t)i * M;
            const int x = __ldg(y + c);
            const double o = __ldg(off + c);
            double l0 = nb_logpmf(x, P.b0[0] + o, 0, P, tab), l1 = nb_logpmf(x, P.b0[1] + o, 1, P, tab);
            if (PLUS_MZ) { l0 = plus_min_zero(l0, P); l1 = plus_min_zero(l1, P); }
            if (ZI) l0 = zinb_cell(x, l0, P.z0 + o);
            a0 += l0;
            a1 += l1;
        }
        LL[j] = a0;
        LL[j + M] = a1;
    }
}

// ---- dictionary-encoded path ------------------------------------------------------------------
// dt$Group maps every (bin, sample) cell to a row of dtUnique (distinct (count, offset[, control])
// tuples; offsets are rounded to 3 decimals upstream, so G << M*N).  The densities are evaluated
// once per group into a table and each bin gathers them: 4N B/bin of ids instead of 12N B/bin of
// counts+offsets, and no transcendental work in the M*N stream.  Values are bit-identical to the
// per-cell kernels (same function of the same inputs, same summation order).
template <bool PLUS_MZ, bool ZI = false>
__global__ void __launch_bounds__(EM_NT)
    group_table_consensus_kernel(int64_t G, const double *__restrict__ uy, const double *__restrict__ uoff,
                                 const __grid_constant__ EmisParams P, const double *__restrict__ tab,
                                 double2 *__restrict__ E) {
    const int64_t g = (int64_t)blockIdx.x * EM_NT + threadIdx.x;
    if (g > G) return;
    if (g == 0) { E[0] = make_double2(0.0, 0.0); return; }
    const int x = (int)uy[g];
    const double o = uoff[g];
    double l0 = nb_logpmf(x, P.b0[0] + o, 0, P, tab), l1 = nb_logpmf(x, P.b0[1] + o, 1, P, tab);
    if (PLUS_MZ) { l0 = plus_min_zero(l0, P); l1 = plus_min_zero(l1, P); }
    if (ZI) l0 = zinb_cell(x, l0, P.z0 + o);
    E[g] = make_double2(l0, l1);
}

__global__ void __launch_bounds__(EM_NT)
    gather_consensus_kernel(int64_t M, int N, const int32_t *__restrict__ group, const double2 *__restrict__ E,
                            double *__restrict__ LL) {
    for (int64_t j = (int64_t)blockIdx.x * EM_NT + threadIdx.x; j < M; j += (int64_t)gridDim.x * EM_NT) {
        double a0 = 0.0, a1 = 0.0;
        for (int i = 0; i < N; i++) {
            const double2 e = __ldg(E + __ldg(group + j + (int64_t)i * M));
            a0 += e.x;
            a1 += e.y;
        }
        LL[j] = a0;
        LL[j + M] = a1;
    }
}

// differential model: E[g] = (l0, l1) = log NB under D = 0 / D = 1
__global__ void __launch_bounds__(EM_NT)
    group_table_diff_kernel(int64_t G, const double *__restrict__ uy, const double *__restrict__ uoff,
                            const __grid_constant__ EmisParams P, const double *__restrict__ tab,
                            double2 *__restrict__ E) {
    const int64_t g = (int64_t)blockIdx.x * EM_NT + threadIdx.x;
    if (g > G) return;
    if (g == 0) { E[0] = make_double2(0.0, 0.0); return; }
    const int x = (int)uy[g];
    const double o = uoff[g];
    double l0 = nb_logpmf(x, P.b0[0] + o, 0, P, tab);
    if (P.zi_diff) l0 = dzinb_log(x, l0, P);
    E[g] = make_double2(l0, nb_logpmf(x, P.b0[1] + o, 1, P, tab));
}

// per-bin pieces of the differential model: LL0 = sum_i l0_i, LL2 = sum_i l1_i, d_i = l1_i - l0_i
// (y == nullptr selects the dictionary-encoded source: `off` then points at the group table and
// `tab` at the int32 group ids)
template <int NMAX, bool GROUPED>
__device__ __forceinline__ void diff_cells(int64_t j, int64_t M, int N, const int32_t *__restrict__ y,
                                           const double *__restrict__ off, const EmisParams &P,
                                           const double *__restrict__ tab, double &LL0, double &LL2, double *d) {
    LL0 = 0.0;
    LL2 = 0.0;
#pragma unroll
    for (int i = 0; i < NMAX; i++) {
        if (i < N) {
            const int64_t c = j + (int64_t)i * M;
            double l0, l1;
            if (GROUPED) {
                const double2 e = __ldg(reinterpret_cast<const double2 *>(off) +
                                        __ldg(reinterpret_cast<const int32_t *>(tab) + c));
                l0 = e.x;
                l1 = e.y;
            } else {
                const int x = __ldg(y + c);
                const double o = __ldg(off + c);
                l0 = nb_logpmf(x, P.b0[0] + o, 0, P, tab);
                if (P.zi_diff) l0 = dzinb_log(x, l0, P);
                l1 = nb_logpmf(x, P.b0[1] + o, 1, P, tab);
            }
            LL0 += l0;
            LL2 += l1;
            d[i] = l1 - l0;
        } else {
            d[i] = 0.0;
        }
    }
}

// dictionary-encoded source with the group ids already in registers (the caller fetched them one strip ahead, so
// the table gathers of this strip do not wait for an id load first)
template <int NMAX>
__device__ __forceinline__ void diff_cells_ids(const int (&gid)[NMAX], int N, const double2 *__restrict__ E, double &LL0,
                                               double &LL2, double *d) {
    double2 e[NMAX];
#pragma unroll
    for (int i = 0; i < NMAX; i++) e[i] = (i < N) ? __ldg(E + gid[i]) : make_double2(0.0, 0.0);
    LL0 = 0.0;
    LL2 = 0.0;
#pragma unroll
    for (int i = 0; i < NMAX; i++) {
        if (i < N) {
            LL0 += e[i].x;
            LL2 += e[i].y;
            d[i] = e[i].y - e[i].x;
        } else {
            d[i] = 0.0;
        }
    }
}

template <int NMAX>
__device__ __forceinline__ double mix_ll(int b, double LL0, const double *d, const MixParams &X) {
    double s = 0.0;
    const uint64_t mk = X.mask[b];
#pragma unroll
    for (int i = 0; i < NMAX; i++)
        if ((mk >> i) & 1ull) s += d[i];  // mk is uniform: a predicated add (adding 0.0 instead is the same sum)
    return X.ldelta[b] + (LL0 + s);
}

// loglikDifferentialHMM, R/base-loglikelihood.R:45-96.  GROUPED: the dictionary-encoded source (the
// kernel then holds no density code at all, which keeps it small enough for the instruction cache)
template <int NMAX, bool GROUPED>
__global__ void __launch_bounds__(EM_NT)
    emis_diff_hmm_kernel(int64_t M, int N, const int32_t *__restrict__ y, const double *__restrict__ off,
                         const __grid_constant__ EmisParams P, const __grid_constant__ MixParams X,
                         const double *__restrict__ tab, double *__restrict__ LL) {
    for (int64_t j = (int64_t)blockIdx.x * EM_NT + threadIdx.x; j < M; j += (int64_t)gridDim.x * EM_NT) {
        double LL0, LL2, d[NMAX];
        diff_cells<NMAX, GROUPED>(j, M, N, y, off, P, tab, LL0, LL2, d);
        double s = 0.0;
        for (int b = 0; b < X.B; b++) s += exp(mix_ll<NMAX>(b, LL0, d, X));  // not max-shifted, as in the reference
        double LL1 = log(s);
        LL[j] = isinf(LL0) ? P.lminZero : LL0;
        LL[j + M] = isinf(LL1) ? P.lminZero : LL1;
        LL[j + 2 * M] = isinf(LL2) ? P.lminZero : LL2;
    }
}

// innerExpStep, R/base-EM.R:302-314: eta[j,b] = exp(ll_b)/sum_b exp(ll_b) with exp(ll_b)==0 -> minZero.
// Two fusions, both optional (lp1 == nullptr switches them off):
//  * innerMaxStepProb's sums (src/innerMaxStepProb.cpp:33-35): with the posterior of the differential state
//    at hand, sum_j P1[j,1]*eta[j,b] and sum_j P1[j,1] need no extra pass (part[blk*(B+1) + b], reduced
//    by colsum_kernel);
//  * COUNT: the rejection control's first pass (rc.cu).  The weights it thins are P1[j,0], P1[j,1]*eta[j,b]
//    and P1[j,2]; all are in registers here, so the number of weights below the cut `pcut` is counted per
//    (column, warp tile of 256 bins) on the way -- the bins are walked in warp tiles for that.
template <int NMAX, int BMAX, bool GROUPED, bool COUNT>
__global__ void __launch_bounds__(EM_NT, (GROUPED && BMAX <= 14) ? 3 : 1)
    emis_inner_kernel(int64_t M, int N, const int32_t *__restrict__ y, const double *__restrict__ off,
                      const __grid_constant__ EmisParams P, const __grid_constant__ MixParams X,
                      const double *__restrict__ tab, double *__restrict__ eta, const double *__restrict__ lp1, int lin,
                      double *__restrict__ part, double pcut, int64_t nwt, int *__restrict__ counts) {
    double acc[BMAX + 1];
#pragma unroll
    for (int b = 0; b <= BMAX; b++) acc[b] = 0.0;
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (EM_NT / 32);
    for (int64_t tile = ((int64_t)blockIdx.x * EM_NT + threadIdx.x) >> 5; tile < nwt; tile += nwarps) {
        int cnt0 = 0, cnt1 = 0;  // lane (c & 31) counts column c (cnt1: columns 32..63)
        int gnext[NMAX];         // GROUPED: the group ids of the coming strip
        if (GROUPED) {
            const int64_t j0 = tile * RC_WT + lane;
#pragma unroll
            for (int i = 0; i < NMAX; i++)
                gnext[i] = (i < N && j0 < M) ? __ldg(reinterpret_cast<const int32_t *>(tab) + j0 + (int64_t)i * M) : 0;
        }
#pragma unroll 1
        for (int r = 0; r < RC_STRIPS; r++) {
            const int64_t j = tile * RC_WT + r * 32 + lane;
            const bool valid = j < M;
            double v[BMAX], s = 0.0, p1 = 0.0;
            double LL0 = 0.0, LL2 = 0.0, d[NMAX];
            if (GROUPED) {
                int gcur[NMAX];
#pragma unroll
                for (int i = 0; i < NMAX; i++) gcur[i] = gnext[i];
                diff_cells_ids<NMAX>(gcur, N, reinterpret_cast<const double2 *>(off), LL0, LL2, d);  // row 0 of the table for idle lanes
                const int64_t jn = j + 32;
                const bool more = (r + 1 < RC_STRIPS) && jn < M;
#pragma unroll
                for (int i = 0; i < NMAX; i++)
                    gnext[i] = (i < N && more) ? __ldg(reinterpret_cast<const int32_t *>(tab) + jn + (int64_t)i * M) : 0;
            }
            if (valid) {
                if (!GROUPED) diff_cells<NMAX, GROUPED>(j, M, N, y, off, P, tab, LL0, LL2, d);
#pragma unroll
                for (int b = 0; b < BMAX; b++) {
                    if (b < X.B) {
                        v[b] = exp(mix_ll<NMAX>(b, LL0, d, X));
                        if (v[b] == 0.0) v[b] = P.minZero;
                        s += v[b];
                    }
                }
                p1 = lp1 ? rc_post(__ldg(lp1 + M + j), lin) : 0.0;
                acc[BMAX] += p1;
            }
            const double rs = valid ? 1.0 / s : 0.0;
            if (COUNT) {
                const double w0 = valid ? rc_post(__ldg(lp1 + j), lin) : 2.0;  // 2.0: never below a cut <= 1
                const int n = __popc(__ballot_sync(0xffffffffu, rc_flagged(w0, pcut)));
                if (lane == 0) cnt0 += n;
            }
#pragma unroll
            for (int b = 0; b < BMAX; b++) {
                if (b < X.B) {
                    double e = 0.0;
                    if (valid) {
                        // exp(ll) / sumExpLL as in the reference: the correctly rounded quotient from one reciprocal per
                        // bin and Markstein's correction (rc_div_cut; exact unless the significand of the sum is all ones)
                        e = rc_div_cut(v[b], s, rs);
                        eta[j + (int64_t)b * M] = e;
                        acc[b] = fma(p1, e, acc[b]);
                    }
                    if (COUNT) {
                        const int n = __popc(__ballot_sync(0xffffffffu, valid && rc_flagged(rc_mix_weight(p1, e), pcut)));
                        if (lane == ((b + 1) & 31)) { if (b + 1 < 32) cnt0 += n; else cnt1 += n; }
                    }
                }
            }
            if (COUNT) {
                const double w2 = valid ? rc_post(__ldg(lp1 + 2 * M + j), lin) : 2.0;
                const int n = __popc(__ballot_sync(0xffffffffu, rc_flagged(w2, pcut)));
                const int c = X.B + 1;
                if (lane == (c & 31)) { if (c < 32) cnt0 += n; else cnt1 += n; }
            }
        }
        if (COUNT) {
            if (lane < X.B + 2) counts[(int64_t)lane * nwt + tile] = cnt0;
            if (lane + 32 < X.B + 2) counts[(int64_t)(lane + 32) * nwt + tile] = cnt1;
        }
    }
    if (lp1) {
        __shared__ double sh[(EM_NT / 32) * (BMAX + 1)];
        const int w = threadIdx.x >> 5;
#pragma unroll
        for (int b = 0; b <= BMAX; b++) {
#pragma unroll
            for (int dd = 16; dd > 0; dd >>= 1) acc[b] += __shfl_down_sync(0xffffffffu, acc[b], dd);
            if (lane == 0) sh[w * (BMAX + 1) + b] = acc[b];
        }
        __syncthreads();
        for (int b = threadIdx.x; b <= X.B; b += EM_NT) {
            const int src = (b == X.B) ? BMAX : b;
            double t = sh[src];
            for (int i = 1; i < EM_NT / 32; i++) t += sh[i * (BMAX + 1) + src];
            part[(int64_t)blockIdx.x * (X.B + 1) + b] = t;
        }
    }
}

// innerMaxStepProb partial sums: part[blk*(B+1) + b] = sum_j p1[j]*eta[j,b], [.. + B] = sum_j p1[j]
__global__ void __launch_bounds__(256) inner_mstep_kernel(int64_t M, int B, const double *__restrict__ lp1_col1, int lin,
                                                          const double *__restrict__ eta, double *__restrict__ part) {
    __shared__ double sh[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b = 0; b <= B; b++) {
        double s = 0.0;
        for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < M; j += (int64_t)gridDim.x * 256) {
            const double p = rc_post(__ldg(lp1_col1 + j), lin);
            s += (b < B) ? p * __ldg(eta + j + (int64_t)b * M) : p;
        }
        for (int dd = 16; dd > 0; dd >>= 1) s += __shfl_down_sync(0xffffffffu, s, dd);
        if (lane == 0) sh[w] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = sh[0];
            for (int i = 1; i < 8; i++) t += sh[i];
            part[(int64_t)blockIdx.x * (B + 1) + b] = t;
        }
        __syncthreads();
    }
}

// out[c] = sum_i part[i*ncol + c]: one CTA per column, fixed order
__global__ void __launch_bounds__(256) colsum_kernel(const double *__restrict__ part, int nblk, int ncol,
                                                     double *__restrict__ out) {
    __shared__ double sh[256];
    const int c = blockIdx.x;
    double s = 0.0;
    for (int i = threadIdx.x; i < nblk; i += 256) s += part[(int64_t)i * ncol + c];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[c] = sh[0];
}

int grid_for(int64_t M) {
    int64_t g = (M + EM_NT - 1) / EM_NT;
    const int64_t cap = 148 * 8 * 4;  // a few waves of 8 resident CTAs per SM
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

int require_data(ehmm_session *s, const char *who) {
    EHMM_REQUIRE(s->N > 0 && (s->encoded || (s->d_counts && s->d_offsets)), EHMM_ERR_STATE,
                 "%s: ehmm_set_data has not been called", who);
    return EHMM_OK;
}

void fill_density(EmisParams &P, int d, double b0, double size) {
    P.b0[d] = b0;
    P.size[d] = size;
    P.lsize[d] = log(size);
    P.lgs[d] = lgamma(size);
}

int prepare_tab(ehmm_session *s, EmisParams &P, int nd, double minZero) {
    P.minZero = minZero;
    P.lminZero = log(minZero);
    P.thr = P.lminZero + 40.0;  // exp(l) + minZero == exp(l) in fp64 above this
    int64_t tabn = s->count_max + 1;
    if (tabn < 1) tabn = 1;
    if (tabn > TAB_MAX) tabn = TAB_MAX;
    P.tabn = (int)tabn;
    EHMM_TRY(s->lgtab.reserve(sizeof(double) * MAX_DENS * TAB_MAX));
    int n = nd * P.tabn;
    PROF(s, KID_EMIS_TAB), lgtab_kernel<<<(n + 255) / 256, 256, 0, s->stream>>>(P, nd, s->lgtab.as<double>());
    EHMM_CK(cudaGetLastError());
    return EHMM_OK;
}

bool use_groups(const ehmm_session *s, int K) {
    return s->groups_consistent && s->G > 0 && (size_t)(s->G + 1) * 16 <= ((size_t)64 << 20) && K <= 3;
}

int fill_mix(ehmm_session *s, MixParams &X, const double *delta) {
    EHMM_REQUIRE(s->have_design && s->B > 0, EHMM_ERR_STATE, "ehmm_set_design has not been called");
    EHMM_REQUIRE(s->N <= 64, EHMM_ERR_ARG, "differential model supports at most 64 samples (N=%d)", s->N);
    X.B = s->B;
    for (int b = 0; b < s->B; b++) {
        uint64_t m = 0;
        for (int i = 0; i < s->N; i++)
            if (s->design[b * s->N + i]) m |= (1ull << i);
        X.mask[b] = m;
        X.ldelta[b] = log(delta[b]);
    }
    return EHMM_OK;
}

// dictionary-encoded source for the differential kernels: evaluates the per-group (l0, l1) table
// and redirects (y, off, tab) to (nullptr, table, group ids); leaves them untouched otherwise
int diff_source(ehmm_session *s, const EmisParams &P, const int32_t *&y, const double *&off, const double *&tab) {
    if (!use_groups(s, 3)) return EHMM_OK;
    EHMM_TRY(s->gtab.reserve(sizeof(double2) * (size_t)(s->G + 1)));
    const unsigned gb = (unsigned)((s->G + 1 + EM_NT - 1) / EM_NT);
    PROF(s, KID_EMIS_TAB), group_table_diff_kernel<<<gb, EM_NT, 0, s->stream>>>(
        s->G, s->u_chip.as<double>(), s->u_off.as<double>(), P, s->lgtab.as<double>(), s->gtab.as<double2>());
    EHMM_CK(cudaGetLastError());
    y = nullptr;
    off = s->gtab.as<double>();
    tab = reinterpret_cast<const double *>(s->group.as<int32_t>());
    return EHMM_OK;
}

void fill_diff(EmisParams &P, const double *psi, int zinb) {
    // nb:   psi = (b_bg, l_bg, db, dl): mu_D = exp(psi1 + psi3*D + o), size_D = exp(psi2 + psi4*D)
    // zinb: psi = (lambda, b_bg, l_bg, db, dl), the D = 0 density zero-inflated with zip = logistic(lambda)
    //       (R/base-loglikelihood.R:23-36,72-95)
    if (zinb) {
        P.zi_diff = 1;
        P.zip = exp(psi[0]) / (1.0 + exp(psi[0]));
        psi += 1;
    }
    fill_density(P, 0, psi[0], exp(psi[1]));
    fill_density(P, 1, psi[0] + psi[2], exp(psi[1] + psi[3]));
}

}  // namespace

namespace {
template <bool ZI> int launch_consensus(ehmm_session *s, const EmisParams &P) {
    const int g = grid_for(s->M);
    if (use_groups(s, 2)) {
        EHMM_TRY(s->gtab.reserve(sizeof(double2) * (size_t)(s->G + 1)));
        const unsigned gb = (unsigned)((s->G + 1 + EM_NT - 1) / EM_NT);
        PROF(s, KID_EMIS_TAB), group_table_consensus_kernel<true, ZI><<<gb, EM_NT, 0, s->stream>>>(
            s->G, s->u_chip.as<double>(), s->u_off.as<double>(), P, s->lgtab.as<double>(), s->gtab.as<double2>());
        EHMM_CK(cudaGetLastError());
        PROF(s, KID_EMIS_GATHER), gather_consensus_kernel<<<g, EM_NT, 0, s->stream>>>(
            s->M, s->N, s->group.as<int32_t>(), s->gtab.as<double2>(), s->logf.as<double>());
        EHMM_CK(cudaGetLastError());
        return EHMM_OK;
    }
    PROF(s, KID_EMIS_CONSENSUS), emis_consensus_kernel<true, ZI><<<g, EM_NT, 0, s->stream>>>(
        s->M, s->N, s->d_counts, s->d_offsets, P, s->lgtab.as<double>(), s->logf.as<double>());
    EHMM_CK(cudaGetLastError());
    return EHMM_OK;
}
}  // namespace

// zi == nullptr: NB (th1 = (beta0[, beta_ctrl], size)); else zi = (lambda0[, lambda_ctrl]) of the
// zero-inflated background (R/base-loglikelihood.R:123-141).
// With controls (P_ == 3) every linear predictor gets coefficient * controls[1] -- the control of the
// first cell of the whole long table (s->ctrl_first), not the cell's own: see emis_consensus_kernel.
int emis_consensus(ehmm_session *s, const double *th1, const double *th2, int P_, double minZero, const double *zi) {
    EHMM_TRY(require_data(s, "loglikConsensusHMM"));
    EHMM_REQUIRE(s->K == 2, EHMM_ERR_ARG, "loglikConsensusHMM needs a K=2 session (K=%d)", s->K);
    EHMM_REQUIRE(P_ == 2 || P_ == 3, EHMM_ERR_ARG, "P must be 2 (beta0,size) or 3 (beta0,beta_ctrl,size)");
    EHMM_REQUIRE(P_ == 2 || s->have_ctrl_first, EHMM_ERR_STATE, "P=3 needs a controls matrix");
    const double c1 = s->ctrl_first;
    EmisParams P = {};
    // R evaluates theta[1] + theta[2]*controls[1] first and adds the offsets to that sum
    fill_density(P, 0, P_ == 3 ? th1[0] + th1[1] * c1 : th1[0], th1[P_ - 1]);
    fill_density(P, 1, P_ == 3 ? th2[0] + th2[1] * c1 : th2[0], th2[P_ - 1]);
    if (zi) P.z0 = P_ == 3 ? zi[0] + zi[1] * c1 : zi[0];
    EHMM_TRY(prepare_tab(s, P, 2, minZero));
    EHMM_TRY(logf_for_write(s, sizeof(double) * s->M * 2));
    EHMM_TRY(zi ? launch_consensus<true>(s, P) : launch_consensus<false>(s, P));
    s->valid |= DS_LOGF;
    return EHMM_OK;
}

int emis_init(ehmm_session *s, const double *psi1, const double *psi2) {
    EHMM_TRY(require_data(s, "emInitialization"));
    EHMM_REQUIRE(s->K == 2 && s->N == 1, EHMM_ERR_ARG, "the initializer supports one sample and K=2 (N=%d, K=%d)",
                 s->N, s->K);
    EmisParams P = {};
    fill_density(P, 0, psi1[0], psi1[1]);
    fill_density(P, 1, psi2[0], psi2[1]);
    EHMM_TRY(prepare_tab(s, P, 2, 2.2250738585072014e-308));
    EHMM_TRY(logf_for_write(s, sizeof(double) * s->M * 2));
    if (use_groups(s, 2)) {
        EHMM_TRY(s->gtab.reserve(sizeof(double2) * (size_t)(s->G + 1)));
        const unsigned gb = (unsigned)((s->G + 1 + EM_NT - 1) / EM_NT);
        PROF(s, KID_EMIS_TAB), group_table_consensus_kernel<false><<<gb, EM_NT, 0, s->stream>>>(
            s->G, s->u_chip.as<double>(), s->u_off.as<double>(), P, s->lgtab.as<double>(), s->gtab.as<double2>());
        EHMM_CK(cudaGetLastError());
        PROF(s, KID_EMIS_GATHER), gather_consensus_kernel<<<grid_for(s->M), EM_NT, 0, s->stream>>>(
            s->M, 1, s->group.as<int32_t>(), s->gtab.as<double2>(), s->logf.as<double>());
        EHMM_CK(cudaGetLastError());
        s->valid |= DS_LOGF;
        return EHMM_OK;
    }
    PROF(s, KID_EMIS_CONSENSUS), emis_consensus_kernel<false><<<grid_for(s->M), EM_NT, 0, s->stream>>>(
        s->M, 1, s->d_counts, s->d_offsets, P, s->lgtab.as<double>(), s->logf.as<double>());
    EHMM_CK(cudaGetLastError());
    s->valid |= DS_LOGF;
    return EHMM_OK;
}

int emis_differential_hmm(ehmm_session *s, const double *delta, const double *psi, double minZero, int zinb) {
    EHMM_TRY(require_data(s, "loglikDifferentialHMM"));
    EHMM_REQUIRE(s->K == 3, EHMM_ERR_ARG, "loglikDifferentialHMM needs a K=3 session (K=%d)", s->K);
    EmisParams P = {};
    MixParams X = {};
    fill_diff(P, psi, zinb);
    EHMM_TRY(fill_mix(s, X, delta));
    EHMM_TRY(prepare_tab(s, P, 2, minZero));
    EHMM_TRY(logf_for_write(s, sizeof(double) * s->M * 3));
    const int g = grid_for(s->M);
    const int32_t *src_y = s->d_counts;
    const double *src_off = s->d_offsets, *src_tab = s->lgtab.as<double>();
    EHMM_TRY(diff_source(s, P, src_y, src_off, src_tab));
    const bool grouped = (src_y == nullptr);
#define EHMM_DIFF(NM, GR)                                                                                             \
    PROF(s, KID_EMIS_DIFF), emis_diff_hmm_kernel<NM, GR><<<g, EM_NT, 0, s->stream>>>(s->M, s->N, src_y, src_off, P, X, \
                                                                                     src_tab, s->logf.as<double>())
    if (grouped) {  // tight unroll bounds for the shapes of BASELINE configs 3-5 (N = 6 / 10 / 12)
        if (s->N <= 6) EHMM_DIFF(6, true);
        else if (s->N <= 12) EHMM_DIFF(12, true);
        else if (s->N <= 16) EHMM_DIFF(16, true);
        else EHMM_DIFF(64, true);
    } else {
        if (s->N <= 16) EHMM_DIFF(16, false);
        else EHMM_DIFF(64, false);
    }
#undef EHMM_DIFF
    EHMM_CK(cudaGetLastError());
    s->valid |= DS_LOGF;
    return EHMM_OK;
}

int emis_inner_estep(ehmm_session *s, const double *delta, const double *psi, double minZero, int zinb) {
    EHMM_TRY(require_data(s, "innerExpStep"));
    EmisParams P = {};
    MixParams X = {};
    fill_diff(P, psi, zinb);
    EHMM_TRY(fill_mix(s, X, delta));
    EHMM_TRY(prepare_tab(s, P, 2, minZero));
    EHMM_TRY(s->eta.reserve(sizeof(double) * s->M * s->B));
    // one warp tile per warp while that keeps the grid within a few waves
    const int g = (int)std::min<int64_t>((rc_num_wtiles(s->M) + EM_NT / 32 - 1) / (EM_NT / 32), 148 * 64);
    const int32_t *src_y = s->d_counts;
    const double *src_off = s->d_offsets, *src_tab = s->lgtab.as<double>();
    EHMM_TRY(diff_source(s, P, src_y, src_off, src_tab));
    // fuse innerMaxStepProb's sums when the posteriors of the current E-step are in place, and the
    // rejection control's flagged counts when the cut its next call will use is known (ehmm_rc_hint, or
    // the previous call's cut: max(0.9^iter, probCut) is constant once it has reached probCut)
    const bool fuse = has_post(s) && s->K == 3;
    const P1Src ps = p1_source(s);
    const double pcut = s->rc_hint_p;
    const bool count = fuse && pcut > 0.0 && s->G > 0 && s->B + 2 <= 64;
    const double *lp1 = fuse ? ps.p : nullptr;
    const int64_t nwt = rc_num_wtiles(s->M);
    if (count) EHMM_TRY(s->rc_flag.reserve(sizeof(int) * nwt * (s->B + 2)));
    EHMM_TRY(s->inner_part.reserve(sizeof(double) * ((size_t)g * (s->B + 1) + s->B + 1)));
    double *part = s->inner_part.as<double>();
    const bool grouped = (src_y == nullptr);
#define EHMM_INNER2(NM, BM, GR, CT)                                                                                   \
    PROF(s, KID_EMIS_INNER), emis_inner_kernel<NM, BM, GR, CT><<<g, EM_NT, 0, s->stream>>>(                           \
        s->M, s->N, src_y, src_off, P, X, src_tab, s->eta.as<double>(), lp1, ps.lin, part, pcut, nwt, s->rc_flag.as<int>())
#define EHMM_INNER(NM, BM, GR)                        \
    do {                                              \
        if (count) EHMM_INNER2(NM, BM, GR, true);     \
        else EHMM_INNER2(NM, BM, GR, false);          \
    } while (0)
    // the loops over samples and components are unrolled to the template bounds: tight bounds for the
    // dictionary-encoded source (B = 6 / 30 / 14 in BASELINE configs 3-5), wide ones for the per-cell fallback
#define EHMM_INNER_B(NM)                               \
    do {                                               \
        if (s->B <= 6) EHMM_INNER(NM, 6, true);        \
        else if (s->B <= 14) EHMM_INNER(NM, 14, true); \
        else if (s->B <= 30) EHMM_INNER(NM, 30, true); \
        else EHMM_INNER(NM, 62, true);                 \
    } while (0)
    if (grouped) {
        if (s->N <= 6) EHMM_INNER_B(6);
        else if (s->N <= 12) EHMM_INNER_B(12);
        else if (s->N <= 16) EHMM_INNER_B(16);
        else EHMM_INNER_B(64);
    } else if (s->N <= 16) {  // per-cell fallback (no dictionary): few, wide instantiations keep the build short
        if (s->B <= 16) EHMM_INNER(16, 16, false);
        else EHMM_INNER(16, 62, false);
    } else {
        if (s->B <= 16) EHMM_INNER(64, 16, false);
        else EHMM_INNER(64, 62, false);
    }
#undef EHMM_INNER_B
#undef EHMM_INNER
#undef EHMM_INNER2
    EHMM_CK(cudaGetLastError());
    s->post_gen++;  // mixtureProb changed
    s->inner_part_blocks = fuse ? g : 0;
    if (count) {
        s->rc_counts_gen = s->post_gen;
        s->rc_counts_p = pcut;
        s->rc_counts_cols = s->B + 2;
    }
    s->valid |= DS_ETA;
    return EHMM_OK;
}

// sums[b] = sum_j P1[j,1]*eta[j,b] (b < B), sums[B] = sum_j P1[j,1]  (additive across bin blocks)
int inner_mstep_sums(ehmm_session *s, double *sums) {
    EHMM_REQUIRE(has(s, DS_ETA) && has_post(s), EHMM_ERR_STATE, "innerMaxStepProb needs mixtureProb and logProb1");
    EHMM_REQUIRE(s->K == 3, EHMM_ERR_ARG, "innerMaxStepProb needs K=3");
    const int B = s->B;
    int nblk = 148 * 4;
    double *part, *out;
    if (s->inner_part_blocks > 0) {  // sums were accumulated by the innerExpStep kernel
        nblk = s->inner_part_blocks;
        part = s->inner_part.as<double>();
        out = part + (size_t)nblk * (B + 1);
    } else {
        EHMM_TRY(s->misc.reserve(sizeof(double) * ((size_t)nblk * (B + 1) + B + 1)));
        part = s->misc.as<double>();
        out = part + (size_t)nblk * (B + 1);
        PROF(s, KID_INNER_MSTEP), inner_mstep_kernel<<<nblk, 256, 0, s->stream>>>(s->M, B, p1_source(s).p + s->M, p1_source(s).lin, s->eta.as<double>(), part);
        EHMM_CK(cudaGetLastError());
    }
    PROF(s, KID_MISC), colsum_kernel<<<B + 1, 256, 0, s->stream>>>(part, nblk, B + 1, out);
    EHMM_CK(cudaGetLastError());
    EHMM_CK(cudaMemcpyAsync(s->h_pinned + 64, out, sizeof(double) * (B + 1), cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    for (int b = 0; b <= B; b++) sums[b] = s->h_pinned[64 + b];
    return EHMM_OK;
}

int inner_maxstepprob(ehmm_session *s, double *delta_out) {
    double sums[EHMM_MAX_B + 2];
    EHMM_TRY(inner_mstep_sums(s, sums));
    for (int b = 0; b < s->B; b++) delta_out[b] = sums[b] / sums[s->B];
    return EHMM_OK;
}

}  // namespace ehmm
