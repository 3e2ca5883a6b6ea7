// fb.cu -- E-step of the HMM: forward-backward + posteriors as a parallel scan of
// scaled K x K transition operators (replaces src/expStep.cpp:45-123 of the
// reference, whose single loop over all M bins is sequential and unscaled).
//
// Operator of bin j:  T_j = Gamma * diag(e_j),  e_j[k] = exp(logf[j,k] - m_j),  m_j = max_k logf[j,k]
// (bin 0 of the chain: T_0 = diag(e_0), start vector pi).  Forward: alpha_j = alpha_{j-1} T_j,
// backward: beta_{j-1} = T_j beta_j.  Every operator / vector carries a log-scale so the mantissas
// stay O(1) (exact power-of-two renormalisation), which is the log semiring under exp() but costs
// K^3 FMAs per combine instead of K^3 exp/log.
//
// Three launches per E-step over tiles of TB = NT*L bins:
//   fb_reduce_kernel : tile -> product of its operators                       (reads 8K B/bin)
//   fb_carry_kernel  : one CTA, scan over tile operators -> forward carry into and backward carry
//                      out of every tile, log-likelihood
//   fb_apply_kernel  : tile -> logFP, logBP, logProb1, logProb2 + fused sufficient statistics
//                      (reads 8K, writes 24K + 8K^2 B/bin)
// followed by a deterministic reduction of the per-tile statistics.
#include "common.cuh"
#include "logtab.h"
#include "rcweights.cuh"

namespace ehmm {

namespace {

// 64-bit literals that are not encodable as instruction immediates are materialised with two moves
// per use; kept in constant memory they are operands (or loaded once per thread).
__constant__ double kC[5] = {0.693147180559945309417232121458,             // ln 2
                             1.0 / 3.0, 1.0 / 5.0, -1.0 / 6.0, 1.0 / 7.0};  // log1p series (log_norm)
#define LN2 (kC[0])

template <int K> struct Op {
    double m[K * K];  // row-major mantissa, m[i*K+l]
    double ls;        // natural-log scale
};
template <int K> struct Vec {
    double v[K];
    double ls;
};
struct CarryIn {
    double start[EHMM_MAX_K + 1];  // forward vector entering the block + log scale
    double end[EHMM_MAX_K + 1];    // backward vector behind the block + log scale
};
struct FBParams {
    double g[EHMM_MAX_K * EHMM_MAX_K];   // Gamma row-major g[i*K+l] = gamma(i,l)
    double lg[EHMM_MAX_K * EHMM_MAX_K];  // log Gamma (host libm, same values the reference would use)
};

template <int K> struct FBCfg;
// L odd keeps the chunk-strided shared-memory accesses conflict-free (L = 6 measured 8 % slower);
// APPLY_CTAS = resident apply CTAs per SM the register budget is set for (4 spills and is slower)
template <> struct FBCfg<2> { static constexpr int NT = 256, L = 5, APPLY_CTAS = 3; };
template <> struct FBCfg<3> { static constexpr int NT = 192, L = 5, APPLY_CTAS = 3; };
// 256 threads: a larger CTA does not fit next to the side stream's resident CTAs and waits for them to drain; the scan
// is over at most 592 range operators anyway
template <int K> struct CarryCfg { static constexpr int NT = 256; };

// 2^-e for the binary exponent e encoded in the high word `hi` of a non-negative double
// (normal numbers only; e = 0 for zero / subnormal / inf / nan)
__device__ __forceinline__ double inv_pow2(int hi, int &e) {
    e = (hi >> 20) - 1023;
    if (hi < 0x00100000 || hi >= 0x7fe00000) e = 0;
    return __hiloint2double((1023 - e) << 20, 0);
}

// 1/x for normal positive x: hardware seed (2^-23) + two Newton steps (~1 ulp)
__device__ __forceinline__ double rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}

// log(x) for positive normal x from a 128-interval table (tools/gen_logtab.py): the mantissa is
// rotated into [sqrt(2)/2, sqrt(2)) as above, x = 2^e * c_i * (1 + r) with |r| <= 2^-8, and
// log1p(r) is its series to r^7 (remainder < 2^-67).  11 fp64 instructions instead of ~28; error
// below 2 ulp.  sT is the table staged in shared memory.
__device__ __forceinline__ bool normal_pos(double x) {
    return (unsigned)(__double2hiint(x) - 0x00100000) < 0x7fe00000u;
}
constexpr int HI_TINY = 0x03c00000;  // high word of 2^-963 ~ 1.3e-290: below it the linear forward value is rebuilt in log space
__device__ __forceinline__ bool not_tiny(double x) {
    return (unsigned)(__double2hiint(x) - HI_TINY) < (unsigned)(0x7ff00000 - HI_TINY);
}
// x must be a positive normal number (normal_pos)
__device__ __forceinline__ double log_norm(double x, const double2 *sT) {
    const int h2 = __double2hiint(x) + 0x95f62;  // 0x3ff00000 - 0x3fe6a09e
    const int e = (h2 >> 20) - 1023;
    const double2 t = sT[(h2 >> 13) & 127];
    const double m = __hiloint2double((h2 & 0x000fffff) + 0x3fe6a09e, __double2loint(x));
    const double r = fma(m, t.x, -1.0);
    double q = fma(r, kC[4], kC[3]);
    q = fma(r, q, kC[2]);
    q = fma(r, q, -0.25);
    q = fma(r, q, kC[1]);
    q = fma(r, q, -0.5);
    return fma((double)e, LN2, t.y) + fma(r * r, q, r);
}

// e_k = exp(f_k - max_l f_l): the entry of the maximum is exp(0) = 1 exactly, so K - 1 exponentials
template <int K> __device__ __forceinline__ void shifted_exp(const double (&f)[K], double &m, double (&e)[K]) {
    if constexpr (K == 2) {
        const double d = f[0] - f[1], x = exp(-fabs(d));
        m = fmax(f[0], f[1]);
        e[0] = (d >= 0.0) ? 1.0 : x;
        e[1] = (d >= 0.0) ? x : 1.0;
        if (d != d) { e[0] = x; e[1] = x; }  // -inf - -inf: NaN, as exp(f - m) would give
    } else if constexpr (K == 3) {
        const int am = (f[0] >= f[1] && f[0] >= f[2]) ? 0 : (f[1] >= f[2] ? 1 : 2);
        m = (am == 0) ? f[0] : (am == 1 ? f[1] : f[2]);
        const double a = (am == 0) ? f[1] : f[0], b = (am == 2) ? f[1] : f[2];  // the two others, in order
        const double ea = exp(a - m), eb = exp(b - m);
        e[0] = (am == 0) ? 1.0 : ea;
        e[1] = (am == 1) ? 1.0 : (am == 0 ? ea : eb);
        e[2] = (am == 2) ? 1.0 : eb;
    } else {
        m = f[0];
#pragma unroll
        for (int k = 1; k < K; k++) m = fmax(m, f[k]);
#pragma unroll
        for (int k = 0; k < K; k++) e[k] = exp(f[k] - m);
    }
}

template <int K> __device__ __forceinline__ void op_identity(Op<K> &X) {
#pragma unroll
    for (int i = 0; i < K; i++)
#pragma unroll
        for (int l = 0; l < K; l++) X.m[i * K + l] = (i == l) ? 1.0 : 0.0;
    X.ls = 0.0;
}

template <int K> __device__ __forceinline__ void op_renorm(Op<K> &X) {
    int mx = __double2hiint(X.m[0]);  // entries are >= 0: the high words order them (to 2^-20)
#pragma unroll
    for (int i = 1; i < K * K; i++) mx = max(mx, __double2hiint(X.m[i]));
    int e;
    double sc = inv_pow2(mx, e);
#pragma unroll
    for (int i = 0; i < K * K; i++) X.m[i] *= sc;
    X.ls += e * LN2;
}

template <int K> __device__ __forceinline__ void vec_renorm(Vec<K> &x) {
    int mx = __double2hiint(x.v[0]);
#pragma unroll
    for (int i = 1; i < K; i++) mx = max(mx, __double2hiint(x.v[i]));
    int e;
    double sc = inv_pow2(mx, e);
#pragma unroll
    for (int i = 0; i < K; i++) x.v[i] *= sc;
    x.ls += e * LN2;
}

// A applied first, then B (time order): C = A * B
template <int K> __device__ __forceinline__ Op<K> op_mul(const Op<K> &A, const Op<K> &B) {
    Op<K> C;
#pragma unroll
    for (int i = 0; i < K; i++)
#pragma unroll
        for (int l = 0; l < K; l++) {
            double acc = A.m[i * K] * B.m[l];
#pragma unroll
            for (int q = 1; q < K; q++) acc = fma(A.m[i * K + q], B.m[q * K + l], acc);
            C.m[i * K + l] = acc;
        }
    C.ls = A.ls + B.ls;
    op_renorm(C);
    return C;
}

template <int K> __device__ __forceinline__ Vec<K> vec_mul_op(const Vec<K> &x, const Op<K> &A) {
    Vec<K> r;
#pragma unroll
    for (int l = 0; l < K; l++) {
        double acc = x.v[0] * A.m[l];
#pragma unroll
        for (int i = 1; i < K; i++) acc = fma(x.v[i], A.m[i * K + l], acc);
        r.v[l] = acc;
    }
    r.ls = x.ls + A.ls;
    vec_renorm(r);
    return r;
}

template <int K> __device__ __forceinline__ Vec<K> op_mul_vec(const Op<K> &A, const Vec<K> &x) {
    Vec<K> r;
#pragma unroll
    for (int i = 0; i < K; i++) {
        double acc = A.m[i * K] * x.v[0];
#pragma unroll
        for (int l = 1; l < K; l++) acc = fma(A.m[i * K + l], x.v[l], acc);
        r.v[i] = acc;
    }
    r.ls = x.ls + A.ls;
    vec_renorm(r);
    return r;
}

template <int K> __device__ __forceinline__ Op<K> op_shfl_up(const Op<K> &X, int d) {
    Op<K> R;
#pragma unroll
    for (int i = 0; i < K * K; i++) R.m[i] = __shfl_up_sync(0xffffffffu, X.m[i], d);
    R.ls = __shfl_up_sync(0xffffffffu, X.ls, d);
    return R;
}
template <int K> __device__ __forceinline__ Op<K> op_shfl_down(const Op<K> &X, int d) {
    Op<K> R;
#pragma unroll
    for (int i = 0; i < K * K; i++) R.m[i] = __shfl_down_sync(0xffffffffu, X.m[i], d);
    R.ls = __shfl_down_sync(0xffffffffu, X.ls, d);
    return R;
}

template <int K> __device__ __forceinline__ void op_store(double *dst, const Op<K> &X) {
#pragma unroll
    for (int i = 0; i < K * K; i++) dst[i] = X.m[i];
    dst[K * K] = X.ls;
}
template <int K> __device__ __forceinline__ Op<K> op_load(const double *src) {
    Op<K> X;
#pragma unroll
    for (int i = 0; i < K * K; i++) X.m[i] = src[i];
    X.ls = src[K * K];
    return X;
}

// Striped load of one tile: thread t takes bins t, t+NT, ... (coalesced); stores
// f (optional), m_j and e_j = exp(f - m_j) column-wise in shared memory.
template <int K, int NT, int L, bool KEEP_F>
__device__ __forceinline__ void load_tile(const double *__restrict__ logf, int64_t M, int64_t s0, int nvalid,
                                          double *sF, double *sM, double *sE) {
    constexpr int TB = NT * L;
    double f[L][K];
#pragma unroll
    for (int r = 0; r < L; r++) {
        int b = r * NT + threadIdx.x;
#pragma unroll
        for (int k = 0; k < K; k++) f[r][k] = (b < nvalid) ? __ldg(logf + (int64_t)k * M + s0 + b) : 0.0;
    }
#pragma unroll
    for (int r = 0; r < L; r++) {
        int b = r * NT + threadIdx.x;
        double m, e[K];
        shifted_exp<K>(f[r], m, e);
        sM[b] = m;
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (KEEP_F) sF[k * TB + b] = f[r][k];
            sE[k * TB + b] = e[k];
        }
    }
}

// Product of the operators of thread t's chunk of L consecutive bins.
template <int K, int L, int TB>
__device__ __forceinline__ Op<K> chunk_product(const double *sE, const double *sM, int t, int nvalid, bool special0,
                                               const FBParams &P) {
    Op<K> X;
    X.ls = 0.0;
    bool has = false;
#pragma unroll
    for (int i = 0; i < L; i++) {
        int b = t * L + i;
        if (b < nvalid) {
            double e[K];
#pragma unroll
            for (int k = 0; k < K; k++) e[k] = sE[k * TB + b];
            if (!has) {
                if (special0 && b == 0) {
#pragma unroll
                    for (int r = 0; r < K; r++)
#pragma unroll
                        for (int c = 0; c < K; c++) X.m[r * K + c] = (r == c) ? e[c] : 0.0;
                } else {
#pragma unroll
                    for (int r = 0; r < K; r++)
#pragma unroll
                        for (int c = 0; c < K; c++) X.m[r * K + c] = P.g[r * K + c] * e[c];
                }
                has = true;
            } else {
                double T[K * K];
#pragma unroll
                for (int r = 0; r < K; r++)
#pragma unroll
                    for (int c = 0; c < K; c++) {
                        double acc = X.m[r * K] * P.g[c];
#pragma unroll
                        for (int q = 1; q < K; q++) acc = fma(X.m[r * K + q], P.g[q * K + c], acc);
                        T[r * K + c] = acc;
                    }
#pragma unroll
                for (int r = 0; r < K; r++)
#pragma unroll
                    for (int c = 0; c < K; c++) X.m[r * K + c] = T[r * K + c] * e[c];
            }
            X.ls += sM[b];
        }
    }
    if (!has) op_identity(X);
    op_renorm(X);
    return X;
}

// ---- kernel A: tile -> operator product ---------------------------------------------------------
template <int K, int NT, int L>
__global__ void __launch_bounds__(NT) fb_reduce_kernel(const double *__restrict__ logf, int64_t M, int first_is_start,
                                                       const __grid_constant__ FBParams P,
                                                       double *__restrict__ tileops) {
    constexpr int TB = NT * L, NW = NT / 32;
    extern __shared__ double smem[];
    double *sM = smem;
    double *sE = smem + TB;
    double *sWarp = sE + K * TB;  // NW * (K*K+1)
    const int64_t s0 = (int64_t)blockIdx.x * TB;
    const int nvalid = (int)min((int64_t)TB, M - s0);
    load_tile<K, NT, L, false>(logf, M, s0, nvalid, nullptr, sM, sE);
    __syncthreads();
    Op<K> X = chunk_product<K, L, TB>(sE, sM, threadIdx.x, nvalid, first_is_start && blockIdx.x == 0, P);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Op<K> O = op_shfl_down<K>(X, d);
        X = op_mul<K>(X, O);  // lane 0 ends with the ordered product of all 32 lanes
    }
    if (lane == 0) op_store<K>(sWarp + w * (K * K + 1), X);
    __syncthreads();
    if (threadIdx.x == 0) {
        Op<K> T = op_load<K>(sWarp);
        for (int i = 1; i < NW; i++) T = op_mul<K>(T, op_load<K>(sWarp + i * (K * K + 1)));
        op_store<K>(tileops + (int64_t)blockIdx.x * (K * K + 1), T);
    }
}

// ---- kernel B: scan over tile operators ---------------------------------------------------------
// mode 0: only the product of all tile operators -> scal[1 .. K*K+1]          (sharded phase 1)
// mode 1: carries[t] = { A_t[K], lfT, B_t[K], lbT } for every tile, scal[0] = log-likelihood,
//         scal[K*K+2 .. ] = forward vector after the last tile (K + 1 doubles)
template <int K>
__global__ void __launch_bounds__(CarryCfg<K>::NT) fb_carry_kernel(const double *__restrict__ tileops, int64_t nT, int mode,
                                                            const __grid_constant__ CarryIn cin,
                                                            double *__restrict__ carries, double *__restrict__ scal) {
    const double *startvec = cin.start, *endvec = cin.end;
    constexpr int CARRY_NT = CarryCfg<K>::NT;
    constexpr int OPN = K * K + 1, NW = CARRY_NT / 32;
    __shared__ double sWarp[NW * OPN];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t q = (nT + CARRY_NT - 1) / CARRY_NT;
    const int64_t t0 = min(nT, (int64_t)tid * q), t1 = min(nT, t0 + q);
    Op<K> X;
    op_identity(X);
    for (int64_t t = t0; t < t1; t++) X = op_mul<K>(X, op_load<K>(tileops + t * OPN));
    Op<K> Pi = X, Si = X;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Op<K> O = op_shfl_up<K>(Pi, d);
        if (lane >= d) Pi = op_mul<K>(O, Pi);
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Op<K> O = op_shfl_down<K>(Si, d);
        if (lane + d < 32) Si = op_mul<K>(Si, O);
    }
    if (lane == 31) op_store<K>(sWarp + w * OPN, Pi);
    Op<K> Pe = op_shfl_up<K>(Pi, 1), Se = op_shfl_down<K>(Si, 1);
    if (lane == 0) op_identity(Pe);
    if (lane == 31) op_identity(Se);
    __syncthreads();
    if (mode == 0) {
        if (tid == 0) {
            Op<K> T = op_load<K>(sWarp);
            for (int i = 1; i < NW; i++) T = op_mul<K>(T, op_load<K>(sWarp + i * OPN));
            op_store<K>(scal + 1, T);
        }
        return;
    }
    // forward carries
    Vec<K> v;
#pragma unroll
    for (int k = 0; k < K; k++) v.v[k] = startvec[k];
    v.ls = startvec[K];
    vec_renorm(v);
    for (int i = 0; i < w; i++) v = vec_mul_op<K>(v, op_load<K>(sWarp + i * OPN));
    v = vec_mul_op<K>(v, Pe);
    for (int64_t t = t0; t < t1; t++) {
        double *c = carries + t * (2 * K + 2);
#pragma unroll
        for (int k = 0; k < K; k++) c[k] = v.v[k];
        c[K] = v.ls;
        v = vec_mul_op<K>(v, op_load<K>(tileops + t * OPN));
    }
    if (t1 == nT && t0 < nT) {  // owner of the last tile: v = alpha at the last bin
        double sum = v.v[0];
#pragma unroll
        for (int k = 1; k < K; k++) sum += v.v[k];
        scal[0] = v.ls + log(sum);
#pragma unroll
        for (int k = 0; k < K; k++) scal[OPN + 1 + k] = v.v[k];
        scal[OPN + 1 + K] = v.ls;
    }
    // backward carries
    Vec<K> u;
#pragma unroll
    for (int k = 0; k < K; k++) u.v[k] = endvec[k];
    u.ls = endvec[K];
    vec_renorm(u);
    for (int i = NW - 1; i > w; i--) u = op_mul_vec<K>(op_load<K>(sWarp + i * OPN), u);
    u = op_mul_vec<K>(Se, u);
    for (int64_t t = t1 - 1; t >= t0; t--) {
        double *c = carries + t * (2 * K + 2) + K + 1;
#pragma unroll
        for (int k = 0; k < K; k++) c[k] = u.v[k];
        c[K] = u.ls;
        u = op_mul_vec<K>(op_load<K>(tileops + t * OPN), u);
    }
    if (t0 == 0 && t1 > 0) {  // backward vector in front of the block (for the previous bin block)
#pragma unroll
        for (int k = 0; k < K; k++) scal[OPN + 2 + K + k] = u.v[k];
        scal[OPN + 2 + 2 * K] = u.ls;
    }
}

// ---- two-level carry scan ------------------------------------------------------------------------
// The tile operators (nT ~ 1e4 .. 6e4) are scanned in two levels so that no single CTA has to pull
// them through one SM's load/store unit: R ranges of q consecutive tiles are reduced in parallel
// (fb_range_reduce_kernel), fb_carry_kernel scans the R range operators, and
// fb_range_apply_kernel expands the range carries to per-tile carries.
constexpr int RANGE_NT = 128;

template <int K>
__global__ void __launch_bounds__(RANGE_NT) fb_range_reduce_kernel(const double *__restrict__ tileops, int64_t nT, int q,
                                                                   double *__restrict__ rangeops) {
    constexpr int OPN = K * K + 1, NW = RANGE_NT / 32;
    extern __shared__ double sOps[];
    __shared__ double sWarp[NW * OPN];
    const int64_t t0 = (int64_t)blockIdx.x * q;
    const int n = (int)max((int64_t)0, min((int64_t)q, nT - t0));
    for (int i = threadIdx.x; i < n * OPN; i += RANGE_NT) sOps[i] = tileops[t0 * OPN + i];
    __syncthreads();
    const int per = (n + RANGE_NT - 1) / RANGE_NT;
    const int a = min(n, (int)threadIdx.x * per), b = min(n, a + per);
    Op<K> X;
    op_identity(X);
    for (int t = a; t < b; t++) X = op_mul<K>(X, op_load<K>(sOps + t * OPN));
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Op<K> O = op_shfl_down<K>(X, d);
        if (lane + d < 32) X = op_mul<K>(X, O);
    }
    if (lane == 0) op_store<K>(sWarp + w * OPN, X);
    __syncthreads();
    if (threadIdx.x == 0) {
        Op<K> T = op_load<K>(sWarp);
        for (int i = 1; i < NW; i++) T = op_mul<K>(T, op_load<K>(sWarp + i * OPN));
        op_store<K>(rangeops + (int64_t)blockIdx.x * OPN, T);
    }
}

// carries[t] = { A_t[K], lfT, B_t[K], lbT } from the carries of the range the tile belongs to
template <int K>
__global__ void __launch_bounds__(64) fb_range_apply_kernel(const double *__restrict__ tileops, int64_t nT, int q,
                                                            const double *__restrict__ rangecarry,
                                                            double *__restrict__ carries) {
    constexpr int OPN = K * K + 1, CN = 2 * K + 2;
    extern __shared__ double sm[];
    double *sOps = sm;              // q * OPN
    double *sOut = sm + q * OPN;    // q * CN
    const int64_t t0 = (int64_t)blockIdx.x * q;
    const int n = (int)max((int64_t)0, min((int64_t)q, nT - t0));
    for (int i = threadIdx.x; i < n * OPN; i += 64) sOps[i] = tileops[t0 * OPN + i];
    __syncthreads();
    const double *rc = rangecarry + (int64_t)blockIdx.x * CN;
    if (threadIdx.x == 0) {  // forward, sequential over the range
        Vec<K> v;
#pragma unroll
        for (int k = 0; k < K; k++) v.v[k] = rc[k];
        v.ls = rc[K];
        for (int t = 0; t < n; t++) {
#pragma unroll
            for (int k = 0; k < K; k++) sOut[t * CN + k] = v.v[k];
            sOut[t * CN + K] = v.ls;
            v = vec_mul_op<K>(v, op_load<K>(sOps + t * OPN));
        }
    } else if (threadIdx.x == 32) {  // backward, on another warp
        Vec<K> u;
#pragma unroll
        for (int k = 0; k < K; k++) u.v[k] = rc[K + 1 + k];
        u.ls = rc[2 * K + 1];
        for (int t = n - 1; t >= 0; t--) {
#pragma unroll
            for (int k = 0; k < K; k++) sOut[t * CN + K + 1 + k] = u.v[k];
            sOut[t * CN + 2 * K + 1] = u.ls;
            u = op_mul_vec<K>(op_load<K>(sOps + t * OPN), u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * CN; i += 64) carries[t0 * CN + i] = sOut[i];
}

// what the EM-internal form of the kernel writes instead of the four log-domain datasets
struct LeanOut {
    double *prob1;     // M x K posteriors P1 (linear)
    int *counts;       // K = 2: weights below the rejection cut per (state, warp tile of RC_WT bins), or nullptr
    double pcut;       // the cut
    int64_t nwt;       // warp tiles in the block
};

// posterior of one state where the linear product al * be / Z left the normal range (or an operand of it
// did): exp of the log-domain value the full kernel stores, so that "is it exactly 0" -- which decides
// whether the rejection control draws a uniform for it -- has the reference's meaning
__device__ __noinline__ double post_slow(double al, double be, double Z, double a, double f, double dscale) {
    const double lf = not_tiny(al) ? log(al) : (log(a) + f + dscale);
    return exp(fmin((lf + log(be)) - log(Z), 0.0));
}

// ---- kernel C: tile -> all outputs + statistics -------------------------------------------------
// LEAN (the E-step inside the EM loop): nothing there reads logFP, logBP or logProb2 -- maxStepProb, Q and
// BIC come from the fused statistics, the rejection control and innerMaxStepProb want exp(logProb1)
// (R/base-EM.R:113-158, 191-225) -- so only the posteriors P1 are written, linear, and for K = 2 the
// rejection control's counting pass is done on the way.  The datasets are produced by the full form of the
// kernel when somebody asks for them (fb_materialize).
template <int K, int NT, int L, bool LEAN>
__global__ void __launch_bounds__(NT, FBCfg<K>::APPLY_CTAS)
    fb_apply_kernel(const double *__restrict__ logf, int64_t M, int first_is_start, const __grid_constant__ FBParams P,
                    const double *__restrict__ carries, double *__restrict__ logFP, double *__restrict__ logBP,
                    double *__restrict__ logProb1, double *__restrict__ logProb2, double *__restrict__ tilestats,
                    double *__restrict__ row0, const LeanOut lean) {
    constexpr int TB = NT * L, NW = NT / 32, OPN = K * K + 1, NS = K * K + 1;
    extern __shared__ double smem[];
    double2 *sT = reinterpret_cast<double2 *>(smem);  // 128 entries of the log table (not in the lean form: 4 CTAs fit then)
    double *sM = smem + (LEAN ? 0 : 256);  // TB    per-bin max
    double *sA = sM + TB;          // K*TB  e_j, then alpha~_j (post-emission, chunk scale)
    double *sB = sA + K * TB;      // K*TB  beta~_j
    double *sScF = sM;             // the forward pass replaces m_j by the tile-local log scale of alpha~_j
    double *sU = sB + K * TB;      // NT    per chunk: log scale of beta~ behind it + forward scale at its end
    double *sRow = sU + NT;        // (TB/32)*K  log alpha~ of the bin before every 32-bin row
    double *sWarp = sRow + (TB / 32) * K;  // NW*OPN, later NW*NS for the statistics
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t s0 = (int64_t)blockIdx.x * TB;
    const int nvalid = (int)min((int64_t)TB, M - s0);
    const bool special0 = first_is_start && blockIdx.x == 0;
    const double *cr = carries + (int64_t)blockIdx.x * (2 * K + 2);

    __shared__ int sCnt[L * K];
    if (!LEAN && tid < 128) sT[tid] = kLogTab[tid];
    if (LEAN && tid < L * K) sCnt[tid] = 0;
    load_tile<K, NT, L, false>(logf, M, s0, nvalid, nullptr, sM, sA);
    __syncthreads();
    {
        Op<K> X = chunk_product<K, L, TB>(sA, sM, tid, nvalid, special0, P);
        Op<K> Pi = X, Si = X;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            Op<K> O = op_shfl_up<K>(Pi, d);
            if (lane >= d) Pi = op_mul<K>(O, Pi);
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            Op<K> O = op_shfl_down<K>(Si, d);
            if (lane + d < 32) Si = op_mul<K>(Si, O);
        }
        if (lane == 31) op_store<K>(sWarp + w * OPN, Pi);
        Op<K> Pe = op_shfl_up<K>(Pi, 1), Se = op_shfl_down<K>(Si, 1);
        if (lane == 0) op_identity(Pe);
        if (lane == 31) op_identity(Se);
        __syncthreads();
        Vec<K> va, ub;
#pragma unroll
        for (int k = 0; k < K; k++) { va.v[k] = cr[k]; ub.v[k] = cr[K + 1 + k]; }
        va.ls = 0.0;
        ub.ls = 0.0;
        for (int i = 0; i < w; i++) va = vec_mul_op<K>(va, op_load<K>(sWarp + i * OPN));
        va = vec_mul_op<K>(va, Pe);
        for (int i = NW - 1; i > w; i--) ub = op_mul_vec<K>(op_load<K>(sWarp + i * OPN), ub);
        ub = op_mul_vec<K>(Se, ub);

        // backward pass over the chunk (reads e_j from sA before the forward pass overwrites it)
        {
            double u[K], s = ub.ls;
#pragma unroll
            for (int k = 0; k < K; k++) u[k] = ub.v[k];
#pragma unroll
            for (int i = L - 1; i >= 0; i--) {
                int b = tid * L + i;
                if (b < nvalid) {
                    double wv[K];
#pragma unroll
                    for (int k = 0; k < K; k++) { sB[k * TB + b] = u[k]; wv[k] = sA[k * TB + b] * u[k]; }
#pragma unroll
                    for (int r = 0; r < K; r++) {
                        double acc = P.g[r * K] * wv[0];
#pragma unroll
                        for (int c = 1; c < K; c++) acc = fma(P.g[r * K + c], wv[c], acc);
                        u[r] = acc;
                    }
                    s += sM[b];
                }
            }
        }
        // forward pass over the chunk (overwrites e_j with alpha~_j)
        {
            double v[K], s = va.ls;
#pragma unroll
            for (int k = 0; k < K; k++) v[k] = va.v[k];
#pragma unroll
            for (int i = 0; i < L; i++) {
                int b = tid * L + i;
                if (b < nvalid) {
                    double a[K];
                    if (special0 && b == 0) {
#pragma unroll
                        for (int k = 0; k < K; k++) a[k] = v[k];
                    } else {
#pragma unroll
                        for (int c = 0; c < K; c++) {
                            double acc = v[0] * P.g[c];
#pragma unroll
                            for (int r = 1; r < K; r++) acc = fma(v[r], P.g[r * K + c], acc);
                            a[c] = acc;
                        }
                    }
                    s += sM[b];
#pragma unroll
                    for (int k = 0; k < K; k++) { v[k] = a[k] * sA[k * TB + b]; sA[k * TB + b] = v[k]; }
                    sScF[b] = s;
                }
            }
            // scale of beta~_b = ub.ls + (sum of m over the later bins of the chunk) = sU - sScF[b]
            sU[tid] = ub.ls + s;
        }
    }
    __syncthreads();

    // ---- epilogue: striped over bins, everything coalesced, no block-level synchronisation ----
    const double lfT = cr[K], lbT = cr[2 * K + 1];
    double acc[NS];
#pragma unroll
    for (int i = 0; i < NS; i++) acc[i] = 0.0;
    const double lKK = -log((double)(K * K));
    // log alpha~ of tile-local bin bb in the scale sScF[bb] from the log-domain recursion: used where
    // the linear value underflowed (an emission below exp(-667) relative to the bin's largest)
    auto slow_lf = [&](int bb, int k) -> double {
        double ap[K], scp = 0.0, av;
        if (bb > 0) {
#pragma unroll
            for (int i = 0; i < K; i++) ap[i] = sA[i * TB + bb - 1];
            scp = sScF[bb - 1];
        } else {
#pragma unroll
            for (int i = 0; i < K; i++) ap[i] = cr[i];
        }
        if (special0 && bb == 0) {
            av = ap[k];
        } else {
            av = ap[0] * P.g[k];
#pragma unroll
            for (int i = 1; i < K; i++) av = fma(ap[i], P.g[i * K + k], av);
        }
        return log(av) + __ldg(logf + (int64_t)k * M + s0 + bb) + (scp - sScF[bb]);
    };
    if constexpr (LEAN) {
        const bool count = (K == 2) && lean.counts != nullptr;
#pragma unroll 1
        for (int r = 0; r < L; r++) {
            const int b = r * NT + tid;
            const bool valid = b < nvalid;
            const int64_t j = s0 + b;
            const bool isfirst = special0 && b == 0;
            double al[K], be[K], f[K], alp[K], a[K], p1[K];
            if (valid) {
#pragma unroll
                for (int k = 0; k < K; k++) { al[k] = sA[k * TB + b]; be[k] = sB[k * TB + b]; f[k] = __ldg(logf + (int64_t)k * M + j); }
                if (b > 0) {
#pragma unroll
                    for (int k = 0; k < K; k++) alp[k] = sA[k * TB + b - 1];
                } else {
#pragma unroll
                    for (int k = 0; k < K; k++) alp[k] = cr[k];
                }
                if (isfirst) {
#pragma unroll
                    for (int k = 0; k < K; k++) a[k] = alp[k];
                } else {
#pragma unroll
                    for (int c = 0; c < K; c++) {
                        double t = alp[0] * P.g[c];
#pragma unroll
                        for (int i = 1; i < K; i++) t = fma(alp[i], P.g[i * K + c], t);
                        a[c] = t;
                    }
                }
                double Z = 0.0;
#pragma unroll
                for (int k = 0; k < K; k++) Z = fma(al[k], be[k], Z);
                bool fast = normal_pos(Z);
#pragma unroll
                for (int k = 0; k < K; k++) fast = fast && not_tiny(al[k]) && normal_pos(be[k]);
                const double rZ = rcp_fast(Z);
#pragma unroll
                for (int k = 0; k < K; k++) {
                    p1[k] = fmin(al[k] * be[k] * rZ, 1.0);
                    fast = fast && normal_pos(p1[k]);
                }
                double x[K];
                if (fast) {
#pragma unroll
                    for (int k = 0; k < K; k++) x[k] = p1[k];
                } else {
                    const double d = (b > 0 ? sScF[b - 1] : 0.0) - sScF[b];
#pragma unroll
                    for (int k = 0; k < K; k++) x[k] = post_slow(al[k], be[k], Z, a[k], f[k], d);
                }
#pragma unroll
                for (int k = 0; k < K; k++) {
                    lean.prob1[(int64_t)k * M + j] = x[k];
                    acc[K * K] = fma(p1[k], f[k], acc[K * K]);
                }
                if (blockIdx.x == 0 && b == 0) {  // logProb1[0,] for maxStepProb's pi and Q (the block's first bin)
                    const double d = -sScF[0];
                    const double lZ = log(Z);
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        const double lf = not_tiny(al[k]) ? log(al[k]) : (log(a[k]) + f[k] + d);
                        row0[k] = fmin((lf + log(be[k])) - lZ, 0.0);
                    }
                }
                if (!isfirst) {
#pragma unroll
                    for (int l = 0; l < K; l++) {
                        const double ra = p1[l] * rcp_fast(a[l]);
#pragma unroll
                        for (int i = 0; i < K; i++) acc[i * K + l] = fma(alp[i] * P.g[i * K + l], ra, acc[i * K + l]);
                    }
                }
                if (count) {
#pragma unroll
                    for (int k = 0; k < K; k++) p1[k] = x[k];
                }
            }
            if (count) {
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const int n = __popc(__ballot_sync(0xffffffffu, valid && rc_flagged(p1[k], lean.pcut)));
                    if (lane == 0 && n) atomicAdd(&sCnt[r * K + k], n);
                }
            }
        }
        if (count) {
            __syncthreads();
            if (tid < L * K) {  // row r of the tile is warp tile L * blockIdx.x + r of the rejection control
                const int r = tid / K, k = tid - r * K;
                const int64_t wt = (int64_t)blockIdx.x * L + r;
                if (wt < lean.nwt) lean.counts[(int64_t)k * lean.nwt + wt] = sCnt[tid];
            }
        }
    } else {
    // lane 0 of every 32-bin row needs log alpha~ of the bin before the row (another warp's bin)
    for (int idx = tid; idx < (TB / 32) * K; idx += NT) {
        const int row = idx / K, k = idx - row * K, bb = 32 * row - 1;
        double v = 0.0;
        if (row == 0) {
            v = log(fmax(cr[k], 4.9406564584124654e-324));
        } else if (bb < nvalid) {
            const double x = sA[k * TB + bb];
            v = not_tiny(x) ? log_norm(x, sT) : slow_lf(bb, k);
        }
        sRow[idx] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int r = 0; r < L; r++) {
        const int b = r * NT + tid;
        const bool valid = b < nvalid;
        const int64_t j = s0 + b;
        const bool isfirst = special0 && b == 0;
        double al[K], be[K], f[K], alp[K], lf[K], lb[K], lZ;
        double scf = 0, scfp = 0, scb = 0;
        if (valid) {
#pragma unroll
            for (int k = 0; k < K; k++) { al[k] = sA[k * TB + b]; be[k] = sB[k * TB + b]; f[k] = __ldg(logf + (int64_t)k * M + j); }
            scf = sScF[b];
            scb = sU[b / L] - scf;
            if (b > 0) {
#pragma unroll
                for (int k = 0; k < K; k++) alp[k] = sA[k * TB + b - 1];
                scfp = sScF[b - 1];
            } else {
#pragma unroll
                for (int k = 0; k < K; k++) alp[k] = cr[k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < K; k++) { al[k] = 1.0; be[k] = 1.0; f[k] = 0.0; alp[k] = 1.0; }
        }
        // pre-emission forward vector a = alpha~_{j-1} Gamma (in the scale of bin j-1)
        double a[K];
        if (isfirst) {
#pragma unroll
            for (int k = 0; k < K; k++) a[k] = alp[k];
        } else {
#pragma unroll
            for (int c = 0; c < K; c++) {
                double s = alp[0] * P.g[c];
#pragma unroll
                for (int i = 1; i < K; i++) s = fma(alp[i], P.g[i * K + c], s);
                a[c] = s;
            }
        }
        double Z = 0.0;
#pragma unroll
        for (int k = 0; k < K; k++) Z = fma(al[k], be[k], Z);
        // one test for the whole bin: every logarithm has a positive normal argument (the rule)
        bool fast = normal_pos(Z);
#pragma unroll
        for (int k = 0; k < K; k++) fast = fast && not_tiny(al[k]) && normal_pos(be[k]);
        if (fast) {
#pragma unroll
            for (int k = 0; k < K; k++) { lf[k] = log_norm(al[k], sT); lb[k] = log_norm(be[k], sT); }
            lZ = log_norm(Z, sT);
        } else {
#pragma unroll
            for (int k = 0; k < K; k++) {
                lf[k] = not_tiny(al[k]) ? log(al[k]) : (log(a[k]) + f[k] + (scfp - scf));
                lb[k] = log(be[k]);
            }
            lZ = log(Z);
        }
        // log alpha~ of the previous bin: the neighbouring lane has just computed it
        double lfp[K];
#pragma unroll
        for (int k = 0; k < K; k++) lfp[k] = __shfl_up_sync(0xffffffffu, lf[k], 1);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < K; k++) lfp[k] = sRow[(b >> 5) * K + k];
        }
        if (valid) {
            const double rZ = rcp_fast(Z);
            double lp1[K], p1[K];
#pragma unroll
            for (int k = 0; k < K; k++) {
                lp1[k] = fmin((lf[k] + lb[k]) - lZ, 0.0);  // a log-probability; > 0 only by rounding
                p1[k] = al[k] * be[k] * rZ;
                logFP[(int64_t)k * M + j] = lfT + (scf + lf[k]);
                logBP[(int64_t)k * M + j] = lbT + (scb + lb[k]);
                logProb1[(int64_t)k * M + j] = lp1[k];
                acc[K * K] = fma(p1[k], f[k], acc[K * K]);
            }
            if (blockIdx.x == 0 && b == 0) {
#pragma unroll
                for (int k = 0; k < K; k++) row0[k] = lp1[k];
            }
            if (isfirst) {
#pragma unroll
                for (int c = 0; c < K * K; c++) logProb2[(int64_t)c * M + j] = lKK;
            } else {
                const double d = scfp - scf;
#pragma unroll
                for (int l = 0; l < K; l++) {
                    const double ra = p1[l] * rcp_fast(a[l]);
#pragma unroll
                    for (int i = 0; i < K; i++) {
                        // (not clamped: may exceed 0 by the rounding of the O(1e4) scale terms, ~1e-12)
                        logProb2[(int64_t)(i * K + l) * M + j] =
                            lp1[l] + ((((lfp[i] + d) + P.lg[i * K + l]) + f[l]) - lf[l]);
                        acc[i * K + l] = fma(alp[i] * P.g[i * K + l], ra, acc[i * K + l]);
                    }
                }
            }
        }
    }
    }  // !LEAN
    // ---- block reduction of the statistics (fixed order -> deterministic) ----
#pragma unroll
    for (int i = 0; i < NS; i++) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc[i] += __shfl_down_sync(0xffffffffu, acc[i], d);
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NS; i++) sWarp[w * NS + i] = acc[i];
    }
    __syncthreads();
    if (tid < NS) {
        double s = sWarp[tid];
        for (int i = 1; i < NW; i++) s += sWarp[i * NS + tid];
        tilestats[(int64_t)tid * gridDim.x + blockIdx.x] = s;  // [statistic][tile]: coalesced for fb_stats_kernel
    }
}

// First bin of every tile but the first: its joint posterior needs log alpha of the previous bin,
// which the apply kernel only has as the linear carry.  Where that carry underflowed (an emission
// below exp(-690) in the previous tile's last bin) the entry is rebuilt from the log-domain logFP.
template <int K>
__global__ void fb_boundary_fix_kernel(int64_t M, int64_t nT, int TB, const __grid_constant__ FBParams P,
                                       const double *__restrict__ carries, const double *__restrict__ logf,
                                       const double *__restrict__ logFP, const double *__restrict__ logProb1,
                                       double *__restrict__ logProb2) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (t >= nT) return;
    const double *cr = carries + t * (2 * K + 2);
    const int64_t j = t * TB;
#pragma unroll
    for (int i = 0; i < K; i++) {
        if (!not_tiny(cr[i])) {
#pragma unroll
            for (int l = 0; l < K; l++)
                logProb2[(int64_t)(i * K + l) * M + j] =
                    logProb1[(int64_t)l * M + j] +
                    (((logFP[(int64_t)i * M + j - 1] - logFP[(int64_t)l * M + j]) + P.lg[i * K + l]) + logf[(int64_t)l * M + j]);
        }
    }
}

// deterministic sum of the per-tile statistics: out[i] = sum_t tilestats[i*nT + t]; one CTA per statistic
__global__ void __launch_bounds__(256) fb_stats_kernel(const double *__restrict__ tilestats, int64_t nT, int NS,
                                                       double *__restrict__ out) {
    __shared__ double sh[256];
    const int i = blockIdx.x;
    double s = 0.0;
    for (int64_t t = threadIdx.x; t < nT; t += 256) s += tilestats[(int64_t)i * nT + t];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[i] = sh[0];
    (void)NS;
}

// statistics straight from the datasets (used when logProb1/logProb2 were stored from the host,
// i.e. maxStepProb / computeQFunction called on their own inputs as the reference allows)
template <int K>
__global__ void __launch_bounds__(256)
    fb_stats_from_datasets_kernel(int64_t M, int skip_first, const double *__restrict__ logf, const double *__restrict__ lp1,
                                  const double *__restrict__ lp2, double *__restrict__ part) {
    constexpr int NS = K * K + 1;
    __shared__ double sh[8 * NS];
    double acc[NS];
#pragma unroll
    for (int i = 0; i < NS; i++) acc[i] = 0.0;
    for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < M; j += (int64_t)gridDim.x * 256) {
        if (j >= skip_first) {  // the chain's first bin has no transition into it; a later block's first bin has
#pragma unroll
            for (int c = 0; c < K * K; c++) acc[c] += exp(__ldg(lp2 + (int64_t)c * M + j));
        }
        if (logf) {
#pragma unroll
            for (int k = 0; k < K; k++)
                acc[K * K] = fma(exp(__ldg(lp1 + (int64_t)k * M + j)), __ldg(logf + (int64_t)k * M + j), acc[K * K]);
        }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NS; i++) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc[i] += __shfl_down_sync(0xffffffffu, acc[i], d);
        if (lane == 0) sh[w * NS + i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < NS) {
        double t = sh[threadIdx.x];
        for (int i = 1; i < 8; i++) t += sh[i * NS + threadIdx.x];
        part[(int64_t)threadIdx.x * gridDim.x + blockIdx.x] = t;
    }
}

template <int K> size_t reduce_smem() {
    using C = FBCfg<K>;
    return sizeof(double) * ((size_t)(K + 1) * C::NT * C::L + (C::NT / 32) * (K * K + 1));
}
template <int K> size_t apply_smem(bool lean = false) {
    using C = FBCfg<K>;
    size_t TB = (size_t)C::NT * C::L;
    return sizeof(double) * ((lean ? 0 : 256) + (2 * K + 1) * TB + C::NT + (TB / 32) * K + (C::NT / 32) * (K * K + 1));
}

void fill_params(const ehmm_session *s, FBParams &P) {
    const int K = s->K;
    for (int i = 0; i < K; i++)
        for (int l = 0; l < K; l++) {
            double g = s->h_gamma[i + l * K];
            P.g[i * K + l] = g;
            P.lg[i * K + l] = log(g);
        }
}

// scal layout (doubles): [0] ll | [1 .. OPN] block operator | [OPN+1 .. OPN+1+K] fwd vector after block
// | [OPN+2+K .. OPN+2+2K] bwd vector in front of block | [OPN+3+2K ..] statistics (NS) | logProb1[0,] (K)
template <int K> constexpr int scal_stats_off() { return (K * K + 1) + 3 + 2 * K; }
template <int K> constexpr int scal_row0_off() { return scal_stats_off<K>() + K * K + 1; }
template <int K> constexpr int scal_len() { return scal_row0_off<K>() + K; }

template <int K> int64_t num_tiles(int64_t M) {
    constexpr int64_t TB = (int64_t)FBCfg<K>::NT * FBCfg<K>::L;
    return (M + TB - 1) / TB;
}

constexpr int MAX_RANGES = 592;  // 4 per SM
constexpr int MAX_Q = 1024;      // tiles per range staged in shared memory

// R ranges of q tiles each
inline void range_split(int64_t nT, int &R, int &q) {
    int64_t r = nT < 296 ? nT : 296;
    int64_t qq = (nT + r - 1) / r;
    if (qq > MAX_Q) { qq = MAX_Q; }
    r = (nT + qq - 1) / qq;
    R = (int)r;
    q = (int)qq;
}

template <int K> int ensure_scratch(ehmm_session *s) {
    const int64_t nT = num_tiles<K>(s->M);
    EHMM_TRY(s->tileops.reserve(sizeof(double) * nT * (K * K + 1)));
    EHMM_TRY(s->carries.reserve(sizeof(double) * nT * (2 * K + 2)));
    EHMM_TRY(s->tilestats.reserve(sizeof(double) * nT * (K * K + 1)));
    EHMM_TRY(s->fbscal.reserve(sizeof(double) * scal_len<K>()));
    EHMM_TRY(s->rangeops.reserve(sizeof(double) * MAX_RANGES * (K * K + 1)));
    EHMM_TRY(s->rangecarry.reserve(sizeof(double) * MAX_RANGES * (2 * K + 2)));
    static bool attr_done[EHMM_MAX_K + 1] = {false};
    if (!attr_done[K]) {
        using C = FBCfg<K>;
        EHMM_CK(cudaFuncSetAttribute(fb_reduce_kernel<K, C::NT, C::L>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)reduce_smem<K>()));
        EHMM_CK(cudaFuncSetAttribute(fb_apply_kernel<K, C::NT, C::L, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)apply_smem<K>()));
        EHMM_CK(cudaFuncSetAttribute(fb_apply_kernel<K, C::NT, C::L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)apply_smem<K>()));
        // the range kernels stage up to MAX_Q tile operators in shared memory
        EHMM_CK(cudaFuncSetAttribute(fb_range_reduce_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(sizeof(double) * MAX_Q * (K * K + 1))));
        EHMM_CK(cudaFuncSetAttribute(fb_range_apply_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(sizeof(double) * MAX_Q * (K * K + 1 + 2 * K + 2))));
        attr_done[K] = true;
    }
    return EHMM_OK;
}

template <int K> int launch_reduce(ehmm_session *s, const double *d_logf, int mode0_total) {
    using C = FBCfg<K>;
    FBParams P;
    fill_params(s, P);
    const int64_t nT = num_tiles<K>(s->M);
    const int first = (s->m0 == 0) ? 1 : 0;
    PROF(s, KID_FB_REDUCE), fb_reduce_kernel<K, C::NT, C::L><<<(unsigned)nT, C::NT, reduce_smem<K>(), s->stream>>>(
        d_logf, s->M, first, P, s->tileops.as<double>());
    EHMM_CK(cudaGetLastError());
    int R, q;
    range_split(nT, R, q);
    EHMM_REQUIRE(R <= MAX_RANGES, EHMM_ERR_ARG, "expStep: %lld tiles exceed the carry scan's capacity", (long long)nT);
    PROF(s, KID_FB_CARRY), fb_range_reduce_kernel<K><<<R, RANGE_NT, sizeof(double) * q * (K * K + 1), s->stream>>>(
        s->tileops.as<double>(), nT, q, s->rangeops.as<double>());
    EHMM_CK(cudaGetLastError());
    if (mode0_total) {
        CarryIn cin = {};
        PROF(s, KID_FB_CARRY), fb_carry_kernel<K><<<1, CarryCfg<K>::NT, 0, s->stream>>>(s->rangeops.as<double>(), R, 0, cin, nullptr,
                                                         s->fbscal.as<double>());
        EHMM_CK(cudaGetLastError());
    }
    return EHMM_OK;
}

template <int K> int launch_apply(ehmm_session *s, const double *d_logf, const CarryIn &cin, bool lean) {
    using C = FBCfg<K>;
    FBParams P;
    fill_params(s, P);
    const int64_t nT = num_tiles<K>(s->M);
    const int first = (s->m0 == 0) ? 1 : 0;
    double *scal = s->fbscal.as<double>();
    int R, q;
    range_split(nT, R, q);
    PROF(s, KID_FB_CARRY), fb_carry_kernel<K><<<1, CarryCfg<K>::NT, 0, s->stream>>>(s->rangeops.as<double>(), R, 1, cin,
                                                                          s->rangecarry.as<double>(), scal);
    EHMM_CK(cudaGetLastError());
    PROF(s, KID_FB_CARRY), fb_range_apply_kernel<K><<<R, 64, sizeof(double) * q * (K * K + 1 + 2 * K + 2), s->stream>>>(
        s->tileops.as<double>(), nT, q, s->rangecarry.as<double>(), s->carries.as<double>());
    EHMM_CK(cudaGetLastError());
    double *row0 = scal + scal_row0_off<K>();
    if (lean) {
        EHMM_TRY(s->prob1.reserve(sizeof(double) * s->M * K));
        LeanOut lo = {s->prob1.as<double>(), nullptr, 0.0, rc_num_wtiles(s->M)};
        s->fb_counted = false;
        // K = 2: the tile's rows are the rejection control's warp tiles, so its counting pass rides along when
        // the cut of the coming call is known (ehmm_rc_hint, or the previous call's)
        if (K == 2 && C::NT == RC_WT && s->rc_hint_p > 0.0 && s->rc_hint_p <= 1.0) {
            EHMM_TRY(s->rc_flag.reserve(sizeof(int) * lo.nwt * K));
            lo.counts = s->rc_flag.as<int>();
            lo.pcut = s->rc_hint_p;
            s->fb_counted = true;
        }
        PROF(s, KID_FB_APPLY_LEAN), fb_apply_kernel<K, C::NT, C::L, true><<<(unsigned)nT, C::NT, apply_smem<K>(true), s->stream>>>(
            d_logf, s->M, first, P, s->carries.as<double>(), nullptr, nullptr, nullptr, nullptr, s->tilestats.as<double>(), row0, lo);
        EHMM_CK(cudaGetLastError());
    } else {
        EHMM_TRY(s->logFP.reserve(sizeof(double) * s->M * K));
        EHMM_TRY(s->logBP.reserve(sizeof(double) * s->M * K));
        EHMM_TRY(s->logProb1.reserve(sizeof(double) * s->M * K));
        EHMM_TRY(s->logProb2.reserve(sizeof(double) * s->M * K * K));
        const LeanOut lo = {nullptr, nullptr, 0.0, 0};
        PROF(s, KID_FB_APPLY), fb_apply_kernel<K, C::NT, C::L, false><<<(unsigned)nT, C::NT, apply_smem<K>(), s->stream>>>(
            d_logf, s->M, first, P, s->carries.as<double>(), s->logFP.as<double>(), s->logBP.as<double>(),
            s->logProb1.as<double>(), s->logProb2.as<double>(), s->tilestats.as<double>(), row0, lo);
        EHMM_CK(cudaGetLastError());
        if (nT > 1) {
            PROF(s, KID_FB_FIX), fb_boundary_fix_kernel<K><<<(unsigned)((nT - 1 + 255) / 256), 256, 0, s->stream>>>(
                s->M, nT, C::NT * C::L, P, s->carries.as<double>(), d_logf, s->logFP.as<double>(), s->logProb1.as<double>(),
                s->logProb2.as<double>());
            EHMM_CK(cudaGetLastError());
        }
    }
    PROF(s, KID_FB_STATS), fb_stats_kernel<<<K * K + 1, 256, 0, s->stream>>>(s->tilestats.as<double>(), nT, K * K + 1, scal + scal_stats_off<K>());
    EHMM_CK(cudaGetLastError());
    // what fb_materialize needs to produce the datasets of this E-step later
    s->estep_logf = d_logf;
    for (int k = 0; k <= K; k++) { s->estep_cin[0][k] = cin.start[k]; s->estep_cin[1][k] = cin.end[k]; }
    s->lazy_full = lean;
    return EHMM_OK;
}

template <int K> int expstep_impl(ehmm_session *s, const double *d_logf) {
    EHMM_TRY(ensure_scratch<K>(s));
    EHMM_TRY(launch_reduce<K>(s, d_logf, 0));
    CarryIn cin = {};  // start vector (pi, 0), end vector (1, 0)
    for (int k = 0; k < K; k++) { cin.start[k] = s->h_pi[k]; cin.end[k] = 1.0; }
    EHMM_TRY(launch_apply<K>(s, d_logf, cin, s->lean_mode));
    return EHMM_OK;
}

// the four log-domain datasets of the last (lean) E-step: same logf, same carries, same kernel as a full E-step
template <int K> int materialize_impl(ehmm_session *s) {
    CarryIn cin = {};
    for (int k = 0; k <= K; k++) { cin.start[k] = s->estep_cin[0][k]; cin.end[k] = s->estep_cin[1][k]; }
    EHMM_TRY(launch_apply<K>(s, s->estep_logf, cin, false));
    return EHMM_OK;
}

template <int K> int reduce_impl(ehmm_session *s, const double *d_logf, double *op_host) {
    EHMM_TRY(ensure_scratch<K>(s));
    EHMM_TRY(launch_reduce<K>(s, d_logf, 1));
    EHMM_CK(cudaMemcpyAsync(s->h_pinned, s->fbscal.as<double>() + 1, sizeof(double) * (K * K + 1),
                            cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    memcpy(op_host, s->h_pinned, sizeof(double) * (K * K + 1));
    return EHMM_OK;
}

template <int K> int apply_impl(ehmm_session *s, const double *d_logf, const double *fwd, const double *bwd,
                                double *fwd_out) {
    double *scal = s->fbscal.as<double>();
    CarryIn cin = {};
    for (int k = 0; k <= K; k++) { cin.start[k] = fwd[k]; cin.end[k] = bwd[k]; }
    EHMM_TRY(launch_apply<K>(s, d_logf, cin, s->lean_mode));
    if (fwd_out) {
        EHMM_CK(cudaMemcpyAsync(s->h_pinned + 16, scal + (K * K + 1) + 1, sizeof(double) * (K + 1),
                                cudaMemcpyDeviceToHost, s->stream));
        EHMM_CK(cudaStreamSynchronize(s->stream));
        memcpy(fwd_out, s->h_pinned + 16, sizeof(double) * (K + 1));
    }
    return EHMM_OK;
}

template <int K> int fetch_stats_impl(ehmm_session *s) {
    constexpr int NS = K * K + 1;
    double *hp = s->h_pinned + 32;
    if (s->stats_stale) {
        const int nblk = 148 * 4;
        EHMM_TRY(s->fbscal.reserve(sizeof(double) * scal_len<K>()));
        EHMM_TRY(s->tilestats.reserve(sizeof(double) * nblk * NS));
        const bool p12 = has(s, DS_P1 | DS_P2);
        if (p12) {
            PROF(s, KID_MISC), fb_stats_from_datasets_kernel<K><<<nblk, 256, 0, s->stream>>>(
                s->M, s->m0 == 0 ? 1 : 0, has(s, DS_LOGF) ? s->logf.as<double>() : nullptr, s->logProb1.as<double>(),
                s->logProb2.as<double>(), s->tilestats.as<double>());
            EHMM_CK(cudaGetLastError());
            PROF(s, KID_FB_STATS), fb_stats_kernel<<<NS, 256, 0, s->stream>>>(s->tilestats.as<double>(), nblk, NS,
                                                     s->fbscal.as<double>() + scal_stats_off<K>());
            EHMM_CK(cudaGetLastError());
            EHMM_CK(cudaMemcpyAsync(hp + 1, s->fbscal.as<double>() + scal_stats_off<K>(), sizeof(double) * NS,
                                    cudaMemcpyDeviceToHost, s->stream));
            for (int k = 0; k < K; k++)
                EHMM_CK(cudaMemcpyAsync(hp + 1 + NS + k, s->logProb1.as<double>() + (int64_t)k * s->M, sizeof(double),
                                        cudaMemcpyDeviceToHost, s->stream));
        }
        if (has(s, DS_FP))
            for (int k = 0; k < K; k++)
                EHMM_CK(cudaMemcpyAsync(hp + 1 + NS + K + k, s->logFP.as<double>() + (int64_t)k * s->M + (s->M - 1),
                                        sizeof(double), cudaMemcpyDeviceToHost, s->stream));
        EHMM_CK(cudaStreamSynchronize(s->stream));
        if (p12) {
            for (int i = 0; i < NS; i++) s->h_stats[i] = hp[1 + i];
            for (int k = 0; k < K; k++) s->h_lp1_row0[k] = hp[1 + NS + k];
        }
        if (has(s, DS_FP)) {  // lse(logFP[M-1,]) as src/computeBIC.cpp:23-26
            const double *d = hp + 1 + NS + K;
            double C = d[0], sum = 0;
            for (int k = 1; k < K; k++) C = d[k] > C ? d[k] : C;
            for (int k = 0; k < K; k++) sum += exp(d[k] - C);
            s->h_ll = C + log(sum);
        }
        s->stats_on_host = true;
        return EHMM_OK;
    }
    double *scal = s->fbscal.as<double>();
    EHMM_CK(cudaMemcpyAsync(hp, scal, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    // statistics and logProb1[0,] sit next to each other
    EHMM_CK(cudaMemcpyAsync(hp + 1, scal + scal_stats_off<K>(), sizeof(double) * (K * K + 1 + K), cudaMemcpyDeviceToHost,
                            s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    s->h_ll = hp[0];
    for (int i = 0; i < K * K + 1; i++) s->h_stats[i] = hp[1 + i];
    for (int k = 0; k < K; k++) s->h_lp1_row0[k] = hp[1 + K * K + 1 + k];
    s->stats_on_host = true;
    return EHMM_OK;
}

}  // namespace

int fb_expstep(ehmm_session *s, const double *d_logf) {
    s->stats_on_host = false;
    if (s->K == 2) return expstep_impl<2>(s, d_logf);
    if (s->K == 3) return expstep_impl<3>(s, d_logf);
    set_error("expStep: K=%d unsupported (K must be 2 or 3)", s->K);
    return EHMM_ERR_ARG;
}

int fb_reduce(ehmm_session *s, const double *d_logf, double *op_host) {
    s->stats_on_host = false;
    s->cur_logf = d_logf;
    if (s->K == 2) return reduce_impl<2>(s, d_logf, op_host);
    if (s->K == 3) return reduce_impl<3>(s, d_logf, op_host);
    set_error("expStep: K=%d unsupported (K must be 2 or 3)", s->K);
    return EHMM_ERR_ARG;
}

int fb_apply(ehmm_session *s, const double *fwd_carry, const double *bwd_carry, double *fwd_out) {
    const double *d_logf = s->cur_logf;
    if (s->K == 2) return apply_impl<2>(s, d_logf, fwd_carry, bwd_carry, fwd_out);
    if (s->K == 3) return apply_impl<3>(s, d_logf, fwd_carry, bwd_carry, fwd_out);
    set_error("expStep: K=%d unsupported (K must be 2 or 3)", s->K);
    return EHMM_ERR_ARG;
}

int fb_materialize(ehmm_session *s) {
    if (!s->lazy_full) return EHMM_OK;
    EHMM_REQUIRE(s->estep_logf != nullptr, EHMM_ERR_STATE, "the E-step's emission is no longer available");
    if (s->K == 2) EHMM_TRY(materialize_impl<2>(s));
    else EHMM_TRY(materialize_impl<3>(s));
    s->valid |= DS_ESTEP;
    return EHMM_OK;
}

int fb_fetch_stats(ehmm_session *s) {
    if (s->stats_on_host) return EHMM_OK;
    if (s->K == 2) return fetch_stats_impl<2>(s);
    if (s->K == 3) return fetch_stats_impl<3>(s);
    return EHMM_ERR_ARG;
}

}  // namespace ehmm
