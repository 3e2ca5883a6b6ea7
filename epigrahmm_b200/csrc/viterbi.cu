// viterbi.cu -- computeViterbiSequence (src/computeViterbiSequence.cpp:12-54).
//
// The reference scores each bin with logFP (the forward log-probabilities), so LOGV grows like M^2
// (1e15 at 15 M bins) and ties are decided at ulp granularity: the (max,+) recursion
//     V[j+1,k] = fl( max_i fl(V[j,i] + log gamma[i,k]) + logFP[j+1,k] )
// is not re-associable in floating point.  It IS re-associable where it is integer arithmetic:
// while every LOGV value and every candidate stays inside one binade [2^e, 2^(e+1)) all of them are
// integer multiples of u = 2^(e-52), fl(n u + x) = (n + rint(x / u)) u unless x / u lies exactly
// half-way between two integers, and max is exact.  With a_ik = rint(log gamma[i,k] / u) and
// e_jk = rint(logFP[j,k] / u) the recursion is n[j+1,k] = max_i(n[j,i] + a_ik) + e_{j+1,k}: a
// product of K x K integer matrices over (max,+), which is associative and exact in 64-bit integers.
//
//   vit_predict_kernel : per tile of 512 bins, sum_j max_k logFP[j,k] and the largest spread between
//                        states (bounds on LOGV: all log gamma <= 0, all logFP <= 0)
//   vit_plan_kernel    : prefix of those sums -> for every tile either the binade e that all its LOGV
//                        values provably share ("clean") or "break" (a power of two may be crossed)
//   vit_scan_kernel    : one pass with decoupled look-back.  A clean tile without half-way cases
//                        publishes the (max,+) product of its bins, obtains LOGV in front of it from
//                        its predecessors, and replays its bins in integers (back-pointers, range
//                        check of every value and candidate).  A break tile (binade crossing, a
//                        half-way case, the chain's first bin) waits for LOGV in front of it and runs
//                        the reference's fp64 operation order on one thread.
//   vit_bt_*_kernel    : the backtrace S[j] = psi[j+1][S[j+1]] as a suffix scan of maps {0..K-1} -> {0..K-1}
//
// Whatever the integer model cannot represent (a value outside the predicted binade, non-finite or
// positive logFP, mixed binades behind one look-back) raises a flag and the whole chain is re-run by
// the sequential kernel below, which executes the reference's loop operation by operation
// (__dadd_rn, no FMA, first-index argmax) on one thread.  The reference's backtrace re-evaluates
// argmax_i(V[j,i] + log gamma[i,S[j+1]]), the same expression as the forward step's argmax, so
// storing psi[j+1][k] = argmax_i(...) (2 bits per state) is equivalent.
#include "common.cuh"
#include <limits.h>

namespace ehmm {

namespace {

typedef long long i64;

constexpr int VT_NT = 256;
constexpr int VT_TILE = 2048;  // bins per shared-memory tile (sequential kernel, backtrace)

struct VitPar {
    double lpi[EHMM_MAX_K];
    double lg[EHMM_MAX_K * EHMM_MAX_K];  // lg[i*K+k] = log gamma(i,k)
    double vin[EHMM_MAX_K];              // LOGV of the bin in front of this block (bin blocks only)
    int has_vin;                         // 0: this block starts the chain
};

// ---- the reference's loop on one thread (fallback; also the oracle of the parallel path's tests) ----
template <int K>
__device__ __forceinline__ unsigned vit_step_fp(double (&V)[K], const double *lg, const double (&E)[K]) {
    double mc[K];
    unsigned code = 0;
#pragma unroll
    for (int k = 0; k < K; k++) {
        double mx = __dadd_rn(V[0], lg[k]);
        unsigned arg = 0;
#pragma unroll
        for (int i = 1; i < K; i++) {
            const double c = __dadd_rn(V[i], lg[i * K + k]);
            if (c > mx) { mx = c; arg = i; }
        }
        mc[k] = mx;
        code |= arg << (2 * k);
    }
#pragma unroll
    for (int k = 0; k < K; k++) V[k] = __dadd_rn(mc[k], E[k]);
    return code;
}

template <int K>
__global__ void __launch_bounds__(VT_NT)
    viterbi_forward_kernel(const double *__restrict__ logFP, int64_t M, VitPar P, uint8_t *__restrict__ bp,
                           double *__restrict__ vlast) {
    extern __shared__ double sm[];
    double *tile[2] = {sm, sm + K * VT_TILE};
    uint8_t *sbp = reinterpret_cast<uint8_t *>(sm + 2 * K * VT_TILE);
    const int tid = threadIdx.x;
    const int64_t nT = (M + VT_TILE - 1) / VT_TILE;
    double V[K];
#pragma unroll
    for (int k = 0; k < K; k++) V[k] = P.vin[k];
    {
        const int n0 = (int)min((int64_t)VT_TILE, M);
        for (int i = tid; i < n0 * K; i += VT_NT) {
            int k = i / n0, b = i - k * n0;
            tile[0][k * VT_TILE + b] = logFP[(int64_t)k * M + b];
        }
    }
    __syncthreads();
    for (int64_t t = 0; t < nT; t++) {
        const int64_t s0 = t * VT_TILE;
        const int n = (int)min((int64_t)VT_TILE, M - s0);
        const double *cur = tile[t & 1];
        double *nxt = tile[(t + 1) & 1];
        if (tid == 0) {
            for (int b = 0; b < n; b++) {
                double E[K];
#pragma unroll
                for (int k = 0; k < K; k++) E[k] = cur[k * VT_TILE + b];
                if (s0 + b == 0 && !P.has_vin) {
#pragma unroll
                    for (int k = 0; k < K; k++) V[k] = __dadd_rn(P.lpi[k], E[k]);
                    sbp[0] = 0;
                } else {
                    sbp[b] = (uint8_t)vit_step_fp<K>(V, P.lg, E);
                }
            }
        } else if (t + 1 < nT) {
            // the other threads stream the next tile in while thread 0 walks this one
            const int64_t s1 = s0 + VT_TILE;
            const int n1 = (int)min((int64_t)VT_TILE, M - s1);
            for (int i = tid - 1; i < n1 * K; i += VT_NT - 1) {
                int k = i / n1, b = i - k * n1;
                nxt[k * VT_TILE + b] = logFP[(int64_t)k * M + s1 + b];
            }
        }
        __syncthreads();
        for (int b = tid; b < n; b += VT_NT) bp[s0 + b] = sbp[b];
        __syncthreads();
    }
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) vlast[k] = V[k];
    }
}

// ---- parallel forward pass ------------------------------------------------------------------------
constexpr int VS_NT = 128, VS_L = 4, VS_TB = VS_NT * VS_L, VS_NW = VS_NT / 32;
constexpr int VS_BREAK = INT_MIN;
constexpr i64 VS_NEG = -(1LL << 60);  // "minus infinity" of the integer (max,+) matrices
constexpr i64 TWO52 = 1LL << 52, TWO53 = 1LL << 53;
enum { VS_ERR_RANGE = 1, VS_ERR_INPUT = 2, VS_ERR_MIXED = 4 };

// shared-memory index of tile-local bin b: a thread's 4 consecutive bins sit 5 words from the next thread's
__device__ __forceinline__ int vs_idx(int b) { return b + (b >> 2); }
constexpr int VS_ROW = VS_TB + VS_TB / 4;

template <int K>
__global__ void __launch_bounds__(VS_NT) vit_predict_kernel(const double *__restrict__ logFP, int64_t M,
                                                            double *__restrict__ tsum, double *__restrict__ tspread,
                                                            int *__restrict__ err) {
    __shared__ double sh[2 * VS_NW];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t s0 = (int64_t)blockIdx.x * VS_TB;
    double sum = 0.0, spread = 0.0;
    bool bad = false;
#pragma unroll
    for (int r = 0; r < VS_L; r++) {
        const int64_t j = s0 + r * VS_NT + tid;
        if (j < M) {
            double mx = __ldg(logFP + j), mn = mx;
            bad = bad || !(mx <= 0.0 && mx > -INFINITY);
#pragma unroll
            for (int k = 1; k < K; k++) {
                const double v = __ldg(logFP + (int64_t)k * M + j);
                bad = bad || !(v <= 0.0 && v > -INFINITY);
                mx = fmax(mx, v);
                mn = fmin(mn, v);
            }
            sum += mx;
            spread = fmax(spread, mx - mn);
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, d);
        spread = fmax(spread, __shfl_down_sync(0xffffffffu, spread, d));
    }
    if (lane == 0) { sh[w] = sum; sh[VS_NW + w] = spread; }
    if (bad) atomicOr(err, VS_ERR_INPUT);
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i < VS_NW; i++) { sum += sh[i]; spread = fmax(spread, sh[VS_NW + i]); }
        tsum[blockIdx.x] = sum;
        tspread[blockIdx.x] = spread;
    }
}

// plan[t] = binade exponent e when every LOGV value and candidate of tile t (and the vector entering
// it) provably lies in [2^e, 2^(e+1)) and log gamma / u has no half-way case; VS_BREAK otherwise.
// Bounds (all terms <= 0): vmax + P_excl >= V >= vmax + P_incl + (bins so far + 2) * min log gamma - spread,
// P = prefix sums of max_k logFP, vmax = the largest entry of the vector in front of the block.
__global__ void __launch_bounds__(1024) vit_plan_kernel(const double *__restrict__ tsum, const double *__restrict__ tspread,
                                                        int64_t nT, int64_t M, VitPar P, int K, int *__restrict__ plan,
                                                        int *__restrict__ info) {
    __shared__ double sh[1024];
    __shared__ int sbreak;
    const int tid = threadIdx.x;
    const int64_t q = (nT + 1023) / 1024;
    const int64_t t0 = min(nT, (int64_t)tid * q), t1 = min(nT, t0 + q);
    double loc = 0.0;
    for (int64_t t = t0; t < t1; t++) loc += tsum[t];
    sh[tid] = loc;
    if (tid == 0) sbreak = 0;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const double v = (tid >= d) ? sh[tid - d] : 0.0;
        __syncthreads();
        sh[tid] += v;
        __syncthreads();
    }
    double run = sh[tid] - loc;  // exclusive prefix of this thread's first tile
    double vmax = P.has_vin ? P.vin[0] : P.lpi[0], vmin = vmax, lgabs = 0.0;
    bool lgok = true;
    for (int k = 1; k < K; k++) {
        vmax = fmax(vmax, P.has_vin ? P.vin[k] : P.lpi[k]);
        vmin = fmin(vmin, P.has_vin ? P.vin[k] : P.lpi[k]);
    }
    for (int i = 0; i < K * K; i++) {
        lgabs = fmax(lgabs, fabs(P.lg[i]));
        lgok = lgok && (P.lg[i] <= 0.0) && (P.lg[i] > -INFINITY);
    }
    lgok = lgok && (vmax <= 0.0) && (vmax > -INFINITY);
    int nb = 0;
    for (int64_t t = t0; t < t1; t++) {
        const double excl = run, incl = run + tsum[t];
        run = incl;
        const double lo = -(vmax + excl) * (1.0 - 1e-9);
        const int64_t jend = min((t + 1) * VS_TB, M) + 2;
        // the vector entering the tile belongs to the previous tile's last bin (or is the block's input)
        const double spread = fmax(tspread[t], t > 0 ? tspread[t - 1] : vmax - vmin);
        const double hi = (-(vmax + incl) + (double)jend * lgabs + spread) * (1.0 + 1e-9);
        int e = VS_BREAK;
        if (lgok && lo > 0.0 && hi < 1.0e300 && !(t == 0 && !P.has_vin)) {
            const int ee = ilogb(lo);
            if (ee >= -900 && ee <= 1000 && hi < ldexp(1.0, ee + 1)) {
                const double sc = ldexp(1.0, 52 - ee);
                bool ok = true;
                for (int i = 0; i < K * K; i++) {
                    const double x = P.lg[i] * sc;
                    ok = ok && (fabs(x) < 9007199254740992.0) && (fabs(x - rint(x)) != 0.5);
                }
                if (ok) e = ee;
            }
        }
        plan[t] = e;
        nb += (e == VS_BREAK);
    }
    if (nb) atomicAdd(&sbreak, nb);
    __syncthreads();
    if (tid == 0) info[1] = sbreak;
}

template <int K> struct TM { i64 m[K * K]; };

template <int K> __device__ __forceinline__ void tm_identity(TM<K> &X) {
#pragma unroll
    for (int i = 0; i < K; i++)
#pragma unroll
        for (int l = 0; l < K; l++) X.m[i * K + l] = (i == l) ? 0 : VS_NEG;
}
// A first, then B
template <int K> __device__ __forceinline__ TM<K> tm_mul(const TM<K> &A, const TM<K> &B) {
    TM<K> C;
#pragma unroll
    for (int r = 0; r < K; r++)
#pragma unroll
        for (int c = 0; c < K; c++) {
            i64 mx = A.m[r * K] + B.m[c];
#pragma unroll
            for (int q = 1; q < K; q++) mx = max(mx, A.m[r * K + q] + B.m[q * K + c]);
            C.m[r * K + c] = mx;
        }
    return C;
}
template <int K> __device__ __forceinline__ void tv_mul(i64 (&v)[K], const TM<K> &B) {
    i64 r[K];
#pragma unroll
    for (int c = 0; c < K; c++) {
        i64 mx = v[0] + B.m[c];
#pragma unroll
        for (int q = 1; q < K; q++) mx = max(mx, v[q] + B.m[q * K + c]);
        r[c] = mx;
    }
#pragma unroll
    for (int c = 0; c < K; c++) v[c] = r[c];
}
template <int K> __device__ __forceinline__ TM<K> tm_shfl_up(const TM<K> &X, int d) {
    TM<K> R;
#pragma unroll
    for (int i = 0; i < K * K; i++) R.m[i] = __shfl_up_sync(0xffffffffu, X.m[i], d);
    return R;
}

// LOGV (doubles) -> integers in units of 2^(e-52); false unless every entry is an integer in (-2^53, -2^52]
template <int K> __device__ __forceinline__ bool vs_to_int(const double *V, int e, i64 (&n)[K]) {
    const double sc = ldexp(1.0, 52 - e);
    bool ok = true;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const double x = V[k] * sc;
        ok = ok && (x == rint(x)) && (x <= -4503599627370496.0) && (x > -9007199254740992.0);
        n[k] = ok ? (i64)x : -TWO52;
    }
    return ok;
}

struct VSBuf {
    const int *plan;
    int *status;        // 0: nothing yet, 1: aggregate published, 2: LOGV behind the tile published
    i64 *agg;           // nT x K*K
    double *incl;       // nT x K
    unsigned *ticket;
    int *err;
    int *info;          // [0] path, [1] break tiles planned, [2] tiles that found a half-way case
};

template <int K>
__global__ void __launch_bounds__(VS_NT)
    vit_scan_kernel(const double *__restrict__ logFP, int64_t M, int64_t nT, VitPar P, VSBuf B, uint8_t *__restrict__ bp,
                    double *__restrict__ vlast) {
    __shared__ double sE[K * VS_ROW];
    __shared__ uint8_t sbp[VS_TB];
    __shared__ i64 sWarp[VS_NW * K * K];
    __shared__ i64 sAgg[K * K];
    __shared__ double sV[K];
    __shared__ int sTile;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) sTile = (int)atomicAdd(B.ticket, 1u);
    __syncthreads();
    const int64_t t = sTile;
    if (t >= nT) return;
    const int64_t s0 = t * VS_TB;
    const int n = (int)min((int64_t)VS_TB, M - s0);
#pragma unroll
    for (int k = 0; k < K; k++)
        for (int b = tid; b < n; b += VS_NT) sE[k * VS_ROW + vs_idx(b)] = __ldg(logFP + (int64_t)k * M + s0 + b);
    __syncthreads();
    const int e = B.plan[t];
    bool clean = e != VS_BREAK;
    i64 ev[VS_L][K];
    i64 A[K * K];
    if (clean) {
        const double sc = ldexp(1.0, 52 - e);
        bool tie = false, big = false;
#pragma unroll
        for (int i = 0; i < VS_L; i++) {
            const int b = tid * VS_L + i;
#pragma unroll
            for (int k = 0; k < K; k++) {
                const double x = (b < n) ? sE[k * VS_ROW + vs_idx(b)] * sc : 0.0;
                const double r = rint(x);
                tie = tie || (fabs(x - r) == 0.5);
                const bool in = fabs(x) < 9007199254740992.0;
                big = big || !in;
                ev[i][k] = in ? (i64)r : 0;
            }
        }
#pragma unroll
        for (int i = 0; i < K * K; i++) A[i] = (i64)rint(P.lg[i] * sc);
        if (big) atomicOr(B.err, VS_ERR_RANGE);
        const int anytie = __syncthreads_or(tie ? 1 : 0);
        if (anytie) {
            clean = false;
            if (tid == 0) atomicAdd(B.info + 2, 1);
        }
    }
    TM<K> Pe;  // product of the bins of the threads in front of this one (clean tiles)
    if (clean) {
        TM<K> X;
        bool has = false;
#pragma unroll
        for (int i = 0; i < VS_L; i++) {
            if (tid * VS_L + i < n) {
                if (!has) {
#pragma unroll
                    for (int r = 0; r < K; r++)
#pragma unroll
                        for (int c = 0; c < K; c++) X.m[r * K + c] = A[r * K + c] + ev[i][c];
                    has = true;
                } else {
                    TM<K> T;
#pragma unroll
                    for (int r = 0; r < K; r++)
#pragma unroll
                        for (int c = 0; c < K; c++) {
                            i64 mx = X.m[r * K] + A[c];
#pragma unroll
                            for (int q = 1; q < K; q++) mx = max(mx, X.m[r * K + q] + A[q * K + c]);
                            T.m[r * K + c] = mx + ev[i][c];
                        }
                    X = T;
                }
            }
        }
        if (!has) tm_identity(X);
        TM<K> Pi = X;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            TM<K> O = tm_shfl_up<K>(Pi, d);
            if (lane >= d) Pi = tm_mul<K>(O, Pi);
        }
        if (lane == 31) {
#pragma unroll
            for (int i = 0; i < K * K; i++) sWarp[w * K * K + i] = Pi.m[i];
        }
        Pe = tm_shfl_up<K>(Pi, 1);
        if (lane == 0) tm_identity(Pe);
        __syncthreads();
        TM<K> W;
        tm_identity(W);
        for (int i = 0; i < w; i++) {
            TM<K> Y;
#pragma unroll
            for (int c = 0; c < K * K; c++) Y.m[c] = sWarp[i * K * K + c];
            W = tm_mul<K>(W, Y);
        }
        Pe = tm_mul<K>(W, Pe);
        if (tid == 0) {
            TM<K> T;
#pragma unroll
            for (int c = 0; c < K * K; c++) T.m[c] = sWarp[c];
            for (int i = 1; i < VS_NW; i++) {
                TM<K> Y;
#pragma unroll
                for (int c = 0; c < K * K; c++) Y.m[c] = sWarp[i * K * K + c];
                T = tm_mul<K>(T, Y);
            }
#pragma unroll
            for (int c = 0; c < K * K; c++) {
                sAgg[c] = T.m[c];
                B.agg[t * K * K + c] = T.m[c];
            }
            __threadfence();
            atomicExch(B.status + t, 1);
        }
    }
    // ---- LOGV in front of the tile: look back over published aggregates to the nearest published vector ----
    if (tid == 0) {
        double base[K];
        TM<K> acc;
        bool has_acc = false;
        int acc_e = 0;
        int64_t p = t - 1;
        for (;;) {
            if (p < 0) {
#pragma unroll
                for (int k = 0; k < K; k++) base[k] = P.vin[k];  // (unused by the chain's first bin)
                break;
            }
            int st;
            while ((st = *reinterpret_cast<volatile int *>(B.status + p)) == 0) __nanosleep(40);
            __threadfence();
            if (st == 2) {
#pragma unroll
                for (int k = 0; k < K; k++) base[k] = *reinterpret_cast<volatile double *>(B.incl + p * K + k);
                break;
            }
            TM<K> Y;
#pragma unroll
            for (int c = 0; c < K * K; c++) Y.m[c] = *reinterpret_cast<volatile i64 *>(B.agg + p * K * K + c);
            const int ep = B.plan[p];
            if (has_acc) {
                if (ep != acc_e) atomicOr(B.err, VS_ERR_MIXED);
                acc = tm_mul<K>(Y, acc);
            } else {
                acc = Y;
                acc_e = ep;
                has_acc = true;
            }
            p--;
        }
        if (has_acc) {
            i64 nb[K];
            if (!vs_to_int<K>(base, acc_e, nb)) atomicOr(B.err, VS_ERR_RANGE);
            tv_mul<K>(nb, acc);
            const double us = ldexp(1.0, acc_e - 52);
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (!(nb[k] <= -TWO52 && nb[k] > -TWO53)) atomicOr(B.err, VS_ERR_RANGE);
                base[k] = (double)nb[k] * us;
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) sV[k] = base[k];
        if (clean) {  // LOGV behind the tile goes out before the bins are replayed
            i64 nb[K];
            if (!vs_to_int<K>(base, e, nb)) atomicOr(B.err, VS_ERR_RANGE);
            TM<K> T;
#pragma unroll
            for (int c = 0; c < K * K; c++) T.m[c] = sAgg[c];
            tv_mul<K>(nb, T);
            const double us = ldexp(1.0, e - 52);
#pragma unroll
            for (int k = 0; k < K; k++) {
                if (!(nb[k] <= -TWO52 && nb[k] > -TWO53)) atomicOr(B.err, VS_ERR_RANGE);
                const double v = (double)nb[k] * us;
                B.incl[t * K + k] = v;
                if (t == nT - 1) vlast[k] = v;
            }
            __threadfence();
            atomicExch(B.status + t, 2);
        } else {
            double V[K];
#pragma unroll
            for (int k = 0; k < K; k++) V[k] = base[k];
            for (int b = 0; b < n; b++) {
                double E[K];
#pragma unroll
                for (int k = 0; k < K; k++) E[k] = sE[k * VS_ROW + vs_idx(b)];
                if (s0 + b == 0 && !P.has_vin) {
#pragma unroll
                    for (int k = 0; k < K; k++) V[k] = __dadd_rn(P.lpi[k], E[k]);
                    sbp[0] = 0;
                } else {
                    sbp[b] = (uint8_t)vit_step_fp<K>(V, P.lg, E);
                }
            }
#pragma unroll
            for (int k = 0; k < K; k++) {
                B.incl[t * K + k] = V[k];
                if (t == nT - 1) vlast[k] = V[k];
            }
            __threadfence();
            atomicExch(B.status + t, 2);
        }
    }
    __syncthreads();
    if (clean) {
        i64 v[K];
        bool ok = vs_to_int<K>(sV, e, v);
        tv_mul<K>(v, Pe);
        const i64 lo = -TWO52;  // every value and every candidate must lie in (-2^53, -2^52]
#pragma unroll
        for (int i = 0; i < VS_L; i++) {
            const int b = tid * VS_L + i;
            if (b < n) {
                i64 nv[K];
                unsigned code = 0;
#pragma unroll
                for (int k = 0; k < K; k++) {
                    i64 mx = v[0] + A[k];
                    unsigned arg = 0;
                    i64 cmin = mx;
#pragma unroll
                    for (int q = 1; q < K; q++) {
                        const i64 c = v[q] + A[q * K + k];
                        cmin = min(cmin, c);
                        if (c > mx) { mx = c; arg = q; }
                    }
                    code |= arg << (2 * k);
                    nv[k] = mx + ev[i][k];
                    // all candidates and the new value inside (-2^53, -2^52]
                    ok = ok && (mx <= lo) && (cmin > -TWO53) && (nv[k] <= lo) && (nv[k] > -TWO53);
                }
#pragma unroll
                for (int k = 0; k < K; k++) { ok = ok && (v[k] <= lo) && (v[k] > -TWO53); v[k] = nv[k]; }
                sbp[b] = (uint8_t)code;
            }
        }
        if (!ok) atomicOr(B.err, VS_ERR_RANGE);
    }
    __syncthreads();
    for (int b = tid; b < n; b += VS_NT) bp[s0 + b] = sbp[b];
}

// ---- backtrace: S[M-1] = argmax V[M-1,:] (first maximum), S[j] = psi[j+1][S[j+1]] ------------------
// A map {0..K-1} -> {0..K-1} is a byte holding 2 bits per argument (the back-pointer byte itself).
constexpr int BT_NT = 256, BT_L = 8, BT_TB = BT_NT * BT_L, BT_NW = BT_NT / 32;
constexpr unsigned BT_ID = 0x24;  // identity: 0 -> 0, 1 -> 1, 2 -> 2

__device__ __forceinline__ unsigned bt_apply(unsigned f, unsigned s) { return (f >> (2 * s)) & 3u; }
// g first, then f
__device__ __forceinline__ unsigned bt_compose(unsigned f, unsigned g) {
    return bt_apply(f, bt_apply(g, 0)) | (bt_apply(f, bt_apply(g, 1)) << 2) | (bt_apply(f, bt_apply(g, 2)) << 4);
}

// map of thread `tid`'s chunk: state of bin hi+1 (its upper neighbour) -> state of its first bin;
// psi of bin M (past the block's end) is the identity, so "the state of bin M" is the state of bin M-1
__device__ __forceinline__ unsigned bt_chunk_map(const uint8_t *__restrict__ bp, int64_t M, int64_t a) {
    unsigned f = BT_ID;
#pragma unroll
    for (int i = BT_L - 1; i >= 0; i--) {
        const int64_t j = a + i;  // S[j] = psi[j+1][S[j+1]]
        if (j < M) {
            const unsigned psi = (j + 1 < M) ? (unsigned)__ldg(bp + j + 1) : BT_ID;
            f = bt_compose(psi, f);
        }
    }
    return f;
}

__global__ void __launch_bounds__(BT_NT) vit_bt_reduce_kernel(const uint8_t *__restrict__ bp, int64_t M, uint8_t *__restrict__ tilemap) {
    __shared__ unsigned sh[BT_NW];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    unsigned f = bt_chunk_map(bp, M, (int64_t)blockIdx.x * BT_TB + (int64_t)tid * BT_L);
    // ordered composition: lower threads are applied last
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_down_sync(0xffffffffu, f, d);
        if (lane + d < 32) f = bt_compose(f, o);
    }
    if (lane == 0) sh[w] = f;
    __syncthreads();
    if (tid == 0) {
        unsigned g = sh[BT_NW - 1];
        for (int i = BT_NW - 2; i >= 0; i--) g = bt_compose(sh[i], g);
        tilemap[blockIdx.x] = (uint8_t)g;
    }
}

// entry[t] = state of the bin behind tile t (for the last tile: the state of bin M-1 itself)
__global__ void __launch_bounds__(1024) vit_bt_scan_kernel(const uint8_t *__restrict__ tilemap, int64_t nT, int K,
                                                           const double *__restrict__ vlast, int carry_in,
                                                           const uint8_t *__restrict__ bp, uint8_t *__restrict__ entry,
                                                           int *__restrict__ carry_out) {
    __shared__ unsigned sh[1024];
    __shared__ unsigned sEntry[1024];
    const int tid = threadIdx.x;
    const int64_t q = (nT + 1023) / 1024;
    const int64_t t0 = min(nT, (int64_t)tid * q), t1 = min(nT, t0 + q);
    unsigned f = BT_ID;
    for (int64_t t = t1 - 1; t >= t0; t--) f = bt_compose(tilemap[t], f);
    sh[tid] = f;
    __syncthreads();
    if (tid == 0) {
        int best = 0;
        for (int k = 1; k < K; k++)
            if (vlast[k] > vlast[best]) best = k;
        unsigned st = (carry_in >= 0) ? (unsigned)carry_in : (unsigned)best;
        for (int i = 1023; i >= 0; i--) {
            sEntry[i] = st;
            st = bt_apply(sh[i], st);
        }
        // st = S[0]; the state of the bin in front of the block is psi[0][S[0]]
        carry_out[0] = (int)bt_apply(bp[0], st);
    }
    __syncthreads();
    unsigned st = sEntry[tid];
    for (int64_t t = t1 - 1; t >= t0; t--) {
        entry[t] = (uint8_t)st;
        st = bt_apply(tilemap[t], st);
    }
}

__global__ void __launch_bounds__(BT_NT) vit_bt_apply_kernel(const uint8_t *__restrict__ bp, int64_t M,
                                                             const uint8_t *__restrict__ entry, double *__restrict__ states) {
    __shared__ unsigned sh[BT_NW];
    __shared__ uint8_t sst[BT_TB];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t s0 = (int64_t)blockIdx.x * BT_TB;
    const int64_t a = s0 + (int64_t)tid * BT_L;
    const unsigned f = bt_chunk_map(bp, M, a);
    // suffix composition of the chunks behind this thread (exclusive)
    unsigned g = f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_down_sync(0xffffffffu, g, d);
        if (lane + d < 32) g = bt_compose(g, o);
    }
    if (lane == 0) sh[w] = g;  // the whole warp
    unsigned behind = __shfl_down_sync(0xffffffffu, g, 1);
    if (lane == 31) behind = BT_ID;
    __syncthreads();
    unsigned st = entry[blockIdx.x];
    for (int i = BT_NW - 1; i > w; i--) st = bt_apply(sh[i], st);
    st = bt_apply(behind, st);  // state of the bin behind this thread's chunk
#pragma unroll
    for (int i = BT_L - 1; i >= 0; i--) {
        const int64_t j = a + i;
        if (j < M) {
            const unsigned psi = (j + 1 < M) ? (unsigned)__ldg(bp + j + 1) : BT_ID;
            st = bt_apply(psi, st);
            sst[tid * BT_L + i] = (uint8_t)st;
        }
    }
    __syncthreads();
    const int n = (int)min((int64_t)BT_TB, M - s0);
    for (int b = tid; b < n; b += BT_NT) states[s0 + b] = (double)sst[b];
}

// ---- host side -------------------------------------------------------------------------------------
struct VitScratch {
    double *tsum, *tspread, *incl, *vlast;
    i64 *agg;
    int *plan, *status, *err, *info, *carry;
    unsigned *ticket;
    uint8_t *tilemap, *entry;
};

// layout of s->vit_work; the small words (vlast, err, info, ticket, carry) first
VitScratch vit_scratch(ehmm_session *s, int K, int64_t nT, int64_t nBT) {
    char *p = s->vit_work.as<char>();
    VitScratch w;
    w.vlast = reinterpret_cast<double *>(p);            // 8 doubles
    w.err = reinterpret_cast<int *>(p + 64);            // 1
    w.info = w.err + 1;                                 // 3
    w.ticket = reinterpret_cast<unsigned *>(w.err + 4);  // 1
    w.carry = w.err + 5;                                // 1
    p += 128;
    w.tsum = reinterpret_cast<double *>(p); p += sizeof(double) * nT;
    w.tspread = reinterpret_cast<double *>(p); p += sizeof(double) * nT;
    w.incl = reinterpret_cast<double *>(p); p += sizeof(double) * nT * K;
    w.agg = reinterpret_cast<i64 *>(p); p += sizeof(i64) * nT * K * K;
    w.plan = reinterpret_cast<int *>(p); p += sizeof(int) * nT;
    w.status = reinterpret_cast<int *>(p); p += sizeof(int) * nT;
    w.tilemap = reinterpret_cast<uint8_t *>(p); p += (nBT + 7) / 8 * 8;
    w.entry = reinterpret_cast<uint8_t *>(p);
    return w;
}
size_t vit_scratch_bytes(int K, int64_t nT, int64_t nBT) {
    return 128 + sizeof(double) * nT * (2 + K) + sizeof(i64) * nT * K * K + sizeof(int) * nT * 2 + 2 * ((nBT + 7) / 8 * 8) + 64;
}

inline int64_t vs_tiles(int64_t M) { return (M + VS_TB - 1) / VS_TB; }
inline int64_t bt_tiles(int64_t M) { return (M + BT_TB - 1) / BT_TB; }

template <int K> int forward_seq(ehmm_session *s, const VitPar &P, double *vlast) {
    const size_t smem = sizeof(double) * 2 * K * VT_TILE + VT_TILE;
    static bool attr[EHMM_MAX_K + 1] = {false};
    if (!attr[K]) {
        EHMM_CK(cudaFuncSetAttribute(viterbi_forward_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr[K] = true;
    }
    PROF(s, KID_VIT_FWD), viterbi_forward_kernel<K><<<1, VT_NT, smem, s->stream>>>(s->logFP.as<double>(), s->M, P,
                                                                              s->vit_bp.as<uint8_t>(), vlast);
    EHMM_CK(cudaGetLastError());
    return EHMM_OK;
}

// forward pass; on return (the stream has been waited for) h_pinned[320..] holds LOGV of the block's last bin
template <int K> int forward(ehmm_session *s, const VitPar &P) {
    const int64_t M = s->M, nT = vs_tiles(M), nBT = bt_tiles(M);
    EHMM_TRY(s->viterbi.reserve(sizeof(double) * M));
    EHMM_TRY(s->vit_bp.reserve((size_t)M + 64));
    EHMM_TRY(s->vit_work.reserve(vit_scratch_bytes(K, nT, nBT)));
    const VitScratch w = vit_scratch(s, K, nT, nBT);
    int *h = reinterpret_cast<int *>(s->h_pinned + 330);
    s->vit_info[0] = 1;
    s->vit_info[1] = s->vit_info[2] = 0;
    s->vit_info[3] = nT;
    if (s->vit_mode != 1) {
        EHMM_CK(cudaMemsetAsync(w.err, 0, 64, s->stream));
        EHMM_CK(cudaMemsetAsync(w.status, 0, sizeof(int) * nT, s->stream));
        PROF(s, KID_VIT_FWD), vit_predict_kernel<K><<<(unsigned)nT, VS_NT, 0, s->stream>>>(s->logFP.as<double>(), M, w.tsum, w.tspread, w.err);
        EHMM_CK(cudaGetLastError());
        PROF(s, KID_VIT_FWD), vit_plan_kernel<<<1, 1024, 0, s->stream>>>(w.tsum, w.tspread, nT, M, P, K, w.plan, w.info);
        EHMM_CK(cudaGetLastError());
        const VSBuf B = {w.plan, w.status, w.agg, w.incl, w.ticket, w.err, w.info};
        PROF(s, KID_VIT_FWD), vit_scan_kernel<K><<<(unsigned)nT, VS_NT, 0, s->stream>>>(s->logFP.as<double>(), M, nT, P, B,
                                                                                 s->vit_bp.as<uint8_t>(), w.vlast);
        EHMM_CK(cudaGetLastError());
        EHMM_CK(cudaMemcpyAsync(h, w.err, sizeof(int) * 4, cudaMemcpyDeviceToHost, s->stream));
        EHMM_CK(cudaMemcpyAsync(s->h_pinned + 320, w.vlast, sizeof(double) * K, cudaMemcpyDeviceToHost, s->stream));
        EHMM_CK(cudaStreamSynchronize(s->stream));
        s->vit_info[1] = h[2];
        s->vit_info[2] = h[3];
        if (h[0] == 0) {
            s->vit_info[0] = 0;
            return EHMM_OK;
        }
        s->vit_info[0] = 2;  // the integer model did not hold somewhere: the sequential kernel decides
    }
    EHMM_TRY(forward_seq<K>(s, P, w.vlast));
    EHMM_CK(cudaMemcpyAsync(s->h_pinned + 320, w.vlast, sizeof(double) * K, cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    return EHMM_OK;
}

int backtrace(ehmm_session *s, int carry_in, int *carry_out, double *states_out) {
    const int64_t M = s->M, nT = vs_tiles(M), nBT = bt_tiles(M);
    const int K = s->K;
    const VitScratch w = vit_scratch(s, K, nT, nBT);
    const uint8_t *bp = s->vit_bp.as<uint8_t>();
    PROF(s, KID_VIT_BWD), vit_bt_reduce_kernel<<<(unsigned)nBT, BT_NT, 0, s->stream>>>(bp, M, w.tilemap);
    EHMM_CK(cudaGetLastError());
    PROF(s, KID_VIT_BWD), vit_bt_scan_kernel<<<1, 1024, 0, s->stream>>>(w.tilemap, nBT, K, w.vlast, carry_in, bp, w.entry, w.carry);
    EHMM_CK(cudaGetLastError());
    PROF(s, KID_VIT_BWD), vit_bt_apply_kernel<<<(unsigned)nBT, BT_NT, 0, s->stream>>>(bp, M, w.entry, s->viterbi.as<double>());
    EHMM_CK(cudaGetLastError());
    if (states_out)
        EHMM_CK(cudaMemcpyAsync(states_out, s->viterbi.p, sizeof(double) * M, cudaMemcpyDeviceToHost, s->stream));
    int *hc = reinterpret_cast<int *>(s->h_pinned + 340);
    EHMM_CK(cudaMemcpyAsync(hc, w.carry, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    if (carry_out) *carry_out = *hc;
    s->valid |= DS_VIT;
    return EHMM_OK;
}

void fill(VitPar &P, int K, const double *pi, const double *gamma, const double *vin) {
    // host libm log(), i.e. the very values the reference's log(pi.t()) / log(gamma.col(k)) produce
    for (int k = 0; k < K; k++) P.lpi[k] = log(pi[k]);
    for (int i = 0; i < K; i++)
        for (int k = 0; k < K; k++) P.lg[i * K + k] = log(gamma[i + k * K]);
    P.has_vin = vin != nullptr;
    for (int k = 0; k < K; k++) P.vin[k] = vin ? vin[k] : 0.0;
}

}  // namespace

int viterbi(ehmm_session *s, const double *pi, const double *gamma, double *states_out) {
    EHMM_REQUIRE(s->m0 == 0 && s->Mtot == s->M, EHMM_ERR_STATE,
                 "computeViterbiSequence on a bin block: use ehmm_viterbi_forward / ehmm_viterbi_backtrace");
    VitPar P = {};
    fill(P, s->K, pi, gamma, nullptr);
    if (s->K == 2) EHMM_TRY(forward<2>(s, P));
    else if (s->K == 3) EHMM_TRY(forward<3>(s, P));
    else { set_error("viterbi: K=%d unsupported", s->K); return EHMM_ERR_ARG; }
    return backtrace(s, -1, nullptr, states_out);
}

// bin-block form: LOGV enters a block from the block in front of it, so the blocks run one after the other;
// v_in = LOGV of the bin in front of this block (NULL for the first block), v_out = LOGV of its last bin
int viterbi_forward(ehmm_session *s, const double *pi, const double *gamma, const double *v_in, double *v_out) {
    EHMM_REQUIRE((s->m0 == 0) == (v_in == nullptr), EHMM_ERR_ARG, "viterbi_forward: v_in is required exactly for blocks with m0 > 0");
    VitPar P = {};
    fill(P, s->K, pi, gamma, v_in);
    if (s->K == 2) EHMM_TRY(forward<2>(s, P));
    else if (s->K == 3) EHMM_TRY(forward<3>(s, P));
    else { set_error("viterbi: K=%d unsupported", s->K); return EHMM_ERR_ARG; }
    for (int k = 0; k < s->K; k++) v_out[k] = s->h_pinned[320 + k];
    s->vit_forward_done = true;
    return EHMM_OK;
}

// state_last: state of this block's last bin as decided by the next block, -1 for the chain's end;
// state_prev: state of the bin in front of this block (to hand to the previous block)
int viterbi_backtrace(ehmm_session *s, int state_last, int *state_prev, double *states_out) {
    EHMM_REQUIRE(s->vit_forward_done, EHMM_ERR_STATE, "viterbi_backtrace without viterbi_forward");
    EHMM_REQUIRE(state_last >= -1 && state_last < s->K, EHMM_ERR_ARG, "bad state %d", state_last);
    return backtrace(s, state_last, state_prev, states_out);
}

}  // namespace ehmm
