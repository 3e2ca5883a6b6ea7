// rc.cu -- rejection-controlled reweighting and aggregation by group (replaces
// src/consensusRejectionControlled.cpp, src/differentialRejectionControlled.cpp, src/reweight.cpp,
// src/rbinomVectorized.cpp and src/aggregate.cpp, which loop sequentially over bins drawing from
// R's global Mersenne-Twister).
//
// The r-th bin (state-major, then ascending bin) whose posterior weight x satisfies 0 < x < p
// consumes the r-th uniform of the stream.  That order is recovered with a stream compaction whose
// unit is a WARP TILE of 256 consecutive bins (8 strips of 32), so ranks come from ballots alone --
// no shared memory, no block barrier -- and every warp walks ALL columns of its bins (the first
// version gave every column its own CTAs; the background column, which keeps a weight in almost
// every bin, then ran on 1/columns of the machine):
//   rc_count_kernel : flagged weights per (column, warp tile)
//   rc_scan_kernel  : one CTA per column, exclusive scan of its tile counts -> offsets, column total
//   rc_apply_kernel : ranks, Bernoulli thinning exactly as R's rbinom(1, x/p), and the sum of the
//                     surviving weights by group
//   rc_finalize_kernel : exact accumulators -> the fp64 tables the NB-GLM kernels read
//
// Sums by group are EXACT and therefore independent of the order in which the hardware performs
// them: a weight w in [2^-9, 2) is the integer w * 2^62, which is split into two 31-bit limbs that
// are added with native 64-bit integer atomics (REDG.E.ADD.64; 33 bits of headroom per limb, i.e.
// 8e9 addends).  The table entry is that integer rounded once to fp64: bit-identical from run to
// run, for any bin-block sharding (the limbs of all blocks are added before the rounding) and at
// most half an ulp from the true sum -- the reference's own in-order fp64 sum
// (src/aggregate.cpp:23-30) carries up to n roundings.  Every weight that survives a rejection cut
// p >= 2^-9 qualifies (it is p itself or a posterior >= p; the reference's rejectionCut is
// max(0.9^iter, probCut), probCut = 0.05 by default); for smaller cuts (and p = 0: no thinning,
// which controlEM() cannot produce) the kernel falls back to fp64 atomics, whose result depends on
// the order of the additions in the last bits.
// Uniforms come either from a caller-supplied buffer (parity harness: the oracle consumes the same
// buffer) or from R's MT19937 stream generated on the device from the session's .Random.seed.
#include "common.cuh"
#include "rcweights.cuh"
#include <algorithm>
#include <chrono>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

namespace ehmm {

namespace {

constexpr int RC_NT = 256;                  // threads per CTA
constexpr int RC_NW = RC_NT / 32;           // warps = warp tiles per CTA
constexpr int RC_TB = RC_NT * RC_STRIPS;    // bins per block of the exported reweight() helper
constexpr int RC_MAXCOL = EHMM_MAX_B + 2;   // 64
constexpr int64_t RC_UNIF_SLACK = 32 * RC_STRIPS;  // the apply kernel stages up to this many words past a column's last draw
constexpr double RC_EXACT_MIN_P = 0.001953125;  // 2^-9: smallest cut whose survivors are exact 62-bit fixed-point numbers

// ---- R's unif_rand() (Mersenne-Twister) ------------------------------------------------------
// MT_genrand() * 2^-32 followed by fixup() into (0,1)
__device__ __forceinline__ double mt_to_unif(uint32_t tempered) {
    const double i2_32m1 = 2.328306437080797e-10;
    double x = (double)tempered * 2.3283064365386963e-10;
    if (x <= 0.0) return 0.5 * i2_32m1;
    if ((1.0 - x) <= 0.0) return 1.0 - 0.5 * i2_32m1;
    return x;
}

// ---- weights -----------------------------------------------------------------------------------
// column c of the consensus model: exp(logProb1[:,c]); of the differential model: background,
// B mixture components P1[:,1]*eta[:,b], enrichment (src/differentialRejectionControlled.cpp:34-45)
struct RCSrc {
    const double *lp1;  // logProb1, or P1 itself when lin is set
    const double *eta;
    int64_t M;
    int B;         // 0 = consensus
    int columns;
    int lin;
};
struct RCBase {
    int64_t ubase[RC_MAXCOL];  // where the draws of each column start in the uniform buffer
};

// posteriors of bin j (exp once per bin; the count and the apply kernel must see the same bits, so
// both go through these two functions and nothing here can contract into an FMA)
template <bool DIFF> __device__ __forceinline__ void rc_posteriors(const RCSrc &S, int64_t j, double (&p1)[3]) {
    p1[0] = rc_post(__ldg(S.lp1 + j), S.lin);
    p1[1] = rc_post(__ldg(S.lp1 + S.M + j), S.lin);
    p1[2] = DIFF ? rc_post(__ldg(S.lp1 + 2 * S.M + j), S.lin) : 0.0;
}
template <bool DIFF> __device__ __forceinline__ double rc_weight(const RCSrc &S, const double (&p1)[3], int c, int64_t j) {
    if (!DIFF) return c == 0 ? p1[0] : p1[1];
    if (c == 0) return p1[0];
    if (c == S.B + 1) return p1[2];
    return rc_mix_weight(p1[1], __ldg(S.eta + (int64_t)(c - 1) * S.M + j));
}

// lane (c & 31) of the warp keeps the running count of column c (two registers cover 64 columns)
template <bool DIFF>
__global__ void __launch_bounds__(RC_NT) rc_count_kernel(RCSrc S, double p, int64_t nwt, int *__restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int64_t tile = (int64_t)blockIdx.x * RC_NW + (threadIdx.x >> 5);
    if (tile >= nwt) return;
    int cnt0 = 0, cnt1 = 0;
#pragma unroll 1
    for (int r = 0; r < RC_STRIPS; r++) {
        const int64_t j = tile * RC_WT + r * 32 + lane;
        const bool valid = j < S.M;
        double p1[3] = {2.0, 2.0, 2.0};  // 2.0: never below p <= 1
        if (valid) rc_posteriors<DIFF>(S, j, p1);
#pragma unroll 1
        for (int c = 0; c < S.columns; c++) {
            const double x = valid ? rc_weight<DIFF>(S, p1, c, j) : 2.0;
            const int n = __popc(__ballot_sync(0xffffffffu, rc_flagged(x, p)));
            if (lane == (c & 31)) { if (c < 32) cnt0 += n; else cnt1 += n; }
        }
    }
    if (lane < S.columns) counts[(int64_t)lane * nwt + tile] = cnt0;
    if (lane + 32 < S.columns) counts[(int64_t)(lane + 32) * nwt + tile] = cnt1;
}

// one CTA per column: exclusive scan of the column's n tile counts -> offsets (uint32: a column of one
// bin block holds < 2^32 bins), coltot[column] = the column's total.  Each thread owns 16 consecutive
// counts per round (warp-shuffle scan of the thread sums, then of the warp sums).  256 threads: a CTA of
// 1024 does not fit next to the side stream's resident CTAs and waited for them to drain (0.2 ms per call).
constexpr int SCAN_NT = 256;
__global__ void __launch_bounds__(SCAN_NT) rc_scan_kernel(const int *__restrict__ counts_all, int64_t n,
                                                          uint32_t *__restrict__ offsets_all, int64_t *__restrict__ coltot) {
    constexpr int IT = 16, NW = SCAN_NT / 32;
    __shared__ int64_t swarp[NW];
    __shared__ int64_t carry;
    const int *counts = counts_all + (int64_t)blockIdx.x * n;
    uint32_t *offsets = offsets_all + (int64_t)blockIdx.x * n;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += SCAN_NT * IT) {
        const int64_t i0 = base + (int64_t)threadIdx.x * IT;
        int v[IT];
        int64_t tsum = 0;
#pragma unroll
        for (int k = 0; k < IT; k++) {
            v[k] = (i0 + k < n) ? counts[i0 + k] : 0;
            tsum += v[k];
        }
        int64_t incl = tsum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) swarp[w] = incl;
        __syncthreads();
        int64_t before = 0, all = 0;  // the warps in front of this one; the whole round
#pragma unroll
        for (int i = 0; i < NW; i++) {
            const int64_t t = swarp[i];
            if (i < w) before += t;
            all += t;
        }
        int64_t run = carry + before + (incl - tsum);
#pragma unroll
        for (int k = 0; k < IT; k++) {
            if (i0 + k < n) offsets[i0 + k] = (uint32_t)run;
            run += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += all;
        __syncthreads();
    }
    if (threadIdx.x == 0) coltot[blockIdx.x] = carry;
}

// exclusive scan of n ints into int64 offsets; offsets[n] = total (the exported reweight() helper)
__global__ void __launch_bounds__(1024) rw_scan_kernel(const int *__restrict__ counts, int64_t n,
                                                       int64_t *__restrict__ offsets) {
    __shared__ int64_t sh[1024];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = (i < n) ? counts[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            const int64_t t = ((int)threadIdx.x >= d) ? sh[threadIdx.x - d] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < n) offsets[i] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n] = carry;
}

// R::rbinom(1, pp) by inversion given the uniform it would draw (nmath/rbinom.c, n*min(p,1-p) < 30).
// For n = 1 the inner loop ends at ix = 0 when u < q and at ix = 1 otherwise (u - q < q*(p/q) holds
// for every u unif_rand() can return), so no second uniform is ever consumed.
__device__ __forceinline__ double rbinom1(double pp, double u) {
    const double pm = fmin(pp, 1.0 - pp), q = 1.0 - pm;
    int ix = (u < q) ? 0 : 1;
    if (pp > 0.5) ix = 1 - ix;
    return (double)ix;
}

template <typename UT> __device__ __forceinline__ double get_unif(const UT *u, int64_t i);
template <> __device__ __forceinline__ double get_unif<double>(const double *u, int64_t i) { return __ldg(u + i); }
template <> __device__ __forceinline__ double get_unif<uint32_t>(const uint32_t *u, int64_t i) {
    return mt_to_unif(__ldg(u + i));
}

// exact accumulation: w * 2^62 (an integer for w in [2^-9, 2)) as two 31-bit limbs in 64-bit words.
// Layout of an accumulator set for C columns and G groups (cells = C * (G + 1) entries, 64-bit words):
//   [0, cells) low limbs | [cells, 2 cells) high limbs | RC_REP replicas of the low limbs of groups < H of
//   every column | the same for the high limbs        (H = min(RC_HOT, G + 1))
// * the two limbs of an entry live in separate arrays: atomics on one 32-byte sector serialise in L2
//   (3.4 clocks each), so adjacent limbs would double the cost of a frequent group
//   (tools/probes/red_probe.cu: 11.0 ms against 4.8 ms for 15 M adds over 200 zipf-distributed groups);
// * data.table numbers groups by first appearance, so the frequent ones have the small ids (ids <= 4096 take
//   82 % of the adds of BASELINE configs[1]; real tracks with all-zero offsets have a few hundred groups in
//   all): their adds are spread over RC_REP replicas chosen by the warp tile, and the finalisation adds
//   the replicas up -- integer additions, so the split changes nothing in the result.
constexpr int RC_HOT = 4096;
constexpr int RC_REP = 32;

struct RCAcc {
    unsigned long long *w;  // the accumulator set
    int64_t cells;          // C * (G + 1)
    int64_t hot;            // H
    int64_t hotcells;       // RC_REP * C * H
};
inline int64_t rc_hot_groups(int64_t G) { return std::min<int64_t>(RC_HOT, G + 1); }
inline size_t rc_acc_words(int columns, int64_t G) {
    return 2 * ((size_t)columns * (G + 1) + (size_t)RC_REP * columns * rc_hot_groups(G));
}
inline RCAcc rc_acc(unsigned long long *w, int columns, int64_t G) {
    return RCAcc{w, (int64_t)columns * (G + 1), rc_hot_groups(G), (int64_t)RC_REP * columns * rc_hot_groups(G)};
}

// entry (column c, group g), replica rep in [0, RC_REP)
__device__ __forceinline__ void limb_add(const RCAcc &A, int c, int64_t G, int g, int rep, int C, double w) {
    const unsigned long long v = __double2ull_rz(w * 4611686018427387904.0);
    unsigned long long *lo;
    int64_t hi_off;
    if (g < A.hot) {
        lo = A.w + 2 * A.cells + ((int64_t)rep * C + c) * A.hot + g;
        hi_off = A.hotcells;
    } else {
        lo = A.w + (int64_t)c * (G + 1) + g;
        hi_off = A.cells;
    }
    atomicAdd(lo, v & 0x7fffffffull);
    atomicAdd(lo + hi_off, v >> 31);
}

// nrep = 1 for the consensus model (only f[0..M-1] is read, Q6 in SURVEY.md); nrep = N for the
// differential model (repmat(w, N, 1) against the M*N group ids).
// EXACT: integer limbs (acc = limbs: lo[cells] then hi[cells]); else fp64 atomics (acc = buckets).
// SKIPST (differential, exact): the weights >= p of columns 0 and B+1 were summed once for this E-step
// (rc_static_kernel) and are already in the limbs; only the survivors of the thinning are added here.
// CMAX > 0: at most CMAX columns, whose next uniform index (warp-uniform) is held in registers and the
// column loop is unrolled; CMAX = 0: any number of columns, lane (c & 31) holds the state of column c.
struct RCCut {
    double p, rp;  // the cut and RN(1 / p)
};

template <typename UT> __device__ __forceinline__ double rc_thin(double x, const RCCut &cut, const UT *unif, int64_t at) {
    return rbinom1(rc_div_cut(x, cut.p, cut.rp), get_unif<UT>(unif, at)) * cut.p;
}

template <bool EXACT>
__device__ __forceinline__ void rc_scatter(const RCAcc &A, int c, int C, int64_t G, int rep, const int32_t *__restrict__ group,
                                           int64_t j, int64_t M, int nrep, double x) {
    for (int i = 0; i < nrep; i++) {
        const int g = __ldg(group + j + (int64_t)i * M);
        if (EXACT) limb_add(A, c, G, g, rep, C, x);
        else atomicAdd(reinterpret_cast<double *>(A.w) + (int64_t)c * (G + 1) + g, x);
    }
}

// strips per staging group: the uniforms a group of strips can consume (at most 32 per strip and column) are
// copied to shared memory with all loads in flight at once, 4 KB per warp whatever the column count
template <int CMAX, typename UT> struct RCStage {
    static constexpr int SG = (CMAX <= 2 ? 8 : CMAX <= 8 ? 4 : 2) / (sizeof(UT) == 8 ? 2 : 1);
    static constexpr int WORDS = CMAX * 32 * SG;  // per warp
};

// what a bin's weights are made of, as loaded (the next strip's are fetched while the current one is worked on).
// CHUNK: the launch handles columns [c0, c0 + CM) of a call with more columns than fit the registers (B = 30: 32
// columns in four launches of 8); et[c] then belongs to the launch's column c, else to mixture component c - 1.
template <int CM, bool CHUNK> struct RCRaw {
    double lp[3];                                        // logProb1[j, 0..2] (or P1 itself)
    double et[CHUNK ? CM : (CM > 2 ? CM - 2 : 1)];       // mixtureProb[j, b]
    int g0;                                              // consensus: the bin's one group id
};

template <bool DIFF, int CM, bool CHUNK>
__device__ __forceinline__ void rc_load_raw(const RCSrc &S, const int32_t *__restrict__ group, int64_t j, int C, int c0, int CT,
                                            RCRaw<CM, CHUNK> &w) {
    const bool valid = j < S.M;
    w.lp[0] = valid ? __ldg(S.lp1 + j) : 2.0;  // 2 and exp(2) > 1 >= p: never flagged, never added (valid is re-checked)
    w.lp[1] = valid ? __ldg(S.lp1 + S.M + j) : 2.0;
    w.lp[2] = (DIFF && valid) ? __ldg(S.lp1 + 2 * S.M + j) : 2.0;
    w.g0 = (!DIFF && valid) ? __ldg(group + j) : 0;
    if (DIFF) {
        if (CHUNK) {
#pragma unroll
            for (int c = 0; c < CM; c++) {
                const int cg = c0 + c;
                w.et[c] = (valid && c < C && cg >= 1 && cg <= CT - 2) ? __ldg(S.eta + (int64_t)(cg - 1) * S.M + j) : 0.0;
            }
        } else {
#pragma unroll
            for (int b = 0; b < CM - 2; b++) w.et[b] = (valid && b < C - 2) ? __ldg(S.eta + (int64_t)b * S.M + j) : 0.0;
        }
    }
}

// FULL: the call has exactly CMAX columns (B = 6 -> 8, B = 14 -> 16, the consensus model's 2), so the column loop's
// bounds and the "last column" tests are compile-time constants.
template <bool DIFF, typename UT, bool EXACT, bool SKIPST, int CMAX, bool FULL, bool CHUNK = false>
__global__ void __launch_bounds__(RC_NT, CMAX == 0 ? 1 : (CMAX <= 8 ? 4 : 3))
    rc_apply_kernel(RCSrc S, RCCut cut, int64_t nwt, const uint32_t *__restrict__ tileoff, const __grid_constant__ RCBase ub,
                    const UT *__restrict__ unif, const int32_t *__restrict__ group, int nrep, int64_t G,
                    const RCAcc A, const int *__restrict__ counts, int *__restrict__ err, int c0, int Ctot) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const double p = cut.p;
    const bool thin = p > 0.0;
    const int C = FULL ? CMAX : (CHUNK ? min(CMAX, Ctot - c0) : S.columns);  // columns of this launch
    const int CT = CHUNK ? Ctot : C;                                             // columns of the call
    // Warp tiles are handed out by a counter (err[1], zeroed with the error flag): tiles inside peaks scatter many more
    // weights than background tiles, and with a fixed tile per warp a CTA kept its registers and shared memory until
    // its slowest warp was done (ncu: 15.8 of 32 resident warps active on average).
    for (;;) {
    int64_t tile;
    {
        int t = 0;
        if (lane == 0) t = atomicAdd(err + 1, 1);
        tile = __shfl_sync(0xffffffffu, t, 0);
    }
    if (tile >= nwt) return;
    const int rep = (int)(tile & (RC_REP - 1));
    if (CMAX > 0) {
        constexpr int CM = CMAX > 0 ? CMAX : 1;
        constexpr int SG = RCStage<CM, UT>::SG;
        extern __shared__ unsigned char rc_smem[];
        UT *stage = reinterpret_cast<UT *>(rc_smem) + (size_t)(threadIdx.x >> 5) * RCStage<CM, UT>::WORDS;
        int64_t next[CM];    // index of the column's next uniform (warp-uniform)
        int used[CM];        // ... and how many of the staged ones the current group has consumed
#pragma unroll
        for (int c = 0; c < CM; c++) {
            const int cg = CHUNK ? c0 + c : c;
            next[c] = (thin && c < C) ? ub.ubase[cg] + tileoff[(int64_t)cg * nwt + tile] : 0;
        }
        RCRaw<CM, CHUNK> cur, nxt;
        rc_load_raw<DIFF, CM, CHUNK>(S, group, tile * RC_WT + lane, C, c0, CT, cur);
#pragma unroll 1
        for (int r0 = 0; r0 < RC_STRIPS; r0 += SG) {
            if (thin) {  // the next 32 * SG uniforms of every column (the buffer has that much slack behind its end)
                __syncwarp();
#pragma unroll
                for (int c = 0; c < CM; c++)
                    if (c < C) {
#pragma unroll
                        for (int k = 0; k < SG; k++) stage[(c * SG + k) * 32 + lane] = __ldg(unif + next[c] + k * 32 + lane);
                    }
                __syncwarp();
            }
#pragma unroll
            for (int c = 0; c < CM; c++) used[c] = 0;
#pragma unroll 1
            for (int r = r0; r < r0 + SG; r++) {
                const int64_t j = tile * RC_WT + r * 32 + lane;
                const bool valid = j < S.M;
                // the next strip's loads go out before this strip's arithmetic
                rc_load_raw<DIFF, CM, CHUNK>(S, group, (r + 1 < RC_STRIPS) ? j + 32 : S.M, C, c0, CT, nxt);
                double p1[3], x[CM];
                p1[0] = rc_post(cur.lp[0], S.lin);
                p1[1] = rc_post(cur.lp[1], S.lin);
                p1[2] = DIFF ? rc_post(cur.lp[2], S.lin) : 0.0;
#pragma unroll
                for (int c = 0; c < CM; c++) {
                    if (!DIFF) x[c] = c == 0 ? p1[0] : p1[1];
                    else if (CHUNK) x[c] = (c0 + c == 0) ? p1[0] : (c0 + c == CT - 1 ? p1[2] : rc_mix_weight(p1[1], cur.et[c]));
                    else x[c] = c == 0 ? p1[0] : (c == C - 1 ? p1[2] : rc_mix_weight(p1[1], cur.et[c > 0 ? c - 1 : 0]));
                    if (!(valid && c < C)) x[c] = 2.0;
                }
                unsigned need = 0;
#pragma unroll
                for (int c = 0; c < CM; c++) {
                    if (c < C) {
                        const bool flag = rc_flagged(x[c], p);
                        const unsigned bal = __ballot_sync(0xffffffffu, flag);
                        {   // branch-free: every lane reads a staged word (its own rank's, or the first) and thins; most are flagged
                            const UT uw = stage[c * SG * 32 + (flag ? used[c] + __popc(bal & lt) : 0)];
                            double u;
                            if (sizeof(UT) == 8) u = *reinterpret_cast<const double *>(&uw);
                            else u = mt_to_unif(*reinterpret_cast<const uint32_t *>(&uw));
                            const double thinned = rbinom1(rc_div_cut(x[c], cut.p, cut.rp), u) * cut.p;
                            x[c] = flag ? thinned : x[c];
                        }
                        used[c] += __popc(bal);
                        // x < p without a flag is x == 0: stays 0
                        const int cg = CHUNK ? c0 + c : c;
                        const bool cached = SKIPST && !flag && (cg == 0 || cg == CT - 1);
                        if (valid && x[c] != 0.0 && !cached) need |= 1u << c;
                    }
                }
                if (need) {  // the bin's group ids once, whatever the number of columns it adds to
                    for (int i = 0; i < (DIFF ? nrep : 1); i++) {
                        const int g = DIFF ? __ldg(group + j + (int64_t)i * S.M) : cur.g0;
#pragma unroll
                        for (int c = 0; c < CM; c++)
                            if ((need >> c) & 1u) {
                                const int cg = CHUNK ? c0 + c : c;
                                if (EXACT) limb_add(A, cg, G, g, rep, CT, x[c]);
                                else atomicAdd(reinterpret_cast<double *>(A.w) + (int64_t)cg * (G + 1) + g, x[c]);
                            }
                    }
                }
                cur = nxt;
            }
#pragma unroll
            for (int c = 0; c < CM; c++) next[c] += used[c];
        }
        // the ranks handed out must be the counts the offsets were scanned from
        if (thin) {
            bool bad = false;
#pragma unroll
            for (int c = 0; c < CM; c++) {
                const int cg = CHUNK ? c0 + c : c;
                if (c < C) bad = bad || (next[c] - (ub.ubase[cg] + tileoff[(int64_t)cg * nwt + tile]) != counts[(int64_t)cg * nwt + tile]);
            }
            if (bad && lane == 0) atomicOr(err, 1);
        }
        continue;
    }
    // lane (c & 31) holds the running rank and the uniform offset of column c (and of column c + 32)
    int cnt0 = 0, cnt1 = 0;
    int64_t off0 = 0, off1 = 0;
    if (thin) {
        if (lane < C) off0 = ub.ubase[lane] + tileoff[(int64_t)lane * nwt + tile];
        if (lane + 32 < C) off1 = ub.ubase[lane + 32] + tileoff[(int64_t)(lane + 32) * nwt + tile];
    }
#pragma unroll 1
    for (int r = 0; r < RC_STRIPS; r++) {
        const int64_t j = tile * RC_WT + r * 32 + lane;
        const bool valid = j < S.M;
        double p1[3] = {2.0, 2.0, 2.0};
        if (valid) rc_posteriors<DIFF>(S, j, p1);
#pragma unroll 4
        for (int c = 0; c < C; c++) {
            double x = valid ? rc_weight<DIFF>(S, p1, c, j) : 2.0;
            const bool flag = rc_flagged(x, p);
            const unsigned bal = __ballot_sync(0xffffffffu, flag);
            if (bal) {
                const int run = __shfl_sync(0xffffffffu, c < 32 ? cnt0 : cnt1, c & 31);
                const int64_t base = __shfl_sync(0xffffffffu, c < 32 ? off0 : off1, c & 31);
                if (flag) x = rc_thin<UT>(x, cut, unif, base + run + __popc(bal & lt));
                if (lane == (c & 31)) { if (c < 32) cnt0 += __popc(bal); else cnt1 += __popc(bal); }
            }
            const bool cached = SKIPST && !flag && (c == 0 || c == C - 1);
            if (valid && x != 0.0 && !cached) rc_scatter<EXACT>(A, c, C, G, rep, group, j, S.M, nrep, x);
        }
    }
    if (thin) {
        bool bad = false;
        if (lane < C) bad = counts[(int64_t)lane * nwt + tile] != cnt0;
        if (lane + 32 < C) bad = bad || counts[(int64_t)(lane + 32) * nwt + tile] != cnt1;
        if (bad) atomicOr(err, 1);
    }
    }  // next tile
}

// differential model: the weights >= p of the background and the enrichment column (P1[,0], P1[,2]) do not
// change through the inner iterations that follow one E-step (only mixtureProb and the draws do), and
// exact sums can be split any way: their limbs are accumulated once (st: lo[2*(G+1)] then hi[2*(G+1)],
// column 0 first) and copied under every inner iteration's tables.
__global__ void __launch_bounds__(RC_NT)
    rc_static_kernel(RCSrc S, double p, const int32_t *__restrict__ group, int nrep, int64_t G, const RCAcc A) {
    const int rep = (int)((((int64_t)blockIdx.x * RC_NT + threadIdx.x) >> 5) & (RC_REP - 1));
    for (int64_t j = (int64_t)blockIdx.x * RC_NT + threadIdx.x; j < S.M; j += (int64_t)gridDim.x * RC_NT) {
        double p1[3];
        rc_posteriors<true>(S, j, p1);
        const bool k0 = p1[0] >= p, k2 = p1[2] >= p;   // not below the cut: kept as they are
        if (k0 || k2) {
            for (int i = 0; i < nrep; i++) {
                const int g = __ldg(group + j + (int64_t)i * S.M);
                if (k0) limb_add(A, 0, G, g, rep, 2, p1[0]);
                if (k2) limb_add(A, 1, G, g, rep, 2, p1[2]);
            }
        }
    }
}

// fold the replicas of the frequent groups into the main limb arrays (and clear them): the folded set can then be
// copied under other accumulator sets as two plain arrays
__global__ void rc_fold_kernel(const RCAcc A, int C, int64_t G) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)C * A.hot) return;
    const int c = (int)(i / A.hot);
    const int64_t g = i - (int64_t)c * A.hot;
    unsigned long long lo = 0, hi = 0;
    for (int r = 0; r < RC_REP; r++) {
        unsigned long long *q = A.w + 2 * A.cells + ((int64_t)r * C + c) * A.hot + g;
        lo += q[0];
        hi += q[A.hotcells];
        q[0] = 0;
        q[A.hotcells] = 0;
    }
    A.w[(int64_t)c * (G + 1) + g] += lo;
    A.w[A.cells + (int64_t)c * (G + 1) + g] += hi;
}

// limbs -> fp64 table entries: ((hi + carry) * 2^31 + lo) * 2^-62, rounded once (the integer is exact);
// a rank/count mismatch in the apply kernel poisons the tables instead of leaving them plausible
__global__ void __launch_bounds__(256) rc_finalize_kernel(const RCAcc A, int C, int64_t G, double *__restrict__ buckets,
                                                          const int *__restrict__ err) {
    const bool bad = err && *err != 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < A.cells; i += (int64_t)gridDim.x * 256) {
        unsigned long long lo = A.w[i], hi = A.w[i + A.cells];
        const int c = (int)(i / (G + 1));
        const int64_t g = i - (int64_t)c * (G + 1);
        if (g < A.hot) {
            for (int r = 0; r < RC_REP; r++) {
                const unsigned long long *q = A.w + 2 * A.cells + ((int64_t)r * C + c) * A.hot + g;
                lo += q[0];
                hi += q[A.hotcells];
            }
        }
        const unsigned long long a = hi + (lo >> 31), l = lo & 0x7fffffffull;
        const double v = __dadd_rn((double)a * 4.656612873077392578125e-10, (double)l * 2.16840434497100886801e-19);
        buckets[i] = bad ? __longlong_as_double(0x7ff8000000000000ll) : v;
    }
}

// exported aggregate(): the elements of every group are listed in ascending index order (a stable
// counting sort done on the host, which has to walk f anyway to validate it) and one thread adds a
// group's elements in that order -- the order of the reference's loop (src/aggregate.cpp:23-30), so
// the sums carry the reference's own roundings bit for bit.
__global__ void aggregate_kernel(const double *__restrict__ x, const int64_t *__restrict__ start,
                                 const int64_t *__restrict__ perm, int64_t G, double *__restrict__ buckets) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g > G) return;
    double acc = 0.0;
    for (int64_t k = start[g]; k < start[g + 1]; k++) acc = __dadd_rn(acc, x[perm[k]]);
    buckets[g] = acc;
}

// x -> exp-free variant for the exported reweight(): weights are given directly
__global__ void __launch_bounds__(RC_NT) rw_count_kernel(const double *__restrict__ x, int64_t n, double p,
                                                         int *__restrict__ counts) {
    __shared__ int sh[RC_NT / 32];
    const int64_t base = (int64_t)blockIdx.x * RC_TB;
    int cnt = 0;
    for (int r = 0; r < RC_STRIPS; r++) {
        const int64_t j = base + r * RC_NT + threadIdx.x;
        if (j < n) cnt += (x[j] < p && x[j] / p != 0.0) ? 1 : 0;
    }
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, d);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < RC_NT / 32; i++) t += sh[i];
        counts[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(RC_NT) rw_apply_kernel(double *__restrict__ x, int64_t n, double p,
                                                         const int64_t *__restrict__ offsets,
                                                         const double *__restrict__ unif) {
    __shared__ int sh[RC_NT / 32];
    __shared__ int running;
    const int64_t base = (int64_t)blockIdx.x * RC_TB;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t ubase = offsets[blockIdx.x];
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (int r = 0; r < RC_STRIPS; r++) {
        const int64_t j = base + r * RC_NT + threadIdx.x;
        double v = (j < n) ? x[j] : 1.0e300;
        const bool flag = (j < n) && (v < p && v / p != 0.0);
        const unsigned bal = __ballot_sync(0xffffffffu, flag);
        if (lane == 0) sh[w] = __popc(bal);
        __syncthreads();
        int before = running;
        for (int i = 0; i < w; i++) before += sh[i];
        const int rank = before + __popc(bal & ((1u << lane) - 1u));
        if (j < n && v < p) x[j] = flag ? rbinom1(v / p, __ldg(unif + ubase + rank)) * p : 0.0;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int i = 0; i < RC_NT / 32; i++) t += sh[i];
            running += t;
        }
        __syncthreads();
    }
}

// ---- host side -----------------------------------------------------------------------------------
inline int64_t num_wtiles(int64_t M) { return (M + RC_WT - 1) / RC_WT; }
inline unsigned tile_grid(int64_t nwt) { return (unsigned)((nwt + RC_NW - 1) / RC_NW); }

// rc_coltot layout: int64 [RC_MAXCOL] column totals, then one int: rank/count mismatch flag
inline int *err_flag(ehmm_session *s) { return reinterpret_cast<int *>(s->rc_coltot.as<int64_t>() + RC_MAXCOL); }

int reserve_rc(ehmm_session *s, const RCSrc &S) {
    const int64_t nwt = num_wtiles(s->M);
    EHMM_REQUIRE(S.columns <= RC_MAXCOL, EHMM_ERR_ARG, "too many columns (%d)", S.columns);
    EHMM_REQUIRE(s->M < ((int64_t)1 << 32), EHMM_ERR_ARG, "a bin block holds at most 2^32 bins");
    EHMM_TRY(s->rc_flag.reserve(sizeof(int) * nwt * S.columns));
    EHMM_TRY(s->rc_scan.reserve(sizeof(uint32_t) * nwt * S.columns));
    EHMM_TRY(s->rc_coltot.reserve(sizeof(int64_t) * RC_MAXCOL + 64));
    EHMM_TRY(s->rc_buckets.reserve(sizeof(double) * (size_t)S.columns * (s->G + 1)));
    EHMM_TRY(s->rc_limbs.reserve(sizeof(unsigned long long) * rc_acc_words(S.columns, s->G)));
    return EHMM_OK;
}

// phase 1: flagged counts per (column, warp tile), per-column scan; the column totals come to the host
// (it has to know how many uniforms to generate), which is the call's one synchronisation
template <bool DIFF> int count_and_scan(ehmm_session *s, const RCSrc &S, double p, int64_t *coltot) {
    const int64_t nwt = num_wtiles(s->M);
    s->rc_hint_p = p;  // the next call most likely uses the same cut (max(0.9^iter, probCut) settles at probCut)
    if (!(p > 0.0)) {
        for (int c = 0; c < S.columns; c++) coltot[c] = 0;
        return EHMM_OK;
    }
    // the kernel that produced these posteriors may have counted already (innerExpStep / the E-step, told the cut)
    const bool counted = s->rc_counts_gen == s->post_gen && s->rc_counts_p == p && s->rc_counts_cols == S.columns;
    if (!counted) {
        PROF(s, KID_RC_COUNT), rc_count_kernel<DIFF><<<tile_grid(nwt), RC_NT, 0, s->stream>>>(S, p, nwt, s->rc_flag.as<int>());
        EHMM_CK(cudaGetLastError());
        s->rc_counts_gen = s->post_gen;
        s->rc_counts_p = p;
        s->rc_counts_cols = S.columns;
    }
    PROF(s, KID_RC_SCAN), rc_scan_kernel<<<S.columns, SCAN_NT, 0, s->stream>>>(s->rc_flag.as<int>(), nwt, s->rc_scan.as<uint32_t>(),
                                                                           s->rc_coltot.as<int64_t>());
    EHMM_CK(cudaGetLastError());
    int64_t *hp = reinterpret_cast<int64_t *>(s->h_pinned + 200);
    EHMM_CK(cudaMemcpyAsync(hp, s->rc_coltot.p, sizeof(int64_t) * S.columns, cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    for (int c = 0; c < S.columns; c++) coltot[c] = hp[c];
    return EHMM_OK;
}

// limbs of the weights >= p of columns 0 and B+1 for the current logProb1 (no-op when they are in place)
int rc_static_build(ehmm_session *s, const RCSrc &S, double p, int nrep, cudaStream_t st) {
    if (s->rc_static_gen == s->lp1_gen && s->rc_static_p == p) return EHMM_OK;
    const size_t LW = sizeof(unsigned long long);
    EHMM_TRY(s->rc_static.reserve(LW * rc_acc_words(2, s->G)));
    EHMM_CK(cudaMemsetAsync(s->rc_static.p, 0, LW * rc_acc_words(2, s->G), st));
    const RCAcc A = rc_acc(s->rc_static.as<unsigned long long>(), 2, s->G);
    const unsigned grid = (unsigned)std::min<int64_t>((s->M + RC_NT - 1) / RC_NT, 148 * 16);
    if (st == s->stream) {
        PROF(s, KID_RC_STATIC), rc_static_kernel<<<grid, RC_NT, 0, st>>>(S, p, s->group.as<int32_t>(), nrep, s->G, A);
    } else {
        rc_static_kernel<<<grid, RC_NT, 0, st>>>(S, p, s->group.as<int32_t>(), nrep, s->G, A);
        s->launches[KID_RC_STATIC]++;
    }
    EHMM_CK(cudaGetLastError());
    rc_fold_kernel<<<(unsigned)((2 * A.hot + 255) / 256), 256, 0, st>>>(A, 2, s->G);
    EHMM_CK(cudaGetLastError());
    s->rc_static_gen = s->lp1_gen;
    s->rc_static_p = p;
    return EHMM_OK;
}

// phase 2: thin and sum by group.  ub: where each column's draws start in `unif`.
template <bool DIFF, typename UT, bool EXACT, bool SKIPST, int CM, bool FULL, bool CHUNK = false>
int launch_apply_cm(ehmm_session *s, const RCSrc &S, const RCCut &cut, int nrep, const RCBase &ub, const UT *unif, void *accp, int c0 = 0) {
    const RCAcc acc = rc_acc(reinterpret_cast<unsigned long long *>(accp), S.columns, s->G);  // fp64 mode: only .w is used
    const int64_t nwt = num_wtiles(s->M);
    constexpr int CS = CM > 0 ? CM : 1;
    const size_t smem = CM > 0 ? sizeof(UT) * RCStage<CS, UT>::WORDS * RC_NW : 0;
    // as many CTAs as can be resident (4 per SM where the kernel's bounds allow it), each warp fetching tiles until none is left
    const unsigned grid = std::min<unsigned>(tile_grid(nwt), 148u * (CM == 0 ? 1u : (CM <= 8 ? 4u : 3u)));
    if (CHUNK && c0 > 0) EHMM_CK(cudaMemsetAsync(err_flag(s) + 1, 0, sizeof(int), s->stream));  // the tile counter, not the error flag
    PROF(s, KID_RC_APPLY), rc_apply_kernel<DIFF, UT, EXACT, SKIPST, CM, FULL, CHUNK><<<grid, RC_NT, smem, s->stream>>>(
        S, cut, nwt, s->rc_scan.as<uint32_t>(), ub, unif, s->group.as<int32_t>(), nrep, s->G, acc, s->rc_flag.as<int>(), err_flag(s), c0,
        S.columns);
    EHMM_CK(cudaGetLastError());
    return EHMM_OK;
}

template <bool DIFF, typename UT, bool EXACT, bool SKIPST>
int launch_apply_kernel(ehmm_session *s, const RCSrc &S, double p, int nrep, const RCBase &ub, const UT *unif, void *acc) {
    const RCCut cut = {p, p > 0.0 ? 1.0 / p : 0.0};
    if (!DIFF) return launch_apply_cm<DIFF, UT, EXACT, SKIPST, 2, true>(s, S, cut, nrep, ub, unif, acc);  // K = 2 columns, always
    if (S.columns == 8) return launch_apply_cm<DIFF, UT, EXACT, SKIPST, 8, true>(s, S, cut, nrep, ub, unif, acc);    // B = 6: BASELINE configs[2]
    if (S.columns < 8) return launch_apply_cm<DIFF, UT, EXACT, SKIPST, 8, false>(s, S, cut, nrep, ub, unif, acc);
    if (S.columns == 16) return launch_apply_cm<DIFF, UT, EXACT, SKIPST, 16, true>(s, S, cut, nrep, ub, unif, acc);  // B = 14: configs[4]
    if (S.columns < 16) return launch_apply_cm<DIFF, UT, EXACT, SKIPST, 16, false>(s, S, cut, nrep, ub, unif, acc);
    if (DIFF && S.columns <= RC_MAXCOL) {
        // more columns than one launch can keep in registers (B = 30: BASELINE configs[3]): eight at a time; every launch
        // walks all tiles for its columns (the generic kernel below keeps per-column state in lanes, one CTA per SM,
        // and took ten times longer per column)
        for (int c0 = 0; c0 < S.columns; c0 += 8) {
            if (c0 + 8 <= S.columns) EHMM_TRY((launch_apply_cm<DIFF, UT, EXACT, SKIPST, 8, true, true>(s, S, cut, nrep, ub, unif, acc, c0)));
            else EHMM_TRY((launch_apply_cm<DIFF, UT, EXACT, SKIPST, 8, false, true>(s, S, cut, nrep, ub, unif, acc, c0)));
        }
        return EHMM_OK;
    }
    return launch_apply_cm<DIFF, UT, EXACT, SKIPST, 0, false>(s, S, cut, nrep, ub, unif, acc);
}

template <bool DIFF, typename UT>
int launch_apply(ehmm_session *s, const RCSrc &S, double p, int nrep, const RCBase &ub, const UT *unif) {
    const size_t cells = (size_t)S.columns * (s->G + 1);
    const bool exact = p >= RC_EXACT_MIN_P;
    const size_t LW = sizeof(unsigned long long);
    EHMM_CK(cudaMemsetAsync(err_flag(s), 0, 2 * sizeof(int), s->stream));  // error flag and the kernel's tile counter
    if (exact) {
        unsigned long long *limbs = s->rc_limbs.as<unsigned long long>();
        EHMM_CK(cudaMemsetAsync(limbs, 0, LW * rc_acc_words(S.columns, s->G), s->stream));
        if (DIFF) {
            const size_t G1 = (size_t)s->G + 1;
            EHMM_TRY(rc_static_build(s, S, p, nrep, s->stream));
            if (s->static_pending) {  // built on the side stream right after the E-step (rc_static_prefetch)
                EHMM_CK(cudaStreamWaitEvent(s->stream, s->ev_static, 0));
                s->static_pending = false;
            }
            // static limbs (lo: column 0, column B+1; hi: the same) under this call's tables
            const unsigned long long *st = s->rc_static.as<unsigned long long>();
            const size_t last = (size_t)(S.columns - 1) * G1;
            EHMM_CK(cudaMemcpyAsync(limbs, st, LW * G1, cudaMemcpyDeviceToDevice, s->stream));
            EHMM_CK(cudaMemcpyAsync(limbs + last, st + G1, LW * G1, cudaMemcpyDeviceToDevice, s->stream));
            EHMM_CK(cudaMemcpyAsync(limbs + cells, st + 2 * G1, LW * G1, cudaMemcpyDeviceToDevice, s->stream));
            EHMM_CK(cudaMemcpyAsync(limbs + cells + last, st + 3 * G1, LW * G1, cudaMemcpyDeviceToDevice, s->stream));
            EHMM_TRY((launch_apply_kernel<DIFF, UT, true, DIFF>(s, S, p, nrep, ub, unif, limbs)));
        } else {
            EHMM_TRY((launch_apply_kernel<DIFF, UT, true, false>(s, S, p, nrep, ub, unif, limbs)));
        }
    } else {
        EHMM_CK(cudaMemsetAsync(s->rc_buckets.p, 0, sizeof(double) * cells, s->stream));
        EHMM_TRY((launch_apply_kernel<DIFF, UT, false, false>(s, S, p, nrep, ub, unif, s->rc_buckets.p)));
    }
    s->rc_exact = exact;
    s->rc_final = !exact;
    s->rc_columns = S.columns;
    s->rc_N = nrep;
    return EHMM_OK;
}

int upload_state_if_needed(ehmm_session *s) {
    EHMM_TRY(s->mt_raw.reserve(sizeof(uint32_t) * 625));
    if (!s->rng_dev_valid) {
        memcpy(s->h_pinned + 512, s->rng_state, sizeof(uint32_t) * 625);
        EHMM_CK(cudaMemcpyAsync(s->mt_raw.p, s->h_pinned + 512, sizeof(uint32_t) * 625, cudaMemcpyHostToDevice, s->stream));
        s->rng_dev_valid = true;
        s->mtj_valid = 0;
        s->mtj_niv = 0;
        s->mtj_base = false;
    }
    return EHMM_OK;
}

template <bool DIFF>
int run_rc(ehmm_session *s, const RCSrc &S, double p, int nrep, const double *uniforms, int64_t n_uniforms,
           int64_t *consumed) {
    EHMM_TRY(reserve_rc(s, S));
    int64_t coltot[RC_MAXCOL];
    EHMM_TRY(count_and_scan<DIFF>(s, S, p, coltot));
    RCBase ub;
    int64_t total = 0;
    for (int c = 0; c < S.columns; c++) { ub.ubase[c] = total; total += coltot[c]; }
    if (uniforms) {
        EHMM_REQUIRE(total <= n_uniforms, EHMM_ERR_RNG, "rejection control needs %lld uniforms, buffer holds %lld",
                     (long long)total, (long long)n_uniforms);
        EHMM_TRY(s->rc_unif.reserve_growing(sizeof(double) * (size_t)(total + RC_UNIF_SLACK)));
        if (total > 0)
            EHMM_CK(cudaMemcpyAsync(s->rc_unif.p, uniforms, sizeof(double) * total, cudaMemcpyHostToDevice, s->stream));
        EHMM_TRY((launch_apply<DIFF, double>(s, S, p, nrep, ub, s->rc_unif.as<double>())));
    } else {
        EHMM_REQUIRE(s->rng_set || total == 0, EHMM_ERR_RNG,
                     "rejection control needs %lld uniforms: pass a buffer or call ehmm_rng_set_state", (long long)total);
        if (total > 0) {
            EHMM_TRY(upload_state_if_needed(s));
            if (s->spec_n >= total) {
                // generated ahead on the side stream while the posteriors were being computed
                EHMM_TRY(mt_wait_side(s));
                std::swap(s->rc_unif, s->rc_unif_spec);
            } else {
                EHMM_TRY(mt_choose_stride(s, total));
                EHMM_TRY(s->rc_unif.reserve_growing(sizeof(uint32_t) * (size_t)(total + RC_UNIF_SLACK)));
                EHMM_TRY(mt_generate_range(s, 0, total, s->rc_unif.as<uint32_t>()));
            }
            s->spec_n = 0;
            // side stream from here on: the state after this call's draws (R's .Random.seed), the windows of the
            // new position, and the next call's uniforms -- sized from this call's consumption with 25% headroom
            EHMM_TRY(mt_advance_side(s, total, s->h_pinned + 512));
            rng_note_advance(s, total);
            int64_t est = total + total / 4 + 2 * 131072;
            if (est > (int64_t)S.columns * s->M) est = (int64_t)S.columns * s->M;
            EHMM_TRY(mt_choose_stride(s, est));
            EHMM_TRY(s->rc_unif_spec.reserve_growing(sizeof(uint32_t) * (size_t)(est + RC_UNIF_SLACK)));
            EHMM_TRY(mt_speculate(s, est, s->rc_unif_spec.as<uint32_t>()));
            s->spec_n = est;
        } else {
            EHMM_TRY(s->rc_unif.reserve(sizeof(uint32_t) * (size_t)RC_UNIF_SLACK));
        }
        EHMM_TRY((launch_apply<DIFF, uint32_t>(s, S, p, nrep, ub, s->rc_unif.as<uint32_t>())));
    }
    EHMM_TRY(rc_finalize(s));
    // the host does not wait for the tables: the next reader of the stream does
    if (consumed) *consumed = total;
    return EHMM_OK;
}

// ---- sharded form: the caller owns the ordering across bin blocks --------------------------------
// phase 1: flagged counts of this block per column (count + scan kernels; offsets stay in rc_scan)
template <bool DIFF> int rc_count_block(ehmm_session *s, const RCSrc &S, double p, int64_t *counts) {
    EHMM_TRY(reserve_rc(s, S));
    EHMM_TRY(count_and_scan<DIFF>(s, S, p, counts));
    for (int c = 0; c < S.columns; c++) s->rc_colcount[c] = counts[c];
    s->rc_count_p = p;
    s->rc_count_cols = S.columns;
    return EHMM_OK;
}

// phase 2: this block's draws of column c start at stream offset col_offsets[c] (relative to the
// current RNG position); the stream advances by `total` (the draws of all blocks).  The tables are
// left as exact limbs: the caller adds the limbs of all blocks (ehmm_rc_buckets_device) and then
// calls ehmm_rc_finalize, so the rounded table is the same for every sharding.
template <bool DIFF>
int rc_apply_block(ehmm_session *s, const RCSrc &S, double p, int nrep, const int64_t *col_offsets, int64_t total) {
    EHMM_REQUIRE(s->rc_count_cols == S.columns && s->rc_count_p == p, EHMM_ERR_STATE,
                 "rc_apply_at must follow rc_count with the same p");
    EHMM_REQUIRE(s->rng_set || total == 0, EHMM_ERR_RNG, "rejection control needs the RNG state (ehmm_rng_set_state)");
    int64_t local = 0;
    for (int c = 0; c < S.columns; c++) local += s->rc_colcount[c];
    static const bool trace = getenv("EHMM_TRACE") != nullptr;
    auto t_prev = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!trace) return;
        const auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[rc_apply_at] %-18s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count());
        t_prev = t;
    };
    RCBase ub;
    const uint32_t *unif = nullptr;
    // generated ahead on the side stream?  every column's slice (plus what the kernel stages past its end) must lie
    // inside what was generated for it, relative to the same stream position
    bool hit = total > 0 && s->specr_cols == S.columns && s->specr_epoch == s->rng_epoch;
    for (int c = 0; hit && c < S.columns; c++)
        hit = col_offsets[c] >= s->specr_lo[c] && col_offsets[c] + s->rc_colcount[c] + RC_UNIF_SLACK <= s->specr_lo[c] + s->specr_n[c];
    if (total > 0) {
        EHMM_TRY(upload_state_if_needed(s));
        for (int c = 0; c < S.columns; c++)
            EHMM_REQUIRE(col_offsets[c] >= 0 && col_offsets[c] + s->rc_colcount[c] <= total, EHMM_ERR_ARG,
                         "column %d: draws [%lld, %lld) outside [0, %lld)", c, (long long)col_offsets[c],
                         (long long)(col_offsets[c] + s->rc_colcount[c]), (long long)total);
    }
    if (hit) {
        EHMM_TRY(mt_wait_side(s));
        unif = s->rc_unif_specr[s->specr_next ^ 1].as<uint32_t>();
        for (int c = 0; c < S.columns; c++) ub.ubase[c] = s->specr_off[c] + (col_offsets[c] - s->specr_lo[c]);
    } else {
        EHMM_TRY(s->rc_unif.reserve_growing(sizeof(uint32_t) * (size_t)(local + RC_UNIF_SLACK)));
        int64_t at = 0;
        for (int c = 0; c < S.columns; c++) { ub.ubase[c] = at; at += s->rc_colcount[c]; }
        if (total > 0) {
            // sub-stream length from THIS block's draws: its column slices together should keep a few hundred generating
            // CTAs busy (sized from the global total, a slice was a handful of sub-streams and the launch a handful of CTAs)
            EHMM_TRY(mt_choose_stride(s, local));
            int64_t out_off[RC_MAXCOL];
            for (int c = 0; c < S.columns; c++) out_off[c] = ub.ubase[c];
            EHMM_TRY(mt_generate_ranges(s, S.columns, col_offsets, s->rc_colcount, out_off, s->rc_unif.as<uint32_t>()));
        }
        unif = s->rc_unif.as<uint32_t>();
    }
    s->specr_cols = 0;
    lap(hit ? "uniforms (hit)" : "uniforms (miss)");
    EHMM_TRY((launch_apply<DIFF, uint32_t>(s, S, p, nrep, ub, unif)));
    lap("launch_apply");
    if (total > 0) {
        // side stream: the stream's state after all ranks' draws, then the windows of the next call's
        // slice, which sits where this one's did, give or take a few thousand draws
        EHMM_TRY(mt_advance_side(s, total, s->h_pinned + 512));
        rng_note_advance(s, total);
        EHMM_TRY(mt_choose_stride(s, local));
        lap("advance_side");
        // ... and the slices themselves, with a margin of 1/16 of a slice (at least 2^17 words) either side
        int64_t lo[RC_MAXCOL], n[RC_MAXCOL], off[RC_MAXCOL], words = 0;
        for (int c = 0; c < S.columns; c++) {
            const int64_t m = std::max<int64_t>(131072, s->rc_colcount[c] / 16);
            lo[c] = std::max<int64_t>(0, col_offsets[c] - m);
            n[c] = (col_offsets[c] - lo[c]) + s->rc_colcount[c] + m + RC_UNIF_SLACK;
            off[c] = words;
            words += n[c];
        }
        DevBuf &buf = s->rc_unif_specr[s->specr_next];
        EHMM_TRY(buf.reserve_growing(sizeof(uint32_t) * (size_t)words));
        lap("reserve");
        EHMM_TRY(mt_prefetch_ranges(s, S.columns, lo, n, total, 1));
        lap("prefetch_ranges");
        EHMM_TRY(mt_speculate_ranges(s, S.columns, lo, n, off, buf.as<uint32_t>()));
        lap("speculate_ranges");
        for (int c = 0; c < S.columns; c++) { s->specr_lo[c] = lo[c]; s->specr_n[c] = n[c]; s->specr_off[c] = off[c]; }
        s->specr_cols = S.columns;
        s->specr_epoch = s->rng_epoch;
        s->specr_next ^= 1;
    }
    s->rc_count_cols = 0;
    return EHMM_OK;
}

}  // namespace

void rng_note_advance(ehmm_session *s, int64_t n) {
    s->rng_epoch++;
    const int64_t E = (int64_t)s->rng_state[0] + n;  // R: mti counts the words used of the current block of 624
    s->rng_state[0] = (uint32_t)(E <= 624 ? E : (E - 1) % 624 + 1);
    s->rng_pending = true;
}

int rng_sync_host(ehmm_session *s) {
    if (!s->rng_pending) return EHMM_OK;
    EHMM_CK(cudaStreamSynchronize(s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream2));  // the sharded form copies the state on the side stream
    const uint32_t *h = reinterpret_cast<const uint32_t *>(s->h_pinned + 512);
    EHMM_REQUIRE(h[0] == s->rng_state[0], EHMM_ERR_RNG, "RNG position on the device (%u) differs from the booked one (%u)", h[0],
                 s->rng_state[0]);
    memcpy(s->rng_state, h, sizeof(uint32_t) * 625);
    s->rng_pending = false;
    return EHMM_OK;
}

namespace {

struct TmpDev {
    void *p = nullptr;
    ~TmpDev() { if (p) cudaFree(p); }
    int alloc(size_t n) {
        cudaError_t e = cudaMalloc(&p, n ? n : 8);
        if (e != cudaSuccess) { set_error("cudaMalloc(%zu): %s", n, cudaGetErrorString(e)); return EHMM_ERR_NOMEM; }
        return EHMM_OK;
    }
};

}  // namespace

int64_t rc_limb_words(const ehmm_session *s) { return (int64_t)rc_acc_words(s->rc_columns, s->G); }

// the replicas of the frequent groups added into the main limb arrays (and cleared): what is left to exchange
// between bin blocks is the two plain arrays, 2 * columns * (G + 1) words
int rc_fold(ehmm_session *s, int64_t *words) {
    const RCAcc A = rc_acc(s->rc_limbs.as<unsigned long long>(), s->rc_columns, s->G);
    PROF(s, KID_RC_FINAL), rc_fold_kernel<<<(unsigned)((s->rc_columns * A.hot + 255) / 256), 256, 0, s->stream>>>(A, s->rc_columns, s->G);
    EHMM_CK(cudaGetLastError());
    *words = 2 * A.cells;
    return EHMM_OK;
}

// limbs -> fp64 tables (no-op when the tables were accumulated in fp64 or are already final)
int rc_finalize(ehmm_session *s) {
    if (s->rc_final || !s->rc_exact) { s->rc_final = true; return EHMM_OK; }
    const int64_t n = (int64_t)s->rc_columns * (s->G + 1);
    const unsigned grid = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8);
    PROF(s, KID_RC_FINAL), rc_finalize_kernel<<<grid, 256, 0, s->stream>>>(
        rc_acc(s->rc_limbs.as<unsigned long long>(), s->rc_columns, s->G), s->rc_columns, s->G, s->rc_buckets.as<double>(),
        reinterpret_cast<const int *>(s->rc_coltot.as<int64_t>() + RC_MAXCOL));
    EHMM_CK(cudaGetLastError());
    s->rc_final = true;
    return EHMM_OK;
}

static RCSrc rc_source(ehmm_session *s, int differential) {
    const P1Src ps = p1_source(s);
    if (differential) return RCSrc{ps.p, s->eta.as<double>(), s->M, s->B, s->B + 2, ps.lin};
    return RCSrc{ps.p, nullptr, s->M, 0, s->K, ps.lin};
}

// Differential model, right after an E-step: the part of the rejection tables that the coming inner
// iterations share (rc_static_kernel) depends only on logProb1 and the cut, so when the cut is known
// (ehmm_rc_hint, or the previous call's) it is summed on the side stream while the main stream goes on with
// maxStepProb and the first innerExpStep.
int rc_static_prefetch(ehmm_session *s) {
    const double p = s->rc_hint_p;
    if (!(s->K == 3 && s->B > 0 && s->G > 0 && s->group.p && s->N > 0 && p >= RC_EXACT_MIN_P && p <= 1.0 && has_post(s)))
        return EHMM_OK;
    if (s->rc_static_gen == s->lp1_gen && s->rc_static_p == p) return EHMM_OK;
    EHMM_CK(cudaEventRecord(s->ev2, s->stream));
    EHMM_CK(cudaStreamWaitEvent(s->stream2, s->ev2, 0));
    EHMM_TRY(rc_static_build(s, rc_source(s, 1), p, s->N, s->stream2));
    EHMM_CK(cudaEventRecord(s->ev_static, s->stream2));
    s->static_pending = true;
    return EHMM_OK;
}

int rc_consensus(ehmm_session *s, double p, const double *uniforms, int64_t n_uniforms, int64_t *consumed) {
    EHMM_REQUIRE(s->K == 2, EHMM_ERR_ARG, "consensusRejectionControlled needs K=2");
    return run_rc<false>(s, rc_source(s, 0), p, 1, uniforms, n_uniforms, consumed);
}

int rc_differential(ehmm_session *s, double p, int N, const double *uniforms, int64_t n_uniforms, int64_t *consumed) {
    EHMM_REQUIRE(s->K == 3 && s->B > 0, EHMM_ERR_ARG, "differentialRejectionControlled needs K=3 and mixture components");
    return run_rc<true>(s, rc_source(s, 1), p, N, uniforms, n_uniforms, consumed);
}

int rc_count(ehmm_session *s, int differential, double p, int64_t *counts) {
    EHMM_REQUIRE(differential ? (s->K == 3 && s->B > 0) : s->K == 2, EHMM_ERR_ARG, "rc_count: model / session mismatch");
    return differential ? rc_count_block<true>(s, rc_source(s, 1), p, counts) : rc_count_block<false>(s, rc_source(s, 0), p, counts);
}

int rc_apply_at(ehmm_session *s, int differential, double p, const int64_t *col_offsets, int64_t total) {
    return differential ? rc_apply_block<true>(s, rc_source(s, 1), p, s->N, col_offsets, total)
                        : rc_apply_block<false>(s, rc_source(s, 0), p, 1, col_offsets, total);
}

static int fetch_column(ehmm_session *s, int column, std::vector<double> &h) {
    EHMM_TRY(rc_finalize(s));
    h.resize((size_t)s->G + 1);
    EHMM_CK(cudaMemcpyAsync(h.data(), s->rc_buckets.as<double>() + (size_t)column * (s->G + 1), sizeof(double) * (s->G + 1),
                            cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    return EHMM_OK;
}

int rc_table_rows(ehmm_session *s, int column, int64_t *rows) {
    std::vector<double> h;
    EHMM_TRY(fetch_column(s, column, h));
    int64_t n = 0;
    for (double v : h) n += (v != 0.0);
    *rows = n;
    return EHMM_OK;
}

int rc_table_fetch(ehmm_session *s, int column, double *sums, double *ids) {
    std::vector<double> h;
    EHMM_TRY(fetch_column(s, column, h));
    int64_t n = 0;
    for (size_t g = 0; g < h.size(); g++)
        if (h[g] != 0.0) { sums[n] = h[g]; ids[n] = (double)g; n++; }
    return EHMM_OK;
}

int rc_reweight(int device, double *x, int64_t n, double p, const double *uniforms, int64_t n_uniforms, int64_t *consumed) {
    EHMM_CK(cudaSetDevice(device));
    if (consumed) *consumed = 0;
    if (n == 0 || !(p > 0.0)) return EHMM_OK;
    const int64_t nblk = (n + RC_TB - 1) / RC_TB;
    TmpDev dx, dc, doff, du;
    EHMM_TRY(dx.alloc(sizeof(double) * n));
    EHMM_TRY(dc.alloc(sizeof(int) * nblk));
    EHMM_TRY(doff.alloc(sizeof(int64_t) * (nblk + 1)));
    EHMM_CK(cudaMemcpy(dx.p, x, sizeof(double) * n, cudaMemcpyHostToDevice));
    rw_count_kernel<<<(unsigned)nblk, RC_NT>>>((const double *)dx.p, n, p, (int *)dc.p);
    rw_scan_kernel<<<1, 1024>>>((const int *)dc.p, nblk, (int64_t *)doff.p);
    EHMM_CK(cudaGetLastError());
    int64_t total = 0;
    EHMM_CK(cudaMemcpy(&total, (int64_t *)doff.p + nblk, sizeof(int64_t), cudaMemcpyDeviceToHost));
    EHMM_REQUIRE(uniforms || total == 0, EHMM_ERR_RNG, "reweight needs %lld uniforms", (long long)total);
    EHMM_REQUIRE(total <= n_uniforms, EHMM_ERR_RNG, "reweight needs %lld uniforms, buffer holds %lld", (long long)total,
                 (long long)n_uniforms);
    EHMM_TRY(du.alloc(sizeof(double) * total));
    if (total) EHMM_CK(cudaMemcpy(du.p, uniforms, sizeof(double) * total, cudaMemcpyHostToDevice));
    rw_apply_kernel<<<(unsigned)nblk, RC_NT>>>((double *)dx.p, n, p, (const int64_t *)doff.p, (const double *)du.p);
    EHMM_CK(cudaGetLastError());
    EHMM_CK(cudaMemcpy(x, dx.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (consumed) *consumed = total;
    return EHMM_OK;
}

int rc_aggregate(int device, const double *x, int64_t n, const int32_t *f, int64_t G, double *buckets) {
    EHMM_CK(cudaSetDevice(device));
    std::vector<int64_t> start((size_t)G + 2, 0), perm((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        EHMM_REQUIRE(f[i] >= 0 && f[i] <= G, EHMM_ERR_ARG, "aggregate: group id %d at %lld outside [0, %lld]", f[i],
                     (long long)i, (long long)G);
        start[(size_t)f[i] + 1]++;
    }
    for (int64_t g = 0; g <= G; g++) start[(size_t)g + 1] += start[(size_t)g];
    {
        std::vector<int64_t> at(start.begin(), start.end() - 1);
        for (int64_t i = 0; i < n; i++) perm[(size_t)at[(size_t)f[i]]++] = i;  // ascending i inside every group
    }
    TmpDev dx, dp, ds, db;
    EHMM_TRY(dx.alloc(sizeof(double) * n));
    EHMM_TRY(dp.alloc(sizeof(int64_t) * n));
    EHMM_TRY(ds.alloc(sizeof(int64_t) * (G + 2)));
    EHMM_TRY(db.alloc(sizeof(double) * (G + 1)));
    EHMM_CK(cudaMemcpy(dx.p, x, sizeof(double) * n, cudaMemcpyHostToDevice));
    EHMM_CK(cudaMemcpy(dp.p, perm.data(), sizeof(int64_t) * n, cudaMemcpyHostToDevice));
    EHMM_CK(cudaMemcpy(ds.p, start.data(), sizeof(int64_t) * (G + 2), cudaMemcpyHostToDevice));
    aggregate_kernel<<<(unsigned)((G + 1 + 255) / 256), 256>>>((const double *)dx.p, (const int64_t *)ds.p, (const int64_t *)dp.p, G,
                                                              (double *)db.p);
    EHMM_CK(cudaGetLastError());
    EHMM_CK(cudaMemcpy(buckets, db.p, sizeof(double) * (G + 1), cudaMemcpyDeviceToHost));
    return EHMM_OK;
}

}  // namespace ehmm
