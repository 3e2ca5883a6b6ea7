// rc.cu -- rejection-controlled reweighting and aggregation by group (replaces
// src/consensusRejectionControlled.cpp, src/differentialRejectionControlled.cpp, src/reweight.cpp,
// src/rbinomVectorized.cpp and src/aggregate.cpp, which loop sequentially over bins drawing from
// R's global Mersenne-Twister).
//
// The r-th bin (state-major, then ascending bin) whose posterior weight x satisfies 0 < x < p
// consumes the r-th uniform of the stream.  That order is recovered with a stream compaction:
//   rc_count_kernel : flags per (column, block)                      (reads 8 B/bin/column)
//   rc_scan_kernel  : exclusive scan of the block counts (column-major) -> uniform offsets
//   rc_apply_kernel : in-block ranks, Bernoulli thinning exactly as R's rbinom(1, x/p), and
//                     scatter-add of the surviving weights into the group buckets
// Uniforms come either from a caller-supplied buffer (parity harness: the oracle consumes the same
// buffer) or from R's MT19937 stream generated on the device from the session's .Random.seed.
#include "common.cuh"
#include <vector>

namespace ehmm {

namespace {

constexpr int RC_NT = 256;
constexpr int RC_STRIPS = 8;
constexpr int RC_TB = RC_NT * RC_STRIPS;  // bins per block

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
    const double *lp1;
    const double *eta;
    int64_t M;
    int B;         // 0 = consensus
    int columns;
};

__device__ __forceinline__ double rc_weight(const RCSrc &S, int c, int64_t j) {
    if (S.B == 0) return exp(__ldg(S.lp1 + (int64_t)c * S.M + j));
    if (c == 0) return exp(__ldg(S.lp1 + j));
    if (c == S.B + 1) return exp(__ldg(S.lp1 + 2 * S.M + j));
    return exp(__ldg(S.lp1 + S.M + j)) * __ldg(S.eta + (int64_t)(c - 1) * S.M + j);
}

__global__ void __launch_bounds__(RC_NT) rc_count_kernel(RCSrc S, double p, int64_t nblk, int *__restrict__ counts) {
    __shared__ int sh[RC_NT / 32];
    const int c = blockIdx.y;
    const int64_t base = (int64_t)blockIdx.x * RC_TB;
    int cnt = 0;
#pragma unroll
    for (int r = 0; r < RC_STRIPS; r++) {
        const int64_t j = base + r * RC_NT + threadIdx.x;
        if (j < S.M) {
            const double x = rc_weight(S, c, j);
            cnt += (x < p && x / p != 0.0) ? 1 : 0;
        }
    }
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, d);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < RC_NT / 32; i++) t += sh[i];
        counts[(int64_t)c * nblk + blockIdx.x] = t;
    }
}

// exclusive scan of n ints into int64 offsets; offsets[n] = total.  One CTA; each thread owns 16
// consecutive counts per round (warp-shuffle scan of the thread sums, then the warp sums).
__global__ void __launch_bounds__(1024) rc_scan_kernel(const int *__restrict__ counts, int64_t n,
                                                       int64_t *__restrict__ offsets) {
    constexpr int IT = 16;
    __shared__ int64_t swarp[32];
    __shared__ int64_t carry;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024 * IT) {
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
        if (w == 0) {
            int64_t x = swarp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int64_t t = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += t;
            }
            swarp[lane] = x;  // inclusive sums of the warps
        }
        __syncthreads();
        int64_t run = carry + (w > 0 ? swarp[w - 1] : 0) + (incl - tsum);
#pragma unroll
        for (int k = 0; k < IT; k++) {
            if (i0 + k < n) offsets[i0 + k] = run;
            run += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += swarp[31];
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

// nrep = 1 for the consensus model (only f[0..M-1] is read, Q6 in SURVEY.md); nrep = N for the
// differential model (repmat(w, N, 1) against the M*N group ids).
// Each thread owns RC_STRIPS consecutive bins, so ranks inside a tile come from one warp scan and
// one block barrier; the weight loads of a warp still cover one contiguous 2 KB span.
// Group ids are numbered by first appearance (data.table .GRP), so small ids are the frequent
// (count, offset) combinations: ids < RC_HOT accumulate in a shared-memory histogram of a persistent
// CTA (flushed once at the end) and only the long tail goes to global atomics.
constexpr int RC_HOT = 4096;

template <typename UT>
__global__ void __launch_bounds__(RC_NT)
    rc_apply_kernel(RCSrc S, double p, int64_t nblk, const int64_t *__restrict__ offsets, const UT *__restrict__ unif,
                    const int32_t *__restrict__ group, int nrep, int64_t G, double *__restrict__ buckets,
                    double *__restrict__ wout) {
    __shared__ double hot[RC_HOT];
    __shared__ int sh[RC_NT / 32];
    const int c = blockIdx.y;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double *bk = buckets + (int64_t)c * (G + 1);
    for (int i = threadIdx.x; i < RC_HOT; i += RC_NT) hot[i] = 0.0;
    __syncthreads();
    for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const int64_t j0 = blk * RC_TB + (int64_t)threadIdx.x * RC_STRIPS;
        double x[RC_STRIPS];
        unsigned flags = 0;
#pragma unroll
        for (int r = 0; r < RC_STRIPS; r++) {
            const int64_t j = j0 + r;
            x[r] = (j < S.M) ? rc_weight(S, c, j) : 2.0;  // 2.0: never below p <= 1, never stored
            if (x[r] < p && x[r] / p != 0.0) flags |= 1u << r;
        }
        const int cnt = __popc(flags);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) sh[w] = incl;
        __syncthreads();
        int before = incl - cnt;
        for (int i = 0; i < w; i++) before += sh[i];
        int64_t u = offsets[(int64_t)c * nblk + blk] + before;
#pragma unroll
        for (int r = 0; r < RC_STRIPS; r++) {
            const int64_t j = j0 + r;
            if (j < S.M) {
                double wt = x[r];
                if (wt < p) wt = ((flags >> r) & 1u) ? rbinom1(wt / p, get_unif<UT>(unif, u++)) * p : 0.0;
                if (wout) wout[(int64_t)c * S.M + j] = wt;
                if (wt != 0.0) {
                    for (int i = 0; i < nrep; i++) {
                        const int g = __ldg(group + j + (int64_t)i * S.M);
                        if (g < RC_HOT) atomicAdd(&hot[g], wt);
                        else atomicAdd(bk + g, wt);
                    }
                }
            }
        }
        __syncthreads();  // sh[] is reused by the next tile
    }
    __syncthreads();
    const int lim = (int)min((int64_t)RC_HOT, G + 1);
    for (int i = threadIdx.x; i < lim; i += RC_NT)
        if (hot[i] != 0.0) atomicAdd(bk + i, hot[i]);
}

__global__ void aggregate_kernel(const double *__restrict__ x, int64_t n, const int32_t *__restrict__ f,
                                 double *__restrict__ buckets) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double v = x[i];
        if (v != 0.0) atomicAdd(buckets + f[i], v);
    }
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

// persistent CTAs: about four per SM in total (32 KB of shared memory each), split over the columns
dim3 apply_grid(int64_t nblk, int columns) {
    int64_t per = (148 * 4 + columns - 1) / columns;
    if (per > nblk) per = nblk;
    if (per < 1) per = 1;
    return dim3((unsigned)per, (unsigned)columns);
}

int run_rc(ehmm_session *s, const RCSrc &S, double p, int nrep, const double *uniforms, int64_t n_uniforms,
           int64_t *consumed) {
    const int64_t nblk = (s->M + RC_TB - 1) / RC_TB;
    const int64_t ncnt = nblk * S.columns;
    EHMM_TRY(s->rc_flag.reserve(sizeof(int) * ncnt));
    EHMM_TRY(s->rc_scan.reserve(sizeof(int64_t) * (ncnt + 1)));
    EHMM_TRY(s->rc_buckets.reserve(sizeof(double) * (size_t)S.columns * (s->G + 1)));
    EHMM_CK(cudaMemsetAsync(s->rc_buckets.p, 0, sizeof(double) * (size_t)S.columns * (s->G + 1), s->stream));
    dim3 grid((unsigned)nblk, (unsigned)S.columns);
    int64_t total = 0;
    if (p > 0.0) {
        PROF(s, KID_RC_COUNT), rc_count_kernel<<<grid, RC_NT, 0, s->stream>>>(S, p, nblk, s->rc_flag.as<int>());
        EHMM_CK(cudaGetLastError());
        PROF(s, KID_RC_SCAN), rc_scan_kernel<<<1, 1024, 0, s->stream>>>(s->rc_flag.as<int>(), ncnt, s->rc_scan.as<int64_t>());
        EHMM_CK(cudaGetLastError());
        EHMM_CK(cudaMemcpyAsync(s->h_pinned + 200, s->rc_scan.as<int64_t>() + ncnt, sizeof(int64_t), cudaMemcpyDeviceToHost,
                                s->stream));
        EHMM_CK(cudaStreamSynchronize(s->stream));
        memcpy(&total, s->h_pinned + 200, sizeof(int64_t));
    } else {
        EHMM_CK(cudaMemsetAsync(s->rc_scan.p, 0, sizeof(int64_t) * (ncnt + 1), s->stream));
    }
    const int32_t *grp = s->group.as<int32_t>();
    if (uniforms) {
        EHMM_REQUIRE(total <= n_uniforms, EHMM_ERR_RNG, "rejection control needs %lld uniforms, buffer holds %lld",
                     (long long)total, (long long)n_uniforms);
        EHMM_TRY(s->rc_unif.reserve(sizeof(double) * (size_t)(total > 0 ? total : 1)));
        if (total > 0)
            EHMM_CK(cudaMemcpyAsync(s->rc_unif.p, uniforms, sizeof(double) * total, cudaMemcpyHostToDevice, s->stream));
        PROF(s, KID_RC_APPLY), rc_apply_kernel<double><<<apply_grid(nblk, S.columns), RC_NT, 0, s->stream>>>(S, p, nblk, s->rc_scan.as<int64_t>(),
                                                               s->rc_unif.as<double>(), grp, nrep, s->G,
                                                               s->rc_buckets.as<double>(), nullptr);
    } else {
        EHMM_REQUIRE(s->rng_set || total == 0, EHMM_ERR_RNG,
                     "rejection control needs %lld uniforms: pass a buffer or call ehmm_rng_set_state", (long long)total);
        EHMM_TRY(s->rc_unif.reserve(sizeof(uint32_t) * (size_t)(total > 0 ? total : 1)));
        EHMM_TRY(s->mt_raw.reserve(sizeof(uint32_t) * 625));
        if (total > 0) {
            if (!s->rng_dev_valid) {
                memcpy(s->h_pinned + 512, s->rng_state, sizeof(uint32_t) * 625);
                EHMM_CK(cudaMemcpyAsync(s->mt_raw.p, s->h_pinned + 512, sizeof(uint32_t) * 625, cudaMemcpyHostToDevice, s->stream));
                s->rng_dev_valid = true;
                s->mtj_valid = 0;
                s->mtj_niv = 0;
                s->mtj_base = false;
            }
            EHMM_TRY(mt_generate(s, total, s->rc_unif.as<uint32_t>()));
            EHMM_CK(cudaMemcpyAsync(s->h_pinned + 512, s->mt_raw.p, sizeof(uint32_t) * 625, cudaMemcpyDeviceToHost, s->stream));
            rng_note_advance(s, total);
        }
        PROF(s, KID_RC_APPLY), rc_apply_kernel<uint32_t><<<apply_grid(nblk, S.columns), RC_NT, 0, s->stream>>>(S, p, nblk, s->rc_scan.as<int64_t>(),
                                                                 s->rc_unif.as<uint32_t>(), grp, nrep, s->G,
                                                                 s->rc_buckets.as<double>(), nullptr);
    }
    EHMM_CK(cudaGetLastError());
    // the next call's sub-stream windows depend only on the RNG state: prepare them on the side stream
    // (sized from this call's consumption with 25% headroom; a larger need falls back to preparing
    // the windows on the main stream)
    if (!uniforms && total > 0) {
        int64_t est = total + total / 4 + 2 * 131072;
        if (est > (int64_t)S.columns * s->M) est = (int64_t)S.columns * s->M;
        EHMM_TRY(mt_prefetch(s, est));
    }
    // the host does not wait for the tables: the next reader of the stream does
    s->rc_columns = S.columns;
    s->rc_N = nrep;
    if (consumed) *consumed = total;
    return EHMM_OK;
}

// ---- sharded form: the caller owns the ordering across bin blocks --------------------------------
// phase 1: flagged counts of this block per column (count + scan kernels; results stay in rc_scan)
int rc_count_block(ehmm_session *s, const RCSrc &S, double p, int64_t *counts) {
    const int64_t nblk = (s->M + RC_TB - 1) / RC_TB;
    const int64_t ncnt = nblk * S.columns;
    EHMM_REQUIRE(S.columns <= 64, EHMM_ERR_ARG, "too many columns");
    EHMM_TRY(s->rc_flag.reserve(sizeof(int) * ncnt));
    EHMM_TRY(s->rc_scan.reserve(sizeof(int64_t) * (ncnt + 1)));
    dim3 grid((unsigned)nblk, (unsigned)S.columns);
    if (p > 0.0) {
        PROF(s, KID_RC_COUNT), rc_count_kernel<<<grid, RC_NT, 0, s->stream>>>(S, p, nblk, s->rc_flag.as<int>());
        EHMM_CK(cudaGetLastError());
        PROF(s, KID_RC_SCAN), rc_scan_kernel<<<1, 1024, 0, s->stream>>>(s->rc_flag.as<int>(), ncnt, s->rc_scan.as<int64_t>());
        EHMM_CK(cudaGetLastError());
    } else {
        EHMM_CK(cudaMemsetAsync(s->rc_scan.p, 0, sizeof(int64_t) * (ncnt + 1), s->stream));
    }
    int64_t *hp = reinterpret_cast<int64_t *>(s->h_pinned + 200);
    for (int c = 0; c <= S.columns; c++)
        EHMM_CK(cudaMemcpyAsync(hp + c, s->rc_scan.as<int64_t>() + (int64_t)c * nblk, sizeof(int64_t), cudaMemcpyDeviceToHost,
                                s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    for (int c = 0; c < S.columns; c++) { counts[c] = hp[c + 1] - hp[c]; s->rc_colcount[c] = counts[c]; }
    s->rc_count_p = p;
    s->rc_count_cols = S.columns;
    return EHMM_OK;
}

// phase 2: this block's draws of column c start at stream offset col_offsets[c] (relative to the
// current RNG position); the stream advances by `total` (the draws of all blocks).
int rc_apply_block(ehmm_session *s, const RCSrc &S, double p, int nrep, const int64_t *col_offsets, int64_t total) {
    EHMM_REQUIRE(s->rc_count_cols == S.columns && s->rc_count_p == p, EHMM_ERR_STATE,
                 "rc_apply_at must follow rc_count with the same p");
    EHMM_REQUIRE(s->rng_set || total == 0, EHMM_ERR_RNG, "rejection control needs the RNG state (ehmm_rng_set_state)");
    const int64_t nblk = (s->M + RC_TB - 1) / RC_TB;
    int64_t local = 0;
    for (int c = 0; c < S.columns; c++) local += s->rc_colcount[c];
    EHMM_TRY(s->rc_buckets.reserve(sizeof(double) * (size_t)S.columns * (s->G + 1)));
    EHMM_CK(cudaMemsetAsync(s->rc_buckets.p, 0, sizeof(double) * (size_t)S.columns * (s->G + 1), s->stream));
    EHMM_TRY(s->rc_unif.reserve(sizeof(uint32_t) * (size_t)(local > 0 ? local : 1)));
    EHMM_TRY(s->mt_raw.reserve(sizeof(uint32_t) * 625));
    if (total > 0) {
        if (!s->rng_dev_valid) {
            memcpy(s->h_pinned + 512, s->rng_state, sizeof(uint32_t) * 625);
            EHMM_CK(cudaMemcpyAsync(s->mt_raw.p, s->h_pinned + 512, sizeof(uint32_t) * 625, cudaMemcpyHostToDevice, s->stream));
            s->rng_dev_valid = true;
            s->mtj_valid = 0;
            s->mtj_niv = 0;
            s->mtj_base = false;
        }
        int64_t at = 0;
        for (int c = 0; c < S.columns; c++) {
            EHMM_REQUIRE(col_offsets[c] >= 0 && col_offsets[c] + s->rc_colcount[c] <= total, EHMM_ERR_ARG,
                         "column %d: draws [%lld, %lld) outside [0, %lld)", c, (long long)col_offsets[c],
                         (long long)(col_offsets[c] + s->rc_colcount[c]), (long long)total);
            EHMM_TRY(mt_generate_range(s, col_offsets[c], s->rc_colcount[c], s->rc_unif.as<uint32_t>() + at));
            at += s->rc_colcount[c];
        }
    }
    dim3 grid((unsigned)nblk, (unsigned)S.columns);
    PROF(s, KID_RC_APPLY), rc_apply_kernel<uint32_t><<<apply_grid(nblk, S.columns), RC_NT, 0, s->stream>>>(
        S, p, nblk, s->rc_scan.as<int64_t>(), s->rc_unif.as<uint32_t>(), s->group.as<int32_t>(), nrep, s->G,
        s->rc_buckets.as<double>(), nullptr);
    EHMM_CK(cudaGetLastError());
    if (total > 0) {
        // side stream: the stream's state after all ranks' draws, then the windows of the next call's
        // slice, which sits where this one's did, give or take a few thousand draws
        EHMM_TRY(mt_advance_side(s, total, s->h_pinned + 512));
        rng_note_advance(s, total);
        EHMM_TRY(mt_prefetch_ranges(s, S.columns, col_offsets, s->rc_colcount, total, 2));
    }
    s->rc_columns = S.columns;
    s->rc_N = nrep;
    s->rc_count_cols = 0;
    return EHMM_OK;
}

}  // namespace

void rng_note_advance(ehmm_session *s, int64_t n) {
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

int rc_consensus(ehmm_session *s, double p, const double *uniforms, int64_t n_uniforms, int64_t *consumed) {
    RCSrc S = {s->logProb1.as<double>(), nullptr, s->M, 0, s->K};
    return run_rc(s, S, p, 1, uniforms, n_uniforms, consumed);
}

int rc_differential(ehmm_session *s, double p, int N, const double *uniforms, int64_t n_uniforms, int64_t *consumed) {
    EHMM_REQUIRE(s->K == 3 && s->B > 0, EHMM_ERR_ARG, "differentialRejectionControlled needs K=3 and mixture components");
    RCSrc S = {s->logProb1.as<double>(), s->eta.as<double>(), s->M, s->B, s->B + 2};
    return run_rc(s, S, p, N, uniforms, n_uniforms, consumed);
}

static RCSrc rc_source(ehmm_session *s, int differential) {
    if (differential) return RCSrc{s->logProb1.as<double>(), s->eta.as<double>(), s->M, s->B, s->B + 2};
    return RCSrc{s->logProb1.as<double>(), nullptr, s->M, 0, s->K};
}

int rc_count(ehmm_session *s, int differential, double p, int64_t *counts) {
    return rc_count_block(s, rc_source(s, differential), p, counts);
}

int rc_apply_at(ehmm_session *s, int differential, double p, const int64_t *col_offsets, int64_t total) {
    return rc_apply_block(s, rc_source(s, differential), p, differential ? s->N : 1, col_offsets, total);
}

static int fetch_column(ehmm_session *s, int column, std::vector<double> &h) {
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
    rc_scan_kernel<<<1, 1024>>>((const int *)dc.p, nblk, (int64_t *)doff.p);
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
    for (int64_t i = 0; i < n; i++)
        EHMM_REQUIRE(f[i] >= 0 && f[i] <= G, EHMM_ERR_ARG, "aggregate: group id %d at %lld outside [0, %lld]", f[i],
                     (long long)i, (long long)G);
    TmpDev dx, df, db;
    EHMM_TRY(dx.alloc(sizeof(double) * n));
    EHMM_TRY(df.alloc(sizeof(int32_t) * n));
    EHMM_TRY(db.alloc(sizeof(double) * (G + 1)));
    EHMM_CK(cudaMemcpy(dx.p, x, sizeof(double) * n, cudaMemcpyHostToDevice));
    EHMM_CK(cudaMemcpy(df.p, f, sizeof(int32_t) * n, cudaMemcpyHostToDevice));
    EHMM_CK(cudaMemset(db.p, 0, sizeof(double) * (G + 1)));
    if (n) aggregate_kernel<<<148 * 4, 256>>>((const double *)dx.p, n, (const int32_t *)df.p, (double *)db.p);
    EHMM_CK(cudaGetLastError());
    EHMM_CK(cudaMemcpy(buckets, db.p, sizeof(double) * (G + 1), cudaMemcpyDeviceToHost));
    return EHMM_OK;
}

}  // namespace ehmm
