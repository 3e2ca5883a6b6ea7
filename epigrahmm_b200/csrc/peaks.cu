// peaks.cu -- what callPeaks does with the fit's posteriors (R/callPeaks.R:61-68): the FDR thresholding of
// fdrControl (R/base-utils.R:11-25), the merge of adjacent called bins into peaks (GenomicRanges::reduce
// on the bins' ranges), and the hand-over of all datasets to the caller for the HDF5 file.
//
// fdrControl sorts notpp = 1 - P(enriched) ascending (data.table's order(): stable, ties stay in window
// order), forms cumsum(notpp) / 1..n and calls the windows where that running mean is below the level.
// On the device: a stable LSD radix sort of (bits of notpp, window) -- non-negative doubles order like
// their bit patterns -- and a prefix sum in sorted order.  R's cumsum accumulates in long double and stores
// each partial sum rounded to double, so the comparison sees two roundings; a plain fp64 prefix sum decides
// every window identically unless a running mean lies within the rounding error bounds of the level.  The
// kernel measures that margin for every window; if any is too small the whole column is redone by one
// thread that adds in emulated x87 arithmetic (ldsum.h) -- bit for bit R's numbers, just slowly.
#include "common.cuh"
#include "ldsum.h"
#include <algorithm>
#include <vector>

namespace ehmm {

namespace {

constexpr int RS_NT = 256, RS_IPT = 16, RS_TILE = RS_NT * RS_IPT, RS_NW = RS_NT / 32;

// notpp[j] = 1 - exp(logProb1[j, 1]) as the reference forms it (R/callPeaks.R:61, R/base-utils.R:18), its bit
// pattern as the sort key; bad: a probability outside [0, 1] (fdrControl stops on those)
__global__ void notpp_kernel(const double *__restrict__ lp1_col, int64_t M, uint64_t *__restrict__ keys, int *__restrict__ bad) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
        const double prob = exp(lp1_col[j]);
        if (!(prob >= 0.0 && prob <= 1.0)) atomicOr(bad, 1);
        keys[j] = (uint64_t)__double_as_longlong(1.0 - prob);
    }
}

__global__ void __launch_bounds__(RS_NT) rs_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, int shift,
                                                        unsigned *__restrict__ hist, int nblk) {
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_IPT; r++) {
        const int64_t i = base + r * RS_NT + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of `total` counters in place (one CTA): digit-major, block-minor = the order of the output
__global__ void __launch_bounds__(1024) rs_scan_kernel(unsigned *__restrict__ hist, int64_t total) {
    __shared__ unsigned sh[1024];
    const int tid = threadIdx.x;
    const int64_t q = (total + 1023) / 1024;
    const int64_t a = min(total, (int64_t)tid * q), b = min(total, a + q);
    unsigned s = 0;
    for (int64_t i = a; i < b; i++) s += hist[i];
    sh[tid] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const unsigned v = (tid >= d) ? sh[tid - d] : 0u;
        __syncthreads();
        sh[tid] += v;
        __syncthreads();
    }
    unsigned run = sh[tid] - s;
    for (int64_t i = a; i < b; i++) {
        const unsigned c = hist[i];
        hist[i] = run;
        run += c;
    }
}

// stable scatter of one digit: warp w owns 512 consecutive items of the tile, walked in 16 rounds of 32
__global__ void __launch_bounds__(RS_NT)
    rs_scatter_kernel(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin, uint64_t *__restrict__ kout,
                      uint32_t *__restrict__ vout, int64_t n, int shift, const unsigned *__restrict__ offs, int nblk) {
    __shared__ unsigned cnt[RS_NW][256];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    for (int i = tid; i < RS_NW * 256; i += RS_NT) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)w * (32 * RS_IPT);
    uint64_t k[RS_IPT];
    uint32_t v[RS_IPT];
#pragma unroll
    for (int r = 0; r < RS_IPT; r++) {
        const int64_t i = base + r * 32 + lane;
        const bool valid = i < n;
        k[r] = valid ? kin[i] : 0;
        v[r] = valid ? (vin ? vin[i] : (uint32_t)i) : 0u;
        const unsigned d = valid ? (unsigned)((k[r] >> shift) & 255u) : 256u;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid && (peers & lt) == 0) cnt[w][d] += __popc(peers);  // the lowest lane of each digit adds its group
        __syncwarp();
    }
    __syncthreads();
    {   // digit tid: where each warp's items of that digit start in the output
        unsigned run = offs[(size_t)tid * nblk + blockIdx.x];
        for (int ww = 0; ww < RS_NW; ww++) {
            const unsigned c = cnt[ww][tid];
            cnt[ww][tid] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_IPT; r++) {
        const int64_t i = base + r * 32 + lane;
        const bool valid = i < n;
        const unsigned d = valid ? (unsigned)((k[r] >> shift) & 255u) : 256u;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        unsigned pos = 0;
        if (valid) pos = cnt[w][d] + __popc(peers & lt);
        __syncwarp();
        if (valid && (peers & lt) == 0) cnt[w][d] += __popc(peers);
        __syncwarp();
        if (valid) { kout[pos] = k[r]; vout[pos] = v[r]; }
    }
}

// ---- prefix sum in sorted order + the comparison ------------------------------------------------------
__global__ void __launch_bounds__(RS_NT) ps_tile_kernel(const uint64_t *__restrict__ keys, int64_t n, double *__restrict__ tsum) {
    __shared__ double sh[RS_NW];
    const int64_t a = (int64_t)blockIdx.x * RS_TILE + (int64_t)threadIdx.x * RS_IPT;
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < RS_IPT; r++)
        if (a + r < n) s += __longlong_as_double((long long)keys[a + r]);
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = sh[0];
        for (int i = 1; i < RS_NW; i++) t += sh[i];
        tsum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) ps_scan_kernel(double *__restrict__ tsum, int64_t nblk) {
    __shared__ double sh[1024];
    const int tid = threadIdx.x;
    const int64_t q = (nblk + 1023) / 1024;
    const int64_t a = min(nblk, (int64_t)tid * q), b = min(nblk, a + q);
    double s = 0.0;
    for (int64_t i = a; i < b; i++) s += tsum[i];
    sh[tid] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const double v = (tid >= d) ? sh[tid - d] : 0.0;
        __syncthreads();
        sh[tid] += v;
        __syncthreads();
    }
    double run = sh[tid] - s;
    for (int64_t i = a; i < b; i++) {
        const double c = tsum[i];
        tsum[i] = run;
        run += c;
    }
}

struct FdrOut {
    uint8_t *calls;               // by window
    unsigned long long *ncalled;
    int *ambiguous;
    unsigned long long *last;     // largest sorted position called (+1)
};

// Relative error of everything that separates the fp64 running mean from R's: long double accumulation
// ((i) 2^-64 per term), its rounding to double and the division (2^-53 each), this kernel's own summation
// order (a few hundred roundings at most: 2^-43 is generous).
__device__ __forceinline__ double fdr_tolerance(int64_t i) { return (double)(i + 1) * 5.43e-20 + 1.2e-13; }

__global__ void __launch_bounds__(RS_NT)
    ps_apply_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ idx, int64_t n,
                    const double *__restrict__ tbase, double fdr, FdrOut O) {
    __shared__ double sh[RS_NT];
    __shared__ unsigned scount[RS_NW];
    const int tid = threadIdx.x;
    const int64_t a = (int64_t)blockIdx.x * RS_TILE + (int64_t)tid * RS_IPT;
    double x[RS_IPT], s = 0.0;
#pragma unroll
    for (int r = 0; r < RS_IPT; r++) {
        x[r] = (a + r < n) ? __longlong_as_double((long long)keys[a + r]) : 0.0;
        s += x[r];
    }
    sh[tid] = s;
    __syncthreads();
    for (int d = 1; d < RS_NT; d <<= 1) {
        const double v = (tid >= d) ? sh[tid - d] : 0.0;
        __syncthreads();
        sh[tid] += v;
        __syncthreads();
    }
    double run = tbase[blockIdx.x] + (sh[tid] - s);
    unsigned called = 0;
    bool amb = false;
    int64_t last = 0;
#pragma unroll
    for (int r = 0; r < RS_IPT; r++) {
        const int64_t i = a + r;
        if (i < n) {
            run += x[r];
            const double mean = run / (double)(i + 1);
            const bool c = mean < fdr;
            amb = amb || (fabs(mean - fdr) <= fdr_tolerance(i) * fdr);
            O.calls[idx[i]] = c ? 1 : 0;
            called += c;
            if (c) last = i + 1;
        }
    }
    if (amb) atomicOr(O.ambiguous, 1);
    if (last) atomicMax(O.last, (unsigned long long)last);
    for (int d = 16; d > 0; d >>= 1) called += __shfl_down_sync(0xffffffffu, called, d);
    if ((tid & 31) == 0) scount[tid >> 5] = called;
    __syncthreads();
    if (tid == 0) {
        unsigned t = 0;
        for (int i = 0; i < RS_NW; i++) t += scount[i];
        if (t) atomicAdd(O.ncalled, (unsigned long long)t);
    }
}

// R's arithmetic, one addition after the other
__global__ void fdr_exact_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ idx, int64_t n, double fdr, FdrOut O) {
    LdAcc acc = {0, 0};
    unsigned long long called = 0, last = 0;
    for (int64_t i = 0; i < n; i++) {
        acc = ld_add(acc, ld_from_double(__longlong_as_double((long long)keys[i])));
        const bool c = __ddiv_rn(ld_to_double(acc), (double)(i + 1)) < fdr;
        O.calls[idx[i]] = c ? 1 : 0;
        called += c;
        if (c) last = (unsigned long long)i + 1;
    }
    *O.ncalled = called;
    *O.last = last;
}

__global__ void viterbi_calls_kernel(const double *__restrict__ vit, int64_t M, uint8_t *__restrict__ calls,
                                     unsigned long long *__restrict__ ncalled) {
    unsigned c = 0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < M; j += (int64_t)gridDim.x * blockDim.x) {
        const bool p = vit[j] == 1.0;  // R/callPeaks.R:62
        calls[j] = p;
        c += p;
    }
    for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(ncalled, (unsigned long long)c);
}

// ---- peaks = maximal runs of called bins inside one chromosome ----------------------------------------
__device__ __forceinline__ bool is_break(const int64_t *__restrict__ breaks, int nb, int64_t j) {
    int lo = 0, hi = nb;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (breaks[mid] < j) lo = mid + 1; else hi = mid;
    }
    return lo < nb && breaks[lo] == j;
}
__device__ __forceinline__ void run_edges(const uint8_t *__restrict__ calls, int64_t M, const int64_t *__restrict__ breaks, int nb,
                                          int64_t j, bool &start, bool &end) {
    const bool c = calls[j] != 0;
    start = c && (j == 0 || !calls[j - 1] || is_break(breaks, nb, j));
    end = c && (j == M - 1 || !calls[j + 1] || is_break(breaks, nb, j + 1));
}

__global__ void __launch_bounds__(RS_NT) runs_count_kernel(const uint8_t *__restrict__ calls, int64_t M, const int64_t *__restrict__ breaks,
                                                           int nb, unsigned *__restrict__ cnt) {
    __shared__ unsigned sh[RS_NW];
    const int64_t a = (int64_t)blockIdx.x * RS_TILE + (int64_t)threadIdx.x * RS_IPT;
    unsigned c = 0;
    for (int r = 0; r < RS_IPT; r++)
        if (a + r < M) {
            bool st, en;
            run_edges(calls, M, breaks, nb, a + r, st, en);
            c += st;
        }
    for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int i = 0; i < RS_NW; i++) t += sh[i];
        cnt[blockIdx.x] = t;
    }
}

// runs are written in ascending order: a run's start and its end have the same rank among starts / ends
__global__ void __launch_bounds__(RS_NT) runs_write_kernel(const uint8_t *__restrict__ calls, int64_t M, const int64_t *__restrict__ breaks,
                                                           int nb, const unsigned *__restrict__ off, int64_t cap,
                                                           int64_t *__restrict__ starts, int64_t *__restrict__ ends,
                                                           const unsigned *__restrict__ end_off) {
    __shared__ unsigned shs[RS_NT], she[RS_NT];
    const int tid = threadIdx.x;
    const int64_t a = (int64_t)blockIdx.x * RS_TILE + (int64_t)tid * RS_IPT;
    unsigned ns = 0, ne = 0;
    unsigned ms = 0, me = 0;
    for (int r = 0; r < RS_IPT; r++)
        if (a + r < M) {
            bool st, en;
            run_edges(calls, M, breaks, nb, a + r, st, en);
            if (st) { ms |= 1u << r; ns++; }
            if (en) { me |= 1u << r; ne++; }
        }
    shs[tid] = ns;
    she[tid] = ne;
    __syncthreads();
    for (int d = 1; d < RS_NT; d <<= 1) {
        const unsigned vs = (tid >= d) ? shs[tid - d] : 0u, ve = (tid >= d) ? she[tid - d] : 0u;
        __syncthreads();
        shs[tid] += vs;
        she[tid] += ve;
        __syncthreads();
    }
    int64_t ps = (int64_t)off[blockIdx.x] + (shs[tid] - ns), pe = (int64_t)end_off[blockIdx.x] + (she[tid] - ne);
    for (int r = 0; r < RS_IPT; r++) {
        if ((ms >> r) & 1u) { if (ps < cap) starts[ps] = a + r; ps++; }
        if ((me >> r) & 1u) { if (pe < cap) ends[pe] = a + r; pe++; }
    }
}

__global__ void __launch_bounds__(RS_NT) ends_count_kernel(const uint8_t *__restrict__ calls, int64_t M, const int64_t *__restrict__ breaks,
                                                           int nb, unsigned *__restrict__ cnt) {
    __shared__ unsigned sh[RS_NW];
    const int64_t a = (int64_t)blockIdx.x * RS_TILE + (int64_t)threadIdx.x * RS_IPT;
    unsigned c = 0;
    for (int r = 0; r < RS_IPT; r++)
        if (a + r < M) {
            bool st, en;
            run_edges(calls, M, breaks, nb, a + r, st, en);
            c += en;
        }
    for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int i = 0; i < RS_NW; i++) t += sh[i];
        cnt[blockIdx.x] = t;
    }
}

struct PeakBufs {
    uint64_t *k0, *k1;
    uint32_t *v0, *v1;
    unsigned *hist;
    double *tsum;
    unsigned long long *scal;  // [0] called, [1] last position, then int ambiguous, int bad
};

int reserve_peaks(ehmm_session *s, PeakBufs &B) {
    const int64_t M = s->M;
    const int64_t nblk = (M + RS_TILE - 1) / RS_TILE;
    const size_t kb = sizeof(uint64_t) * (size_t)M, vb = (sizeof(uint32_t) * (size_t)M + 7) / 8 * 8;
    const size_t hb = sizeof(unsigned) * 256 * (size_t)nblk, tb = sizeof(double) * (size_t)nblk;
    EHMM_TRY(s->peak_work.reserve(2 * kb + 2 * vb + hb + tb + 64));
    EHMM_TRY(s->peak_calls.reserve((size_t)M + 8));
    char *p = s->peak_work.as<char>();
    B.k0 = reinterpret_cast<uint64_t *>(p); p += kb;
    B.k1 = reinterpret_cast<uint64_t *>(p); p += kb;
    B.v0 = reinterpret_cast<uint32_t *>(p); p += vb;
    B.v1 = reinterpret_cast<uint32_t *>(p); p += vb;
    B.hist = reinterpret_cast<unsigned *>(p); p += hb;
    B.tsum = reinterpret_cast<double *>(p); p += tb;
    B.scal = reinterpret_cast<unsigned long long *>(p);
    return EHMM_OK;
}

}  // namespace

// Stable LSD radix sort of n (64-bit key, 32-bit value) pairs on the session stream: 8 passes of 8 bits, ping-pong
// between (k0, v0) and (k1, v1); the keys start in k0, the values in vinit (nullptr: the indices 0..n-1) -- vinit may be
// v1 but not v0.  hist: 256 * ceil(n / 4096) counters.  Also used to number dt$Group by first appearance (groups.cu).
int radix_sort_pairs(ehmm_session *s, uint64_t *k0, uint64_t *k1, uint32_t *v0, uint32_t *v1, unsigned *hist, int64_t n,
                     const uint32_t *vinit, uint64_t **k_sorted, uint32_t **v_sorted) {
    const int nblk = (int)((n + RS_TILE - 1) / RS_TILE);
    uint64_t *kin = k0, *kout = k1;
    const uint32_t *vin = vinit;
    uint32_t *vout = v0, *vother = v1;
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 8 * pass;
        PROF(s, KID_PEAKS), rs_hist_kernel<<<nblk, RS_NT, 0, s->stream>>>(kin, n, shift, hist, nblk);
        PROF(s, KID_PEAKS), rs_scan_kernel<<<1, 1024, 0, s->stream>>>(hist, (int64_t)256 * nblk);
        PROF(s, KID_PEAKS), rs_scatter_kernel<<<nblk, RS_NT, 0, s->stream>>>(kin, vin, kout, vout, n, shift, hist, nblk);
        EHMM_CK(cudaGetLastError());
        std::swap(kin, kout);
        vin = vout;
        std::swap(vout, vother);
    }
    *k_sorted = kin;                          // back in k0 after an even number of passes
    *v_sorted = const_cast<uint32_t *>(vin);  // the buffer written last
    return EHMM_OK;
}

// fdrControl(prob = exp(logProb1[, 2]), fdr) -> calls[M] (0/1, by window; also kept on the device for peak_runs)
int call_peaks_fdr(ehmm_session *s, double fdr, uint8_t *calls_out, int64_t *n_called, double *cutoff) {
    EHMM_REQUIRE(fdr > 0.0 && fdr < 1.0, EHMM_ERR_ARG, "fdr must be between 0 and 1");  // R/base-utils.R:15-16
    EHMM_REQUIRE(s->M < ((int64_t)1 << 32), EHMM_ERR_ARG, "callPeaks: at most 2^32 windows per block");
    EHMM_REQUIRE(s->m0 == 0 && s->Mtot == s->M, EHMM_ERR_STATE, "callPeaks ranks the windows of the whole chain: gather the blocks' logProb1 first");
    const int64_t M = s->M;
    const int nblk = (int)((M + RS_TILE - 1) / RS_TILE);
    PeakBufs B;
    EHMM_TRY(reserve_peaks(s, B));
    int *flags = reinterpret_cast<int *>(B.scal + 2);
    EHMM_CK(cudaMemsetAsync(B.scal, 0, 32, s->stream));
    PROF(s, KID_PEAKS), notpp_kernel<<<148 * 8, 256, 0, s->stream>>>(s->logProb1.as<double>() + M, M, B.k0, flags + 1);
    EHMM_CK(cudaGetLastError());
    uint64_t *kin = nullptr;
    uint32_t *vin = nullptr;
    EHMM_TRY(radix_sort_pairs(s, B.k0, B.k1, B.v0, B.v1, B.hist, M, nullptr, &kin, &vin));
    // after 8 passes the sorted keys are back in k0 and the windows in the buffer written last
    const FdrOut O = {s->peak_calls.as<uint8_t>(), B.scal, flags, B.scal + 1};
    if (s->peaks_exact) {
        PROF(s, KID_PEAKS), fdr_exact_kernel<<<1, 1, 0, s->stream>>>(kin, vin, M, fdr, O);
    } else {
        PROF(s, KID_PEAKS), ps_tile_kernel<<<nblk, RS_NT, 0, s->stream>>>(kin, M, B.tsum);
        PROF(s, KID_PEAKS), ps_scan_kernel<<<1, 1024, 0, s->stream>>>(B.tsum, nblk);
        PROF(s, KID_PEAKS), ps_apply_kernel<<<nblk, RS_NT, 0, s->stream>>>(kin, vin, M, B.tsum, fdr, O);
    }
    EHMM_CK(cudaGetLastError());
    unsigned long long *h = reinterpret_cast<unsigned long long *>(s->h_pinned + 400);
    EHMM_CK(cudaMemcpyAsync(h, B.scal, 32, cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    const int *hf = reinterpret_cast<const int *>(h + 2);
    EHMM_REQUIRE(hf[1] == 0, EHMM_ERR_ARG, "Posterior probabilities must be between 0 and 1");
    s->peaks_path = s->peaks_exact ? 1 : 0;
    if (!s->peaks_exact && hf[0]) {  // a running mean within rounding distance of the level: R's own arithmetic decides
        EHMM_CK(cudaMemsetAsync(B.scal, 0, 32, s->stream));
        PROF(s, KID_PEAKS), fdr_exact_kernel<<<1, 1, 0, s->stream>>>(kin, vin, M, fdr, O);
        EHMM_CK(cudaGetLastError());
        EHMM_CK(cudaMemcpyAsync(h, B.scal, 32, cudaMemcpyDeviceToHost, s->stream));
        EHMM_CK(cudaStreamSynchronize(s->stream));
        s->peaks_path = 2;
    }
    if (n_called) *n_called = (int64_t)h[0];
    if (cutoff) {  // the largest notpp among the called windows (0 when none is called)
        *cutoff = 0.0;
        if (h[1] > 0) {
            uint64_t kb = 0;
            EHMM_CK(cudaMemcpy(&kb, kin + (h[1] - 1), sizeof(uint64_t), cudaMemcpyDeviceToHost));
            memcpy(cutoff, &kb, sizeof(double));
        }
    }
    if (calls_out) {
        EHMM_CK(cudaMemcpyAsync(calls_out, s->peak_calls.p, (size_t)M, cudaMemcpyDeviceToHost, s->stream));
        EHMM_CK(cudaStreamSynchronize(s->stream));
    }
    s->peaks_valid = true;
    return EHMM_OK;
}

// method = 'viterbi' (R/callPeaks.R:62): the windows whose Viterbi state is 1
int call_peaks_viterbi(ehmm_session *s, uint8_t *calls_out, int64_t *n_called) {
    const int64_t M = s->M;
    PeakBufs B;
    EHMM_TRY(reserve_peaks(s, B));
    EHMM_CK(cudaMemsetAsync(B.scal, 0, 32, s->stream));
    PROF(s, KID_PEAKS), viterbi_calls_kernel<<<148 * 8, 256, 0, s->stream>>>(s->viterbi.as<double>(), M, s->peak_calls.as<uint8_t>(), B.scal);
    EHMM_CK(cudaGetLastError());
    unsigned long long *h = reinterpret_cast<unsigned long long *>(s->h_pinned + 400);
    EHMM_CK(cudaMemcpyAsync(h, B.scal, 8, cudaMemcpyDeviceToHost, s->stream));
    if (calls_out) EHMM_CK(cudaMemcpyAsync(calls_out, s->peak_calls.p, (size_t)M, cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    if (n_called) *n_called = (int64_t)h[0];
    s->peaks_valid = true;
    return EHMM_OK;
}

// maximal runs of called windows that do not cross a chromosome start (breaks: ascending window indices at
// which a new chromosome begins): what GenomicRanges::reduce(rowRanges(object)[peakindex]) merges
// (R/callPeaks.R:67-68; the bins tile each chromosome).  starts / ends: 0-based, inclusive.
int peak_runs(ehmm_session *s, const int64_t *breaks, int nbreaks, int64_t cap, int64_t *starts, int64_t *ends, int64_t *nruns) {
    EHMM_REQUIRE(s->peaks_valid, EHMM_ERR_STATE, "peak_runs: callPeaks has not run");
    const int64_t M = s->M;
    const int nblk = (int)((M + RS_TILE - 1) / RS_TILE);
    for (int i = 1; i < nbreaks; i++) EHMM_REQUIRE(breaks[i] > breaks[i - 1], EHMM_ERR_ARG, "peak_runs: chromosome starts must ascend");
    EHMM_TRY(s->misc.reserve(sizeof(int64_t) * (size_t)(nbreaks + 1) + 2 * sizeof(unsigned) * (size_t)(nblk + 1) + 64));
    int64_t *dbreaks = s->misc.as<int64_t>();
    unsigned *cs = reinterpret_cast<unsigned *>(dbreaks + nbreaks + 1), *ce = cs + nblk + 1;
    if (nbreaks) EHMM_CK(cudaMemcpyAsync(dbreaks, breaks, sizeof(int64_t) * nbreaks, cudaMemcpyHostToDevice, s->stream));
    const uint8_t *calls = s->peak_calls.as<uint8_t>();
    PROF(s, KID_PEAKS), runs_count_kernel<<<nblk, RS_NT, 0, s->stream>>>(calls, M, dbreaks, nbreaks, cs);
    PROF(s, KID_PEAKS), ends_count_kernel<<<nblk, RS_NT, 0, s->stream>>>(calls, M, dbreaks, nbreaks, ce);
    // total = last offset + last count: read before the in-place scans
    std::vector<unsigned> hc((size_t)nblk);
    EHMM_CK(cudaMemcpyAsync(hc.data(), cs, sizeof(unsigned) * nblk, cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    int64_t total = 0;
    for (unsigned c : hc) total += c;
    *nruns = total;
    if (total == 0 || cap == 0 || !starts || !ends) return EHMM_OK;
    PROF(s, KID_PEAKS), rs_scan_kernel<<<1, 1024, 0, s->stream>>>(cs, nblk);
    PROF(s, KID_PEAKS), rs_scan_kernel<<<1, 1024, 0, s->stream>>>(ce, nblk);
    const int64_t n = std::min(cap, total);
    EHMM_TRY(s->rc_tmp.reserve(2 * sizeof(int64_t) * (size_t)n));
    int64_t *ds = s->rc_tmp.as<int64_t>(), *de = ds + n;
    PROF(s, KID_PEAKS), runs_write_kernel<<<nblk, RS_NT, 0, s->stream>>>(calls, M, dbreaks, nbreaks, cs, n, ds, de, ce);
    EHMM_CK(cudaGetLastError());
    EHMM_CK(cudaMemcpyAsync(starts, ds, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaMemcpyAsync(ends, de, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    return EHMM_OK;
}

}  // namespace ehmm
