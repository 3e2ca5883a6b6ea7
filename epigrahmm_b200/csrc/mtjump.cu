// mtjump.cu -- R's Mersenne-Twister stream generated in parallel on the device.
//
// rbinom() in the reference consumes unif_rand() draws one by one (src/rbinomVectorized.cpp:17);
// at 15 M bins that is ~1.5e7 draws per EM iteration and a single sequential MT19937 is the
// slowest step of the iteration.  MT19937 is GF(2)-linear, so sub-stream s (words
// [1 + s*J, 1 + (s+1)*J) of the stream, J = 2^17) can start from a window obtained by jumping
// ahead: w[m + J*2^k] = XOR_{i : c_i = 1} w[m + i] with c = x^(J*2^k) mod phi (mtpoly.h).
//   mt_jump_kernel     : round k doubles the number of known windows (window 2^k + c from window c)
//   mt_generate_kernel : one CTA per sub-stream runs the recurrence (227-wide) and writes tempered
//                        words; the CTA that reaches the end also writes R's next .Random.seed
// The windows depend only on the RNG state, not on the data, so they are prepared on the side
// stream while the E-step runs.
#include "common.cuh"
#include "mtpoly.h"
#include <algorithm>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

namespace ehmm {

namespace {

constexpr int LOG2J = 17;
constexpr int64_t JW = (int64_t)1 << LOG2J;  // words per sub-stream
constexpr int LEVELS = 15;                   // up to 2^15 sub-streams = 4.3e9 draws per call
constexpr int MT_N = 624, MT_M = 397, MT_D = MT_N - MT_M;  // 227
constexpr int EXPW = mtpoly::DEG + MT_N - 1;               // words a jump needs (20560)
constexpr int JUMP_NT = 640;                                // one output word per thread (624 used)
constexpr int JUMP_SPLIT = 8;                              // CTAs per window, each takes a slice of the coefficient list
constexpr int ZERO_IDX = EXPW + 8;                         // index whose word is 0 for every output lane
constexpr int BUFW = ZERO_IDX + MT_N + 8;                  // expanded words + zeroed tail
constexpr int MAX_IDX = 20480;                             // index-list capacity per level (uint16)
constexpr int GEN_NT = 256;

__device__ __forceinline__ uint32_t twist(uint32_t a, uint32_t b) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// window 0 = absolute words [1, 625) of the stream whose array `state` holds words [0, 624)
__global__ void mt_base_kernel(const uint32_t *__restrict__ state, uint32_t *__restrict__ S) {
    for (int i = threadIdx.x; i < MT_N; i += blockDim.x)
        S[i] = (i < MT_N - 1) ? state[1 + i + 1] : (state[1 + MT_M] ^ twist(state[1], state[2]));
}

// dst[c] ^= partial jump(src[c]) by the polynomial whose set coefficients are listed in `idx`
// (nidx entries, padded to a multiple of 8 with ZERO_IDX, which addresses a zeroed region).  The
// coefficient list is split over gridDim.y CTAs (a window is latency-bound in one CTA and the
// doubling rounds are serial); the destination windows are zeroed beforehand and accumulate by XOR.
__device__ __forceinline__ void jump_body(const uint16_t *__restrict__ idx, int nidx, const uint32_t *__restrict__ src,
                                          uint32_t *__restrict__ dst, uint32_t *sm) {
    uint32_t *buf = sm;                                                   // EXPW words + zero pad
    uint16_t *sidx = reinterpret_cast<uint16_t *>(sm + BUFW);             // this CTA's slice of the list
    const int tid = threadIdx.x;
    const int per = ((nidx / 8 + gridDim.y - 1) / gridDim.y) * 8;         // entries per slice (multiple of 8)
    const int q0 = min(nidx, (int)blockIdx.y * per), q1 = min(nidx, q0 + per);
    for (int i = tid; i < MT_N; i += JUMP_NT) buf[i] = src[i];
    for (int i = EXPW + tid; i < BUFW; i += JUMP_NT) buf[i] = 0u;
    for (int i = tid; i < (q1 - q0) / 2; i += JUMP_NT)
        reinterpret_cast<uint32_t *>(sidx)[i] = reinterpret_cast<const uint32_t *>(idx + q0)[i];
    __syncthreads();
    for (int b0 = MT_N; b0 < EXPW; b0 += MT_D) {
        const int i = b0 + tid;
        if (tid < MT_D && i < EXPW) buf[i] = buf[i - MT_D] ^ twist(buf[i - MT_N], buf[i - MT_N + 1]);
        __syncthreads();
    }
    if (tid < MT_N) {
        uint32_t a0 = 0;
        const uint32_t *bp = buf + tid;
        for (int q = 0; q < q1 - q0; q += 8) {
            const uint4 pk = *reinterpret_cast<const uint4 *>(sidx + q);  // 8 indices, broadcast
            const unsigned w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
            for (int j = 0; j < 4; j++) a0 ^= bp[w[j] & 0xffffu] ^ bp[w[j] >> 16];
        }
        if (a0) atomicXor(dst + tid, a0);
    }
}

__global__ void __launch_bounds__(JUMP_NT) mt_jump_kernel(const uint16_t *__restrict__ idx, int nidx,
                                                          const uint32_t *__restrict__ srcw, uint32_t *__restrict__ dstw) {
    extern __shared__ uint32_t sm[];
    // window blockIdx.x of the source run jumps to window blockIdx.x of the destination run
    jump_body(idx, nidx, srcw + (size_t)blockIdx.x * MT_N, dstw + (size_t)blockIdx.x * MT_N, sm);
}

// several runs of windows in one launch, each with its own polynomial: entry e takes windows src[e] + c to dst[e] + c
// (c < cta0[e + 1] - cta0[e]) by the polynomial of level lv[e].  The windows a bin block needs lie in several
// separate stretches of the stream (one per column of the rejection control); prepared one stretch after the
// other, the ~100 dependent launches of a call took longer than the call's own kernels.
constexpr int JB_MAX = 72;
struct JumpBatch {
    int n;
    int cta0[JB_MAX + 1];
    int src[JB_MAX], dst[JB_MAX], lv[JB_MAX], nidx[JB_MAX];
};
__global__ void __launch_bounds__(JUMP_NT) mt_jump_multi_kernel(const uint16_t *__restrict__ idx_all, const __grid_constant__ JumpBatch B,
                                                                uint32_t *__restrict__ S) {
    extern __shared__ uint32_t sm[];
    int e = 0;
    while (e + 1 < B.n && (int)blockIdx.x >= B.cta0[e + 1]) e++;
    const int c = (int)blockIdx.x - B.cta0[e];
    jump_body(idx_all + (size_t)B.lv[e] * MAX_IDX, B.nidx[e], S + (size_t)(B.src[e] + c) * MT_N, S + (size_t)(B.dst[e] + c) * MT_N, sm);
}

struct ZeroRuns {
    int n;
    int first[JB_MAX], count[JB_MAX];
};
__global__ void mt_zero_runs_kernel(const __grid_constant__ ZeroRuns Z, uint32_t *__restrict__ S) {
    for (int r = 0; r < Z.n; r++) {
        uint32_t *p = S + (size_t)Z.first[r] * MT_N;
        const int64_t n = (int64_t)Z.count[r] * MT_N;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0u;
    }
}

// One CTA walks sub-stream s.  It writes the tempered words of the absolute range
// [pos + lo_rel, pos + hi_rel) that fall into its sub-stream (out[a - (pos + lo_rel)]) and, when
// n_state > 0, the raw words of R's state array after n_state draws (the aligned 624-word block
// containing the last drawn word, and mti).
__device__ __forceinline__ void gen_substream(const uint32_t *__restrict__ state, const uint32_t *__restrict__ S, int s, int64_t lo_rel,
                                              int64_t hi_rel, uint32_t *__restrict__ out, int64_t n_state,
                                              uint32_t *__restrict__ state_out, int64_t jw, uint32_t (*buf)[MT_N]) {
    const int tid = threadIdx.x;
    const int64_t pos = state[0];
    const int64_t lo_abs = pos + lo_rel, hi_abs = pos + hi_rel;
    const int64_t a0 = (s == 0) ? 0 : 1 + (int64_t)s * jw;
    const int64_t nxt = 1 + (int64_t)(s + 1) * jw;
    const int64_t lo = max(lo_abs, a0), hi = min(hi_abs, nxt);
    int64_t gen_end = (lo < hi) ? hi : a0;
    int64_t st_lo = 0, st_hi = 0, st_end = 0;  // this CTA writes state words [max(a0, st_lo), st_end)
    bool last = false;
    if (n_state > 0) {
        const int64_t E = pos + n_state;
        st_lo = ((E - 1) / MT_N) * MT_N;
        st_hi = st_lo + MT_N;
        last = (E > a0 && E <= nxt);
        if (last) st_end = st_hi;
        else if (nxt < E) st_end = min(nxt, st_hi);
        if (st_end > max(a0, st_lo)) gen_end = max(gen_end, st_end);
        else st_end = 0;
        if (last && tid == 0) state_out[0] = (uint32_t)(E - st_lo);
    }
    if (gen_end <= a0) return;
    const uint32_t *src = (s == 0) ? state + 1 : S + (size_t)s * MT_N;
    for (int i = tid; i < MT_N; i += GEN_NT) buf[0][i] = src[i];
    __syncthreads();
    int cur = 0;
    for (int64_t cb = a0;; cb += MT_N) {
        const uint32_t *bc = buf[cur];
        if (cb + MT_N > lo || cb + MT_N > st_lo) {
            for (int kk = tid; kk < MT_N; kk += GEN_NT) {
                const int64_t a = cb + kk;
                const uint32_t v = bc[kk];
                if (a >= lo && a < hi) out[a - lo_abs] = temper(v);
                if (a >= st_lo && a < st_end) state_out[1 + (a - st_lo)] = v;
            }
        }
        if (cb + MT_N >= gen_end) break;
        uint32_t *nw = buf[cur ^ 1];
        for (int kk = tid; kk < MT_D; kk += GEN_NT) nw[kk] = bc[kk + MT_M] ^ twist(bc[kk], bc[kk + 1]);
        __syncthreads();
        for (int kk = MT_D + tid; kk < 2 * MT_D; kk += GEN_NT) nw[kk] = nw[kk - MT_D] ^ twist(bc[kk], bc[kk + 1]);
        __syncthreads();
        for (int kk = 2 * MT_D + tid; kk < MT_N; kk += GEN_NT)
            nw[kk] = nw[kk - MT_D] ^ ((kk == MT_N - 1) ? twist(bc[kk], nw[0]) : twist(bc[kk], bc[kk + 1]));
        __syncthreads();
        cur ^= 1;
    }
}

// CTA b handles sub-stream s_first + b of one range
__global__ void __launch_bounds__(GEN_NT)
    mt_generate_kernel(const uint32_t *__restrict__ state, const uint32_t *__restrict__ S, int s_first, int64_t lo_rel,
                       int64_t hi_rel, uint32_t *__restrict__ out, int64_t n_state, uint32_t *__restrict__ state_out,
                       int64_t jw) {
    __shared__ uint32_t buf[2][MT_N];
    gen_substream(state, S, s_first + (int)blockIdx.x, lo_rel, hi_rel, out, n_state, state_out, jw, buf);
}

// several ranges of the stream in one launch (a bin block's slices of the rejection control's columns): CTAs
// [cta0[r], cta0[r + 1]) walk the sub-streams of range r, whose words go to out + out_off[r]
constexpr int GEN_MAXR = 64;
struct GenRanges {
    int n;
    int cta0[GEN_MAXR + 1];
    int s_first[GEN_MAXR];
    int64_t lo_rel[GEN_MAXR], hi_rel[GEN_MAXR], out_off[GEN_MAXR];
};

__global__ void __launch_bounds__(GEN_NT)
    mt_generate_multi_kernel(const uint32_t *__restrict__ state, const uint32_t *__restrict__ S, const __grid_constant__ GenRanges R,
                             uint32_t *__restrict__ out, int64_t jw) {
    __shared__ uint32_t buf[2][MT_N];
    int r = 0;
    while (r + 1 < R.n && (int)blockIdx.x >= R.cta0[r + 1]) r++;
    gen_substream(state, S, R.s_first[r] + ((int)blockIdx.x - R.cta0[r]), R.lo_rel[r], R.hi_rel[r], out + R.out_off[r], -1, nullptr, jw, buf);
}

// jump polynomials: computed once per process on the host, uploaded once per device
std::mutex g_mu;
std::vector<uint16_t> g_idx;    // LEVELS * MAX_IDX coefficient positions
int g_nidx[LEVELS] = {0};
std::map<int, uint16_t *> g_dev_idx;
// expanded words + this CTA's slice of the coefficient list (84.8 KB + 5.1 KB: two CTAs per SM)
constexpr size_t JUMP_SMEM = sizeof(uint32_t) * BUFW + sizeof(uint16_t) * (MAX_IDX / JUMP_SPLIT + 16);

int device_polys(int device, const uint16_t **out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_idx.empty()) {
        std::vector<mtpoly::Poly> polys = mtpoly::jump_polys(LOG2J, LEVELS);
        EHMM_REQUIRE((int)polys.size() == LEVELS, EHMM_ERR_RNG, "MT19937 characteristic polynomial has the wrong degree");
        g_idx.assign((size_t)LEVELS * MAX_IDX, (uint16_t)ZERO_IDX);
        for (int k = 0; k < LEVELS; k++) {
            int n = 0;
            for (int i = 0; i < mtpoly::DEG; i++)
                if (mtpoly::get(polys[k], i)) {
                    EHMM_REQUIRE(n < MAX_IDX, EHMM_ERR_RNG, "jump polynomial denser than expected");
                    g_idx[(size_t)k * MAX_IDX + n++] = (uint16_t)i;
                }
            g_nidx[k] = (n + 7) & ~7;
        }
    }
    auto it = g_dev_idx.find(device);
    if (it == g_dev_idx.end()) {
        uint16_t *d = nullptr;
        EHMM_CK(cudaMalloc(&d, sizeof(uint16_t) * g_idx.size()));
        EHMM_CK(cudaMemcpy(d, g_idx.data(), sizeof(uint16_t) * g_idx.size(), cudaMemcpyHostToDevice));
        EHMM_CK(cudaFuncSetAttribute(mt_jump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JUMP_SMEM));
        EHMM_CK(cudaFuncSetAttribute(mt_jump_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)JUMP_SMEM));
        it = g_dev_idx.emplace(device, d).first;
    }
    *out = it->second;
    return EHMM_OK;
}

// Sub-streams are JW << mt_shift words long: a call that draws 1e8 uniforms would otherwise need 800
// windows, and the jump that produces a window costs as much as generating a fifth of a sub-stream.  The
// squaring chain x^(JW * 2^k) already holds the polynomials of the longer strides (level k + mt_shift).
inline int64_t jw_of(const ehmm_session *s) { return JW << s->mt_shift; }

int64_t substreams_for(const ehmm_session *s, int64_t pos, int64_t n) {  // sub-streams touched by draws [pos, pos + n)
    if (n <= 0) return 0;
    const int64_t last = pos + n - 1;
    return (last <= 0 ? 0 : (last - 1) / jw_of(s)) + 1;
}

}  // namespace

namespace {

int launch_jump(ehmm_session *s, cudaStream_t st, const uint16_t *idx, int k, const uint32_t *src, uint32_t *dst, int64_t cnt) {
    EHMM_CK(cudaMemsetAsync(dst, 0, sizeof(uint32_t) * MT_N * (size_t)cnt, st));
    const int lv = k + s->mt_shift;  // level of the polynomial x^(jw_of(s) * 2^k)
    EHMM_REQUIRE(lv < LEVELS, EHMM_ERR_RNG, "draw position beyond the prepared jump polynomials");
    mt_jump_kernel<<<dim3((unsigned)cnt, JUMP_SPLIT), JUMP_NT, JUMP_SMEM, st>>>(idx + (size_t)lv * MAX_IDX, g_nidx[lv], src, dst);
    EHMM_CK(cudaGetLastError());
    s->launches[KID_MT]++;
    return EHMM_OK;
}

// window buffer for sub-streams [0, P) plus scratch windows (two for mt_prepare_range, LEVELS per run for
// mt_prepare_ranges); growing it invalidates what it held
constexpr int SCRATCH_W = 2 + JB_MAX * LEVELS;
int reserve_windows(ehmm_session *s, int64_t P) {
    const size_t need = sizeof(uint32_t) * MT_N * (size_t)(P + SCRATCH_W);
    if (s->mt_windows.bytes < need) {
        if (s->mtj_pending) {  // kernels on the side stream may still write the old buffer
            EHMM_CK(cudaStreamSynchronize(s->stream2));
            s->mtj_pending = false;
        }
        EHMM_TRY(s->mt_windows.reserve(need + need / 2));
        s->mtj_valid = 0;
        s->mtj_niv = 0;
        s->mtj_base = false;
    }
    return EHMM_OK;
}

int base_window(ehmm_session *s, cudaStream_t st) {
    if (s->mtj_base) return EHMM_OK;
    mt_base_kernel<<<1, 256, 0, st>>>(s->mt_raw.as<uint32_t>(), s->mt_windows.as<uint32_t>());
    EHMM_CK(cudaGetLastError());
    s->launches[KID_MT]++;
    s->mtj_base = true;
    return EHMM_OK;
}

bool windows_cover(const ehmm_session *s, int64_t lo, int64_t hi) {
    if (lo > hi) return true;
    if (hi < s->mtj_valid) return true;
    for (int i = 0; i < s->mtj_niv; i++)
        if (s->mtj_iv[i][0] <= lo && hi <= s->mtj_iv[i][1]) return true;
    return false;
}

}  // namespace

// Prepare the windows of sub-streams 1..P-1 for the stream whose state is in s->mt_raw (device).
int mt_prepare(ehmm_session *s, cudaStream_t st, int64_t max_draws) {
    const int64_t P = substreams_for(s, 624, max_draws);
    if (P <= 1) { if (s->mtj_valid < 1) s->mtj_valid = 1; return EHMM_OK; }
    EHMM_REQUIRE(P <= ((int64_t)1 << (LEVELS - s->mt_shift)), EHMM_ERR_RNG, "too many uniforms in one call (%lld)", (long long)max_draws);
    const uint16_t *idx = nullptr;
    EHMM_TRY(device_polys(s->device, &idx));
    EHMM_TRY(reserve_windows(s, P));
    uint32_t *S = s->mt_windows.as<uint32_t>();
    EHMM_TRY(base_window(s, st));
    for (int k = 0; ((int64_t)1 << k) < P; k++) {
        const int64_t base = (int64_t)1 << k;
        const int64_t cnt = (P - base < base) ? (P - base) : base;
        EHMM_TRY(launch_jump(s, st, idx, k, S, S + (size_t)base * MT_N, cnt));
    }
    s->mtj_valid = P;
    return EHMM_OK;
}

// Windows of sub-streams [q_lo, q_hi] only (a bin block's slice of the stream in a sharded fit starts
// far into it): window q_lo from window 0 by one jump per set bit of q_lo, then doubling inside the
// range -- popcount(q_lo) + log2(length) launches and q_hi - q_lo + popcount(q_lo) window jumps instead
// of q_hi.
int mt_prepare_range(ehmm_session *s, cudaStream_t st, int64_t q_lo, int64_t q_hi) {
    if (q_lo < 0) q_lo = 0;
    if (q_hi < q_lo || windows_cover(s, q_lo, q_hi)) return EHMM_OK;
    EHMM_REQUIRE(q_hi < ((int64_t)1 << (LEVELS - s->mt_shift)), EHMM_ERR_RNG, "draw position beyond 2^%d sub-streams", LEVELS - s->mt_shift);
    const uint16_t *idx = nullptr;
    EHMM_TRY(device_polys(s->device, &idx));
    EHMM_TRY(reserve_windows(s, q_hi + 1));
    uint32_t *S = s->mt_windows.as<uint32_t>();
    const int64_t cap = (int64_t)(s->mt_windows.bytes / (sizeof(uint32_t) * MT_N));
    uint32_t *scratch[2] = {S + (size_t)(cap - 2) * MT_N, S + (size_t)(cap - 1) * MT_N};
    EHMM_TRY(base_window(s, st));
    if (q_lo > 0) {
        const uint32_t *cur = S;
        int bits = 0, done = 0;
        for (int k = 0; k < LEVELS - s->mt_shift; k++) bits += (q_lo >> k) & 1;
        for (int k = 0; k < LEVELS - s->mt_shift; k++) {
            if (!((q_lo >> k) & 1)) continue;
            uint32_t *dst = (++done == bits) ? S + (size_t)q_lo * MT_N : scratch[done & 1];
            EHMM_TRY(launch_jump(s, st, idx, k, cur, dst, 1));
            cur = dst;
        }
    }
    const int64_t len = q_hi - q_lo + 1;
    uint32_t *R = S + (size_t)q_lo * MT_N;
    for (int k = 0; ((int64_t)1 << k) < len; k++) {
        const int64_t base = (int64_t)1 << k;
        const int64_t cnt = (len - base < base) ? (len - base) : base;
        EHMM_TRY(launch_jump(s, st, idx, k, R, R + (size_t)base * MT_N, cnt));
    }
    if (s->mtj_niv < 80) {
        s->mtj_iv[s->mtj_niv][0] = q_lo;
        s->mtj_iv[s->mtj_niv][1] = q_hi;
        s->mtj_niv++;
    }
    return EHMM_OK;
}

// Windows of several stretches [q_lo[i], q_hi[i]] of sub-streams at once: the stretches are merged where they touch,
// every run reaches its first window from window 0 by one jump per set bit of its index, and then doubles; all
// runs take their t-th step in the same launch (popcount + log2(length) launches in all instead of per run).
int mt_prepare_ranges(ehmm_session *s, cudaStream_t st, int nr, const int64_t *q_lo_in, const int64_t *q_hi_in) {
    std::vector<std::pair<int64_t, int64_t>> iv;
    int64_t qmax = 0;
    for (int i = 0; i < nr; i++) {
        const int64_t lo = std::max<int64_t>(0, q_lo_in[i]), hi = q_hi_in[i];
        if (hi < lo || hi < 1 || windows_cover(s, lo, hi)) continue;
        iv.push_back({lo, hi});
        qmax = std::max(qmax, hi);
    }
    if (iv.empty()) return EHMM_OK;
    std::sort(iv.begin(), iv.end());
    std::vector<std::pair<int64_t, int64_t>> runs;
    for (const auto &v : iv) {
        if (!runs.empty() && v.first <= runs.back().second + 1) runs.back().second = std::max(runs.back().second, v.second);
        else runs.push_back(v);
    }
    if ((int)runs.size() > JB_MAX) {  // (never with <= 64 columns) one run after the other
        for (const auto &r : runs) EHMM_TRY(mt_prepare_range(s, st, r.first, r.second));
        return EHMM_OK;
    }
    EHMM_REQUIRE(qmax < ((int64_t)1 << (LEVELS - s->mt_shift)), EHMM_ERR_RNG, "draw position beyond 2^%d sub-streams", LEVELS - s->mt_shift);
    const uint16_t *idx = nullptr;
    EHMM_TRY(device_polys(s->device, &idx));
    EHMM_TRY(reserve_windows(s, qmax + 1));
    // (a growing buffer dropped the runs found covered above: they are simply prepared again by the next caller)
    uint32_t *S = s->mt_windows.as<uint32_t>();
    const int64_t cap = (int64_t)(s->mt_windows.bytes / (sizeof(uint32_t) * MT_N));
    const int64_t scratch0 = cap - SCRATCH_W;  // LEVELS scratch windows per run (the last two belong to mt_prepare_range)
    EHMM_TRY(base_window(s, st));
    const int R = (int)runs.size();
    {   // every destination window is written once: zero them all first
        ZeroRuns Z;
        Z.n = 0;
        for (int r = 0; r < R; r++) {
            Z.first[Z.n] = (int)std::max<int64_t>(1, runs[r].first);  // window 0 is the base window
            Z.count[Z.n] = (int)(runs[r].second - Z.first[Z.n] + 1);
            if (Z.count[Z.n] > 0) Z.n++;
        }
        EHMM_CK(cudaMemsetAsync(S + (size_t)scratch0 * MT_N, 0, sizeof(uint32_t) * MT_N * (size_t)(JB_MAX * LEVELS), st));
        mt_zero_runs_kernel<<<148, 256, 0, st>>>(Z, S);
        EHMM_CK(cudaGetLastError());
    }
    const int nlev = LEVELS - s->mt_shift;
    auto launch = [&](const JumpBatch &B) -> int {
        if (B.n == 0) return EHMM_OK;
        mt_jump_multi_kernel<<<dim3((unsigned)B.cta0[B.n], JUMP_SPLIT), JUMP_NT, JUMP_SMEM, st>>>(idx, B, S);
        EHMM_CK(cudaGetLastError());
        s->launches[KID_MT]++;
        return EHMM_OK;
    };
    auto add = [&](JumpBatch &B, int64_t src, int64_t dst, int k, int64_t cnt) {
        const int lv = k + s->mt_shift;
        B.src[B.n] = (int)src;
        B.dst[B.n] = (int)dst;
        B.lv[B.n] = lv;
        B.nidx[B.n] = g_nidx[lv];
        B.cta0[B.n + 1] = B.cta0[B.n] + (int)cnt;
        B.n++;
    };
    // reach the first window of every run
    std::vector<int64_t> cur((size_t)R, 0);
    std::vector<int> done((size_t)R, 0), bits((size_t)R, 0), nextk((size_t)R, 0);
    int maxbits = 0;
    for (int r = 0; r < R; r++) {
        for (int k = 0; k < nlev; k++) bits[(size_t)r] += (int)((runs[r].first >> k) & 1);
        maxbits = std::max(maxbits, bits[(size_t)r]);
    }
    for (int t = 0; t < maxbits; t++) {
        JumpBatch B;
        B.n = 0;
        B.cta0[0] = 0;
        for (int r = 0; r < R; r++) {
            if (done[(size_t)r] >= bits[(size_t)r]) continue;
            int k = nextk[(size_t)r];
            while (!((runs[r].first >> k) & 1)) k++;
            nextk[(size_t)r] = k + 1;
            done[(size_t)r]++;
            const int64_t dst = (done[(size_t)r] == bits[(size_t)r]) ? runs[r].first : scratch0 + (int64_t)r * LEVELS + t;
            add(B, cur[(size_t)r], dst, k, 1);
            cur[(size_t)r] = dst;
        }
        EHMM_TRY(launch(B));
    }
    // double inside every run
    int64_t maxlen = 0;
    for (int r = 0; r < R; r++) maxlen = std::max(maxlen, runs[r].second - runs[r].first + 1);
    for (int k = 0; ((int64_t)1 << k) < maxlen; k++) {
        JumpBatch B;
        B.n = 0;
        B.cta0[0] = 0;
        const int64_t base = (int64_t)1 << k;
        for (int r = 0; r < R; r++) {
            const int64_t len = runs[r].second - runs[r].first + 1;
            if (len <= base) continue;
            add(B, runs[r].first, runs[r].first + base, k, std::min(base, len - base));
        }
        EHMM_TRY(launch(B));
    }
    for (int r = 0; r < R && s->mtj_niv < 80; r++) {
        s->mtj_iv[s->mtj_niv][0] = runs[r].first;
        s->mtj_iv[s->mtj_niv][1] = runs[r].second;
        s->mtj_niv++;
    }
    return EHMM_OK;
}

static int64_t substream_of(const ehmm_session *s, int64_t a) { return a <= 0 ? 0 : (a - 1) / jw_of(s); }

static int wait_pending(ehmm_session *s) {
    if (s->mtj_pending) {  // windows being prepared on the side stream
        EHMM_CK(cudaStreamWaitEvent(s->stream, s->ev, 0));
        s->mtj_pending = false;
    }
    return EHMM_OK;
}

int mt_wait_side(ehmm_session *s) { return wait_pending(s); }

static int ensure_windows(ehmm_session *s, int64_t P) {
    EHMM_TRY(wait_pending(s));
    if (P > 1 && s->mtj_valid < P) EHMM_TRY(mt_prepare(s, s->stream, (P - 1) * jw_of(s) + 2));
    return EHMM_OK;
}

static int ensure_range(ehmm_session *s, int64_t q_lo, int64_t q_hi) {
    EHMM_TRY(wait_pending(s));
    if (q_hi >= 1 && !windows_cover(s, q_lo, q_hi)) EHMM_TRY(mt_prepare_range(s, s->stream, q_lo, q_hi));
    return EHMM_OK;
}

// Generate n tempered words of the stream into `out` (device) and advance s->mt_raw.
int mt_generate(ehmm_session *s, int64_t n, uint32_t *out) {
    if (n <= 0) return EHMM_OK;
    const int64_t pos = s->rng_state[0];
    const int64_t P = substream_of(s, pos + n - 1) + 1;
    EHMM_TRY(s->mt_raw2.reserve(sizeof(uint32_t) * 625));
    EHMM_TRY(ensure_windows(s, P));
    {
        PROF(s, KID_MT), mt_generate_kernel<<<(unsigned)P, GEN_NT, 0, s->stream>>>(
            s->mt_raw.as<uint32_t>(), s->mt_windows.as<uint32_t>(), 0, 0, n, out, n, s->mt_raw2.as<uint32_t>(), jw_of(s));
    }
    EHMM_CK(cudaGetLastError());
    std::swap(s->mt_raw, s->mt_raw2);
    s->mtj_valid = 0;
    s->mtj_niv = 0;
    s->mtj_base = false;
    return EHMM_OK;
}

// Words [lo_rel, lo_rel + n) of the stream (relative to the current position) without advancing it.
int mt_generate_range(ehmm_session *s, int64_t lo_rel, int64_t n, uint32_t *out) {
    if (n <= 0) return EHMM_OK;
    const int64_t pos = s->rng_state[0];
    const int64_t s0 = substream_of(s, pos + lo_rel), s1 = substream_of(s, pos + lo_rel + n - 1);
    EHMM_TRY(ensure_range(s, s0, s1));
    PROF(s, KID_MT), mt_generate_kernel<<<(unsigned)(s1 - s0 + 1), GEN_NT, 0, s->stream>>>(
        s->mt_raw.as<uint32_t>(), s->mt_windows.as<uint32_t>(), (int)s0, lo_rel, lo_rel + n, out, -1, nullptr, jw_of(s));
    EHMM_CK(cudaGetLastError());
    return EHMM_OK;
}

// nr ranges [lo_rel[i], lo_rel[i] + n[i]) of the stream (relative to the current position) in ONE launch, range i to
// out + out_off[i]; nothing is advanced.  A range is walked by one CTA per sub-stream it touches, so launching the
// ranges one by one leaves most of the machine idle (a column's slice of a sharded fit is a few sub-streams long).
int mt_generate_ranges(ehmm_session *s, int nr, const int64_t *lo_rel, const int64_t *n, const int64_t *out_off, uint32_t *out) {
    EHMM_REQUIRE(nr <= GEN_MAXR, EHMM_ERR_ARG, "at most %d ranges per launch", GEN_MAXR);
    const int64_t pos = s->rng_state[0];
    // the window buffer must not grow between the ranges (growing drops what it holds): size it for the farthest one first
    int64_t qmax = 0;
    for (int i = 0; i < nr; i++)
        if (n[i] > 0) qmax = std::max(qmax, substream_of(s, pos + lo_rel[i] + n[i] - 1));
    EHMM_TRY(wait_pending(s));
    if (qmax >= 1) EHMM_TRY(reserve_windows(s, qmax + 1));
    {
        int64_t qlo[GEN_MAXR], qhi[GEN_MAXR];
        int m = 0;
        for (int i = 0; i < nr; i++) {
            if (n[i] <= 0) continue;
            qlo[m] = substream_of(s, pos + lo_rel[i]);
            qhi[m] = substream_of(s, pos + lo_rel[i] + n[i] - 1);
            m++;
        }
        EHMM_TRY(mt_prepare_ranges(s, s->stream, m, qlo, qhi));
    }
    GenRanges R;
    R.n = 0;
    int ctas = 0;
    for (int i = 0; i < nr; i++) {
        if (n[i] <= 0) continue;
        const int64_t s0 = substream_of(s, pos + lo_rel[i]), s1 = substream_of(s, pos + lo_rel[i] + n[i] - 1);
        EHMM_REQUIRE(s1 < 1 || windows_cover(s, s0, s1), EHMM_ERR_RNG, "sub-stream windows [%lld, %lld] are not in place", (long long)s0,
                     (long long)s1);
        R.cta0[R.n] = ctas;
        R.s_first[R.n] = (int)s0;
        R.lo_rel[R.n] = lo_rel[i];
        R.hi_rel[R.n] = lo_rel[i] + n[i];
        R.out_off[R.n] = out_off[i];
        ctas += (int)(s1 - s0 + 1);
        R.n++;
    }
    if (R.n == 0) return EHMM_OK;
    R.cta0[R.n] = ctas;
    PROF(s, KID_MT), mt_generate_multi_kernel<<<(unsigned)ctas, GEN_NT, 0, s->stream>>>(s->mt_raw.as<uint32_t>(), s->mt_windows.as<uint32_t>(), R, out,
                                                                                 jw_of(s));
    EHMM_CK(cudaGetLastError());
    return EHMM_OK;
}

// The same on the side stream, ahead of the call that will consume the words (sharded fits: a block's column slices
// sit where the previous call's did, give or take the change of the counts in between, so the caller asks for the
// previous slices with a margin either side).  Call after mt_advance_side: the ranges are relative to the NEW
// position.  The main stream meets the side stream again at the next wait_pending.
int mt_speculate_ranges(ehmm_session *s, int nr, const int64_t *lo_rel, const int64_t *n, const int64_t *out_off, uint32_t *out) {
    EHMM_REQUIRE(nr <= GEN_MAXR, EHMM_ERR_ARG, "at most %d ranges per launch", GEN_MAXR);
    const int64_t pos = s->rng_state[0];
    int64_t qmax = 0;
    for (int i = 0; i < nr; i++)
        if (n[i] > 0) qmax = std::max(qmax, substream_of(s, pos + lo_rel[i] + n[i] - 1));
    if (qmax >= 1) EHMM_TRY(reserve_windows(s, qmax + 1));
    {
        int64_t qlo[GEN_MAXR], qhi[GEN_MAXR];
        int m = 0;
        for (int i = 0; i < nr; i++) {
            if (n[i] <= 0) continue;
            qlo[m] = substream_of(s, pos + lo_rel[i]);
            qhi[m] = substream_of(s, pos + lo_rel[i] + n[i] - 1);
            m++;
        }
        EHMM_TRY(mt_prepare_ranges(s, s->stream2, m, qlo, qhi));
    }
    GenRanges R;
    R.n = 0;
    int ctas = 0;
    for (int i = 0; i < nr; i++) {
        if (n[i] <= 0) continue;
        const int64_t s0 = substream_of(s, pos + lo_rel[i]), s1 = substream_of(s, pos + lo_rel[i] + n[i] - 1);
        R.cta0[R.n] = ctas;
        R.s_first[R.n] = (int)s0;
        R.lo_rel[R.n] = lo_rel[i];
        R.hi_rel[R.n] = lo_rel[i] + n[i];
        R.out_off[R.n] = out_off[i];
        ctas += (int)(s1 - s0 + 1);
        R.n++;
    }
    if (R.n > 0) {
        R.cta0[R.n] = ctas;
        for (int r = 0; r < R.n; r++)  // (a growing window buffer drops earlier ranges: then nothing is generated and the caller misses)
            EHMM_REQUIRE(substream_of(s, pos + R.hi_rel[r] - 1) < 1 ||
                             windows_cover(s, R.s_first[r], substream_of(s, pos + R.hi_rel[r] - 1)),
                         EHMM_ERR_RNG, "speculative generation: sub-stream windows are not in place");
        mt_generate_multi_kernel<<<(unsigned)ctas, GEN_NT, 0, s->stream2>>>(s->mt_raw.as<uint32_t>(), s->mt_windows.as<uint32_t>(), R, out,
                                                                         jw_of(s));
        EHMM_CK(cudaGetLastError());
        s->launches[KID_MT]++;
    }
    EHMM_CK(cudaEventRecord(s->ev, s->stream2));
    s->mtj_pending = true;
    return EHMM_OK;
}

// Advance the stream by n draws (R's state array afterwards) without producing output.
int mt_advance(ehmm_session *s, int64_t n) {
    if (n <= 0) return EHMM_OK;
    const int64_t pos = s->rng_state[0];
    const int64_t E = pos + n, st_lo = ((E - 1) / 624) * 624;
    const int64_t s0 = substream_of(s, st_lo), s1 = substream_of(s, E - 1);
    EHMM_TRY(s->mt_raw2.reserve(sizeof(uint32_t) * 625));
    EHMM_TRY(ensure_range(s, s0, s1));
    PROF(s, KID_MT), mt_generate_kernel<<<(unsigned)(s1 - s0 + 1), GEN_NT, 0, s->stream>>>(
        s->mt_raw.as<uint32_t>(), s->mt_windows.as<uint32_t>(), (int)s0, 0, 0, nullptr, n, s->mt_raw2.as<uint32_t>(), jw_of(s));
    EHMM_CK(cudaGetLastError());
    std::swap(s->mt_raw, s->mt_raw2);
    s->mtj_valid = 0;
    s->mtj_niv = 0;
    s->mtj_base = false;
    return EHMM_OK;
}

// mt_advance on the side stream, followed by the copy of the new state to the pinned host buffer: the
// single CTA that walks to the end of the last sub-stream (~0.1 ms) then overlaps the table kernels.
// The main stream meets the side stream again at the next wait_pending (mt_prefetch_ranges records it).
int mt_advance_side(ehmm_session *s, int64_t n, void *h_state) {
    if (n <= 0) return EHMM_OK;
    const int64_t pos = s->rng_state[0];
    const int64_t E = pos + n, st_lo = ((E - 1) / 624) * 624;
    const int64_t s0 = substream_of(s, st_lo), s1 = substream_of(s, E - 1);
    EHMM_TRY(s->mt_raw2.reserve(sizeof(uint32_t) * 625));
    EHMM_TRY(wait_pending(s));
    EHMM_CK(cudaEventRecord(s->ev2, s->stream));
    EHMM_CK(cudaStreamWaitEvent(s->stream2, s->ev2, 0));
    if (s1 >= 1) EHMM_TRY(mt_prepare_range(s, s->stream2, s0, s1));
    mt_generate_kernel<<<(unsigned)(s1 - s0 + 1), GEN_NT, 0, s->stream2>>>(
        s->mt_raw.as<uint32_t>(), s->mt_windows.as<uint32_t>(), (int)s0, 0, 0, nullptr, n, s->mt_raw2.as<uint32_t>(), jw_of(s));
    EHMM_CK(cudaGetLastError());
    s->launches[KID_MT]++;
    std::swap(s->mt_raw, s->mt_raw2);
    s->mtj_valid = 0;
    s->mtj_niv = 0;
    s->mtj_base = false;
    EHMM_CK(cudaMemcpyAsync(h_state, s->mt_raw.p, sizeof(uint32_t) * 625, cudaMemcpyDeviceToHost, s->stream2));
    return EHMM_OK;
}

// Choose the sub-stream length for calls that draw about `draws` uniforms: about 150-300 sub-streams (one
// generating CTA each, all resident at once) when there are that many draws, never shorter than JW.
// Changing it drops the prepared windows (they are spaced by the old length).
int mt_choose_stride(ehmm_session *s, int64_t draws) {
    int shift = 0;
    while (shift < 5 && (draws >> (LOG2J + shift + 1)) >= 150) shift++;
    if (shift == s->mt_shift) return EHMM_OK;
    if (s->mtj_pending) {
        EHMM_CK(cudaStreamSynchronize(s->stream2));
        s->mtj_pending = false;
    }
    s->mt_shift = shift;
    s->mtj_valid = 0;
    s->mtj_niv = 0;
    return EHMM_OK;
}

// After a call: start preparing the windows for the next one on the side stream.
int mt_prefetch(ehmm_session *s, int64_t max_draws) {
    if (substreams_for(s, 624, max_draws) <= 1) return EHMM_OK;
    EHMM_CK(cudaEventRecord(s->ev2, s->stream));
    EHMM_CK(cudaStreamWaitEvent(s->stream2, s->ev2, 0));
    EHMM_TRY(mt_prepare(s, s->stream2, max_draws));
    EHMM_CK(cudaEventRecord(s->ev, s->stream2));
    s->mtj_pending = true;
    return EHMM_OK;
}

// After a call (and after mt_advance_side has queued the new state on the side stream): generate the next
// `est` words of the stream into `out` on the side stream, ahead of the call that will consume them -- how many
// it needs is only known once its posteriors have been counted, but a call draws about as many uniforms
// as the one before it.  Nothing is advanced; the consumer takes the words it needs (mt_spec_ready orders the
// main stream behind the generation) and advances the stream itself.
int mt_speculate(ehmm_session *s, int64_t est, uint32_t *out) {
    if (est <= 0) return EHMM_OK;
    const int64_t pos = s->rng_state[0];
    EHMM_TRY(mt_prepare(s, s->stream2, est + MT_N));
    const int64_t P = substream_of(s, pos + est - 1) + 1;
    mt_generate_kernel<<<(unsigned)P, GEN_NT, 0, s->stream2>>>(s->mt_raw.as<uint32_t>(), s->mt_windows.as<uint32_t>(), 0, 0, est, out,
                                                             -1, nullptr, jw_of(s));
    EHMM_CK(cudaGetLastError());
    s->launches[KID_MT]++;
    EHMM_CK(cudaEventRecord(s->ev, s->stream2));
    s->mtj_pending = true;
    return EHMM_OK;
}

// Sharded form: prepare only the windows the next call is expected to touch -- relative draw ranges
// [lo[i], lo[i] + n[i]) with `margin` sub-streams either side, and the words around position `total`
// (the state after the call) -- on the side stream.
int mt_prefetch_ranges(ehmm_session *s, int nr, const int64_t *lo, const int64_t *n, int64_t total, int margin) {
    EHMM_CK(cudaEventRecord(s->ev2, s->stream));
    EHMM_CK(cudaStreamWaitEvent(s->stream2, s->ev2, 0));
    const int64_t pos = s->rng_state[0];
    int64_t qlo[GEN_MAXR + 1], qhi[GEN_MAXR + 1];
    int m = 0;
    for (int i = 0; i < nr && m < GEN_MAXR; i++) {
        if (n[i] <= 0) continue;
        qlo[m] = substream_of(s, pos + lo[i]) - margin;
        qhi[m] = substream_of(s, pos + lo[i] + n[i] - 1) + margin;
        m++;
    }
    if (total > 0) {
        const int64_t E = pos + total, st_lo = ((E - 1) / MT_N) * MT_N;
        qlo[m] = substream_of(s, st_lo) - margin;
        qhi[m] = substream_of(s, E - 1) + margin;
        m++;
    }
    EHMM_TRY(mt_prepare_ranges(s, s->stream2, m, qlo, qhi));
    EHMM_CK(cudaEventRecord(s->ev, s->stream2));
    s->mtj_pending = true;
    return EHMM_OK;
}

}  // namespace ehmm
