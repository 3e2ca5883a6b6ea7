// viterbi.cu -- computeViterbiSequence (src/computeViterbiSequence.cpp:12-54).
//
// The reference scores each bin with logFP (the forward log-probabilities), so LOGV grows like M^2
// (1e15 at 15 M bins) and ties are decided at ulp granularity: the (max,+) recursion
//     V[j+1,k] = fl( max_i fl(V[j,i] + log gamma[i,k]) + logFP[j+1,k] )
// is not re-associable and is run here in exactly that fp64 operation order (__dadd_rn, no FMA,
// first-index argmax) by one thread, while the rest of the CTA streams logFP tiles through shared
// memory and writes bit-packed back-pointers.  The reference's backtrace re-evaluates
// argmax_i(V[j,i] + log gamma[i,S[j+1]]), which is the same expression as the forward step's
// argmax, so storing psi[j+1][k] = argmax_i(...) (2 bits per state) is equivalent.
#include "common.cuh"

namespace ehmm {

namespace {

constexpr int VT_NT = 256;
constexpr int VT_TILE = 2048;  // bins per shared-memory tile

struct VitPar {
    double lpi[EHMM_MAX_K];
    double lg[EHMM_MAX_K * EHMM_MAX_K];  // lg[i*K+k] = log gamma(i,k)
    double vin[EHMM_MAX_K];              // LOGV of the bin in front of this block (bin blocks only)
    int has_vin;                         // 0: this block starts the chain
};

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
    // prefetch tile 0
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
                if (s0 + b == 0 && !P.has_vin) {
#pragma unroll
                    for (int k = 0; k < K; k++) V[k] = __dadd_rn(P.lpi[k], cur[k * VT_TILE]);
                    sbp[0] = 0;
                } else {
                    double mc[K];
                    unsigned code = 0;
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        double mx = __dadd_rn(V[0], P.lg[k]);
                        unsigned arg = 0;
#pragma unroll
                        for (int i = 1; i < K; i++) {
                            const double c = __dadd_rn(V[i], P.lg[i * K + k]);
                            if (c > mx) { mx = c; arg = i; }
                        }
                        mc[k] = mx;
                        code |= arg << (2 * k);
                    }
#pragma unroll
                    for (int k = 0; k < K; k++) V[k] = __dadd_rn(mc[k], cur[k * VT_TILE + b]);
                    sbp[b] = (uint8_t)code;
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

// backtrace: S[M-1] = argmax V[M-1,:] (first maximum), S[j] = psi[j+1][S[j+1]]
// carry_in: state of this block's last bin when a later block decided it, -1 at the chain's end
// (then argmax of V).  carry[0] receives the state of the bin in front of the block.
template <int K>
__global__ void __launch_bounds__(VT_NT) viterbi_backtrace_kernel(const uint8_t *__restrict__ bp, int64_t M,
                                                                  const double *__restrict__ vlast, int carry_in,
                                                                  double *__restrict__ states, int *__restrict__ carry_out) {
    __shared__ uint8_t sbp[VT_TILE + 1];
    __shared__ uint8_t sst[VT_TILE];
    __shared__ int carry;
    const int tid = threadIdx.x;
    const int64_t nT = (M + VT_TILE - 1) / VT_TILE;
    if (tid == 0) {
        int best = 0;
        for (int k = 1; k < K; k++)
            if (vlast[k] > vlast[best]) best = k;
        carry = (carry_in >= 0) ? carry_in : best;  // state of bin M-1
    }
    __syncthreads();
    for (int64_t t = nT - 1; t >= 0; t--) {
        const int64_t s0 = t * VT_TILE;
        const int n = (int)min((int64_t)VT_TILE, M - s0);
        // need psi[j+1] for j in [s0, s0+n): bp[s0+1 .. s0+n]
        for (int b = tid; b < n; b += VT_NT) sbp[b] = (s0 + b + 1 < M) ? bp[s0 + b + 1] : 0;
        __syncthreads();
        if (tid == 0) {
            int st = carry;
            int b = n - 1;
            if (s0 + b == M - 1) { sst[b] = (uint8_t)st; b--; }
            for (; b >= 0; b--) {
                st = (sbp[b] >> (2 * st)) & 3;
                sst[b] = (uint8_t)st;
            }
            carry = st;
        }
        __syncthreads();
        for (int b = tid; b < n; b += VT_NT) states[s0 + b] = (double)sst[b];
        __syncthreads();
    }
    // psi of the block's first bin applied to its state: the state of the bin in front of the block
    if (tid == 0) carry_out[0] = (bp[0] >> (2 * carry)) & 3;
}

template <int K> int forward(ehmm_session *s, const VitPar &P) {
    const int64_t M = s->M;
    EHMM_TRY(s->viterbi.reserve(sizeof(double) * M));
    EHMM_TRY(s->vit_bp.reserve((size_t)M + 64));
    EHMM_TRY(s->misc.reserve(256));
    const size_t smem = sizeof(double) * 2 * K * VT_TILE + VT_TILE;
    static bool attr[EHMM_MAX_K + 1] = {false};
    if (!attr[K]) {
        EHMM_CK(cudaFuncSetAttribute(viterbi_forward_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr[K] = true;
    }
    PROF(s, KID_VIT_FWD), viterbi_forward_kernel<K><<<1, VT_NT, smem, s->stream>>>(s->logFP.as<double>(), M, P,
                                                                              s->vit_bp.as<uint8_t>(), s->misc.as<double>());
    EHMM_CK(cudaGetLastError());
    return EHMM_OK;
}

template <int K> int backtrace(ehmm_session *s, int carry_in, int *carry_out, double *states_out) {
    const int64_t M = s->M;
    int *dcarry = reinterpret_cast<int *>(s->misc.as<double>() + 16);
    PROF(s, KID_VIT_BWD), viterbi_backtrace_kernel<K><<<1, VT_NT, 0, s->stream>>>(
        s->vit_bp.as<uint8_t>(), M, s->misc.as<double>(), carry_in, s->viterbi.as<double>(), dcarry);
    EHMM_CK(cudaGetLastError());
    if (states_out)
        EHMM_CK(cudaMemcpyAsync(states_out, s->viterbi.p, sizeof(double) * M, cudaMemcpyDeviceToHost, s->stream));
    int hc = 0;
    EHMM_CK(cudaMemcpyAsync(&hc, dcarry, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    if (carry_out) *carry_out = hc;
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
    if (s->K == 2) { EHMM_TRY(forward<2>(s, P)); return backtrace<2>(s, -1, nullptr, states_out); }
    if (s->K == 3) { EHMM_TRY(forward<3>(s, P)); return backtrace<3>(s, -1, nullptr, states_out); }
    set_error("viterbi: K=%d unsupported", s->K);
    return EHMM_ERR_ARG;
}

// bin-block form: the (max,+) recursion is serial over the chain, so blocks run one after the other;
// v_in = LOGV of the bin in front of this block (NULL for the first block), v_out = LOGV of its last bin
int viterbi_forward(ehmm_session *s, const double *pi, const double *gamma, const double *v_in, double *v_out) {
    EHMM_REQUIRE((s->m0 == 0) == (v_in == nullptr), EHMM_ERR_ARG, "viterbi_forward: v_in is required exactly for blocks with m0 > 0");
    VitPar P = {};
    fill(P, s->K, pi, gamma, v_in);
    if (s->K == 2) EHMM_TRY(forward<2>(s, P));
    else if (s->K == 3) EHMM_TRY(forward<3>(s, P));
    else { set_error("viterbi: K=%d unsupported", s->K); return EHMM_ERR_ARG; }
    EHMM_CK(cudaMemcpyAsync(s->h_pinned + 320, s->misc.p, sizeof(double) * s->K, cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    for (int k = 0; k < s->K; k++) v_out[k] = s->h_pinned[320 + k];
    s->vit_forward_done = true;
    return EHMM_OK;
}

// state_last: state of this block's last bin as decided by the next block, -1 for the chain's end;
// state_prev: state of the bin in front of this block (to hand to the previous block)
int viterbi_backtrace(ehmm_session *s, int state_last, int *state_prev, double *states_out) {
    EHMM_REQUIRE(s->vit_forward_done, EHMM_ERR_STATE, "viterbi_backtrace without viterbi_forward");
    EHMM_REQUIRE(state_last >= -1 && state_last < s->K, EHMM_ERR_ARG, "bad state %d", state_last);
    if (s->K == 2) return backtrace<2>(s, state_last, state_prev, states_out);
    return backtrace<3>(s, state_last, state_prev, states_out);
}

}  // namespace ehmm
