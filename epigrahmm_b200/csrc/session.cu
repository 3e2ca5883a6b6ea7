// session.cu -- the extern "C" boundary (include/ehmm.h): session registry, data upload,
// dataset access and the small host-side finishing steps of the M-step (maxStepProb's floor and
// renormalisation, Q and BIC from the fused sufficient statistics).
#include "common.cuh"
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>

namespace ehmm {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::mutex g_mu;
static std::map<std::string, ehmm_session *> g_sessions;

namespace {

__global__ void minmax_i32_kernel(const int32_t *__restrict__ x, int64_t n, int *__restrict__ out) {
    int mx = INT_MIN, mn = INT_MAX;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int v = __ldg(x + i);
        mx = max(mx, v);
        mn = min(mn, v);
    }
    for (int d = 16; d > 0; d >>= 1) {
        mx = max(mx, __shfl_down_sync(0xffffffffu, mx, d));
        mn = min(mn, __shfl_down_sync(0xffffffffu, mn, d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out, mx);
        atomicMin(out + 1, mn);
    }
}

// does dtUnique[group] reproduce the data cell by cell?  (it must, for the dictionary-encoded
// emission path; the reference builds Group from exactly these columns, R/base-core.R:97,187-193)
__global__ void group_check_kernel(int64_t n, const int32_t *__restrict__ group, int64_t G,
                                   const int32_t *__restrict__ y, const double *__restrict__ off,
                                   const double *__restrict__ ctrl, const double *__restrict__ uy,
                                   const double *__restrict__ uoff, const double *__restrict__ uctrl,
                                   int *__restrict__ bad) {
    int b = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = group[i];
        if (g < 1 || g > G) { b |= 2; continue; }
        if (uy[g] != (double)y[i] || uoff[g] != off[i] || (ctrl && uctrl && uctrl[g] != ctrl[i])) b |= 1;
    }
    if (b) atomicOr(bad, b);
}

__global__ void exp_kernel(const double *__restrict__ x, int64_t n, double *__restrict__ y) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = exp(x[i]);
}

int scan_counts(ehmm_session *s) {
    int h[2] = {INT_MIN, INT_MAX};
    EHMM_TRY(s->misc.reserve(64));
    int *d = s->misc.as<int>();
    EHMM_CK(cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, s->stream));
    PROF(s, KID_MISC), minmax_i32_kernel<<<148 * 8, 256, 0, s->stream>>>(s->d_counts, s->M * s->N, d);
    EHMM_CK(cudaGetLastError());
    EHMM_CK(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    EHMM_REQUIRE(h[1] >= 0, EHMM_ERR_ARG, "counts must be non-negative (min = %d)", h[1]);
    s->count_max = h[0];
    return EHMM_OK;
}

// dtUnique columns are stored with a leading unused slot so that the 1-based group id indexes them
int upload_unique(ehmm_session *s, int64_t G, const double *u_chip, const double *u_offsets, const double *u_controls,
                  const uint8_t *u_dsg) {
    const size_t gb = (size_t)(G + 1) * sizeof(double);
    EHMM_TRY(s->u_chip.reserve(gb));
    EHMM_TRY(s->u_off.reserve(gb));
    EHMM_CK(cudaMemsetAsync(s->u_chip.p, 0, sizeof(double), s->stream));
    EHMM_CK(cudaMemsetAsync(s->u_off.p, 0, sizeof(double), s->stream));
    EHMM_CK(cudaMemcpyAsync(s->u_chip.as<double>() + 1, u_chip, gb - sizeof(double), cudaMemcpyHostToDevice, s->stream));
    EHMM_CK(cudaMemcpyAsync(s->u_off.as<double>() + 1, u_offsets, gb - sizeof(double), cudaMemcpyHostToDevice, s->stream));
    if (u_controls) {
        EHMM_TRY(s->u_ctrl.reserve(gb));
        EHMM_CK(cudaMemsetAsync(s->u_ctrl.p, 0, sizeof(double), s->stream));
        EHMM_CK(cudaMemcpyAsync(s->u_ctrl.as<double>() + 1, u_controls, gb - sizeof(double), cudaMemcpyHostToDevice, s->stream));
    } else {
        s->u_ctrl.release();
    }
    if (u_dsg) {
        EHMM_REQUIRE(s->B > 0, EHMM_ERR_STATE, "set_groups: Dsg.Mix columns given before ehmm_set_design");
        // host layout G x B column-major -> device (G+1) x B column-major
        EHMM_TRY(s->u_dsg.reserve((size_t)(G + 1) * s->B));
        EHMM_CK(cudaMemsetAsync(s->u_dsg.p, 0, (size_t)(G + 1) * s->B, s->stream));
        for (int b = 0; b < s->B; b++)
            EHMM_CK(cudaMemcpyAsync(s->u_dsg.as<uint8_t>() + (size_t)b * (G + 1) + 1, u_dsg + (size_t)b * G, (size_t)G,
                                    cudaMemcpyHostToDevice, s->stream));
    }
    return EHMM_OK;
}

__global__ void widen_u16_kernel(int64_t n, const uint16_t *__restrict__ in, int32_t *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = in[i];
}

__global__ void group_range_kernel(int64_t n, const int32_t *__restrict__ group, int64_t G, int *__restrict__ bad) {
    int b = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = group[i];
        if (g < 1 || g > G) b = 1;
    }
    if (b) atomicOr(bad, 1);
}

struct DS {
    ehmm::DevBuf *buf;
    int64_t rows, cols;
    unsigned bit;
};

int lookup(ehmm_session *s, const char *name, DS &d) {
    const int64_t M = s->M;
    const int K = s->K;
    if (!strcmp(name, "logLikelihood")) d = {&s->logf, M, K, DS_LOGF};
    else if (!strcmp(name, "logFP")) d = {&s->logFP, M, K, DS_FP};
    else if (!strcmp(name, "logBP")) d = {&s->logBP, M, K, DS_BP};
    else if (!strcmp(name, "logProb1")) d = {&s->logProb1, M, K, DS_P1};
    else if (!strcmp(name, "logProb2")) d = {&s->logProb2, M, K * K, DS_P2};
    else if (!strcmp(name, "mixtureProb")) d = {&s->eta, M, s->B, DS_ETA};
    else if (!strcmp(name, "viterbi")) d = {&s->viterbi, M, 1, DS_VIT};
    else {
        set_error("unknown dataset '%s'", name);
        return EHMM_ERR_ARG;
    }
    return EHMM_OK;
}

int set_theta(ehmm_session *s, const double *pi, const double *gamma) {
    EHMM_REQUIRE(pi && gamma, EHMM_ERR_ARG, "pi/gamma must not be null");
    const int K = s->K;
    for (int k = 0; k < K; k++) {
        EHMM_REQUIRE(pi[k] >= 0 && isfinite(pi[k]), EHMM_ERR_ARG, "pi[%d] = %g is not a probability", k, pi[k]);
        s->h_pi[k] = pi[k];
    }
    for (int i = 0; i < K * K; i++) {
        EHMM_REQUIRE(gamma[i] >= 0 && isfinite(gamma[i]), EHMM_ERR_ARG, "gamma[%d] = %g is not a probability", i, gamma[i]);
        s->h_gamma[i] = gamma[i];
    }
    return EHMM_OK;
}

int upload_logf(ehmm_session *s, const double *logf, const double **d_out) {
    if (logf) {
        EHMM_TRY(logf_for_write(s, sizeof(double) * s->M * s->K));
        EHMM_CK(cudaMemcpyAsync(s->logf.p, logf, sizeof(double) * s->M * s->K, cudaMemcpyHostToDevice, s->stream));
        s->valid |= DS_LOGF;
    }
    EHMM_REQUIRE(has(s, DS_LOGF), EHMM_ERR_STATE, "expStep: no logf given and no emission has been computed");
    *d_out = s->logf.as<double>();
    return EHMM_OK;
}

// bookkeeping behind every E-step: which datasets exist now, what may be reused downstream
int estep_done(ehmm_session *s) {
    if (s->lazy_full) s->valid = (s->valid & ~(unsigned)DS_ESTEP) | DS_P1LIN;   // lean: P1 only, the rest on demand
    else s->valid = (s->valid | DS_ESTEP) & ~(unsigned)DS_P1LIN;
    s->stats_stale = false;
    posteriors_changed(s);
    if (s->lazy_full && s->fb_counted) {  // the kernel counted the weights below the announced cut on the way
        s->rc_counts_gen = s->post_gen;
        s->rc_counts_p = s->rc_hint_p;
        s->rc_counts_cols = s->K;
    }
    return rc_static_prefetch(s);
}

}  // namespace
}  // namespace ehmm

using namespace ehmm;

extern "C" {

const char *ehmm_last_error(void) { return g_err; }
int ehmm_version(void) { return 100; }

int ehmm_device_count(int *n) {
    EHMM_REQUIRE(n, EHMM_ERR_ARG, "null pointer");
    EHMM_CK(cudaGetDeviceCount(n));
    return EHMM_OK;
}

int ehmm_session_open(const char *key, int64_t M, int K, int device, ehmm_session **out) {
    EHMM_REQUIRE(key && out, EHMM_ERR_ARG, "null pointer");
    EHMM_REQUIRE(M >= 1, EHMM_ERR_ARG, "M must be >= 1 (M=%lld)", (long long)M);
    EHMM_REQUIRE(K == 2 || K == 3, EHMM_ERR_ARG, "K must be 2 or 3 (K=%d)", K);
    int ndev = 0;
    EHMM_CK(cudaGetDeviceCount(&ndev));
    EHMM_REQUIRE(device >= 0 && device < ndev, EHMM_ERR_CUDA, "device %d not available (%d CUDA devices)", device, ndev);
    EHMM_CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    EHMM_CK(cudaGetDeviceProperties(&prop, device));
    EHMM_REQUIRE(prop.major == 10, EHMM_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                 prop.major, prop.minor);
    ehmm_session *s = new ehmm_session();
    s->key = key;
    s->device = device;
    s->M = M;
    s->Mtot = M;
    s->K = K;
    {   // the EM iteration's own kernels go first; what runs ahead of need (uniform generation, jump windows, static
        // table parts, staged uploads) fills the machine behind them
        int lo = 0, hi = 0;
        EHMM_CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        const char *e = getenv("EHMM_STREAM_PRIORITY");
        const bool prio = !(e && e[0] == '0');
        EHMM_CK(cudaStreamCreateWithPriority(&s->stream, cudaStreamNonBlocking, prio ? hi : lo));
        EHMM_CK(cudaStreamCreateWithPriority(&s->stream2, cudaStreamNonBlocking, lo));
        EHMM_CK(cudaStreamCreateWithPriority(&s->stream_up, cudaStreamNonBlocking, lo));
    }
    for (int i = 0; i < 2; i++) {
        EHMM_CK(cudaEventCreateWithFlags(&s->ev_up[i], cudaEventDisableTiming));
        EHMM_CK(cudaEventCreateWithFlags(&s->ev_used[i], cudaEventDisableTiming));
    }
    EHMM_CK(cudaEventCreateWithFlags(&s->ev, cudaEventDisableTiming));
    EHMM_CK(cudaEventCreateWithFlags(&s->ev2, cudaEventDisableTiming));
    EHMM_CK(cudaEventCreateWithFlags(&s->ev_static, cudaEventDisableTiming));
    EHMM_CK(cudaMallocHost(&s->h_pinned, sizeof(double) * 1024));
    EHMM_CK(cudaHostAlloc(&s->h_glm, sizeof(double) * 64, cudaHostAllocMapped));
    memset(s->h_glm, 0, sizeof(double) * 64);
    EHMM_CK(cudaHostGetDevicePointer(&s->d_glm, s->h_glm, 0));
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_sessions.find(s->key);
        if (it != g_sessions.end()) {
            // same key re-opened: the reference would overwrite the HDF5 file (hdf5_opts::replace)
            ehmm_session *old = it->second;
            g_sessions.erase(it);
            ehmm_session_close(old);
        }
        g_sessions[s->key] = s;
    }
    *out = s;
    return EHMM_OK;
}

int ehmm_session_find(const char *key, ehmm_session **out) {
    EHMM_REQUIRE(key && out, EHMM_ERR_ARG, "null pointer");
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_sessions.find(key);
    EHMM_REQUIRE(it != g_sessions.end(), EHMM_ERR_STATE, "no session for '%s' (the reference would fail to open the HDF5 file)", key);
    *out = it->second;
    return EHMM_OK;
}

int ehmm_session_close(ehmm_session *s) {
    if (!s) return EHMM_OK;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    if (s->stream2) cudaStreamSynchronize(s->stream2);
    if (s->stream_up) cudaStreamSynchronize(s->stream_up);
    {
        std::unique_lock<std::mutex> lk(g_mu, std::try_to_lock);
        auto it = g_sessions.find(s->key);
        if (it != g_sessions.end() && it->second == s) g_sessions.erase(it);
    }
    DevBuf *bufs[] = {&s->counts_own, &s->offsets_own, &s->controls_own, &s->grp_tags, &s->grp_first, &s->grp_list, &s->grp_rank, &s->grp_misc, &s->grp_sort, &s->group, &s->u_chip, &s->u_off, &s->u_ctrl,
                      &s->u_dsg, &s->logf, &s->logFP, &s->logBP, &s->logProb1, &s->logProb2, &s->prob1, &s->logf_alt, &s->eta, &s->viterbi, &s->vit_bp, &s->vit_work, &s->peak_work, &s->peak_calls,
                      &s->tileops, &s->carries, &s->tilestats, &s->fbscal, &s->rangeops, &s->rangecarry, &s->rc_flag, &s->rc_scan,
                      &s->rc_buckets, &s->rc_limbs, &s->rc_coltot, &s->rc_static, &s->rc_unif, &s->rc_unif_spec, &s->rc_unif_specr[0], &s->rc_unif_specr[1], &s->rc_tmp, &s->mt_raw, &s->mt_raw2, &s->mt_windows, &s->gtab, &s->glm_ticket, &s->inner_part, &s->lgtab, &s->glm_part, &s->misc};
    for (DevBuf *b : bufs) b->release();
    for (int i = 0; i < 2; i++) {
        s->stage[i].release();
        if (s->ev_up[i]) cudaEventDestroy(s->ev_up[i]);
        if (s->ev_used[i]) cudaEventDestroy(s->ev_used[i]);
    }
    if (s->stream_up) cudaStreamDestroy(s->stream_up);
    for (cudaEvent_t e : s->prof_pool) cudaEventDestroy(e);
    if (s->h_pinned) cudaFreeHost(s->h_pinned);
    if (s->h_glm) cudaFreeHost(s->h_glm);
    if (s->ev) cudaEventDestroy(s->ev);
    if (s->ev2) cudaEventDestroy(s->ev2);
    if (s->ev_static) cudaEventDestroy(s->ev_static);
    if (s->stream) cudaStreamDestroy(s->stream);
    if (s->stream2) cudaStreamDestroy(s->stream2);
    delete s;
    return EHMM_OK;
}

int ehmm_session_stream(ehmm_session *s, void **stream) {
    EHMM_REQUIRE(s && stream, EHMM_ERR_ARG, "null pointer");
    *stream = (void *)s->stream;
    return EHMM_OK;
}

int ehmm_session_sync(ehmm_session *s) {
    EHMM_TRY(check_session(s));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    return EHMM_OK;
}

int ehmm_session_set_block(ehmm_session *s, int64_t m0, int64_t Mtot) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(m0 >= 0 && m0 + s->M <= Mtot, EHMM_ERR_ARG, "block [%lld, %lld) outside chain of %lld bins", (long long)m0,
                 (long long)(m0 + s->M), (long long)Mtot);
    s->m0 = m0;
    posteriors_changed(s);
    s->Mtot = Mtot;
    return EHMM_OK;
}

int ehmm_set_data(ehmm_session *s, int N, const int32_t *counts, const double *offsets, const double *controls) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(N >= 1 && counts && offsets, EHMM_ERR_ARG, "set_data: N >= 1, counts and offsets required");
    const size_t n = (size_t)s->M * N;
    EHMM_TRY(s->counts_own.reserve(n * sizeof(int32_t)));
    EHMM_TRY(s->offsets_own.reserve(n * sizeof(double)));
    EHMM_CK(cudaMemcpyAsync(s->counts_own.p, counts, n * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
    EHMM_CK(cudaMemcpyAsync(s->offsets_own.p, offsets, n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    const double *dc = nullptr;
    if (controls) {
        EHMM_TRY(s->controls_own.reserve(n * sizeof(double)));
        EHMM_CK(cudaMemcpyAsync(s->controls_own.p, controls, n * sizeof(double), cudaMemcpyHostToDevice, s->stream));
        dc = s->controls_own.as<double>();
        s->ctrl_first = controls[0];
    }
    s->have_ctrl_first = controls != nullptr;
    s->N = N;
    s->groups_consistent = false;
    s->encoded = false;
    s->d_counts = s->counts_own.as<int32_t>();
    s->d_offsets = s->offsets_own.as<double>();
    s->d_controls = dc;
    return scan_counts(s);
}

int ehmm_set_data_device(ehmm_session *s, int N, const int32_t *counts, const double *offsets, const double *controls) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(N >= 1 && counts && offsets, EHMM_ERR_ARG, "set_data: N >= 1, counts and offsets required");
    s->N = N;
    s->groups_consistent = false;
    s->encoded = false;
    s->d_counts = counts;
    s->d_offsets = offsets;
    s->d_controls = controls;
    s->have_ctrl_first = controls != nullptr;
    if (controls) EHMM_CK(cudaMemcpy(&s->ctrl_first, controls, sizeof(double), cudaMemcpyDeviceToHost));
    return scan_counts(s);
}

int ehmm_controls_first(ehmm_session *s, double *value) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(value, EHMM_ERR_ARG, "null pointer");
    EHMM_REQUIRE(s->have_ctrl_first, EHMM_ERR_STATE, "no controls matrix has been set");
    *value = s->ctrl_first;
    return EHMM_OK;
}

int ehmm_set_controls_first(ehmm_session *s, double value) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(isfinite(value), EHMM_ERR_ARG, "controls[1] = %g is not finite", value);
    s->ctrl_first = value;
    s->have_ctrl_first = true;
    return EHMM_OK;
}

int ehmm_set_groups(ehmm_session *s, const int32_t *group, int64_t G, const double *u_chip, const double *u_offsets,
                    const double *u_controls, const uint8_t *u_dsg) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(s->N > 0 && !s->encoded, EHMM_ERR_STATE, "set_groups: call ehmm_set_data first");
    EHMM_REQUIRE(group && G >= 1 && u_chip && u_offsets, EHMM_ERR_ARG, "set_groups: group, G, ChIP and offsets required");
    const size_t n = (size_t)s->M * s->N;
    EHMM_TRY(s->group.reserve(n * sizeof(int32_t)));
    EHMM_CK(cudaMemcpyAsync(s->group.p, group, n * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
    EHMM_TRY(upload_unique(s, G, u_chip, u_offsets, u_controls, u_dsg));
    // validate the ids and find out whether dtUnique matches the data (enables the gather path)
    EHMM_TRY(s->misc.reserve(64));
    int *bad = s->misc.as<int>();
    EHMM_CK(cudaMemsetAsync(bad, 0, sizeof(int), s->stream));
    PROF(s, KID_MISC), group_check_kernel<<<148 * 8, 256, 0, s->stream>>>(
        (int64_t)n, s->group.as<int32_t>(), G, s->d_counts, s->d_offsets, s->d_controls, s->u_chip.as<double>(),
        s->u_off.as<double>(), u_controls ? s->u_ctrl.as<double>() : nullptr, bad);
    EHMM_CK(cudaGetLastError());
    int hb = 0;
    EHMM_CK(cudaMemcpyAsync(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    EHMM_REQUIRE(!(hb & 2), EHMM_ERR_ARG, "set_groups: group ids must lie in [1, G]");
    s->groups_consistent = (hb == 0) && (!s->d_controls || u_controls);
    s->G = G;
    s->rc_columns = 0;
    posteriors_changed(s);
    return EHMM_OK;
}

int ehmm_build_groups(ehmm_session *s, int64_t *G) {
    EHMM_TRY(check_session(s));
    return build_groups(s, G);
}

int ehmm_groups_fetch(ehmm_session *s, int32_t *group, double *u_chip, double *u_offsets, double *u_controls,
                      uint8_t *u_dsg) {
    EHMM_TRY(check_session(s));
    return groups_fetch(s, group, u_chip, u_offsets, u_controls, u_dsg);
}

int ehmm_groups_first_cell(ehmm_session *s, int64_t *first) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(first, EHMM_ERR_ARG, "null pointer");
    EHMM_REQUIRE(s->G > 0 && (int64_t)s->group_first.size() == s->G, EHMM_ERR_STATE,
                 "groups_first_cell: the groups were not built by ehmm_build_groups");
    memcpy(first, s->group_first.data(), sizeof(int64_t) * s->G);
    return EHMM_OK;
}

int ehmm_groups_remap(ehmm_session *s, const int32_t *new_id, int64_t G_new, const double *u_chip, const double *u_offsets,
                      const double *u_controls, const uint8_t *u_dsg) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(new_id && G_new >= 1 && u_chip && u_offsets, EHMM_ERR_ARG, "groups_remap: bad arguments");
    EHMM_TRY(groups_remap(s, new_id, G_new));
    EHMM_TRY(upload_unique(s, G_new, u_chip, u_offsets, u_controls, u_dsg));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    s->G = G_new;
    posteriors_changed(s);
    s->groups_consistent = true;
    s->rc_columns = 0;
    return EHMM_OK;
}

int ehmm_set_design_n(ehmm_session *s, int N, int B, const uint8_t *design) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(N >= 1 && N <= 64 && B >= 1 && B <= EHMM_MAX_B && design, EHMM_ERR_ARG, "set_design: bad N/B");
    memcpy(s->design, design, (size_t)B * N);
    s->N = N;
    s->B = B;
    s->have_design = true;
    return EHMM_OK;
}

static int set_encoded(ehmm_session *s, int N, const int32_t *group, const uint16_t *group16, int64_t G,
                       const double *u_chip, const double *u_offsets, const double *u_controls, const uint8_t *u_dsg);

int ehmm_set_data_encoded(ehmm_session *s, int N, const int32_t *group, int64_t G, const double *u_chip,
                          const double *u_offsets, const double *u_controls, const uint8_t *u_dsg) {
    return set_encoded(s, N, group, nullptr, G, u_chip, u_offsets, u_controls, u_dsg);
}

int ehmm_set_data_encoded16(ehmm_session *s, int N, const uint16_t *group, int64_t G, const double *u_chip,
                            const double *u_offsets, const double *u_controls, const uint8_t *u_dsg) {
    EHMM_REQUIRE(G <= 65535, EHMM_ERR_ARG, "set_data_encoded16: G = %lld does not fit 16 bits", (long long)G);
    return set_encoded(s, N, nullptr, group, G, u_chip, u_offsets, u_controls, u_dsg);
}

// Start copying a Group column (M*N ids of bytes_per_id = 2 or 4 bytes, pinned host memory for a truly
// asynchronous copy) into a staging buffer on the session's copy stream and return at once.  The next
// ehmm_set_data_encoded[16] call that is given the same host pointer takes the staged copy instead of
// copying again, so the upload of table k+1 overlaps the EM iteration on table k.
int ehmm_prefetch_groups(ehmm_session *s, int N, const void *group, int bytes_per_id) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(group && N >= 1 && (bytes_per_id == 2 || bytes_per_id == 4), EHMM_ERR_ARG, "prefetch_groups: bad arguments");
    const int b = s->stage_next;
    const size_t bytes = (size_t)s->M * N * bytes_per_id;
    EHMM_TRY(s->stage[b].reserve(bytes));
    if (s->stage_used_pending[b]) {  // the previous tenant of this buffer may still be read on the main stream
        EHMM_CK(cudaStreamWaitEvent(s->stream_up, s->ev_used[b], 0));
        s->stage_used_pending[b] = false;
    }
    EHMM_CK(cudaMemcpyAsync(s->stage[b].p, group, bytes, cudaMemcpyHostToDevice, s->stream_up));
    EHMM_CK(cudaEventRecord(s->ev_up[b], s->stream_up));
    s->staged_src[b] = group;
    s->staged_bytes[b] = bytes;
    s->stage_next = b ^ 1;
    return EHMM_OK;
}

// staged copy of `src` (bytes long), or -1
static int staged_slot(ehmm_session *s, const void *src, size_t bytes) {
    for (int b = 0; b < 2; b++)
        if (s->staged_src[b] == src && s->staged_bytes[b] == bytes) return b;
    return -1;
}

static int set_encoded(ehmm_session *s, int N, const int32_t *group, const uint16_t *group16, int64_t G,
                       const double *u_chip, const double *u_offsets, const double *u_controls, const uint8_t *u_dsg) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(N >= 1 && (group || group16) && G >= 1 && u_chip && u_offsets, EHMM_ERR_ARG, "set_data_encoded: bad arguments");
    EHMM_REQUIRE((size_t)(G + 1) * 16 <= ((size_t)64 << 20), EHMM_ERR_ARG,
                 "set_data_encoded: %lld groups exceed the dictionary limit; use ehmm_set_data", (long long)G);
    double mx = 0;
    for (int64_t g = 0; g < G; g++) {
        EHMM_REQUIRE(u_chip[g] >= 0 && u_chip[g] == floor(u_chip[g]) && u_chip[g] < 2147483647.0, EHMM_ERR_ARG,
                     "set_data_encoded: dtUnique ChIP[%lld] = %g is not a count", (long long)g, u_chip[g]);
        if (u_chip[g] > mx) mx = u_chip[g];
    }
    s->N = N;
    s->d_counts = nullptr;
    s->d_offsets = nullptr;
    s->d_controls = nullptr;
    s->encoded = true;
    s->encoded_ctrl = u_controls != nullptr;
    s->have_ctrl_first = u_controls != nullptr;
    if (u_controls) {
        const int64_t g0 = group ? (int64_t)group[0] : (int64_t)group16[0];
        EHMM_REQUIRE(g0 >= 1 && g0 <= G, EHMM_ERR_ARG, "set_data_encoded: group ids must lie in [1, G]");
        s->ctrl_first = u_controls[g0 - 1];
    }
    s->count_max = (int64_t)mx;
    const size_t n = (size_t)s->M * N;
    EHMM_TRY(s->group.reserve(n * sizeof(int32_t)));
    const int slot = group ? staged_slot(s, group, n * sizeof(int32_t)) : staged_slot(s, group16, n * sizeof(uint16_t));
    if (slot >= 0) {  // prefetched (ehmm_prefetch_groups): wait for the copy on the device, not on the host
        EHMM_CK(cudaStreamWaitEvent(s->stream, s->ev_up[slot], 0));
        if (group) {
            EHMM_CK(cudaMemcpyAsync(s->group.p, s->stage[slot].p, n * sizeof(int32_t), cudaMemcpyDeviceToDevice, s->stream));
        } else {
            PROF(s, KID_MISC), widen_u16_kernel<<<148 * 8, 256, 0, s->stream>>>((int64_t)n, s->stage[slot].as<uint16_t>(),
                                                                           s->group.as<int32_t>());
            EHMM_CK(cudaGetLastError());
        }
        EHMM_CK(cudaEventRecord(s->ev_used[slot], s->stream));
        s->stage_used_pending[slot] = true;
        s->staged_src[slot] = nullptr;  // a staged copy is consumed once
    } else if (group) {
        EHMM_CK(cudaMemcpyAsync(s->group.p, group, n * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
    } else {
        EHMM_TRY(s->rc_tmp.reserve(n * sizeof(uint16_t)));
        EHMM_CK(cudaMemcpyAsync(s->rc_tmp.p, group16, n * sizeof(uint16_t), cudaMemcpyHostToDevice, s->stream));
        PROF(s, KID_MISC), widen_u16_kernel<<<148 * 8, 256, 0, s->stream>>>((int64_t)n, s->rc_tmp.as<uint16_t>(),
                                                                       s->group.as<int32_t>());
        EHMM_CK(cudaGetLastError());
    }
    EHMM_TRY(upload_unique(s, G, u_chip, u_offsets, u_controls, u_dsg));
    EHMM_TRY(s->misc.reserve(64));
    int *bad = s->misc.as<int>();
    EHMM_CK(cudaMemsetAsync(bad, 0, sizeof(int), s->stream));
    PROF(s, KID_MISC), group_range_kernel<<<148 * 8, 256, 0, s->stream>>>((int64_t)n, s->group.as<int32_t>(), G, bad);
    EHMM_CK(cudaGetLastError());
    EHMM_CK(cudaMemcpyAsync(s->h_pinned + 300, bad, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    int hb;
    memcpy(&hb, s->h_pinned + 300, sizeof(int));
    EHMM_REQUIRE(hb == 0, EHMM_ERR_ARG, "set_data_encoded: group ids must lie in [1, G]");
    s->groups_consistent = true;
    s->G = G;
    s->rc_columns = 0;
    posteriors_changed(s);
    return EHMM_OK;
}

int ehmm_set_design(ehmm_session *s, int B, const uint8_t *design) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(s->N > 0, EHMM_ERR_STATE, "set_design: call ehmm_set_data first");
    EHMM_REQUIRE(B >= 1 && B <= EHMM_MAX_B && design, EHMM_ERR_ARG, "set_design: 1 <= B <= %d", EHMM_MAX_B);
    EHMM_REQUIRE(s->N <= 64, EHMM_ERR_ARG, "set_design: at most 64 samples");
    memcpy(s->design, design, (size_t)B * s->N);
    s->B = B;
    s->have_design = true;
    return EHMM_OK;
}

int ehmm_loglik_consensus(ehmm_session *s, const double *theta1, const double *theta2, int P, double minZero) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(theta1 && theta2, EHMM_ERR_ARG, "null parameters");
    return emis_consensus(s, theta1, theta2, P, minZero);
}

int ehmm_loglik_consensus_zinb(ehmm_session *s, const double *theta1, const double *theta2, int Q, double minZero) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(theta1 && theta2, EHMM_ERR_ARG, "null parameters");
    EHMM_REQUIRE(Q == 3 || Q == 5, EHMM_ERR_ARG, "theta1 must hold 3 (lambda, beta, size) or 5 (with controls) entries");
    const int nl = (Q - 1) / 2;  // zero-inflation coefficients come first, then the NB part
    return emis_consensus(s, theta1 + nl, theta2, Q - nl, minZero, theta1);
}

int ehmm_loglik_init(ehmm_session *s, const double *psi1, const double *psi2) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(psi1 && psi2, EHMM_ERR_ARG, "null parameters");
    return emis_init(s, psi1, psi2);
}

int ehmm_loglik_differential_hmm_zinb(ehmm_session *s, const double *delta, const double *psi, double minZero) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(delta && psi, EHMM_ERR_ARG, "null parameters");
    return emis_differential_hmm(s, delta, psi, minZero, 1);
}

int ehmm_loglik_differential_hmm(ehmm_session *s, const double *delta, const double *psi, double minZero) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(delta && psi, EHMM_ERR_ARG, "null parameters");
    return emis_differential_hmm(s, delta, psi, minZero);
}

int ehmm_inner_expstep(ehmm_session *s, const double *delta, const double *psi, double minZero) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(delta && psi, EHMM_ERR_ARG, "null parameters");
    return emis_inner_estep(s, delta, psi, minZero);
}

int ehmm_inner_expstep_zinb(ehmm_session *s, const double *delta, const double *psi, double minZero) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(delta && psi, EHMM_ERR_ARG, "null parameters");
    return emis_inner_estep(s, delta, psi, minZero, 1);
}

int ehmm_expstep(ehmm_session *s, const double *pi, const double *gamma, const double *logf) {
    EHMM_TRY(check_session(s));
    EHMM_TRY(set_theta(s, pi, gamma));
    const double *d_logf = nullptr;
    EHMM_TRY(upload_logf(s, logf, &d_logf));
    EHMM_TRY(fb_expstep(s, d_logf));
    return estep_done(s);
}

int ehmm_expstep_device(ehmm_session *s, const double *pi, const double *gamma, const double *logf_device) {
    EHMM_TRY(check_session(s));
    EHMM_TRY(set_theta(s, pi, gamma));
    EHMM_REQUIRE(logf_device, EHMM_ERR_ARG, "null logf");
    // the caller owns this emission: nothing can be owed against it, so the datasets are written at once
    const bool lean = s->lean_mode;
    s->lean_mode = false;
    const int rc = fb_expstep(s, logf_device);
    s->lean_mode = lean;
    EHMM_TRY(rc);
    return estep_done(s);
}

int ehmm_expstep_reduce(ehmm_session *s, const double *pi, const double *gamma, const double *logf, double *op) {
    EHMM_TRY(check_session(s));
    EHMM_TRY(set_theta(s, pi, gamma));
    EHMM_REQUIRE(op, EHMM_ERR_ARG, "null op");
    const double *d_logf = nullptr;
    EHMM_TRY(upload_logf(s, logf, &d_logf));
    EHMM_TRY(fb_reduce(s, d_logf, op));
    s->reduce_pending = true;
    return EHMM_OK;
}

int ehmm_expstep_apply(ehmm_session *s, const double *fwd_carry, const double *bwd_carry, double *fwd_out) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(s->reduce_pending, EHMM_ERR_STATE, "expstep_apply without expstep_reduce");
    EHMM_REQUIRE(fwd_carry && bwd_carry, EHMM_ERR_ARG, "null carry");
    EHMM_TRY(fb_apply(s, fwd_carry, bwd_carry, fwd_out));
    s->reduce_pending = false;
    return estep_done(s);
}

namespace {
// exact power-of-two renormalisation of a small non-negative vector with a natural-log scale
void renorm_host(double *m, int n, double &ls) {
    double mx = m[0];
    for (int i = 1; i < n; i++) mx = m[i] > mx ? m[i] : mx;
    if (!(mx > 0.0) || !std::isfinite(mx)) return;
    int e;
    frexp(mx, &e);
    e -= 1;  // floor(log2(mx))
    const double sc = ldexp(1.0, -e);
    for (int i = 0; i < n; i++) m[i] *= sc;
    ls += e * 0.693147180559945309417232121458;
}
}  // namespace

// The sharded E-step's second half with the carries composed here: ops_all holds the P block
// operators (K*K row-major mantissa + log scale each) gathered from all ranks; this block's forward
// carry is pi * prod(ops[:rank]), its backward carry prod(ops[rank+1:]) * 1.
int ehmm_expstep_apply_ops(ehmm_session *s, const double *ops_all, int P, int rank) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(s->reduce_pending, EHMM_ERR_STATE, "expstep_apply_ops without expstep_reduce");
    EHMM_REQUIRE(ops_all && P >= 1 && rank >= 0 && rank < P, EHMM_ERR_ARG, "bad operator list");
    const int K = s->K, ON = K * K + 1;
    double fwd[EHMM_MAX_K + 1], bwd[EHMM_MAX_K + 1], t[EHMM_MAX_K];
    for (int k = 0; k < K; k++) { fwd[k] = s->h_pi[k]; bwd[k] = 1.0; }
    fwd[K] = 0.0;
    bwd[K] = 0.0;
    for (int r = 0; r < rank; r++) {
        const double *A = ops_all + (size_t)r * ON;
        for (int l = 0; l < K; l++) {
            double acc = 0.0;
            for (int i = 0; i < K; i++) acc += fwd[i] * A[i * K + l];
            t[l] = acc;
        }
        for (int l = 0; l < K; l++) fwd[l] = t[l];
        fwd[K] += A[K * K];
        renorm_host(fwd, K, fwd[K]);
    }
    for (int r = P - 1; r > rank; r--) {
        const double *A = ops_all + (size_t)r * ON;
        for (int i = 0; i < K; i++) {
            double acc = 0.0;
            for (int l = 0; l < K; l++) acc += A[i * K + l] * bwd[l];
            t[i] = acc;
        }
        for (int i = 0; i < K; i++) bwd[i] = t[i];
        bwd[K] += A[K * K];
        renorm_host(bwd, K, bwd[K]);
    }
    EHMM_TRY(fb_apply(s, fwd, bwd, nullptr));
    s->reduce_pending = false;
    return estep_done(s);
}

int ehmm_estep_stats(ehmm_session *s, double *stats) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(avail(s, DS_P1 | DS_P2) && stats, EHMM_ERR_STATE, "estep_stats: expStep has not run");
    EHMM_TRY(fb_fetch_stats(s));
    for (int i = 0; i < s->K * s->K + 1; i++) stats[i] = s->h_stats[i];
    return EHMM_OK;
}

int ehmm_estep_row0(ehmm_session *s, double *row0) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(avail(s, DS_P1) && row0, EHMM_ERR_STATE, "estep_row0: expStep has not run");
    EHMM_TRY(fb_fetch_stats(s));
    for (int k = 0; k < s->K; k++) row0[k] = s->h_lp1_row0[k];
    return EHMM_OK;
}

int ehmm_estep_set_stats(ehmm_session *s, const double *stats, const double *logprob1_row0, double loglik) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(stats && logprob1_row0, EHMM_ERR_ARG, "null pointer");
    for (int i = 0; i < s->K * s->K + 1; i++) s->h_stats[i] = stats[i];
    for (int k = 0; k < s->K; k++) s->h_lp1_row0[k] = logprob1_row0[k];
    s->h_ll = loglik;
    s->stats_on_host = true;
    return EHMM_OK;
}

int ehmm_loglik_value(ehmm_session *s, double *ll) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(avail(s, DS_FP) && ll, EHMM_ERR_STATE, "loglik: expStep has not run");
    EHMM_TRY(fb_fetch_stats(s));
    *ll = s->h_ll;
    return EHMM_OK;
}

// maxStepProb (src/maxStepProb.cpp:58-77): the sums over bins come from the statistics fused into
// the E-step kernel; only the K*K divisions, the 1e-6 floor and the renormalisation run here.
int ehmm_maxstepprob(ehmm_session *s, double *pi_out, double *gamma_out) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(avail(s, DS_P1 | DS_P2), EHMM_ERR_STATE, "maxStepProb: expStep has not run for '%s'", s->key.c_str());
    EHMM_REQUIRE(pi_out && gamma_out, EHMM_ERR_ARG, "null output");
    EHMM_TRY(fb_fetch_stats(s));
    const int K = s->K;
    for (int k1 = 0; k1 < K; k1++) {
        double denom = 0;
        for (int k2 = 0; k2 < K; k2++) denom += s->h_stats[k1 * K + k2];
        for (int k2 = 0; k2 < K; k2++) gamma_out[k1 + k2 * K] = s->h_stats[k1 * K + k2] / denom;
    }
    double sum = 0;
    for (int k = 0; k < K; k++) {
        double p = exp(s->h_lp1_row0[k]);
        if (p < 0.000001) p = 0.000001;
        pi_out[k] = p;
    }
    {   // Armadillo accu order: two running sums
        double v1 = 0, v2 = 0;
        for (int k = 0; k < K; k++) ((k & 1) ? v2 : v1) += pi_out[k];
        sum = v1 + v2;
    }
    for (int k = 0; k < K; k++) pi_out[k] /= sum;
    for (int i = 0; i < K * K; i++)
        if (gamma_out[i] < 0.000001) gamma_out[i] = 0.000001;
    for (int k1 = 0; k1 < K; k1++) {
        double v1 = 0, v2 = 0;
        for (int k2 = 0; k2 < K; k2++) ((k2 & 1) ? v2 : v1) += fabs(gamma_out[k1 + k2 * K]);
        const double nrm = v1 + v2;
        for (int k2 = 0; k2 < K; k2++) gamma_out[k1 + k2 * K] /= nrm;
    }
    return EHMM_OK;
}

// computeQFunction (src/computeQFunction.cpp:30)
int ehmm_qfunction(ehmm_session *s, const double *pi, const double *gamma, double *q) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(avail(s, DS_LOGF | DS_P1 | DS_P2), EHMM_ERR_STATE, "computeQFunction: expStep has not run");
    EHMM_REQUIRE(pi && gamma && q, EHMM_ERR_ARG, "null pointer");
    EHMM_TRY(fb_fetch_stats(s));
    const int K = s->K;
    double q0 = 0, q2 = 0;
    for (int k = 0; k < K; k++) q0 += exp(s->h_lp1_row0[k]) * log(pi[k]);
    for (int i = 0; i < K; i++)
        for (int l = 0; l < K; l++) q2 += s->h_stats[i * K + l] * log(gamma[i + l * K]);
    *q = q0 + s->h_stats[K * K] + q2;
    return EHMM_OK;
}

// computeBIC (src/computeBIC.cpp:23-26)
int ehmm_bic(ehmm_session *s, int numPar, int numSamples, double *bic) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(avail(s, DS_FP) && bic, EHMM_ERR_STATE, "computeBIC: expStep has not run");
    EHMM_TRY(fb_fetch_stats(s));
    *bic = -2 * s->h_ll + numPar * log((double)((int64_t)numSamples * s->Mtot));
    return EHMM_OK;
}

int ehmm_inner_maxstepprob(ehmm_session *s, double *delta_out) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(delta_out, EHMM_ERR_ARG, "null output");
    return inner_maxstepprob(s, delta_out);
}

int ehmm_marginal_probability(ehmm_session *s, double *out) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(avail(s, DS_P1) && out, EHMM_ERR_STATE, "getMarginalProbability: expStep has not run");
    EHMM_TRY(ensure(s, DS_P1));  // exp(logProb1), the reference's expression, not the linear copy
    const int64_t n = s->M * s->K;
    EHMM_TRY(s->rc_tmp.reserve(sizeof(double) * n));
    PROF(s, KID_MISC), exp_kernel<<<148 * 8, 256, 0, s->stream>>>(s->logProb1.as<double>(), n, s->rc_tmp.as<double>());
    EHMM_CK(cudaGetLastError());
    EHMM_CK(cudaMemcpyAsync(out, s->rc_tmp.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    return EHMM_OK;
}

int ehmm_viterbi(ehmm_session *s, const double *pi, const double *gamma, double *states_out) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(avail(s, DS_FP), EHMM_ERR_STATE, "computeViterbiSequence: no logFP dataset (expStep has not run)");
    EHMM_REQUIRE(pi && gamma, EHMM_ERR_ARG, "null pointer");
    EHMM_TRY(ensure(s, DS_FP));
    return viterbi(s, pi, gamma, states_out);
}

int ehmm_viterbi_forward(ehmm_session *s, const double *pi, const double *gamma, const double *v_in, double *v_out) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(avail(s, DS_FP), EHMM_ERR_STATE, "computeViterbiSequence: no logFP dataset (expStep has not run)");
    EHMM_REQUIRE(pi && gamma && v_out, EHMM_ERR_ARG, "null pointer");
    EHMM_TRY(ensure(s, DS_FP));
    return viterbi_forward(s, pi, gamma, v_in, v_out);
}

int ehmm_viterbi_backtrace(ehmm_session *s, int state_last, int *state_prev, double *states_out) {
    EHMM_TRY(check_session(s));
    return viterbi_backtrace(s, state_last, state_prev, states_out);
}

int ehmm_call_peaks(ehmm_session *s, double fdr, uint8_t *calls, int64_t *n_called, double *cutoff) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(avail(s, DS_P1), EHMM_ERR_STATE, "callPeaks: no logProb1 dataset (expStep has not run)");
    EHMM_TRY(ensure(s, DS_P1));
    return call_peaks_fdr(s, fdr, calls, n_called, cutoff);
}

int ehmm_call_peaks_viterbi(ehmm_session *s, uint8_t *calls, int64_t *n_called) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(has(s, DS_VIT), EHMM_ERR_STATE, "callPeaks(method = 'viterbi'): computeViterbiSequence has not run");
    return call_peaks_viterbi(s, calls, n_called);
}

int ehmm_call_peaks_exact(ehmm_session *s, int force) {
    EHMM_REQUIRE(s, EHMM_ERR_ARG, "null session");
    s->peaks_exact = force != 0;
    return EHMM_OK;
}

int ehmm_call_peaks_path(ehmm_session *s, int *path) {
    EHMM_REQUIRE(s && path, EHMM_ERR_ARG, "null pointer");
    *path = s->peaks_path;
    return EHMM_OK;
}

int ehmm_peak_runs(ehmm_session *s, const int64_t *chrom_starts, int n_chrom_starts, int64_t cap, int64_t *starts, int64_t *ends,
                   int64_t *n_runs) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(n_runs && n_chrom_starts >= 0 && (n_chrom_starts == 0 || chrom_starts) && cap >= 0, EHMM_ERR_ARG, "peak_runs: bad arguments");
    return peak_runs(s, chrom_starts, n_chrom_starts, cap, starts, ends, n_runs);
}

// every dataset the reference's HDF5 file holds at the end of a fit, in one call.  Armadillo writes an
// M x K column-major matrix as an HDF5 dataset of dims (K, M) (src/expStep.cpp:116-122): the bytes copied
// here are exactly that dataset's, so the glue hands each buffer to H5Dwrite as it is.  NULL = skip.
int ehmm_flush(ehmm_session *s, double *logLikelihood, double *logFP, double *logBP, double *logProb1, double *logProb2,
               double *mixtureProb, double *viterbi) {
    EHMM_TRY(check_session(s));
    const int64_t M = s->M;
    const int K = s->K;
    struct { double *dst; DevBuf *buf; int64_t n; unsigned bit; const char *name; } d[] = {
        {logLikelihood, &s->logf, M * K, DS_LOGF, "logLikelihood"}, {logFP, &s->logFP, M * K, DS_FP, "logFP"},
        {logBP, &s->logBP, M * K, DS_BP, "logBP"},                  {logProb1, &s->logProb1, M * K, DS_P1, "logProb1"},
        {logProb2, &s->logProb2, M * K * K, DS_P2, "logProb2"},     {mixtureProb, &s->eta, M * s->B, DS_ETA, "mixtureProb"},
        {viterbi, &s->viterbi, M, DS_VIT, "viterbi"}};
    for (auto &e : d) {
        if (!e.dst) continue;
        EHMM_REQUIRE(avail(s, e.bit), EHMM_ERR_STATE, "flush: dataset '%s' has not been written", e.name);
        EHMM_TRY(ensure(s, e.bit));
    }
    for (auto &e : d)
        if (e.dst) EHMM_CK(cudaMemcpyAsync(e.dst, e.buf->p, sizeof(double) * e.n, cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    return EHMM_OK;
}

int ehmm_expstep_mode(ehmm_session *s, int lean) {
    EHMM_REQUIRE(s && (lean == 0 || lean == 1), EHMM_ERR_ARG, "expstep_mode: 0 (all datasets at once) or 1 (EM-internal)");
    s->lean_mode = lean != 0;
    return EHMM_OK;
}

int ehmm_materialize(ehmm_session *s) {
    EHMM_TRY(check_session(s));
    return fb_materialize(s);
}

int ehmm_viterbi_mode(ehmm_session *s, int mode) {
    EHMM_REQUIRE(s && (mode == 0 || mode == 1), EHMM_ERR_ARG, "viterbi_mode: 0 (parallel scan) or 1 (sequential kernel)");
    s->vit_mode = mode;
    return EHMM_OK;
}

int ehmm_viterbi_info(ehmm_session *s, int64_t *info4) {
    EHMM_REQUIRE(s && info4, EHMM_ERR_ARG, "null pointer");
    for (int i = 0; i < 4; i++) info4[i] = s->vit_info[i];
    return EHMM_OK;
}

int ehmm_rng_set_state(ehmm_session *s, const uint32_t *state625) {
    EHMM_REQUIRE(s && state625, EHMM_ERR_ARG, "null pointer");
    EHMM_REQUIRE(state625[0] <= 624, EHMM_ERR_ARG, "mti = %u out of range", state625[0]);
    memcpy(s->rng_state, state625, sizeof(uint32_t) * 625);
    s->rng_set = true;
    s->rng_epoch++;
    s->specr_cols = 0;
    s->rng_dev_valid = false;
    s->rng_pending = false;
    s->spec_n = 0;
    return EHMM_OK;
}

int ehmm_rng_get_state(ehmm_session *s, uint32_t *state625) {
    EHMM_REQUIRE(s && state625, EHMM_ERR_ARG, "null pointer");
    EHMM_TRY(rng_sync_host(s));
    memcpy(state625, s->rng_state, sizeof(uint32_t) * 625);
    return EHMM_OK;
}

int ehmm_rc_consensus(ehmm_session *s, double p, const double *uniforms, int64_t n_uniforms, int64_t *consumed) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(has_post(s), EHMM_ERR_STATE, "consensusRejectionControlled: expStep has not run");
    EHMM_REQUIRE(s->G > 0, EHMM_ERR_STATE, "consensusRejectionControlled: ehmm_set_groups has not been called");
    EHMM_REQUIRE(p >= 0.0 && p <= 1.0, EHMM_ERR_ARG, "the rejection cut p = %g is not a probability", p);
    return rc_consensus(s, p, uniforms, n_uniforms, consumed);
}

int ehmm_rc_differential(ehmm_session *s, double p, int N, const double *uniforms, int64_t n_uniforms, int64_t *consumed) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(has_post(s) && has(s, DS_ETA), EHMM_ERR_STATE, "differentialRejectionControlled: needs logProb1 and mixtureProb");
    EHMM_REQUIRE(s->G > 0, EHMM_ERR_STATE, "differentialRejectionControlled: ehmm_set_groups has not been called");
    EHMM_REQUIRE(p >= 0.0 && p <= 1.0, EHMM_ERR_ARG, "the rejection cut p = %g is not a probability", p);
    EHMM_REQUIRE(N == s->N, EHMM_ERR_ARG, "N=%d does not match the data (N=%d)", N, s->N);
    return rc_differential(s, p, N, uniforms, n_uniforms, consumed);
}

int ehmm_rc_count(ehmm_session *s, int differential, double p, int64_t *counts) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(counts, EHMM_ERR_ARG, "null pointer");
    EHMM_REQUIRE(has_post(s) && (!differential || has(s, DS_ETA)), EHMM_ERR_STATE, "rc_count: posteriors not available");
    EHMM_REQUIRE(s->G > 0, EHMM_ERR_STATE, "rc_count: ehmm_set_groups has not been called");
    EHMM_REQUIRE(p >= 0.0 && p <= 1.0, EHMM_ERR_ARG, "the rejection cut p = %g is not a probability", p);
    return rc_count(s, differential, p, counts);
}

int ehmm_rc_apply_at(ehmm_session *s, int differential, double p, const int64_t *col_offsets, int64_t total) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(col_offsets && total >= 0, EHMM_ERR_ARG, "bad arguments");
    return rc_apply_at(s, differential, p, col_offsets, total);
}

int ehmm_rc_buckets_device(ehmm_session *s, void **ptr, int64_t *n, int *is_integer) {
    EHMM_REQUIRE(s && ptr && n && is_integer, EHMM_ERR_ARG, "null pointer");
    EHMM_REQUIRE(s->rc_columns > 0, EHMM_ERR_STATE, "no rejection tables");
    const int64_t cells = (int64_t)s->rc_columns * (s->G + 1);
    if (s->rc_exact) {
        EHMM_TRY(check_session(s));
        *ptr = s->rc_limbs.p;
        EHMM_TRY(rc_fold(s, n));  // the replicas of the frequent groups are folded in: two plain limb arrays remain
        *is_integer = 1;
        s->rc_final = false;  // the caller is about to change the limbs: round them again
    } else {
        *ptr = s->rc_buckets.p;
        *n = cells;
        *is_integer = 0;
    }
    return EHMM_OK;
}

int ehmm_rc_hint(ehmm_session *s, double p) {
    EHMM_REQUIRE(s && p >= 0.0 && p <= 1.0, EHMM_ERR_ARG, "rc_hint: p must lie in [0, 1]");
    s->rc_hint_p = p;
    return EHMM_OK;
}

int ehmm_rc_finalize(ehmm_session *s) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(s->rc_columns > 0, EHMM_ERR_STATE, "no rejection tables");
    return rc_finalize(s);
}

int ehmm_inner_mstep_sums(ehmm_session *s, double *sums) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(sums, EHMM_ERR_ARG, "null output");
    return inner_mstep_sums(s, sums);
}

int ehmm_rc_table_rows(ehmm_session *s, int column, int64_t *rows) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(rows && column >= 0 && column < s->rc_columns, EHMM_ERR_ARG, "bad column %d (have %d)", column, s->rc_columns);
    return rc_table_rows(s, column, rows);
}

int ehmm_rc_table_fetch(ehmm_session *s, int column, double *sums, double *ids) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(sums && ids && column >= 0 && column < s->rc_columns, EHMM_ERR_ARG, "bad column %d (have %d)", column, s->rc_columns);
    return rc_table_fetch(s, column, sums, ids);
}

int ehmm_rc_buckets_fetch(ehmm_session *s, int column, double *buckets) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(buckets && column >= 0 && column < s->rc_columns, EHMM_ERR_ARG, "bad column %d (have %d)", column, s->rc_columns);
    EHMM_TRY(rc_finalize(s));
    EHMM_CK(cudaMemcpyAsync(buckets, s->rc_buckets.as<double>() + (size_t)column * (s->G + 1), sizeof(double) * (s->G + 1),
                            cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    return EHMM_OK;
}

int ehmm_reweight(int device, double *x, int64_t n, double p, const double *uniforms, int64_t n_uniforms, int64_t *consumed) {
    EHMM_REQUIRE(x && n >= 0, EHMM_ERR_ARG, "null pointer");
    return rc_reweight(device, x, n, p, uniforms, n_uniforms, consumed);
}

int ehmm_aggregate(int device, const double *x, int64_t n, const int32_t *f, int64_t G, double *buckets) {
    EHMM_REQUIRE(x && f && buckets && n >= 0 && G >= 0, EHMM_ERR_ARG, "null pointer");
    return rc_aggregate(device, x, n, f, G, buckets);
}

int ehmm_glm_nb(ehmm_session *s, int column, const double *par, int P, double *value, double *grad) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(par && value, EHMM_ERR_ARG, "null pointer");
    EHMM_REQUIRE(column >= 0 && column < s->rc_columns, EHMM_ERR_STATE, "glmNB: rejection table %d not available", column);
    EHMM_TRY(rc_finalize(s));
    return glm_nb(s, column, par, P, value, grad);
}

int ehmm_glm_zinb(ehmm_session *s, int column, const double *par, int P, double minZero, double *value, double *grad) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(par && value, EHMM_ERR_ARG, "null pointer");
    EHMM_TRY(rc_finalize(s));
    return glm_zinb(s, column, par, P, minZero, value, grad);
}

int ehmm_glm_nb_batch(ehmm_session *s, int n, const int *columns, const double *pars, int P, double *values, double *grads) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(columns && pars && values, EHMM_ERR_ARG, "null pointer");
    EHMM_TRY(rc_finalize(s));
    return glm_nb_batch(s, n, columns, pars, P, values, grads);
}

int ehmm_glm_mix_nb(ehmm_session *s, const double *par, double minZero, double *value, double *grad) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(par && value, EHMM_ERR_ARG, "null pointer");
    EHMM_REQUIRE(s->rc_columns == s->B + 2 && s->B > 0, EHMM_ERR_STATE, "glmMixNB: differential rejection tables not available");
    EHMM_TRY(rc_finalize(s));
    return glm_mix_nb(s, par, minZero, value, grad);
}

int ehmm_glm_mix_zinb(ehmm_session *s, const double *par, double minZero, double *value, double *grad) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(par && value, EHMM_ERR_ARG, "null pointer");
    EHMM_REQUIRE(s->rc_columns == s->B + 2 && s->B > 0, EHMM_ERR_STATE, "glmMixZINB: differential rejection tables not available");
    EHMM_TRY(rc_finalize(s));
    return glm_mix_zinb(s, par, minZero, value, grad);
}

int ehmm_profile_kernels(void) { return KID_COUNT; }

const char *ehmm_profile_name(int kernel_id) {
    return (kernel_id >= 0 && kernel_id < KID_COUNT) ? kKernelNames[kernel_id] : "";
}

int ehmm_profile(ehmm_session *s, int enable) {
    EHMM_TRY(check_session(s));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    for (int i = 0; i < KID_COUNT; i++) s->launches[i] = 0;
    s->prof_recs.clear();
    s->prof_used = 0;
    s->prof_on = enable != 0;
    return EHMM_OK;
}

int ehmm_profile_read(ehmm_session *s, int64_t *launches, double *ms) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(launches && ms, EHMM_ERR_ARG, "null pointer");
    EHMM_CK(cudaStreamSynchronize(s->stream));
    for (int i = 0; i < KID_COUNT; i++) { launches[i] = s->launches[i]; ms[i] = 0.0; }
    for (const auto &r : s->prof_recs) {
        float t = 0.f;
        EHMM_CK(cudaEventElapsedTime(&t, r.e0, r.e1));
        ms[r.kid] += t;
    }
    return EHMM_OK;
}

int ehmm_dataset_dims(ehmm_session *s, const char *name, int64_t *rows, int64_t *cols) {
    EHMM_REQUIRE(s && name && rows && cols, EHMM_ERR_ARG, "null pointer");
    DS d;
    EHMM_TRY(lookup(s, name, d));
    EHMM_REQUIRE(avail(s, d.bit), EHMM_ERR_STATE, "dataset '%s' has not been written", name);
    *rows = d.rows;
    *cols = d.cols;
    return EHMM_OK;
}

int ehmm_fetch(ehmm_session *s, const char *name, double *dst, int64_t n) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(name && dst, EHMM_ERR_ARG, "null pointer");
    DS d;
    EHMM_TRY(lookup(s, name, d));
    EHMM_REQUIRE(avail(s, d.bit), EHMM_ERR_STATE, "dataset '%s' has not been written", name);
    EHMM_TRY(ensure(s, d.bit));
    EHMM_REQUIRE(n == d.rows * d.cols, EHMM_ERR_ARG, "dataset '%s' has %lld elements, buffer has %lld", name,
                 (long long)(d.rows * d.cols), (long long)n);
    EHMM_CK(cudaMemcpyAsync(dst, d.buf->p, sizeof(double) * n, cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    return EHMM_OK;
}

int ehmm_device_ptr(ehmm_session *s, const char *name, void **ptr) {
    EHMM_REQUIRE(s && name && ptr, EHMM_ERR_ARG, "null pointer");
    DS d;
    EHMM_TRY(lookup(s, name, d));
    EHMM_REQUIRE(avail(s, d.bit), EHMM_ERR_STATE, "dataset '%s' has not been written", name);
    EHMM_TRY(check_session(s));
    EHMM_TRY(ensure(s, d.bit));
    *ptr = d.buf->p;
    return EHMM_OK;
}

int ehmm_store(ehmm_session *s, const char *name, const double *src, int64_t rows, int64_t cols) {
    EHMM_TRY(check_session(s));
    EHMM_REQUIRE(name && src, EHMM_ERR_ARG, "null pointer");
    DS d;
    if (!strcmp(name, "mixtureProb") && s->B == 0) s->B = (int)cols;
    EHMM_TRY(lookup(s, name, d));
    EHMM_REQUIRE(rows == d.rows && cols == d.cols, EHMM_ERR_ARG, "dataset '%s' must be %lld x %lld (got %lld x %lld)", name,
                 (long long)d.rows, (long long)d.cols, (long long)rows, (long long)cols);
    // a dataset of the last E-step is about to be replaced: the others must exist as that E-step left them
    if (d.bit & DS_ESTEP) EHMM_TRY(ensure(s, DS_ESTEP));
    if (d.bit == DS_LOGF) EHMM_TRY(logf_for_write(s, sizeof(double) * rows * cols));
    else EHMM_TRY(d.buf->reserve(sizeof(double) * rows * cols));
    if (d.bit & DS_ESTEP) { s->lazy_full = false; s->valid &= ~(unsigned)DS_P1LIN; }
    EHMM_CK(cudaMemcpyAsync(d.buf->p, src, sizeof(double) * rows * cols, cudaMemcpyHostToDevice, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    s->valid |= d.bit;
    posteriors_changed(s);
    if (d.bit & (DS_ESTEP | DS_LOGF)) {
        // a dataset was replaced from the host: the fused statistics no longer describe it
        s->stats_on_host = false;
        s->stats_stale = true;
    }
    return EHMM_OK;
}

}  // extern "C"
