// common.cuh -- session state and helpers shared by the CUDA translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string>
#include <utility>
#include <vector>
#include "../../include/ehmm.h"

namespace ehmm {

void set_error(const char *fmt, ...);

#define EHMM_CK(call)                                                                  \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) {                                                       \
            ehmm::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return (e_ == cudaErrorMemoryAllocation) ? EHMM_ERR_NOMEM : EHMM_ERR_CUDA; \
        }                                                                              \
    } while (0)

#define EHMM_REQUIRE(cond, code, ...)      \
    do {                                   \
        if (!(cond)) {                     \
            ehmm::set_error(__VA_ARGS__);  \
            return (code);                 \
        }                                  \
    } while (0)

#define EHMM_TRY(expr)              \
    do {                            \
        int rc_ = (expr);           \
        if (rc_ != EHMM_OK) return rc_; \
    } while (0)

// grow-only device buffer
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int reserve(size_t n) {
        if (n <= bytes) return EHMM_OK;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu bytes): %s", n, cudaGetErrorString(e));
            return EHMM_ERR_NOMEM;
        }
        bytes = n;
        return EHMM_OK;
    }
    // for buffers whose size follows the data (uniform draws per call): a request beyond the capacity allocates a
    // quarter more than asked, so that the next slightly larger call does not free and allocate again (cudaFree
    // waits for the device: 10 ms per occurrence at a 2 GB buffer)
    int reserve_growing(size_t n) { return n <= bytes ? EHMM_OK : reserve(n + n / 4); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

enum : unsigned {
    DS_LOGF = 1, DS_FP = 2, DS_BP = 4, DS_P1 = 8, DS_P2 = 16, DS_ETA = 32, DS_VIT = 64,
    DS_P1LIN = 128,  // the posteriors P1 themselves (the lean E-step writes these instead of logProb1)
    DS_ESTEP = DS_FP | DS_BP | DS_P1 | DS_P2
};

// kernel ids for the launch counters / optional CUDA-event profile (ehmm_profile*)
enum KernelId : int {
    KID_EMIS_TAB, KID_EMIS_GATHER, KID_EMIS_CONSENSUS, KID_EMIS_DIFF, KID_EMIS_INNER, KID_INNER_MSTEP,
    KID_FB_REDUCE, KID_FB_CARRY, KID_FB_APPLY, KID_FB_APPLY_LEAN, KID_FB_FIX, KID_FB_STATS,
    KID_RC_COUNT, KID_RC_SCAN, KID_MT, KID_RC_APPLY, KID_RC_STATIC, KID_RC_FINAL, KID_GLM, KID_GLM_FINAL,
    KID_VIT_FWD, KID_VIT_BWD, KID_PEAKS, KID_MISC, KID_COUNT
};
static const char *const kKernelNames[KID_COUNT] = {
    "emis_tables", "emis_gather", "emis_consensus", "emis_diff_hmm", "emis_inner", "inner_mstep",
    "fb_reduce", "fb_carry", "fb_apply", "fb_apply_em", "fb_boundary_fix", "fb_stats",
    "rc_count", "rc_scan", "mt_generate", "rc_apply", "rc_static", "rc_finalize", "glm", "glm_final",
    "viterbi_forward", "viterbi_backtrace", "call_peaks", "misc"};

// forward-backward tiling (fb.cu)
constexpr int FB_MAX_STATS = EHMM_MAX_K * EHMM_MAX_K + 1;

}  // namespace ehmm

struct ehmm_session {
    std::string key;
    int device = 0;
    int64_t M = 0;      // bins in this block
    int K = 0;          // hidden states
    int64_t m0 = 0;     // first global bin of this block
    int64_t Mtot = 0;   // bins in the global chain
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // side stream (uniform generation)
    // staged upload of the next table's Group column (ehmm_prefetch_groups): copy stream, two staging
    // buffers used in turn, "copied" / "consumed" events per buffer
    cudaStream_t stream_up = nullptr;
    cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_used[2] = {nullptr, nullptr};
    ehmm::DevBuf stage[2];
    const void *staged_src[2] = {nullptr, nullptr};
    size_t staged_bytes[2] = {0, 0};
    bool stage_used_pending[2] = {false, false};
    int stage_next = 0;
    cudaEvent_t ev = nullptr, ev2 = nullptr;

    // ---- data (dt) ----
    int N = 0;
    int B = 0;
    int64_t G = 0;
    ehmm::DevBuf counts_own, offsets_own, controls_own;  // owned copies (host upload)
    const int32_t *d_counts = nullptr;
    const double *d_offsets = nullptr;
    const double *d_controls = nullptr;
    // controls[1] of the GLOBAL long table (bin 1 of sample 1): the only control the reference's emission
    // uses (its ifelse() quirk, R/base-loglikelihood.R:113); a bin block with m0 > 0 gets it from block 0
    double ctrl_first = 0.0;
    bool have_ctrl_first = false;
    ehmm::DevBuf grp_tags, grp_first, grp_list, grp_rank, grp_misc, grp_sort;  // build_groups work space (kept between calls)
    ehmm::DevBuf group;                     // int32 M x N
    ehmm::DevBuf u_chip, u_off, u_ctrl;     // dtUnique columns, index 0..G (0 unused)
    ehmm::DevBuf u_dsg;                     // uint8 (G+1) x B
    std::vector<int64_t> group_first;       // build_groups: first cell (column-major, local) of groups 1..G
    uint8_t design[EHMM_MAX_B * 64] = {0};  // B x N
    bool have_design = false;

    // ---- datasets ----
    ehmm::DevBuf logf, logFP, logBP, logProb1, logProb2, eta, viterbi, vit_bp;
    // ---- lean (EM-internal) E-step ----
    ehmm::DevBuf prob1;          // M x K posteriors, linear (DS_P1LIN)
    ehmm::DevBuf logf_alt;       // second emission buffer: an emission computed while the datasets of the last lean
                                 // E-step are still owed goes here, so that they can be produced from the old one
    bool lean_mode = false;      // ehmm_expstep_mode: E-steps write P1 + statistics only
    bool lazy_full = false;      // logFP / logBP / logProb1 / logProb2 of the last E-step have not been produced yet
    const double *estep_logf = nullptr;          // emission the last E-step ran on
    double estep_cin[2][EHMM_MAX_K + 1] = {{0}};  // ... and the vectors entering / leaving its block
    bool fb_counted = false;     // the last lean E-step (K = 2) counted the weights below rc_hint_p per warp tile
    // ---- callPeaks ----
    ehmm::DevBuf peak_work, peak_calls;  // sort buffers; the last calls (one byte per window)
    bool peaks_exact = false;            // ehmm_call_peaks_exact: always add in emulated long double
    int peaks_path = 0;                  // last fdr call: 0 fp64 prefix sum (certified), 1 emulated (forced), 2 emulated (margin too small)
    bool peaks_valid = false;
    bool vit_forward_done = false;
    ehmm::DevBuf vit_work;       // tile sums, plan, look-back state and backtrace maps of the Viterbi scan
    int vit_mode = 0;            // 0: parallel scan (sequential kernel where the integer model does not hold), 1: sequential kernel
    int64_t vit_info[4] = {0};   // last forward pass: path (0 scan, 1 sequential, 2 scan abandoned), break tiles, half-way tiles, tiles
    unsigned valid = 0;        // DS_* bits: which datasets hold data
    bool stats_stale = false;  // datasets were replaced from the host; statistics must be recomputed

    // ---- forward-backward scratch ----
    ehmm::DevBuf tileops, carries, tilestats, fbscal, rangeops, rangecarry;  // fbscal: [ll, stats..., total op...]
    double h_pi[EHMM_MAX_K] = {0};
    double h_gamma[EHMM_MAX_K * EHMM_MAX_K] = {0};  // column-major as given
    double h_stats[ehmm::FB_MAX_STATS] = {0};       // K*K joint sums, then sum(P1*logf)
    double h_lp1_row0[EHMM_MAX_K] = {0};
    double h_ll = 0;
    bool stats_on_host = false;
    bool reduce_pending = false;
    const double *cur_logf = nullptr;  // logf the pending sharded E-step was reduced from

    // ---- rejection control ----
    uint32_t rng_state[625] = {624};
    bool rng_set = false;
    ehmm::DevBuf rc_flag, rc_scan, rc_buckets, rc_unif, rc_tmp, mt_raw, mt_raw2, mt_windows;
    ehmm::DevBuf rc_unif_spec;   // uniforms generated ahead of the next call on the side stream (mt_speculate)
    int64_t spec_n = 0;          // ... that many words from the current stream position (0: none)
    // sharded form: this block's column slices generated ahead, with a margin either side (mt_speculate_ranges)
    ehmm::DevBuf rc_unif_specr[2];
    int specr_next = 0;          // buffer the next speculation writes
    int specr_cols = 0;          // 0: nothing generated ahead
    uint64_t specr_epoch = 0;    // rng_epoch the ranges are relative to
    int64_t specr_lo[64] = {0}, specr_n[64] = {0}, specr_off[64] = {0};
    uint64_t rng_epoch = 1;      // bumped whenever the stream position changes
    ehmm::DevBuf rc_limbs, rc_coltot;  // exact accumulators (2 x u64 per table entry); column totals + error flag
    bool rc_exact = false;       // the last rejection tables were accumulated as exact limbs ...
    bool rc_final = false;       // ... and have been rounded into rc_buckets
    int mt_shift = 0;            // sub-streams are 2^(17 + mt_shift) words long (mt_choose_stride)
    int64_t mtj_valid = 0;       // sub-stream windows valid for the state in mt_raw
    bool mtj_pending = false;    // ... being computed on stream2 (event ev)
    bool mtj_base = false;       // window 0 is in place
    int64_t mtj_iv[80][2] = {{0}};  // further valid runs of windows [lo, hi] (sharded fits prepare only their slice)
    int mtj_niv = 0;
    bool rng_dev_valid = false;  // mt_raw holds rng_state
    bool rng_pending = false;    // the state after the last draw is on its way to h_pinned + 512 (rng_state[0] is already current)
    int rc_columns = 0;
    int rc_N = 0;
    int64_t rc_colcount[64] = {0};  // sharded form: flagged counts of this block per column
    double rc_count_p = -1.0;
    int rc_count_cols = 0;

    // ---- emission / GLM scratch ----
    ehmm::DevBuf lgtab, glm_part, glm_ticket, misc, gtab, inner_part;
    int inner_part_blocks = 0;  // > 0: inner_part holds innerMaxStepProb partial sums for the current eta / logProb1
    // ---- what the rejection control may reuse ----
    uint64_t lp1_gen = 1;        // bumped whenever logProb1, the groups or the block change
    uint64_t post_gen = 1;       // bumped whenever logProb1 or mixtureProb change
    double rc_hint_p = -1.0;     // cut announced for the next rejection-control call (ehmm_rc_hint), else the last one used
    uint64_t rc_counts_gen = 0;  // rc_flag holds the flagged counts per (column, warp tile) of the posteriors of
    double rc_counts_p = -1.0;   // generation post_gen for this cut (counted by the kernel that produced them)
    int rc_counts_cols = 0;
    ehmm::DevBuf rc_static;      // differential model: exact limbs of the weights >= p of columns 0 and B+1, which
    uint64_t rc_static_gen = 0;  // stay the same through the inner iterations of one E-step (generation lp1_gen,
    double rc_static_p = -1.0;   // cut rc_static_p)
    cudaEvent_t ev_static = nullptr;  // recorded behind a build of rc_static on the side stream
    bool static_pending = false;
    bool glm_ticket_ready = false;
    bool encoded = false;       // data given as dt$Group + dtUnique only (no count / offset matrices)
    bool encoded_ctrl = false;
    bool groups_consistent = false;  // dtUnique[group] reproduces (counts, offsets[, controls]) cell by cell
    int64_t count_max = -1;
    double *h_pinned = nullptr;  // small pinned staging buffer (1024 doubles; [512..] = RNG state)
    // mapped (zero-copy) result slot of the batched GLM kernel: the kernel's last block stores values and
    // gradients here and then an epoch flag; the host polls the flag instead of copying and synchronising
    double *h_glm = nullptr, *d_glm = nullptr;  // 64 doubles; [63] holds the flag (as unsigned)
    unsigned glm_epoch = 0;

    // ---- launch counters and optional per-kernel CUDA-event timing ----
    int64_t launches[ehmm::KID_COUNT] = {0};
    bool prof_on = false;
    struct ProfRec { int kid; cudaEvent_t e0, e1; };
    std::vector<cudaEvent_t> prof_pool;
    size_t prof_used = 0;
    std::vector<ProfRec> prof_recs;
    cudaEvent_t prof_event() {
        if (prof_used == prof_pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            prof_pool.push_back(e);
        }
        return prof_pool[prof_used++];
    }
};

namespace ehmm {

inline int check_session(ehmm_session *s) {
    EHMM_REQUIRE(s != nullptr, EHMM_ERR_ARG, "null session");
    cudaError_t e = cudaSetDevice(s->device);
    if (e != cudaSuccess) {
        set_error("cudaSetDevice(%d): %s", s->device, cudaGetErrorString(e));
        return EHMM_ERR_CUDA;
    }
    return EHMM_OK;
}

// counts every launch; with profiling on, brackets it with CUDA events on the session stream.
// Used as a temporary in front of the launch: `PROF(s, KID_X), kernel<<<...>>>(...);`
struct ProfScope {
    ehmm_session *s;
    cudaEvent_t e1 = nullptr;
    ProfScope(ehmm_session *s_, int kid) : s(s_) {
        s->launches[kid]++;
        if (s->prof_on) {
            cudaEvent_t e0 = s->prof_event();
            e1 = s->prof_event();
            cudaEventRecord(e0, s->stream);
            s->prof_recs.push_back({kid, e0, e1});
        }
    }
    ~ProfScope() {
        if (e1) cudaEventRecord(e1, s->stream);
    }
};
#define PROF(s, kid) ehmm::ProfScope(s, ehmm::kid)

// fb.cu
int fb_expstep(ehmm_session *s, const double *d_logf);
int fb_reduce(ehmm_session *s, const double *d_logf, double *op_host);
int fb_apply(ehmm_session *s, const double *fwd_carry, const double *bwd_carry, double *fwd_out);
int fb_fetch_stats(ehmm_session *s);
int fb_materialize(ehmm_session *s);
// RNG state on the host: a draw only books the new position (rng_note_advance) and queues the copy
// of the 624 words; whoever needs the words waits for it here.
void rng_note_advance(ehmm_session *s, int64_t n);
int rng_sync_host(ehmm_session *s);
inline bool has(const ehmm_session *s, unsigned bits) { return (s->valid & bits) == bits; }
// datasets that exist or can be produced on demand from the last lean E-step
inline bool avail(const ehmm_session *s, unsigned bits) {
    return ((s->valid | (s->lazy_full ? (unsigned)DS_ESTEP : 0u)) & bits) == bits;
}
// posteriors in either form
inline bool has_post(const ehmm_session *s) { return (s->valid & (DS_P1 | DS_P1LIN)) != 0; }
// produce the datasets `bits` names if a lean E-step still owes them
inline int ensure(ehmm_session *s, unsigned bits) {
    if (s->lazy_full && (bits & DS_ESTEP & ~s->valid)) return fb_materialize(s);
    return EHMM_OK;
}
// where kernels read the posteriors P1[j,k] from: the linear copy when there is one, else exp(logProb1)
struct P1Src {
    const double *p;
    int lin;
};
inline P1Src p1_source(const ehmm_session *s) {
    if (s->valid & DS_P1LIN) return P1Src{s->prob1.as<double>(), 1};
    return P1Src{s->logProb1.as<double>(), 0};
}
// the buffer the next emission is written to (see logf_alt)
inline int logf_for_write(ehmm_session *s, size_t bytes) {
    if (s->lazy_full && s->estep_logf == s->logf.p && s->logf.p) std::swap(s->logf, s->logf_alt);
    return s->logf.reserve(bytes);
}
// logProb1 (or what it is summed against: the groups, the block) changed: nothing derived from it may be reused
inline void posteriors_changed(ehmm_session *s) {
    s->lp1_gen++;
    s->post_gen++;
    s->inner_part_blocks = 0;
}
// emission.cu
int emis_consensus(ehmm_session *s, const double *th1, const double *th2, int P, double minZero, const double *zi = nullptr);
int emis_init(ehmm_session *s, const double *psi1, const double *psi2);
int emis_differential_hmm(ehmm_session *s, const double *delta, const double *psi, double minZero, int zinb = 0);
int emis_inner_estep(ehmm_session *s, const double *delta, const double *psi, double minZero, int zinb = 0);
int inner_maxstepprob(ehmm_session *s, double *delta_out);
int inner_mstep_sums(ehmm_session *s, double *sums);
// rc.cu
int rc_consensus(ehmm_session *s, double p, const double *uniforms, int64_t n_uniforms, int64_t *consumed);
int rc_differential(ehmm_session *s, double p, int N, const double *uniforms, int64_t n_uniforms, int64_t *consumed);
int rc_count(ehmm_session *s, int differential, double p, int64_t *counts);
int rc_apply_at(ehmm_session *s, int differential, double p, const int64_t *col_offsets, int64_t total);
int rc_finalize(ehmm_session *s);
int64_t rc_limb_words(const ehmm_session *s);
int rc_fold(ehmm_session *s, int64_t *words);
int rc_static_prefetch(ehmm_session *s);
int rc_table_rows(ehmm_session *s, int column, int64_t *rows);
int rc_table_fetch(ehmm_session *s, int column, double *sums, double *ids);
int rc_reweight(int device, double *x, int64_t n, double p, const double *uniforms, int64_t n_uniforms, int64_t *consumed);
int rc_aggregate(int device, const double *x, int64_t n, const int32_t *f, int64_t G, double *buckets);
// groups.cu
int build_groups(ehmm_session *s, int64_t *G_out);
int groups_fetch(ehmm_session *s, int32_t *group, double *u_chip, double *u_off, double *u_ctrl, uint8_t *u_dsg);
int groups_remap(ehmm_session *s, const int32_t *new_id, int64_t G_new);
// mtjump.cu
int mt_generate(ehmm_session *s, int64_t n, uint32_t *out);
int mt_generate_range(ehmm_session *s, int64_t lo_rel, int64_t n, uint32_t *out);
int mt_generate_ranges(ehmm_session *s, int nr, const int64_t *lo_rel, const int64_t *n, const int64_t *out_off, uint32_t *out);
int mt_speculate_ranges(ehmm_session *s, int nr, const int64_t *lo_rel, const int64_t *n, const int64_t *out_off, uint32_t *out);
int mt_advance(ehmm_session *s, int64_t n);
int mt_advance_side(ehmm_session *s, int64_t n, void *h_state);
int mt_prefetch(ehmm_session *s, int64_t max_draws);
int mt_choose_stride(ehmm_session *s, int64_t draws);
int mt_speculate(ehmm_session *s, int64_t est, uint32_t *out);
int mt_wait_side(ehmm_session *s);
int mt_prefetch_ranges(ehmm_session *s, int nr, const int64_t *lo, const int64_t *n, int64_t total, int margin);
// glm.cu
int glm_nb(ehmm_session *s, int column, const double *par, int P, double *value, double *grad);
int glm_zinb(ehmm_session *s, int column, const double *par, int P, double minZero, double *value, double *grad);
int glm_nb_batch(ehmm_session *s, int ncols, const int *columns, const double *pars, int P, double *values, double *grads);
int glm_mix_nb(ehmm_session *s, const double *par, double minZero, double *value, double *grad);
int glm_mix_zinb(ehmm_session *s, const double *par, double minZero, double *value, double *grad);
// peaks.cu
int call_peaks_fdr(ehmm_session *s, double fdr, uint8_t *calls_out, int64_t *n_called, double *cutoff);
int call_peaks_viterbi(ehmm_session *s, uint8_t *calls_out, int64_t *n_called);
int radix_sort_pairs(ehmm_session *s, uint64_t *k0, uint64_t *k1, uint32_t *v0, uint32_t *v1, unsigned *hist, int64_t n,
                     const uint32_t *vinit, uint64_t **k_sorted, uint32_t **v_sorted);
int peak_runs(ehmm_session *s, const int64_t *breaks, int nbreaks, int64_t cap, int64_t *starts, int64_t *ends, int64_t *nruns);
// viterbi.cu
int viterbi(ehmm_session *s, const double *pi, const double *gamma, double *states_out);
int viterbi_forward(ehmm_session *s, const double *pi, const double *gamma, const double *v_in, double *v_out);
int viterbi_backtrace(ehmm_session *s, int state_last, int *state_prev, double *states_out);

}  // namespace ehmm
