/*
 * ehmm.h -- C ABI of libehmm_b200.so: the B200 (sm_100a) implementation of
 * epigraHMM's EM hot path.
 *
 * Each entry point replaces one function of the reference (citations are
 * relative to the reference tree, plbaldoni/epigraHMM v1.9.2).  The
 * reference passes state between its native calls through an HDF5 file whose
 * path is the only handle (src/expStep.cpp:50-51,116-122); here the same path
 * string keys a device-resident session and the "datasets" (logLikelihood,
 * logFP, logBP, logProb1, logProb2, mixtureProb, viterbi) are HBM arrays that
 * ehmm_fetch copies out on demand.
 *
 * Conventions
 *   - every function returns EHMM_OK (0) or an EHMM_ERR_* code; the message is
 *     available from ehmm_last_error() (thread local).  Nothing throws or
 *     longjmps across this boundary.
 *   - all matrices are column-major fp64 exactly as R / Armadillo hold them
 *     (gamma[i + k*K] = gamma(i,k)); counts are int32 (the reference holds
 *     as.numeric(counts), R/base-core.R:182), group ids are 1-based int32
 *     (data.table .GRP, R/base-core.R:97,187-193).
 *   - pointers are HOST pointers unless the name ends in _device.
 *   - there is no CPU fallback: every call fails with EHMM_ERR_CUDA when no
 *     sm_100 device is usable.
 */
#ifndef EHMM_H
#define EHMM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EHMM_OK 0
#define EHMM_ERR_ARG 1    /* bad argument (null pointer, bad shape, K unsupported) */
#define EHMM_ERR_CUDA 2   /* CUDA runtime failure (message holds cudaGetErrorString) */
#define EHMM_ERR_STATE 3  /* call order violated (e.g. maxStepProb before expStep) */
#define EHMM_ERR_NOMEM 4  /* device allocation failed */
#define EHMM_ERR_RNG 5    /* uniform buffer exhausted */

#define EHMM_MAX_K 3      /* the reference only ever uses K = 2 (consensus) or 3 (differential) */
#define EHMM_MAX_B 62     /* mixture components (2^G - 2 patterns) */

typedef struct ehmm_session ehmm_session;

const char *ehmm_last_error(void);
int ehmm_version(void);
int ehmm_device_count(int *n);

/* ---- session = the reference's HDF5 path handle ------------------------------- */
/* key: the hdf5 path string the R glue receives (src/expStep.cpp:50-51). */
int ehmm_session_open(const char *key, int64_t M, int K, int device, ehmm_session **out);
int ehmm_session_find(const char *key, ehmm_session **out);
int ehmm_session_close(ehmm_session *s);
/* CUDA stream the session launches on (cudaStream_t as void*), for event timing. */
int ehmm_session_stream(ehmm_session *s, void **stream);
int ehmm_session_sync(ehmm_session *s);

/* This session holds bins [m0, m0+M) of one global chain of Mtot bins (bin-block
 * sharding; the chain is single and global, R/base-core.R:181-183). */
int ehmm_session_set_block(ehmm_session *s, int64_t m0, int64_t Mtot);

/* ---- data: the long table `dt` (R/base-core.R:89-104,181-198) ------------------- */
/* counts M x N int32, offsets M x N fp64, controls (already log1p'd, R/base-core.R:186) or NULL. */
int ehmm_set_data(ehmm_session *s, int N, const int32_t *counts, const double *offsets, const double *controls);
int ehmm_set_data_device(ehmm_session *s, int N, const int32_t *counts, const double *offsets, const double *controls);
/* controls[1] -- the (log1p'd) control of bin 1 of sample 1 of the WHOLE long table.  It is the only
 * control value the reference's emission uses: ifelse(useControl, theta[2]*controls, 0) has a
 * length-1 test and returns one element (R/base-loglikelihood.R:113,119,126-129,137), which R then
 * recycles over every row.  ehmm_set_data* record it; a bin block with m0 > 0 must be given block 0's
 * value (ehmm_controls_first there, ehmm_set_controls_first here). */
int ehmm_controls_first(ehmm_session *s, double *value);
int ehmm_set_controls_first(ehmm_session *s, double value);
/* dt$Group (M x N, 1-based) and dtUnique rows for groups 1..G: ChIP, offsets[, controls]
 * [, Dsg.Mix 0/1 as G x B column-major] (R/base-core.R:99-104,190-198). */
int ehmm_set_groups(ehmm_session *s, const int32_t *group, int64_t G, const double *u_chip, const double *u_offsets,
                    const double *u_controls, const uint8_t *u_dsg);
/* Build dt$Group and dtUnique on the device from the uploaded table (and design), numbered by first
 * appearance in the column-major long table like data.table's .GRP (R/base-core.R:97-104,184-198).
 * Replaces the host-side grouping + ehmm_set_groups.  ehmm_groups_fetch copies them out (any pointer
 * may be NULL; u_dsg is G x B column-major). */
int ehmm_build_groups(ehmm_session *s, int64_t *G);
int ehmm_groups_fetch(ehmm_session *s, int32_t *group, double *u_chip, double *u_offsets, double *u_controls,
                      uint8_t *u_dsg);
/* Bin blocks of one sharded table need ONE dictionary (the chain and dt are global): after
 * ehmm_build_groups on every block, ehmm_groups_first_cell gives the column-major cell index
 * (sample * M + bin, local to the block) at which each of the block's G groups first appears, the
 * caller merges the blocks' dtUnique rows into the global numbering (first appearance in the global
 * column-major table, as .GRP numbers them) and ehmm_groups_remap rewrites the ids in place: old id g
 * becomes new_id[g - 1]; dtUnique is replaced by the G_new global rows. */
int ehmm_groups_first_cell(ehmm_session *s, int64_t *first);
int ehmm_groups_remap(ehmm_session *s, const int32_t *new_id, int64_t G_new, const double *u_chip, const double *u_offsets,
                      const double *u_controls, const uint8_t *u_dsg);
/* Dictionary-encoded form of the same table: only dt$Group and dtUnique are uploaded (4 B per cell
 * instead of 12-20 B); counts/offsets are implied by dtUnique[group].  Replaces ehmm_set_data +
 * ehmm_set_groups.  For the differential model call ehmm_set_design_n(s, N, B, design) first. */
int ehmm_set_data_encoded(ehmm_session *s, int N, const int32_t *group, int64_t G, const double *u_chip,
                          const double *u_offsets, const double *u_controls, const uint8_t *u_dsg);
/* same with 16-bit group ids (G <= 65535): halves the upload again; widened to int32 on the device */
int ehmm_set_data_encoded16(ehmm_session *s, int N, const uint16_t *group, int64_t G, const double *u_chip,
                            const double *u_offsets, const double *u_controls, const uint8_t *u_dsg);
/* Start copying the Group column of the NEXT table (M*N ids of 2 or 4 bytes, pinned host memory)
 * to a staging buffer on a copy stream and return at once; the next ehmm_set_data_encoded[16] call
 * given the same pointer uses the staged copy.  Lets the upload of table k+1 overlap the EM
 * iteration on table k (two staging buffers, used in turn).  The host buffer must stay unchanged
 * until that call, and every prefetch must be followed by it: a staged copy is matched by address
 * and consumed once. */
int ehmm_prefetch_groups(ehmm_session *s, int N, const void *group, int bytes_per_id);
int ehmm_set_design_n(ehmm_session *s, int N, int B, const uint8_t *design);
/* design[b*N + i] = Dsg.Mix_b of sample i (generateMixtureIntercepts, R/base-utils.R:242-264). */
int ehmm_set_design(ehmm_session *s, int B, const uint8_t *design);

/* ---- emission log-likelihoods -> dataset logLikelihood (device) ------------------ */
/* loglikConsensusHMM NB branch, R/base-loglikelihood.R:104-122. P = 2 or 3 parameters per state. */
int ehmm_loglik_consensus(ehmm_session *s, const double *theta1, const double *theta2, int P, double minZero);
/* dist = 'zinb' (R/base-loglikelihood.R:123-141): the background state is zero-inflated,
 * theta1 = (lambda0[, lambda_ctrl], beta0[, beta_ctrl], size) with Q = 3 or 5 entries (the offsets
 * enter the zero-inflation predictor too, l.125); theta2 = (beta0[, beta_ctrl], size) as above. */
int ehmm_loglik_consensus_zinb(ehmm_session *s, const double *theta1, const double *theta2, int Q, double minZero);
/* emInitialization emission, R/base-EM.R:24-30 (N = 1, dnbinom(log=TRUE)). */
int ehmm_loglik_init(ehmm_session *s, const double *psi1, const double *psi2);
/* loglikDifferentialHMM NB branch, R/base-loglikelihood.R:45-71,96. psi = (b, l, db, dl). */
int ehmm_loglik_differential_hmm(ehmm_session *s, const double *delta, const double *psi, double minZero);
/* dist = 'zinb' (R/base-loglikelihood.R:23-36,72-96): psi = (lambda, b_bg, l_bg, db, dl); the
 * D = 0 density is zero-inflated with zip = exp(lambda) / (1 + exp(lambda)). */
int ehmm_loglik_differential_hmm_zinb(ehmm_session *s, const double *delta, const double *psi, double minZero);
/* innerExpStep, R/base-EM.R:302-314 -> dataset mixtureProb. */
int ehmm_inner_expstep(ehmm_session *s, const double *delta, const double *psi, double minZero);
int ehmm_inner_expstep_zinb(ehmm_session *s, const double *delta, const double *psi, double minZero);

/* ---- E-step / M-step (the C++ sources under src/) ------------------------------------ */
/* expStep, src/expStep.cpp:45-123. logf: host M x K matrix, or NULL to use the
 * device-resident logLikelihood written by an ehmm_loglik_* call. */
int ehmm_expstep(ehmm_session *s, const double *pi, const double *gamma, const double *logf);
/* same, logf already in HBM (M x K column-major). */
int ehmm_expstep_device(ehmm_session *s, const double *pi, const double *gamma, const double *logf_device);
/* log-likelihood lse(logFP[M-1,]) of the last expStep (src/expStep.cpp:91-92). */
int ehmm_loglik_value(ehmm_session *s, double *ll);
/* maxStepProb, src/maxStepProb.cpp:37-81. */
int ehmm_maxstepprob(ehmm_session *s, double *pi_out, double *gamma_out);
/* computeQFunction, src/computeQFunction.cpp:12-33. */
int ehmm_qfunction(ehmm_session *s, const double *pi, const double *gamma, double *q);
/* computeBIC, src/computeBIC.cpp:12-29. */
int ehmm_bic(ehmm_session *s, int numPar, int numSamples, double *bic);
/* innerMaxStepProb, src/innerMaxStepProb.cpp:15-38. */
int ehmm_inner_maxstepprob(ehmm_session *s, double *delta_out);
/* additive form for bin blocks: sums[b] = sum_j P1[j,1] eta[j,b] (b < B), sums[B] = sum_j P1[j,1]. */
int ehmm_inner_mstep_sums(ehmm_session *s, double *sums);
/* getMarginalProbability, src/getMarginalProbability.cpp:12-21 (out: M x K). */
int ehmm_marginal_probability(ehmm_session *s, double *out);
/* ---- after the fit: callPeaks (R/callPeaks.R:61-68) ------------------------------------------------ */
/* fdrControl(prob = exp(logProb1[, 2]), fdr), R/base-utils.R:11-25: calls[j] = 1 where the running mean of the
 * ascending-sorted 1 - prob is below fdr (ties in window order, as data.table's order() leaves them); R's cumsum
 * adds in long double and stores doubles -- reproduced exactly (see csrc/peaks.cu).  calls (M bytes, host) may
 * be NULL: the calls stay on the device for ehmm_peak_runs.  cutoff: the largest 1 - prob called. */
int ehmm_call_peaks(ehmm_session *s, double fdr, uint8_t *calls, int64_t *n_called, double *cutoff);
/* method = 'viterbi' (R/callPeaks.R:62): calls[j] = (viterbi[j] == 1). */
int ehmm_call_peaks_viterbi(ehmm_session *s, uint8_t *calls, int64_t *n_called);
/* force = 1: every ehmm_call_peaks adds in emulated long double (the path taken on its own only when a running
 * mean lies within rounding distance of the level).  path: 0 fp64 prefix sum with certified margins,
 * 1 emulated (forced), 2 emulated (margin too small). */
int ehmm_call_peaks_exact(ehmm_session *s, int force);
int ehmm_call_peaks_path(ehmm_session *s, int *path);
/* GenomicRanges::reduce over the called bins (R/callPeaks.R:67-68): maximal runs of consecutive called windows that
 * do not cross a chromosome start (chrom_starts: ascending 0-based window indices where a chromosome begins; may be
 * empty).  starts / ends (cap entries each, inclusive window indices, ascending) may be NULL to ask for the count. */
int ehmm_peak_runs(ehmm_session *s, const int64_t *chrom_starts, int n_chrom_starts, int64_t cap, int64_t *starts, int64_t *ends,
                   int64_t *n_runs);
/* All datasets of the fit in one call, each as the bytes of the reference's HDF5 dataset (an M x K column-major
 * matrix is the (K, M) dataset Armadillo writes, src/expStep.cpp:116-122).  NULL pointers are skipped. */
int ehmm_flush(ehmm_session *s, double *logLikelihood, double *logFP, double *logBP, double *logProb1, double *logProb2,
               double *mixtureProb, double *viterbi);

/* How expStep delivers its results.  lean = 0 (default): the reference's contract -- logFP, logBP, logProb1
 * and logProb2 are written by the E-step (src/expStep.cpp:116-122).  lean = 1, for E-steps inside the EM
 * loop: only the posteriors P1 and the sufficient statistics are produced (nothing in R/base-EM.R:113-158,
 * 191-225 reads more: maxStepProb, Q and BIC use the statistics, the rejection control and innerMaxStepProb
 * use exp(logProb1)); the four datasets of the LAST E-step are produced on demand -- by ehmm_fetch,
 * ehmm_device_ptr, ehmm_store, ehmm_viterbi*, ehmm_marginal_probability or ehmm_materialize -- by the same
 * kernel on the same inputs, so they hold the same bits as with lean = 0. */
int ehmm_expstep_mode(ehmm_session *s, int lean);
int ehmm_materialize(ehmm_session *s);
/* computeViterbiSequence, src/computeViterbiSequence.cpp:12-54 (states_out: M doubles or NULL). */
int ehmm_viterbi(ehmm_session *s, const double *pi, const double *gamma, double *states_out);
/* How the (max,+) recursion is run.  mode 0 (default): parallel scan -- inside one binade of LOGV the
 * reference's fp64 recursion is exact integer arithmetic, hence associative; tiles that cross a binade or
 * hold a half-way rounding case run the reference's operation order on one thread, and the whole chain
 * does if the integer model is ever left.  mode 1: the sequential kernel for every bin.  The states are
 * bit-identical either way.  info4 (last forward pass): path taken (0 scan, 1 sequential, 2 scan abandoned
 * for the sequential kernel), tiles planned sequential, tiles that found a half-way case, tiles. */
int ehmm_viterbi_mode(ehmm_session *s, int mode);
int ehmm_viterbi_info(ehmm_session *s, int64_t *info4);

/* Bin-block form of computeViterbiSequence: the (max,+) recursion is serial over the chain, so the
 * blocks run one after the other handing on K doubles (forward) and one state (backtrace).
 * v_in: LOGV of the bin in front of this block (NULL for the block that starts the chain);
 * state_last: state of this block's last bin decided by the next block (-1 at the chain's end);
 * state_prev: state of the bin in front of this block. */
int ehmm_viterbi_forward(ehmm_session *s, const double *pi, const double *gamma, const double *v_in, double *v_out);
int ehmm_viterbi_backtrace(ehmm_session *s, int state_last, int *state_prev, double *states_out);

/* Sharded E-step (bin blocks across GPUs): phase 1 reduces this block to one
 * scaled K x K operator (op: K*K + 1 doubles = mantissa matrix row-major + log scale);
 * phase 2 applies the carries the host computed from all blocks' operators
 * (fwd/bwd: K + 1 doubles = vector + log scale). */
int ehmm_expstep_reduce(ehmm_session *s, const double *pi, const double *gamma, const double *logf, double *op);
int ehmm_expstep_apply(ehmm_session *s, const double *fwd_carry, const double *bwd_carry, double *fwd_out);
/* phase 2 with the carries composed by the library: ops_all = the P operators gathered in rank
 * order (P * (K*K + 1) doubles); forward carry = pi * prod(ops[:rank]), backward carry =
 * prod(ops[rank+1:]) * 1.  Does not synchronise; ehmm_estep_stats / ehmm_loglik_value do. */
int ehmm_expstep_apply_ops(ehmm_session *s, const double *ops_all, int P, int rank);
/* raw sufficient statistics of the last expStep: K*K joint sums (j >= 1) then sum(P1 * logf). */
int ehmm_estep_stats(ehmm_session *s, double *stats);
/* logProb1[0,] of this block (K doubles): maxStepProb's pi comes from the first block's row. */
int ehmm_estep_row0(ehmm_session *s, double *row0);
int ehmm_estep_set_stats(ehmm_session *s, const double *stats, const double *logprob1_row0, double loglik);

/* ---- rejection control (src/reweight.cpp, rbinomVectorized.cpp, aggregate.cpp) ---- */
/* R's Mersenne-Twister state: state[0] = mti, state[1..624] = mt (.Random.seed[2:626]). */
int ehmm_rng_set_state(ehmm_session *s, const uint32_t *state625);
int ehmm_rng_get_state(ehmm_session *s, uint32_t *state625);
/* Optional: announce the cut p of the NEXT rejection-control call.  The kernels that produce the posterior
 * weights (innerExpStep; the E-step) then count the weights below p on the way and the rejection control
 * skips its own counting pass.  Without a hint the previous call's cut is assumed (the reference's
 * rejectionCut = max(0.9^iter, probCut), R/base-EM.R:119-121, is constant once it has reached probCut); a
 * call with another cut simply counts for itself, so a wrong hint costs time, never correctness. */
int ehmm_rc_hint(ehmm_session *s, double p);
/* consensusRejectionControlled, src/consensusRejectionControlled.cpp:14-31.
 * uniforms: host buffer of unif_rand() draws to consume in order, or NULL to draw
 * from the session's Mersenne-Twister on the device.  consumed: draws used. */
int ehmm_rc_consensus(ehmm_session *s, double p, const double *uniforms, int64_t n_uniforms, int64_t *consumed);
/* differentialRejectionControlled, src/differentialRejectionControlled.cpp:14-48. */
int ehmm_rc_differential(ehmm_session *s, double p, int N, const double *uniforms, int64_t n_uniforms, int64_t *consumed);
/* Sharded form (bin blocks across GPUs; the draw order is global: column-major, then ascending
 * global bin).  ehmm_rc_count returns this block's flagged counts per column; after the caller has
 * combined the counts of all blocks, ehmm_rc_apply_at thins and aggregates this block with its
 * column c draws starting at stream offset col_offsets[c] (relative to the current RNG position)
 * and advances the stream by `total` draws.  differential = 0: K columns; 1: B + 2 columns. */
int ehmm_rc_count(ehmm_session *s, int differential, double p, int64_t *counts);
int ehmm_rc_apply_at(ehmm_session *s, int differential, double p, const int64_t *col_offsets, int64_t total);
/* Device pointer to the accumulators of the rejection tables, for adding the tables of several bin
 * blocks in place (e.g. an NCCL all-reduce ordered on the session stream).  is_integer = 1: n
 * unsigned 64-bit words, the exact fixed-point limbs (two per table entry; the call first folds the
 * replicas of the frequent groups into them ON THE SESSION STREAM -- add on that stream or
 * synchronise it first) -- add them as integers and the rounded tables are bit-identical for every
 * sharding; is_integer = 0 (cuts below 2^-9): n doubles.  ehmm_rc_finalize rounds the limbs into the fp64 tables (the readers below and the GLM
 * entry points do it themselves when it is still pending). */
int ehmm_rc_buckets_device(ehmm_session *s, void **ptr, int64_t *n, int *is_integer);
int ehmm_rc_finalize(ehmm_session *s);
/* the field<mat> result: table `column` has `rows` (sum, id) pairs with non-zero sum, ascending id. */
int ehmm_rc_table_rows(ehmm_session *s, int column, int64_t *rows);
int ehmm_rc_table_fetch(ehmm_session *s, int column, double *sums, double *ids);
/* dense bucket sums (G+1 doubles, index = group id) of one column. */
int ehmm_rc_buckets_fetch(ehmm_session *s, int column, double *buckets);
/* stand-alone exported helpers: reweight (src/reweight.cpp:16-22), aggregate (src/aggregate.cpp:20-45). */
int ehmm_reweight(int device, double *x, int64_t n, double p, const double *uniforms, int64_t n_uniforms, int64_t *consumed);
int ehmm_aggregate(int device, const double *x, int64_t n, const int32_t *f, int64_t G, double *buckets);

/* ---- NB-GLM objective + gradient over the rejection tables (R/base-optim.R) -------- */
/* glmNB / derivNB, R/base-optim.R:109-130, over table `column` joined with dtUnique.
 * par = (beta0[, beta_ctrl], phi); value and grad (P + 1) are the NEGATIVE log-lik as in R. */
int ehmm_glm_nb(ehmm_session *s, int column, const double *par, int P, double *value, double *grad);
/* glmZINB / derivZINB (R/base-optim.R:136-181) over rejection table `column`:
 * par = (beta[P], phi, lambda[P]), grad has 2P + 1 entries in the same order. */
int ehmm_glm_zinb(ehmm_session *s, int column, const double *par, int P, double minZero, double *value, double *grad);
/* n evaluations in one launch (the per-state optimisations of R/base-EM.R:131-141 are independent,
 * so their L-BFGS-B instances can advance in lock-step): columns[q], pars[q*(P+1) ..] -> values[q],
 * grads[q*(P+1) ..]. */
int ehmm_glm_nb_batch(ehmm_session *s, int n, const int *columns, const double *pars, int P, double *values, double *grads);
/* glmMixNB / derivMixNB, R/base-optim.R:187-255, over all B+2 tables. par = (b, db, l, dl). */
int ehmm_glm_mix_nb(ehmm_session *s, const double *par, double minZero, double *value, double *grad);
/* glmMixZINB / derivMixZINB (R/base-optim.R:261-362): par = (lambda, b_bg, l_bg, b_enr, l_enr),
 * grad has 5 entries. */
int ehmm_glm_mix_zinb(ehmm_session *s, const double *par, double minZero, double *value, double *grad);

/* ---- launch accounting and per-kernel timing --------------------------------------------- */
/* Number of distinct kernels the library tracks and their names (index = kernel id). */
int ehmm_profile_kernels(void);
const char *ehmm_profile_name(int kernel_id);
/* enable != 0: reset the counters and bracket every launch with CUDA events on the session stream
 * (for the bench's roofline figures); enable == 0: stop timing, keep counting launches. */
int ehmm_profile(ehmm_session *s, int enable);
/* launches[id] = launches since the last reset; ms[id] = summed event time (0 unless enabled). */
int ehmm_profile_read(ehmm_session *s, int64_t *launches, double *ms);

/* ---- dataset access ------------------------------------------------------------------ */
/* name in {logLikelihood, logFP, logBP, logProb1, logProb2, mixtureProb, viterbi}. */
int ehmm_dataset_dims(ehmm_session *s, const char *name, int64_t *rows, int64_t *cols);
int ehmm_fetch(ehmm_session *s, const char *name, double *dst, int64_t n);
int ehmm_device_ptr(ehmm_session *s, const char *name, void **ptr);
/* load a dataset from the host (lets each exported function be called on its own inputs). */
int ehmm_store(ehmm_session *s, const char *name, const double *src, int64_t rows, int64_t cols);

/* ---- exchange of small host vectors between the ranks of one node -------------------------- */
/* The bin-block sharded fit exchanges K x K block operators, sufficient statistics
 * and flagged counts: hundreds of bytes, produced and consumed on the host.  These go through a
 * POSIX shared-memory segment (name: "/..." unique per job, agreed through the launcher) instead of
 * a device collective; NCCL carries the rejection tables (ehmm_rc_buckets_device).  Collective:
 * every rank of the job calls open / allgather in the same order. */
typedef struct ehmm_comm ehmm_comm;
int ehmm_comm_open(const char *name, int rank, int world, int64_t slot_bytes, double timeout_s, ehmm_comm **out);
/* in: nbytes (<= slot_bytes) from this rank; out: world * nbytes in rank order */
int ehmm_comm_allgather(ehmm_comm *c, const void *in, int64_t nbytes, void *out);
int ehmm_comm_close(ehmm_comm *c);

#ifdef __cplusplus
}
#endif
#endif /* EHMM_H */
