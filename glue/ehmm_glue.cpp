// ehmm_glue.cpp -- what src/*.cpp of plbaldoni/epigraHMM become when libehmm_b200.so is dropped in.
//
// The thirteen [[Rcpp::export]] functions keep the names, argument lists and return shapes of the
// reference (src/RcppExports.cpp:12-171, R stubs R/RcppExports.R:4-116), so R/RcppExports.R, the
// .Call table (src/RcppExports.cpp:174-194) and every R caller stay as they are; their bodies call
// the C ABI of include/ehmm.h instead of Armadillo + HDF5.  The second block holds the new internal
// exports that the R-side numeric functions forward to.
//
// Compiled by R CMD INSTALL with Rcpp only (no RcppArmadillo, no Rhdf5lib).  R is not installed in
// the image this repository is built in: tests/test_glue.py compiles this file against the minimal
// stand-in tests/mock_rcpp/Rcpp.h and runs every export through tests/glue_harness.cpp on the GPU.
#include <Rcpp.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

extern "C" {
#include "ehmm.h"
}

using namespace Rcpp;

namespace {

#define EHMM(call)                                         \
    do {                                                   \
        if ((call) != EHMM_OK) stop(ehmm_last_error());    \
    } while (0)

// the reference reads element 0 of the hdf5 argument only (src/expStep.cpp:50-51)
std::string key_of(const StringVector &hdf5) { return as<std::string>(hdf5[0]); }

ehmm_session *sess(const StringVector &hdf5) {
    ehmm_session *s = nullptr;
    EHMM(ehmm_session_find(key_of(hdf5).c_str(), &s));
    return s;
}

int device_index() {  // controlEM()'s field set is frozen (R/epigraHMM.R:54): the device comes from the environment
    const char *e = std::getenv("EHMM_DEVICE");
    return e ? std::atoi(e) : 0;
}

void dims(ehmm_session *s, const char *name, int64_t &rows, int64_t &cols) { EHMM(ehmm_dataset_dims(s, name, &rows, &cols)); }

// R's Mersenne-Twister state goes to the device generator and comes back advanced by the draws made:
// .Random.seed = (kind, mti, mt[624]); the library takes (mti, mt[624])
struct SeedHandOver {
    ehmm_session *s;
    Environment g;
    IntegerVector seed;
    explicit SeedHandOver(ehmm_session *s_) : s(s_), g(Environment::global_env()), seed(g[".Random.seed"]) {
        if (seed.size() != 626 || seed[0] % 100 != 3) stop("epigraHMM on the GPU needs RNGkind('Mersenne-Twister')");
        EHMM(ehmm_rng_set_state(s, reinterpret_cast<const uint32_t *>(seed.begin() + 1)));
    }
    void take_back() {
        EHMM(ehmm_rng_get_state(s, reinterpret_cast<uint32_t *>(seed.begin() + 1)));
        g[".Random.seed"] = seed;
    }
};

// arma::field<arma::mat>(n, 1) of (sum, group id) tables, as Rcpp wraps it: a list with dim = c(n, 1)
List tables(ehmm_session *s, int n) {
    List out(n);
    for (int c = 0; c < n; c++) {
        int64_t rows = 0;
        EHMM(ehmm_rc_table_rows(s, c, &rows));
        NumericMatrix tab((int)rows, 2);
        EHMM(ehmm_rc_table_fetch(s, c, tab.begin(), tab.begin() + rows));
        out[c] = tab;
    }
    out.attr("dim") = Dimension(n, 1);
    return out;
}

std::vector<int32_t> group_ids(const NumericVector &f) {  // dt$Group arrives as doubles (arma::vec f)
    std::vector<int32_t> g((size_t)f.size());
    for (R_xlen_t i = 0; i < f.size(); i++) g[(size_t)i] = (int32_t)f[i];
    return g;
}

}  // namespace

// ---- the reference's thirteen exports ---------------------------------------------------------------

// src/expStep.cpp:45-123.  logf = NULL: the emission is already in HBM (loglik*HMM below wrote it).
// [[Rcpp::export]]
void expStep(NumericVector pi, NumericMatrix gamma, Nullable<NumericMatrix> logf, StringVector hdf5) {
    const std::string key = key_of(hdf5);
    ehmm_session *s = nullptr;
    const double *lf = nullptr;
    NumericMatrix m;
    if (logf.isNotNull()) {  // exported form (man/expStep.Rd): a host matrix, the "file" may not exist yet
        m = NumericMatrix(logf);
        lf = m.begin();
        if (ehmm_session_find(key.c_str(), &s) != EHMM_OK)
            EHMM(ehmm_session_open(key.c_str(), m.nrow(), (int)pi.size(), device_index(), &s));
    } else {
        s = sess(hdf5);
    }
    EHMM(ehmm_expstep(s, pi.begin(), gamma.begin(), lf));  // R matrices are column-major: no copy
}

// src/maxStepProb.cpp:37-81
// [[Rcpp::export]]
List maxStepProb(StringVector hdf5) {
    ehmm_session *s = sess(hdf5);
    int64_t M, K;
    dims(s, "logProb1", M, K);
    NumericVector pi((int)K);
    NumericMatrix gamma((int)K, (int)K);
    EHMM(ehmm_maxstepprob(s, pi.begin(), gamma.begin()));
    return List::create(Named("pi") = pi, Named("gamma") = gamma);
}

// src/computeViterbiSequence.cpp:12-54 (also leaves the 'viterbi' dataset behind, l.51)
// [[Rcpp::export]]
NumericVector computeViterbiSequence(StringVector hdf5, NumericVector pi, NumericMatrix gamma) {
    ehmm_session *s = sess(hdf5);
    int64_t M, K;
    dims(s, "logFP", M, K);
    NumericVector states((int)M);
    EHMM(ehmm_viterbi(s, pi.begin(), gamma.begin(), states.begin()));
    return states;
}

// src/computeQFunction.cpp:12-33
// [[Rcpp::export]]
double computeQFunction(StringVector hdf5, NumericVector pi, NumericMatrix gamma) {
    double q = 0.0;
    EHMM(ehmm_qfunction(sess(hdf5), pi.begin(), gamma.begin(), &q));
    return q;
}

// src/computeBIC.cpp:12-29
// [[Rcpp::export]]
double computeBIC(StringVector hdf5, int numPar, int numSamples) {
    double bic = 0.0;
    EHMM(ehmm_bic(sess(hdf5), numPar, numSamples, &bic));
    return bic;
}

// src/innerMaxStepProb.cpp:15-38
// [[Rcpp::export]]
NumericVector innerMaxStepProb(StringVector hdf5) {
    ehmm_session *s = sess(hdf5);
    int64_t M, B;
    dims(s, "mixtureProb", M, B);
    NumericVector delta((int)B);
    EHMM(ehmm_inner_maxstepprob(s, delta.begin()));
    return delta;
}

// src/consensusRejectionControlled.cpp:14-31.  dt$Group was handed over once per fit (ehmm_R_set_groups);
// f is accepted for the signature's sake and checked for length only.
// [[Rcpp::export]]
List consensusRejectionControlled(StringVector hdf5, NumericVector f, double p) {
    ehmm_session *s = sess(hdf5);
    RNGScope scope;  // src/RcppExports.cpp:19
    int64_t M, K;
    dims(s, "logProb1", M, K);
    if (f.size() < M) stop("consensusRejectionControlled: f is shorter than the number of bins");
    SeedHandOver rng(s);
    int64_t used = 0;
    EHMM(ehmm_rc_consensus(s, p, nullptr, 0, &used));
    rng.take_back();
    return tables(s, (int)K);
}

// src/differentialRejectionControlled.cpp:14-48
// [[Rcpp::export]]
List differentialRejectionControlled(StringVector hdf5, NumericVector f, double p, int N) {
    ehmm_session *s = sess(hdf5);
    RNGScope scope;
    int64_t M, B;
    dims(s, "mixtureProb", M, B);
    if (f.size() != M * N) stop("differentialRejectionControlled: f must hold M * N group ids");
    SeedHandOver rng(s);
    int64_t used = 0;
    EHMM(ehmm_rc_differential(s, p, N, nullptr, 0, &used));
    rng.take_back();
    return tables(s, (int)B + 2);
}

// src/getMarginalProbability.cpp:12-21
// [[Rcpp::export]]
NumericMatrix getMarginalProbability(StringVector hdf5) {
    ehmm_session *s = sess(hdf5);
    int64_t M, K;
    dims(s, "logProb1", M, K);
    NumericMatrix out((int)M, (int)K);
    EHMM(ehmm_marginal_probability(s, out.begin()));
    return out;
}

// src/aggregate.cpp:20-45: rows (sum, id) of the non-zero sums, ascending id
// [[Rcpp::export]]
NumericMatrix aggregate(NumericVector x, NumericVector f) {
    const std::vector<int32_t> g = group_ids(f);
    int32_t G = 0;
    for (R_xlen_t i = 0; i < x.size(); i++) G = g[(size_t)i] > G ? g[(size_t)i] : G;  // the reference walks x.size() entries (l.22)
    std::vector<double> b((size_t)G + 1);
    EHMM(ehmm_aggregate(device_index(), x.begin(), (int64_t)x.size(), g.data(), G, b.data()));
    int rows = 0;
    for (double v : b) rows += (v != 0.0);
    NumericMatrix out(rows, 2);
    int r = 0;
    for (int32_t id = 0; id <= G; id++)
        if (b[(size_t)id] != 0.0) {
            out(r, 0) = b[(size_t)id];
            out(r, 1) = (double)id;
            r++;
        }
    return out;
}

// src/reweight.cpp:16-22: x[i] < p  ->  rbinom(1, x[i] / p) * p.  R's rbinom draws exactly one uniform unless
// its probability is 0 or 1; the draws are made here, in R's stream, and handed to the device in order.
// [[Rcpp::export]]
NumericVector reweight(NumericVector x, double p) {
    RNGScope scope;
    NumericVector out = clone(x);
    int64_t need = 0;
    for (R_xlen_t i = 0; i < x.size(); i++) need += (x[i] < p && x[i] / p != 0.0);
    std::vector<double> u((size_t)need);
    for (int64_t i = 0; i < need; i++) u[(size_t)i] = unif_rand();
    int64_t used = 0;
    EHMM(ehmm_reweight(device_index(), out.begin(), (int64_t)out.size(), p, u.data(), need, &used));
    if (used != need) stop("reweight: uniform count mismatch");
    return out;
}

// src/rbinomVectorized.cpp:15-19: one R::rbinom(1, prob[i]) per element, left on the host (it is a loop
// over R's own generator and nothing else)
// [[Rcpp::export]]
NumericVector rbinomVectorized(NumericVector prob) {
    RNGScope scope;
    NumericVector out((int)prob.size());
    for (R_xlen_t i = 0; i < prob.size(); i++) out[i] = R::rbinom(1.0, prob[i]);
    return out;
}

// ---- new internal exports: what the R-side numeric functions forward to ---------------------------

// start of a fit (R/base-core.R:61 deletes the old file; :181-198 builds dt): one session per hdf5 path
// [[Rcpp::export]]
void ehmm_R_open(StringVector hdf5, double M, int K) {
    ehmm_session *s = nullptr;
    EHMM(ehmm_session_open(key_of(hdf5).c_str(), (int64_t)M, K, device_index(), &s));
}

// [[Rcpp::export]]
void ehmm_R_close(StringVector hdf5) {
    ehmm_session *s = nullptr;
    if (ehmm_session_find(key_of(hdf5).c_str(), &s) == EHMM_OK) EHMM(ehmm_session_close(s));
}

// dt's columns as matrices: ChIP (integer counts), offsets, controls (log1p'd, R/base-core.R:186) or NULL
// [[Rcpp::export]]
void ehmm_R_set_data(StringVector hdf5, IntegerMatrix counts, NumericMatrix offsets, Nullable<NumericMatrix> controls) {
    ehmm_session *s = sess(hdf5);
    const double *ctrl = nullptr;
    NumericMatrix c;
    if (controls.isNotNull()) {
        c = NumericMatrix(controls);
        ctrl = c.begin();
    }
    EHMM(ehmm_set_data(s, counts.ncol(), reinterpret_cast<const int32_t *>(counts.begin()), offsets.begin(), ctrl));
}

// Dsg.Mix columns of the differential model (generateMixtureIntercepts, R/base-utils.R:242-264): B x N, 0/1
// [[Rcpp::export]]
void ehmm_R_set_design(StringVector hdf5, IntegerMatrix design) {
    const int B = design.nrow(), N = design.ncol();
    std::vector<uint8_t> d((size_t)B * N);
    for (int b = 0; b < B; b++)
        for (int i = 0; i < N; i++) d[(size_t)b * N + i] = (uint8_t)design(b, i);
    EHMM(ehmm_set_design(sess(hdf5), B, d.data()));
}

// dt[, Group := .GRP, by = ...] and dtUnique (R/base-core.R:97-104,187-198) built on the device; returns G
// [[Rcpp::export]]
double ehmm_R_build_groups(StringVector hdf5) {
    int64_t G = 0;
    EHMM(ehmm_build_groups(sess(hdf5), &G));
    return (double)G;
}

// loglikConsensusHMM (R/base-loglikelihood.R:104-143): theta1 / theta2 as the R function receives them
// [[Rcpp::export]]
void ehmm_R_loglik_consensus(StringVector hdf5, NumericVector theta1, NumericVector theta2, double minZero, std::string dist) {
    ehmm_session *s = sess(hdf5);
    if (dist == "zinb") EHMM(ehmm_loglik_consensus_zinb(s, theta1.begin(), theta2.begin(), (int)theta1.size(), minZero));
    else EHMM(ehmm_loglik_consensus(s, theta1.begin(), theta2.begin(), (int)theta1.size(), minZero));
}

// emInitialization's emission (R/base-EM.R:24-30)
// [[Rcpp::export]]
void ehmm_R_loglik_init(StringVector hdf5, NumericVector psi1, NumericVector psi2) {
    EHMM(ehmm_loglik_init(sess(hdf5), psi1.begin(), psi2.begin()));
}

// loglikDifferentialHMM (R/base-loglikelihood.R:45-98): psi = unlist(theta$psi)
// [[Rcpp::export]]
void ehmm_R_loglik_differential(StringVector hdf5, NumericVector delta, NumericVector psi, double minZero, std::string dist) {
    ehmm_session *s = sess(hdf5);
    if (dist == "zinb") EHMM(ehmm_loglik_differential_hmm_zinb(s, delta.begin(), psi.begin(), minZero));
    else EHMM(ehmm_loglik_differential_hmm(s, delta.begin(), psi.begin(), minZero));
}

// innerExpStep (R/base-EM.R:302-314): the mixtureProb dataset stays in HBM
// [[Rcpp::export]]
void ehmm_R_inner_expstep(StringVector hdf5, NumericVector delta, NumericVector psi, double minZero, std::string dist) {
    ehmm_session *s = sess(hdf5);
    if (dist == "zinb") EHMM(ehmm_inner_expstep_zinb(s, delta.begin(), psi.begin(), minZero));
    else EHMM(ehmm_inner_expstep(s, delta.begin(), psi.begin(), minZero));
}

// the cut of the coming rejection-control call (R/base-EM.R:119-121 knows it before the E-step) and the
// EM-internal form of the E-step: set at the top of emConsensus / outerEMDifferential
// [[Rcpp::export]]
void ehmm_R_em_hints(StringVector hdf5, double rejectionCut, bool insideEM) {
    ehmm_session *s = sess(hdf5);
    EHMM(ehmm_rc_hint(s, rejectionCut));
    EHMM(ehmm_expstep_mode(s, insideEM ? 1 : 0));
}

// glmNB + derivNB (R/base-optim.R:109-130) over rejection table `column` (0-based): list(value, gradient);
// optim(fn, gr) calls fn and gr with the same par, so the R wrappers cache the pair
// [[Rcpp::export]]
List ehmm_R_glm_nb(StringVector hdf5, int column, NumericVector par) {
    double v = 0.0;
    NumericVector g((int)par.size());
    EHMM(ehmm_glm_nb(sess(hdf5), column, par.begin(), (int)par.size() - 1, &v, g.begin()));
    return List::create(Named("value") = v, Named("gradient") = g);
}

// glmZINB + derivZINB (R/base-optim.R:136-181): par = (beta[P], phi, lambda[P])
// [[Rcpp::export]]
List ehmm_R_glm_zinb(StringVector hdf5, int column, NumericVector par, double minZero) {
    double v = 0.0;
    NumericVector g((int)par.size());
    EHMM(ehmm_glm_zinb(sess(hdf5), column, par.begin(), ((int)par.size() - 1) / 2, minZero, &v, g.begin()));
    return List::create(Named("value") = v, Named("gradient") = g);
}

// glmMixNB + derivMixNB (R/base-optim.R:187-255) / glmMixZINB + derivMixZINB (:261-362)
// [[Rcpp::export]]
List ehmm_R_glm_mix(StringVector hdf5, NumericVector par, double minZero, std::string dist) {
    double v = 0.0;
    NumericVector g((int)par.size());
    ehmm_session *s = sess(hdf5);
    if (dist == "zinb") EHMM(ehmm_glm_mix_zinb(s, par.begin(), minZero, &v, g.begin()));
    else EHMM(ehmm_glm_mix_nb(s, par.begin(), minZero, &v, g.begin()));
    return List::create(Named("value") = v, Named("gradient") = g);
}

// end of a fit (R/base-core.R:212-226): a dataset as the M x cols matrix rhdf5 would read back; the R side
// h5write()s it so that callPeaks / callPatterns / info keep reading the file
// [[Rcpp::export]]
NumericMatrix ehmm_R_fetch(StringVector hdf5, std::string name) {
    ehmm_session *s = sess(hdf5);
    int64_t r, c;
    dims(s, name.c_str(), r, c);
    NumericMatrix out((int)r, (int)c);
    EHMM(ehmm_fetch(s, name.c_str(), out.begin(), r * c));
    return out;
}

// callPeaks(method = fdr) on the device (R/callPeaks.R:61-68 with fdrControl, R/base-utils.R:11-25):
// logical vector of the bins called
// [[Rcpp::export]]
LogicalVector ehmm_R_call_peaks(StringVector hdf5, double fdr) {
    ehmm_session *s = sess(hdf5);
    int64_t M, K;
    dims(s, "logProb1", M, K);
    std::vector<uint8_t> calls((size_t)M);
    int64_t n = 0;
    double cutoff = 0.0;
    EHMM(ehmm_call_peaks(s, fdr, calls.data(), &n, &cutoff));
    LogicalVector out((int)M);
    for (int64_t j = 0; j < M; j++) out[j] = calls[(size_t)j] != 0;
    return out;
}
