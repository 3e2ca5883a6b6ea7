// glue_harness.cpp -- runs every export of glue/ehmm_glue.cpp (compiled against tests/mock_rcpp/Rcpp.h) on one
// small consensus data set read from a binary file and writes the results to another; tests/test_glue.py
// compares them with the ctypes path of the same library.
//   input : int64 M, N; int32 counts[M*N]; double offsets[M*N]; uint32 seed[625]; double theta[...]
//   output: named arrays, "name n v0 v1 ..." per line (%.17g)
#include <Rcpp.h>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>

using namespace Rcpp;

void expStep(NumericVector, NumericMatrix, Nullable<NumericMatrix>, StringVector);
List maxStepProb(StringVector);
NumericVector computeViterbiSequence(StringVector, NumericVector, NumericMatrix);
double computeQFunction(StringVector, NumericVector, NumericMatrix);
double computeBIC(StringVector, int, int);
NumericVector innerMaxStepProb(StringVector);
List consensusRejectionControlled(StringVector, NumericVector, double);
List differentialRejectionControlled(StringVector, NumericVector, double, int);
NumericMatrix getMarginalProbability(StringVector);
NumericMatrix aggregate(NumericVector, NumericVector);
NumericVector reweight(NumericVector, double);
NumericVector rbinomVectorized(NumericVector);
void ehmm_R_open(StringVector, double, int);
void ehmm_R_close(StringVector);
void ehmm_R_set_data(StringVector, IntegerMatrix, NumericMatrix, Nullable<NumericMatrix>);
double ehmm_R_build_groups(StringVector);
void ehmm_R_loglik_consensus(StringVector, NumericVector, NumericVector, double, std::string);
void ehmm_R_em_hints(StringVector, double, bool);
void ehmm_R_set_design(StringVector, IntegerMatrix);
void ehmm_R_loglik_differential(StringVector, NumericVector, NumericVector, double, std::string);
void ehmm_R_inner_expstep(StringVector, NumericVector, NumericVector, double, std::string);
List ehmm_R_glm_mix(StringVector, NumericVector, double, std::string);
List ehmm_R_glm_nb(StringVector, int, NumericVector);
NumericMatrix ehmm_R_fetch(StringVector, std::string);
LogicalVector ehmm_R_call_peaks(StringVector, double);

static FILE *out;
static void put(const char *name, const double *v, size_t n) {
    std::fprintf(out, "%s %zu", name, n);
    for (size_t i = 0; i < n; i++) std::fprintf(out, " %.17g", v[i]);
    std::fprintf(out, "\n");
}
static void put(const char *name, const NumericVector &v) { put(name, v.begin(), (size_t)v.size()); }
static void put(const char *name, const NumericMatrix &m) { put(name, m.begin(), (size_t)m.nrow() * m.ncol()); }

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    std::ifstream in(argv[1], std::ios::binary);
    int64_t M, N;
    in.read((char *)&M, 8);
    in.read((char *)&N, 8);
    IntegerMatrix counts((int)M, (int)N);
    NumericMatrix offsets((int)M, (int)N);
    in.read((char *)counts.begin(), 4 * M * N);
    in.read((char *)offsets.begin(), 8 * M * N);
    IntegerVector seed(626);
    seed[0] = 10403;   // RNGkind: Mersenne-Twister, Inversion, Rejection
    in.read((char *)(seed.begin() + 1), 4 * 625);
    NumericVector pi(2), t1(2), t2(2);
    NumericMatrix gamma(2, 2);
    in.read((char *)pi.begin(), 16);
    in.read((char *)gamma.begin(), 32);
    in.read((char *)t1.begin(), 16);
    in.read((char *)t2.begin(), 16);
    // differential block: int64 M2, N2, B; counts, offsets, design (B x N2, int32), delta[B], psi[4], pi[3], gamma[9]
    int64_t M2, N2, B;
    in.read((char *)&M2, 8);
    in.read((char *)&N2, 8);
    in.read((char *)&B, 8);
    IntegerMatrix counts2((int)M2, (int)N2), design((int)B, (int)N2);
    NumericMatrix offsets2((int)M2, (int)N2), gamma3(3, 3);
    NumericVector delta((int)B), psi(4), pi3(3);
    in.read((char *)counts2.begin(), 4 * M2 * N2);
    in.read((char *)offsets2.begin(), 8 * M2 * N2);
    in.read((char *)design.begin(), 4 * B * N2);
    in.read((char *)delta.begin(), 8 * B);
    in.read((char *)psi.begin(), 32);
    in.read((char *)pi3.begin(), 24);
    in.read((char *)gamma3.begin(), 72);
    out = std::fopen(argv[2], "w");
    try {
        StringVector h5{"/tmp/glue_harness.h5"};
        Environment g = Environment::global_env();
        const IntegerVector seed0 = clone(seed);   // (vectors are handles: the rejection control advances `seed` in place)
        g[".Random.seed"] = seed;
        ehmm_R_open(h5, (double)M, 2);
        ehmm_R_set_data(h5, counts, offsets, Nullable<NumericMatrix>());
        const double G = ehmm_R_build_groups(h5);
        put("G", &G, 1);
        ehmm_R_em_hints(h5, 0.05, true);
        ehmm_R_loglik_consensus(h5, t1, t2, 2.2250738585072014e-308, "nb");
        expStep(pi, gamma, Nullable<NumericMatrix>(), h5);
        List m = maxStepProb(h5);
        put("pi", m.get<NumericVector>("pi"));
        put("gamma", m.get<NumericMatrix>("gamma"));
        NumericVector f((int)(M * N));   // dt$Group lives in the session: the argument is only length-checked
        List rc = consensusRejectionControlled(h5, f, 0.05);
        for (int k = 0; k < 2; k++) {
            char nm[16];
            std::snprintf(nm, sizeof nm, "rc%d", k);
            put(nm, rc.get<NumericMatrix>(k));
        }
        IntegerVector after = g[".Random.seed"];
        std::vector<double> sd(626);
        for (int i = 0; i < 626; i++) sd[(size_t)i] = (double)(uint32_t)after[i];
        put("seed_after", sd.data(), 626);
        NumericVector par{t2[0], 1.0 / t2[1]};
        List gl = ehmm_R_glm_nb(h5, 1, par);
        const double gv = gl.get<double>("value");
        put("glm_value", &gv, 1);
        put("glm_grad", gl.get<NumericVector>("gradient"));
        const double q = computeQFunction(h5, pi, gamma), bic = computeBIC(h5, 5, (int)N);
        put("q", &q, 1);
        put("bic", &bic, 1);
        put("viterbi", computeViterbiSequence(h5, pi, gamma));
        put("marginal", getMarginalProbability(h5));
        put("logFP", ehmm_R_fetch(h5, "logFP"));
        LogicalVector pk = ehmm_R_call_peaks(h5, 0.05);
        std::vector<double> pkd((size_t)M);
        for (int64_t j = 0; j < M; j++) pkd[(size_t)j] = pk[j];
        put("peaks", pkd.data(), (size_t)M);
        // the exported form of expStep: a host matrix under a path no session exists for yet
        NumericMatrix lf = ehmm_R_fetch(h5, "logLikelihood");
        StringVector h5b{"/tmp/glue_harness_b.h5"};
        expStep(pi, gamma, Nullable<NumericMatrix>(lf), h5b);
        put("logFP_b", ehmm_R_fetch(h5b, "logFP"));
        // the stand-alone helpers
        NumericVector x((int)M), grp((int)M);
        for (int64_t j = 0; j < M; j++) { x[j] = (double)((j * 37) % 101) / 100.0; grp[j] = (double)(1 + (j * 7) % 13); }
        put("aggregate", aggregate(x, grp));
        mock_rng_state() = 42;
        put("reweight", reweight(x, 0.3));
        mock_rng_state() = 42;
        put("rbinom", rbinomVectorized(x));
        ehmm_R_close(h5);
        ehmm_R_close(h5b);
        // the differential model's exports
        {
            StringVector h5d{"/tmp/glue_harness_d.h5"};
            g[".Random.seed"] = clone(seed0);
            ehmm_R_open(h5d, (double)M2, 3);
            ehmm_R_set_data(h5d, counts2, offsets2, Nullable<NumericMatrix>());
            ehmm_R_set_design(h5d, design);
            const double Gd = ehmm_R_build_groups(h5d);
            put("d_G", &Gd, 1);
            ehmm_R_em_hints(h5d, 0.05, true);
            ehmm_R_loglik_differential(h5d, delta, psi, 2.2250738585072014e-308, "nb");
            expStep(pi3, gamma3, Nullable<NumericMatrix>(), h5d);
            List md = maxStepProb(h5d);
            put("d_gamma", md.get<NumericMatrix>("gamma"));
            ehmm_R_inner_expstep(h5d, delta, psi, 2.2250738585072014e-308, "nb");
            put("d_delta", innerMaxStepProb(h5d));
            NumericVector fd((int)(M2 * N2));
            List rcd = differentialRejectionControlled(h5d, fd, 0.05, (int)N2);
            for (int c = 0; c < (int)B + 2; c++) {
                char nm[16];
                std::snprintf(nm, sizeof nm, "d_rc%d", c);
                put(nm, rcd.get<NumericMatrix>(c));
            }
            NumericVector par{psi[0], psi[2], psi[1], psi[3]};   // optimDifferential's order (b, db, l, dl)
            List gm = ehmm_R_glm_mix(h5d, par, 2.2250738585072014e-308, "nb");
            const double gmv = gm.get<double>("value");
            put("d_glm_value", &gmv, 1);
            put("d_glm_grad", gm.get<NumericVector>("gradient"));
            put("d_mixtureProb", ehmm_R_fetch(h5d, "mixtureProb"));
            ehmm_R_close(h5d);
        }
        // errors become R errors (stop): a path nobody opened
        try {
            maxStepProb(StringVector{"/tmp/never_opened.h5"});
            std::fprintf(out, "error_raised 1 0\n");
        } catch (const std::exception &e) {
            std::fprintf(out, "error_raised 1 1\n");
        }
    } catch (const std::exception &e) {
        std::fprintf(out, "FAILED 0\n");
        std::fclose(out);
        std::cerr << "glue harness: " << e.what() << std::endl;
        return 1;
    }
    std::fclose(out);
    return 0;
}
