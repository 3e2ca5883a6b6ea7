// glm.cu -- weighted NB-GLM objective and gradient over the rejection tables (replaces glmNB/derivNB
// and glmMixNB/derivMixNB, R/base-optim.R:109-130,187-255, which R's optim(L-BFGS-B) evaluates over
// the aggregated group table).  The tables stay on the device as dense bucket vectors indexed by
// group id; a zero bucket contributes exactly 0, so summing over all G groups equals summing over
// the reference's non-zero rows.  Block partial sums are reduced in a fixed order (deterministic).
#include "common.cuh"
#include <atomic>
#include <chrono>

namespace ehmm {

namespace {

constexpr int GLM_NT = 256;
constexpr int GLM_MAXOUT = 8;

// digamma for x > 0: recurrence up to x >= 12, then the asymptotic series (same form as the oracle)
__device__ __forceinline__ double digamma_pos(double x) {
    double r = 0.0;
    while (x < 12.0) { r -= 1.0 / x; x += 1.0; }
    const double xi = 1.0 / x, x2 = xi * xi;
    const double s = x2 * (1.0 / 12 - x2 * (1.0 / 120 - x2 * (1.0 / 252 - x2 * (1.0 / 240 -
                     x2 * (1.0 / 132 - x2 * (691.0 / 32760 - x2 * (1.0 / 12)))))));
    return r + (log(x) - 0.5 * xi - s);
}

// log NB(y; size, mu = exp(eta)) with lsize = log(size), lgs = lgamma(size)
__device__ __forceinline__ double nb_logpmf_direct(double y, double eta, double size, double lsize, double lgs) {
    const double t = eta - lsize;
    const double w = log1p(exp(-fabs(t)));
    const double sp = (t > 0.0) ? t + w : w;
    const double spm = (t > 0.0) ? w : w - t;
    const double c = (lgamma(y + size) - lgs) - lgamma(y + 1.0);
    return (c - size * sp) - y * spm;
}

template <int NOUT> __device__ __forceinline__ void block_store(double *acc, double *__restrict__ part) {
    __shared__ double sh[(GLM_NT / 32) * NOUT];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NOUT; i++) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc[i] += __shfl_down_sync(0xffffffffu, acc[i], d);
        if (lane == 0) sh[w * NOUT + i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < NOUT) {
        double t = sh[threadIdx.x];
        for (int i = 1; i < GLM_NT / 32; i++) t += sh[i * NOUT + threadIdx.x];
        part[(int64_t)blockIdx.x * NOUT + threadIdx.x] = t;
    }
}

struct NBPar {
    double b0, b1, phi, size, lsize, lgs, dg_size;
    int has_ctrl;
};

// out per evaluation: [0] = sum w*l, [1] = sum w*r, [2] = sum w*r*ctrl, [3] = sum w*dphi
// Batched form: blockIdx.y selects the (column, parameter) pair, so the optimiser can evaluate all
// states in one launch; the last block to finish (ticket counter) reduces the per-block partials in
// a fixed order, so no second launch is needed and the result is deterministic.
struct NBBatch {
    NBPar par[EHMM_MAX_K];
    int column[EHMM_MAX_K];
};

__global__ void __launch_bounds__(GLM_NT)
    glm_nb_batch_kernel(int64_t G, const double *__restrict__ buckets, const double *__restrict__ y,
                        const double *__restrict__ off, const double *__restrict__ ctrl, NBBatch B,
                        double *__restrict__ part, unsigned *__restrict__ ticket, double *__restrict__ out,
                        double *__restrict__ hout, unsigned epoch) {
    const int q = blockIdx.y;
    const NBPar P = B.par[q];
    const double *wt = buckets + (size_t)B.column[q] * (G + 1);
    double acc[4] = {0, 0, 0, 0};
    for (int64_t g = 1 + (int64_t)blockIdx.x * GLM_NT + threadIdx.x; g <= G; g += (int64_t)gridDim.x * GLM_NT) {
        const double w = wt[g];
        if (w == 0.0) continue;
        const double yy = y[g], cv = P.has_ctrl ? ctrl[g] : 0.0;
        const double eta = (P.has_ctrl ? (P.b0 + cv * P.b1) : P.b0) + off[g];
        const double mu = exp(eta);
        double l = nb_logpmf_direct(yy, eta, P.size, P.lsize, P.lgs);
        if (isinf(l)) l = -690.77552789821368;
        const double den = 1.0 + P.phi * mu;
        const double r = (yy - mu) / den;
        const double dphi = (log(den) + P.phi * (yy - mu) / den - digamma_pos(yy + P.size) + P.dg_size) / (P.phi * P.phi);
        acc[0] += w * l;
        acc[1] += w * r;
        acc[2] += w * r * cv;
        acc[3] += w * dphi;
    }
    double *mypart = part + (size_t)q * gridDim.x * 4;
    {
        __shared__ double sh[(GLM_NT / 32) * 4];
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < 4; i++) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) acc[i] += __shfl_down_sync(0xffffffffu, acc[i], d);
            if (lane == 0) sh[w * 4 + i] = acc[i];
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            double t = sh[threadIdx.x];
            for (int i = 1; i < GLM_NT / 32; i++) t += sh[i * 4 + threadIdx.x];
            mypart[(size_t)blockIdx.x * 4 + threadIdx.x] = t;
        }
    }
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket + q, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last) {
        __shared__ double red[GLM_NT];
        for (int c = 0; c < 4; c++) {
            double t = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += GLM_NT) t += __ldcg(mypart + (size_t)i * 4 + c);
            red[threadIdx.x] = t;
            __syncthreads();
            for (int d = GLM_NT / 2; d > 0; d >>= 1) {
                if ((int)threadIdx.x < d) red[threadIdx.x] += red[threadIdx.x + d];
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                out[q * 4 + c] = red[0];
                hout[q * 4 + c] = red[0];  // zero-copy: straight into the host's mapped result slot
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            ticket[q] = 0u;
            // the evaluation whose last block finishes last publishes the epoch (after all results)
            __threadfence_system();
            if (atomicAdd(ticket + EHMM_MAX_K, 1u) == gridDim.y - 1) {
                ticket[EHMM_MAX_K] = 0u;
                __threadfence_system();
                *reinterpret_cast<volatile unsigned *>(hout + 63) = epoch;
            }
        }
    }
}

// glmZINB / derivZINB (R/base-optim.R:136-181) over the background state's rejection table.
struct ZinbPar {
    double b0, b1, l0, l1, phi, size, lsize, lgs, dg_size, minZero, thr;
    int has_ctrl;
};

// out: [0] = sum w*l, [1] = sum d/dbeta0, [2] = sum d/dbeta_ctrl, [3] = sum d/dphi, [4] = sum d/dlambda0, [5] = d/dlambda_ctrl
__global__ void __launch_bounds__(GLM_NT)
    glm_zinb_kernel(int64_t G, const double *__restrict__ wt, const double *__restrict__ y, const double *__restrict__ off,
                    const double *__restrict__ ctrl, ZinbPar P, double *__restrict__ part) {
    double acc[6] = {0, 0, 0, 0, 0, 0};
    const double iphi2 = 1.0 / (P.phi * P.phi);
    for (int64_t g = 1 + (int64_t)blockIdx.x * GLM_NT + threadIdx.x; g <= G; g += (int64_t)gridDim.x * GLM_NT) {
        const double w = wt[g];
        if (w == 0.0) continue;
        const double yy = y[g], cv = P.has_ctrl ? ctrl[g] : 0.0, o = off[g];
        const double eta = (P.has_ctrl ? (P.b0 + cv * P.b1) : P.b0) + o;
        const double mu = exp(eta);
        const double z = 1.0 / (1.0 + exp(-((P.has_ctrl ? (P.l0 + cv * P.l1) : P.l0) + o)));
        const double l1mz = log(1.0 - z), aux = 1.0 + P.phi * mu, laux = log(aux);
        double d = nb_logpmf_direct(yy, eta, P.size, P.lsize, P.lgs);  // log dnbinom(y); y = 0 gives the reference's d0
        if (!(d > P.thr)) d = log(exp(d) + P.minZero);
        double lval, db, dp, dl;
        if (yy == 0.0) {
            const double P0 = z + exp(l1mz + d), c = (w / P0) * (1.0 - z);
            lval = log(P0);
            db = c * (-mu * pow(aux, -(P.size + 1.0)));
            dp = c * pow(aux, -P.size) * (laux * iphi2 - mu / (P.phi * aux));
            dl = (w / P0) * (1.0 - exp(d)) * z * (1.0 - z);
        } else {
            const double P1 = exp(l1mz + d), ed = exp(d), c = (w / P1) * (1.0 - z) * ed;
            lval = l1mz + d;
            db = c * (yy - mu) / aux;
            dp = c * (laux * iphi2 - mu * (yy + P.size) / aux - digamma_pos(yy + P.size) * iphi2 + P.dg_size * iphi2 + yy / P.phi);
            dl = (w / P1) * (-ed) * z * (1.0 - z);
        }
        acc[0] += w * lval;
        acc[1] += db;
        acc[2] += db * cv;
        acc[3] += dp;
        acc[4] += dl;
        acc[5] += dl * cv;
    }
    block_store<6>(acc, part);
}

struct MixPar {
    double b, db, l, dl, minZero, thr;
    double size[2], lsize[2], lgs[2], dg[2];  // D = 0 / 1
    int B;
};

// out: [0] = value, [1..4] = gradient wrt (b, db, l, dl).
// One thread per dtUnique row: a row's count and offset are shared by all B + 2 tables, and a table entry enters
// with the D = 0 or the D = 1 density only, so the two densities (and their digammas) are evaluated once per row and
// weighted by the row's summed D = 0 / D = 1 weights -- (B + 2) / 2 times less transcendental work than one
// evaluation per table entry.  The last block to finish (ticket) adds the block partials in a fixed order and stores
// the five sums in the host's mapped result slot, then the epoch flag the host polls.
__global__ void __launch_bounds__(GLM_NT)
    glm_mix_kernel(int64_t G, int ncol, const double *__restrict__ buckets, const double *__restrict__ y,
                   const double *__restrict__ off, const uint8_t *__restrict__ dsg, MixPar P, double *__restrict__ part,
                   unsigned *__restrict__ ticket, double *__restrict__ hout, unsigned epoch) {
    double acc[5] = {0, 0, 0, 0, 0};
    for (int64_t g = 1 + (int64_t)blockIdx.x * GLM_NT + threadIdx.x; g <= G; g += (int64_t)gridDim.x * GLM_NT) {
        double W[2] = {buckets[g], buckets[(int64_t)(ncol - 1) * (G + 1) + g]};  // background: D = 0; enrichment: D = 1
        for (int c = 1; c < ncol - 1; c++) {
            const double w = buckets[(int64_t)c * (G + 1) + g];
            if (w != 0.0) W[dsg[g + (int64_t)(c - 1) * (G + 1)] ? 1 : 0] += w;
        }
        if (W[0] == 0.0 && W[1] == 0.0) continue;
        const double yy = y[g], o = off[g];
#pragma unroll
        for (int D = 0; D < 2; D++) {
            const double w = W[D];
            if (w == 0.0) continue;
            // background: b + off; mixture: (b + db*D) + off; enrichment: (b + db) + off  -- same values
            const double eta = (D ? (P.b + P.db) : P.b) + o;
            const double mu = exp(eta), ph = P.size[D];
            const double lpmf = nb_logpmf_direct(yy, eta, ph, P.lsize[D], P.lgs[D]);
            const double lv = (lpmf > P.thr) ? lpmf : log(exp(lpmf) + P.minZero);
            acc[0] += -w * lv;
            const double a = -w * (yy / mu - ((yy + ph) / (mu + ph))) * mu;
            const double cc = -w * (digamma_pos(yy + ph) - P.dg[D] + 1.0 + P.lsize[D] - (yy + ph) / (mu + ph) - log(mu + ph)) * ph;
            acc[1] += a;
            acc[3] += cc;
            if (D) { acc[2] += a; acc[4] += cc; }
        }
    }
    block_store<5>(acc, part);
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (is_last) {
        __shared__ double red[GLM_NT];
        for (int c = 0; c < 5; c++) {
            double t = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += GLM_NT) t += __ldcg(part + (size_t)i * 5 + c);
            red[threadIdx.x] = t;
            __syncthreads();
            for (int d = GLM_NT / 2; d > 0; d >>= 1) {
                if ((int)threadIdx.x < d) red[threadIdx.x] += red[threadIdx.x + d];
                __syncthreads();
            }
            if (threadIdx.x == 0) hout[c] = red[0];  // zero-copy: straight into the host's mapped result slot
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            *ticket = 0u;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned *>(hout + 63) = epoch;
        }
    }
}

// glmMixZINB / derivMixZINB (R/base-optim.R:261-362): rows of column 0 (background) and mixture rows
// with Dsg.Mix = 0 are zero-inflated NB with (zip, mu_z, size_z); rows of the last column and
// mixture rows with Dsg.Mix = 1 are NB with (mu_n, size_n).
struct MixZPar {
    double lam, zip, bz, bn, minZero, thr;
    double size[2], lsize[2], lgs[2], dg[2];  // 0: zero-inflated background, 1: enrichment
};

// out: [0] = value, [1..5] = gradient wrt (lambda, b_z, l_z, b_n, l_n)
__global__ void __launch_bounds__(GLM_NT)
    glm_mix_zinb_kernel(int64_t G, int ncol, const double *__restrict__ buckets, const double *__restrict__ y,
                        const double *__restrict__ off, const uint8_t *__restrict__ dsg, MixZPar P, double *__restrict__ part) {
    double acc[6] = {0, 0, 0, 0, 0, 0};
    const int64_t total = (int64_t)ncol * (G + 1);
    for (int64_t idx = (int64_t)blockIdx.x * GLM_NT + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * GLM_NT) {
        const double w = buckets[idx];
        if (w == 0.0) continue;
        const int c = (int)(idx / (G + 1));
        const int64_t g = idx - (int64_t)c * (G + 1);
        int D;
        if (c == 0) D = 0;
        else if (c == ncol - 1) D = 1;
        else D = dsg[g + (int64_t)(c - 1) * (G + 1)];
        const double yy = y[g], o = off[g];
        const double ph = P.size[D], eta = (D ? P.bn : P.bz) + o, mu = exp(eta);
        const double lpmf = nb_logpmf_direct(yy, eta, ph, P.lsize[D], P.lgs[D]);
        if (D) {  // log(dnbinom + minZero) and the NB score
            const double lv = (lpmf > P.thr) ? lpmf : log(exp(lpmf) + P.minZero);
            acc[0] += -w * lv;
            acc[4] += -w * (yy / mu - ((yy + ph) / (mu + ph))) * mu;
            acc[5] += -w * (digamma_pos(yy + ph) - P.dg[1] + 1.0 + P.lsize[1] - (yy + ph) / (mu + ph) - log(mu + ph)) * ph;
        } else {
            const double zip = P.zip;
            if (yy == 0.0) {
                const double d0 = exp(lpmf);                          // dnbinom(0) = (ph/(mu+ph))^ph
                const double dz = zip + (1.0 - zip) * (d0 + P.minZero);  // dzinb(0, log = FALSE)
                const double q = ph / (mu + ph);
                acc[0] += -w * log(dz);
                acc[1] += -w * (((1.0 - d0) / dz) * zip * (1.0 - zip));
                acc[2] += -w * ((-(1.0 - zip) * pow(q, ph + 1.0) / dz) * mu);
                acc[3] += -w * ((1.0 - zip) * pow(q, ph) * (log(q) + mu / (mu + ph)) * (1.0 / dz) * ph);
            } else {
                acc[0] += -w * log((1.0 - zip) * (exp(lpmf) + P.minZero));
                acc[1] += -w * (-zip);
                acc[2] += -w * ((yy / mu - (yy + ph) / (mu + ph)) * mu);
                acc[3] += -w * ((digamma_pos(yy + ph) - P.dg[0] + 1.0 + P.lsize[0] - (yy + ph) / (mu + ph) - log(mu + ph)) * ph);
            }
        }
    }
    block_store<6>(acc, part);
}

__global__ void __launch_bounds__(256) glm_final_kernel(const double *__restrict__ part, int nblk, int ncol,
                                                        double *__restrict__ out) {
    __shared__ double sh[256];
    for (int c = 0; c < ncol; c++) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nblk; i += 256) s += part[(int64_t)i * ncol + c];
        sh[threadIdx.x] = s;
        __syncthreads();
        for (int d = 128; d > 0; d >>= 1) {
            if ((int)threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[c] = sh[0];
        __syncthreads();
    }
}

double host_digamma(double x) {
    double r = 0.0;
    while (x < 12.0) { r -= 1.0 / x; x += 1.0; }
    const double xi = 1.0 / x, x2 = xi * xi;
    const double s = x2 * (1.0 / 12 - x2 * (1.0 / 120 - x2 * (1.0 / 252 - x2 * (1.0 / 240 -
                     x2 * (1.0 / 132 - x2 * (691.0 / 32760 - x2 * (1.0 / 12)))))));
    return r + (log(x) - 0.5 * xi - s);
}

int blocks_for(int64_t n) {
    int64_t b = (n + GLM_NT - 1) / GLM_NT;
    if (b < 1) b = 1;
    if (b > 148 * 4) b = 148 * 4;
    return (int)b;
}

int finish(ehmm_session *s, int nblk, int nout, double *host_out) {
    double *part = s->glm_part.as<double>();
    double *out = part + (size_t)148 * 4 * GLM_MAXOUT;
    PROF(s, KID_GLM_FINAL), glm_final_kernel<<<1, 256, 0, s->stream>>>(part, nblk, nout, out);
    EHMM_CK(cudaGetLastError());
    EHMM_CK(cudaMemcpyAsync(s->h_pinned + 128, out, sizeof(double) * nout, cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    for (int i = 0; i < nout; i++) host_out[i] = s->h_pinned[128 + i];
    return EHMM_OK;
}

}  // namespace

int glm_nb_batch(ehmm_session *s, int ncols, const int *columns, const double *pars, int P_, double *values,
                 double *grads);

// wait for a kernel's epoch in the mapped result slot (a few microseconds after it ends) instead of a copy + stream
// synchronisation; a kernel that never reports falls back to the synchronising path for the error
static int poll_epoch(ehmm_session *s, unsigned epoch, const char *who) {
    volatile unsigned *flag = reinterpret_cast<volatile unsigned *>(s->h_glm + 63);
    const auto t0 = std::chrono::steady_clock::now();
    unsigned spins = 0;
    while (*flag != epoch) {
        if ((++spins & 0x3ff) == 0) {
            if (cudaStreamQuery(s->stream) != cudaErrorNotReady) break;  // finished (or failed) without the flag
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(30)) break;
        }
    }
    if (*flag != epoch) {
        EHMM_CK(cudaStreamSynchronize(s->stream));
        EHMM_REQUIRE(*flag == epoch, EHMM_ERR_CUDA, "%s: the kernel did not report its result", who);
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return EHMM_OK;
}

static int ensure_ticket(ehmm_session *s) {
    if (!s->glm_ticket_ready) {
        EHMM_TRY(s->glm_ticket.reserve(64));
        EHMM_CK(cudaMemsetAsync(s->glm_ticket.p, 0, 64, s->stream));
        s->glm_ticket_ready = true;
    }
    return EHMM_OK;
}

int glm_nb(ehmm_session *s, int column, const double *par, int P_, double *value, double *grad) {
    return glm_nb_batch(s, 1, &column, par, P_, value, grad);
}

int glm_nb_batch(ehmm_session *s, int ncols, const int *columns, const double *pars, int P_, double *values,
                 double *grads) {
    EHMM_REQUIRE(P_ == 1 || P_ == 2, EHMM_ERR_ARG, "glmNB: P must be 1 (Intercept) or 2 (Intercept, controls)");
    EHMM_REQUIRE(P_ == 1 || s->u_ctrl.p, EHMM_ERR_STATE, "glmNB: dtUnique has no controls column");
    EHMM_REQUIRE(ncols >= 1 && ncols <= EHMM_MAX_K, EHMM_ERR_ARG, "glmNB batch: 1..%d evaluations", EHMM_MAX_K);
    NBBatch B;
    for (int q = 0; q < ncols; q++) {
        const double *par = pars + (size_t)q * (P_ + 1);
        EHMM_REQUIRE(columns[q] >= 0 && columns[q] < s->rc_columns, EHMM_ERR_STATE, "glmNB: rejection table %d not available", columns[q]);
        EHMM_REQUIRE(par[P_] > 0 && isfinite(par[P_]), EHMM_ERR_ARG, "glmNB: dispersion %g must be positive", par[P_]);
        NBPar &P = B.par[q];
        P.b0 = par[0];
        P.b1 = P_ == 2 ? par[1] : 0.0;
        P.phi = par[P_];
        P.size = 1.0 / P.phi;
        P.lsize = log(P.size);
        P.lgs = lgamma(P.size);
        P.dg_size = host_digamma(P.size);
        P.has_ctrl = (P_ == 2);
        B.column[q] = columns[q];
    }
    const int nblk = blocks_for(s->G);
    EHMM_TRY(s->glm_part.reserve(sizeof(double) * ((size_t)EHMM_MAX_K * 148 * 4 * 4 + EHMM_MAX_K * 4) + 64 +
                                 sizeof(double) * ((size_t)148 * 4 * GLM_MAXOUT + GLM_MAXOUT)));
    EHMM_TRY(ensure_ticket(s));
    double *part = s->glm_part.as<double>();
    double *out = part + (size_t)EHMM_MAX_K * 148 * 4 * 4;
    dim3 grid((unsigned)nblk, (unsigned)ncols);
    const unsigned epoch = ++s->glm_epoch ? s->glm_epoch : ++s->glm_epoch;  // never 0
    PROF(s, KID_GLM), glm_nb_batch_kernel<<<grid, GLM_NT, 0, s->stream>>>(
        s->G, s->rc_buckets.as<double>(), s->u_chip.as<double>(), s->u_off.as<double>(), s->u_ctrl.as<double>(), B, part,
        s->glm_ticket.as<unsigned>(), out, s->d_glm, epoch);
    EHMM_CK(cudaGetLastError());
    EHMM_TRY(poll_epoch(s, epoch, "glmNB batch"));
    for (int q = 0; q < ncols; q++) {
        const double *o = s->h_glm + 4 * q;
        values[q] = -o[0];
        if (grads) {
            double *g = grads + (size_t)q * (P_ + 1);
            g[0] = -o[1];
            if (P_ == 2) g[1] = -o[2];
            g[P_] = -o[3];
        }
    }
    return EHMM_OK;
}

// par = (beta[P], phi, lambda[P]), P = 1 (Intercept) or 2 (Intercept, controls)
int glm_zinb(ehmm_session *s, int column, const double *par, int P_, double minZero, double *value, double *grad) {
    EHMM_REQUIRE(P_ == 1 || P_ == 2, EHMM_ERR_ARG, "glmZINB: P must be 1 (Intercept) or 2 (Intercept, controls)");
    EHMM_REQUIRE(P_ == 1 || s->u_ctrl.p, EHMM_ERR_STATE, "glmZINB: dtUnique has no controls column");
    EHMM_REQUIRE(column >= 0 && column < s->rc_columns, EHMM_ERR_STATE, "glmZINB: rejection table %d not available", column);
    EHMM_REQUIRE(par[P_] > 0 && isfinite(par[P_]), EHMM_ERR_ARG, "glmZINB: dispersion %g must be positive", par[P_]);
    ZinbPar P;
    P.b0 = par[0];
    P.b1 = P_ == 2 ? par[1] : 0.0;
    P.phi = par[P_];
    P.l0 = par[P_ + 1];
    P.l1 = P_ == 2 ? par[P_ + 2] : 0.0;
    P.size = 1.0 / P.phi;
    P.lsize = log(P.size);
    P.lgs = lgamma(P.size);
    P.dg_size = host_digamma(P.size);
    P.minZero = minZero;
    P.thr = log(minZero) + 40.0;
    P.has_ctrl = (P_ == 2);
    EHMM_TRY(s->glm_part.reserve(sizeof(double) * ((size_t)EHMM_MAX_K * 148 * 4 * 4 + EHMM_MAX_K * 4) + 64 +
                                 sizeof(double) * ((size_t)148 * 4 * GLM_MAXOUT + GLM_MAXOUT)));
    const int nblk = blocks_for(s->G);
    PROF(s, KID_GLM), glm_zinb_kernel<<<nblk, GLM_NT, 0, s->stream>>>(
        s->G, s->rc_buckets.as<double>() + (size_t)column * (s->G + 1), s->u_chip.as<double>(), s->u_off.as<double>(),
        s->u_ctrl.as<double>(), P, s->glm_part.as<double>());
    EHMM_CK(cudaGetLastError());
    double o[6];
    EHMM_TRY(finish(s, nblk, 6, o));
    *value = -o[0];
    if (grad) {
        grad[0] = -o[1];
        if (P_ == 2) grad[1] = -o[2];
        grad[P_] = -o[3];
        grad[P_ + 1] = -o[4];
        if (P_ == 2) grad[P_ + 2] = -o[5];
    }
    return EHMM_OK;
}

int glm_mix_nb(ehmm_session *s, const double *par, double minZero, double *value, double *grad) {
    EHMM_REQUIRE(s->u_dsg.p, EHMM_ERR_STATE, "glmMixNB: dtUnique has no Dsg.Mix columns (ehmm_set_groups)");
    MixPar P;
    P.b = par[0];
    P.db = par[1];
    P.l = par[2];
    P.dl = par[3];
    P.minZero = minZero;
    P.thr = log(minZero) + 40.0;
    P.B = s->B;
    for (int D = 0; D < 2; D++) {
        const double ls = D ? (par[2] + par[3]) : par[2];
        P.size[D] = exp(ls);
        P.lsize[D] = log(P.size[D]);
        P.lgs[D] = lgamma(P.size[D]);
        P.dg[D] = host_digamma(P.size[D]);
    }
    EHMM_TRY(s->glm_part.reserve(sizeof(double) * ((size_t)EHMM_MAX_K * 148 * 4 * 4 + EHMM_MAX_K * 4) + 64 +
                                 sizeof(double) * ((size_t)148 * 4 * GLM_MAXOUT + GLM_MAXOUT)));
    EHMM_TRY(ensure_ticket(s));
    const int ncol = s->B + 2;
    const int nblk = blocks_for(s->G);
    const unsigned epoch = ++s->glm_epoch ? s->glm_epoch : ++s->glm_epoch;  // never 0
    PROF(s, KID_GLM), glm_mix_kernel<<<nblk, GLM_NT, 0, s->stream>>>(s->G, ncol, s->rc_buckets.as<double>(), s->u_chip.as<double>(),
                                                  s->u_off.as<double>(), s->u_dsg.as<uint8_t>(), P, s->glm_part.as<double>(),
                                                  s->glm_ticket.as<unsigned>() + 8, s->d_glm, epoch);
    EHMM_CK(cudaGetLastError());
    EHMM_TRY(poll_epoch(s, epoch, "glmMixNB"));
    *value = s->h_glm[0];
    if (grad)
        for (int i = 0; i < 4; i++) grad[i] = s->h_glm[1 + i];
    return EHMM_OK;
}

// par = (lambda, b_bg, l_bg, b_enr, l_enr) -- optimDifferential's "newpar" (R/base-optim.R:80-82)
int glm_mix_zinb(ehmm_session *s, const double *par, double minZero, double *value, double *grad) {
    EHMM_REQUIRE(s->u_dsg.p, EHMM_ERR_STATE, "glmMixZINB: dtUnique has no Dsg.Mix columns (ehmm_set_groups)");
    MixZPar P;
    P.lam = par[0];
    P.zip = exp(par[0]) / (1.0 + exp(par[0]));
    P.bz = par[1];
    P.bn = par[3];
    P.minZero = minZero;
    P.thr = log(minZero) + 40.0;
    for (int D = 0; D < 2; D++) {
        P.size[D] = exp(D ? par[4] : par[2]);
        P.lsize[D] = log(P.size[D]);
        P.lgs[D] = lgamma(P.size[D]);
        P.dg[D] = host_digamma(P.size[D]);
    }
    EHMM_TRY(s->glm_part.reserve(sizeof(double) * ((size_t)EHMM_MAX_K * 148 * 4 * 4 + EHMM_MAX_K * 4) + 64 +
                                 sizeof(double) * ((size_t)148 * 4 * GLM_MAXOUT + GLM_MAXOUT)));
    const int ncol = s->B + 2;
    const int nblk = blocks_for((int64_t)ncol * (s->G + 1));
    PROF(s, KID_GLM), glm_mix_zinb_kernel<<<nblk, GLM_NT, 0, s->stream>>>(
        s->G, ncol, s->rc_buckets.as<double>(), s->u_chip.as<double>(), s->u_off.as<double>(), s->u_dsg.as<uint8_t>(), P,
        s->glm_part.as<double>());
    EHMM_CK(cudaGetLastError());
    double o[6];
    EHMM_TRY(finish(s, nblk, 6, o));
    *value = o[0];
    if (grad)
        for (int i = 0; i < 5; i++) grad[i] = o[1 + i];
    return EHMM_OK;
}

}  // namespace ehmm
