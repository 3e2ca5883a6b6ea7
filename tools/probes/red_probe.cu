// red_probe.cu -- how fast can B200 scatter-add?  Measures the building blocks considered for the
// rejection tables (csrc/rc.cu): native 64-bit integer / fp64 global reductions (REDG), one or two
// limbs, warp-level pre-aggregation, for a flat and a peaked distribution of group ids.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_probe red_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
typedef unsigned long long u64;

__global__ void k_read(const int *idx, const double *v, int64_t n, double *out) {
    double acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += v[i] * idx[i];
    if (acc == 1.2345) out[0] = acc;
}
__global__ void k_u64(const int *idx, const double *v, int64_t n, u64 *acc) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(acc + idx[i], __double2ull_rz(v[i] * 4611686018427387904.0));
}
__global__ void k_u64x2(const int *idx, const double *v, int64_t n, u64 *acc) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const u64 x = __double2ull_rz(v[i] * 4611686018427387904.0);
        atomicAdd(acc + 2 * (int64_t)idx[i], x & 0x7fffffffull);
        atomicAdd(acc + 2 * (int64_t)idx[i] + 1, x >> 31);
    }
}
__global__ void k_u64x2_split(const int *idx, const double *v, int64_t n, u64 *acc, int64_t G) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const u64 x = __double2ull_rz(v[i] * 4611686018427387904.0);
        atomicAdd(acc + idx[i], x & 0x7fffffffull);
        atomicAdd(acc + G + idx[i], x >> 31);
    }
}
__global__ void k_f64(const int *idx, const double *v, int64_t n, double *acc) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(acc + idx[i], v[i]);
}
// lanes holding the same id add up in the warp first (integer adds: any order gives the same bits)
__global__ void k_u64x2_match(const int *idx, const double *v, int64_t n, u64 *acc) {
    const int lane = threadIdx.x & 31;
    const int64_t nround = (n + (int64_t)gridDim.x * blockDim.x - 1) / ((int64_t)gridDim.x * blockDim.x);
    for (int64_t r = 0; r < nround; r++) {
        const int64_t i = r * gridDim.x * blockDim.x + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const bool ok = i < n;
        const int g = ok ? idx[i] : -1 - lane;
        u64 x = ok ? __double2ull_rz(v[i] * 4611686018427387904.0) : 0;
        u64 lo = x & 0x7fffffffull, hi = x >> 31;
        const unsigned peers = __match_any_sync(0xffffffffu, g);
        const int leader = __ffs(peers) - 1;
        if (peers != (1u << lane)) {  // someone shares my id: sum over the peer set
            unsigned rest = peers & ~(1u << leader);
            u64 slo = 0, shi = 0;
            // leader gathers (serial over peers; few)
            for (unsigned m = peers; m; m &= m - 1) {
                const int src = __ffs(m) - 1;
                const u64 a = __shfl_sync(peers, lo, src), b = __shfl_sync(peers, hi, src);
                slo += a; shi += b;
            }
            lo = slo; hi = shi; (void)rest;
        }
        if (ok && lane == leader) { atomicAdd(acc + 2 * (int64_t)g, lo); atomicAdd(acc + 2 * (int64_t)g + 1, hi); }
    }
}
// per-warp sort-free aggregation is expensive when ids are distinct; test cost of match alone via flat ids

template <typename F> float timeit(F f, int reps = 5) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char **argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 15000000;
    struct Case { const char *name; int64_t G; double zipf; } cases[] = {
        {"flat G=4e4", 40000, 0.0}, {"flat G=1.1e5", 110000, 0.0}, {"peaked G=200 (zipf 1.2)", 200, 1.2}, {"peaked G=4e4 (zipf 1.0)", 40000, 1.0}};
    std::vector<int> h(n); std::vector<double> hv(n);
    int *d_idx; double *d_v; u64 *d_acc; double *d_out;
    CK(cudaMalloc(&d_idx, n * 4)); CK(cudaMalloc(&d_v, n * 8)); CK(cudaMalloc(&d_acc, 16 * 1200000)); CK(cudaMalloc(&d_out, 64));
    for (auto &c : cases) {
        // inverse-CDF sampling of a zipf(s) over G ids (s = 0: uniform)
        std::vector<double> cdf(c.G); double tot = 0;
        for (int64_t g = 0; g < c.G; g++) { tot += c.zipf == 0 ? 1.0 : pow((double)(g + 1), -c.zipf); cdf[g] = tot; }
        uint64_t st = 88172645463325252ull;
        for (int64_t i = 0; i < n; i++) {
            st ^= st << 13; st ^= st >> 7; st ^= st << 17;
            const double u = (st >> 11) * (1.0 / 9007199254740992.0) * tot;
            h[i] = (int)(std::lower_bound(cdf.begin(), cdf.end(), u) - cdf.begin());
            hv[i] = 0.05 + 0.9 * ((st >> 20) & 0xfffff) / 1048576.0;
        }
        CK(cudaMemcpy(d_idx, h.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_v, hv.data(), n * 8, cudaMemcpyHostToDevice));
        const int grid = 148 * 8, nt = 256;
        printf("== %s, n = %lld\n", c.name, (long long)n);
        printf("  read only          %.4f ms\n", timeit([&] { k_read<<<grid, nt>>>(d_idx, d_v, n, d_out); }));
        printf("  fp64 RED           %.4f ms\n", timeit([&] { k_f64<<<grid, nt>>>(d_idx, d_v, n, (double *)d_acc); }));
        printf("  u64 RED x1         %.4f ms\n", timeit([&] { k_u64<<<grid, nt>>>(d_idx, d_v, n, d_acc); }));
        printf("  u64 RED x2 adjacent %.4f ms\n", timeit([&] { k_u64x2<<<grid, nt>>>(d_idx, d_v, n, d_acc); }));
        printf("  u64 RED x2 split    %.4f ms\n", timeit([&] { k_u64x2_split<<<grid, nt>>>(d_idx, d_v, n, d_acc, c.G); }));
        printf("  u64 x2 + match_any  %.4f ms\n", timeit([&] { k_u64x2_match<<<grid, nt>>>(d_idx, d_v, n, d_acc); }));
        for (int g2 : {148 * 2, 148 * 16, 148 * 32})
            printf("  u64 RED x2 adjacent grid=%d  %.4f ms\n", g2, timeit([&] { k_u64x2<<<g2, nt>>>(d_idx, d_v, n, d_acc); }));
    }
    return 0;
}
