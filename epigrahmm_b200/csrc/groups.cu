// groups.cu -- dt$Group and dtUnique on the device (replaces the data.table step
// `dt[, Group := .GRP, by = c('ChIP'[, 'controls'], 'offsets'[, Dsg.Mix*])]` + `unique(dt, by='Group')`,
// R/base-core.R:97-104,184-198, the largest host-side cost of a fit at genome scale).
//
// Cells of the column-major long table are keyed by (count, offset bits[, control bits], design
// code of the sample).  One pass inserts every cell into an open-addressing table of 64-bit key
// hashes (atomicCAS claims a slot, atomicMin records the first cell of the group); the occupied
// slots are compacted, ordered by first appearance on the host (G entries only) -- which is exactly
// data.table's .GRP numbering -- and a second pass writes the ids.  A final pass compares every
// cell's full key with the key of its group's first cell, so a 64-bit hash collision cannot go
// unnoticed (the call fails and the caller falls back to host-side grouping).
#include "common.cuh"
#include <algorithm>
#include <vector>

namespace ehmm {

namespace {

struct DCode {
    int code[64];  // design code of each sample (samples with identical Dsg.Mix columns share one)
};

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

__device__ __forceinline__ unsigned long long cell_hash(int y, double off, double ctrl, int dc) {
    // +0.0 canonicalises -0.0 so that equal values hash equally
    unsigned long long h = mix64((unsigned long long)(unsigned)y * 0x9e3779b97f4a7c15ULL + (unsigned)dc);
    h = mix64(h ^ (unsigned long long)__double_as_longlong(off + 0.0));
    h = mix64(h ^ (unsigned long long)__double_as_longlong(ctrl + 0.0) * 0xd6e8feb86659fd93ULL);
    return h ? h : 1ULL;
}

__global__ void __launch_bounds__(256)
    grp_insert_kernel(int64_t M, int N, const int32_t *__restrict__ y, const double *__restrict__ off,
                      const double *__restrict__ ctrl, DCode D, unsigned long long *__restrict__ tags,
                      unsigned long long *__restrict__ first, unsigned long long mask, int32_t *__restrict__ cellslot,
                      int *__restrict__ fail) {
    const int64_t n = M * N;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(c / M);
        const unsigned long long h = cell_hash(y[c], off[c], ctrl ? ctrl[c] : 0.0, D.code[i]);
        unsigned long long slot = h & mask;
        int probes = 0;
        while (true) {
            const unsigned long long prev = atomicCAS(tags + slot, 0ULL, h);
            if (prev == 0ULL || prev == h) break;
            slot = (slot + 1) & mask;
            if (++probes > 4096) { atomicOr(fail, 1); break; }
        }
        atomicMin(first + slot, (unsigned long long)c);
        cellslot[c] = (int32_t)slot;
    }
}

// the occupied slots as (first cell, slot) pairs, in no particular order: keys[i] = first cell, slots[i] = slot
__global__ void grp_compact_kernel(unsigned long long cap, const unsigned long long *__restrict__ tags,
                                   const unsigned long long *__restrict__ first, unsigned long long *__restrict__ keys,
                                   uint32_t *__restrict__ slots, unsigned long long maxlist, unsigned long long *__restrict__ count) {
    for (unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; s < cap;
         s += (unsigned long long)gridDim.x * blockDim.x) {
        if (tags[s] != 0ULL) {
            const unsigned long long i = atomicAdd(count, 1ULL);
            if (i < maxlist) { keys[i] = first[s]; slots[i] = (uint32_t)s; }
        }
    }
}

// .GRP numbers the groups by first appearance: the pairs sorted by first cell give slot -> rank
__global__ void grp_rank_kernel(int64_t G, const uint32_t *__restrict__ slots_sorted, int32_t *__restrict__ slotrank) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < G) slotrank[slots_sorted[i]] = (int32_t)i;
}

// group[c] = rank of the cell's slot; verify the full key against the group's first cell
__global__ void __launch_bounds__(256)
    grp_assign_kernel(int64_t M, int N, const int32_t *__restrict__ y, const double *__restrict__ off,
                      const double *__restrict__ ctrl, DCode D, const int32_t *__restrict__ cellslot,
                      const int32_t *__restrict__ slotrank, const unsigned long long *__restrict__ first,
                      int32_t *__restrict__ group, int *__restrict__ fail) {
    const int64_t n = M * N;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
        const int32_t slot = cellslot[c];
        const int64_t f = (int64_t)first[slot];
        const bool same = y[c] == y[f] && off[c] == off[f] && (!ctrl || ctrl[c] == ctrl[f]) &&
                          D.code[(int)(c / M)] == D.code[(int)(f / M)];
        if (!same) atomicOr(fail, 2);
        group[c] = slotrank[slot] + 1;
    }
}

// dtUnique row g (1-based slot g of the arrays): the columns of the group's first cell
__global__ void grp_unique_kernel(int64_t G, int64_t M, const unsigned long long *__restrict__ firstsorted,
                                  const int32_t *__restrict__ y, const double *__restrict__ off,
                                  const double *__restrict__ ctrl, double *__restrict__ uy, double *__restrict__ uoff,
                                  double *__restrict__ uctrl, int32_t *__restrict__ usample) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g > G) return;
    if (g == 0) {
        uy[0] = 0.0;
        uoff[0] = 0.0;
        if (uctrl) uctrl[0] = 0.0;
        usample[0] = 0;
        return;
    }
    const int64_t f = (int64_t)firstsorted[g - 1];
    uy[g] = (double)y[f];
    uoff[g] = off[f];
    if (uctrl) uctrl[g] = ctrl[f];
    usample[g] = (int32_t)(f / M);
}

__global__ void grp_dsg_kernel(int64_t G, int B, int N, const int32_t *__restrict__ usample, const uint8_t *__restrict__ design,
                               uint8_t *__restrict__ udsg) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (G + 1) * B) return;
    const int b = (int)(idx / (G + 1));
    const int64_t g = idx - (int64_t)b * (G + 1);
    udsg[idx] = g == 0 ? 0 : design[b * N + usample[g]];
}

}  // namespace

int build_groups(ehmm_session *s, int64_t *G_out) {
    EHMM_REQUIRE(s->N > 0 && !s->encoded && s->d_counts && s->d_offsets, EHMM_ERR_STATE,
                 "build_groups: ehmm_set_data has not been called");
    const int64_t M = s->M, n = M * s->N;
    DCode D = {};
    if (s->have_design) {  // samples with identical design columns get the same code
        int ncode = 0;
        for (int i = 0; i < s->N; i++) {
            int found = -1;
            for (int j = 0; j < i && found < 0; j++) {
                bool eq = true;
                for (int b = 0; b < s->B; b++) eq = eq && s->design[b * s->N + i] == s->design[b * s->N + j];
                if (eq) found = D.code[j];
            }
            D.code[i] = found >= 0 ? found : ncode++;
        }
    }
    EHMM_TRY(s->group.reserve(sizeof(int32_t) * n));
    EHMM_TRY(s->rc_tmp.reserve(sizeof(int32_t) * n));
    int32_t *cellslot = s->rc_tmp.as<int32_t>();
    // work space kept in the session: allocating and freeing 100 MB of tables per call cost more than the kernels
    DevBuf &tags = s->grp_tags, &first = s->grp_first, &list = s->grp_list, &rank = s->grp_rank, &misc2 = s->grp_misc, &sortb = s->grp_sort;
    // flags / counter in the first 64 bytes, the design matrix behind them: sized once, because a
    // growing reserve() reallocates and would drop the verification flag written in between
    EHMM_TRY(misc2.reserve(64 + (size_t)EHMM_MAX_B * 64));
    unsigned long long cap = 1ULL << 22, G = 0;
    unsigned long long *keys0 = nullptr, *keys1 = nullptr;
    uint32_t *v0 = nullptr, *v1 = nullptr;
    unsigned *hist = nullptr;
    for (int attempt = 0;; attempt++) {
        EHMM_TRY(tags.reserve(sizeof(unsigned long long) * cap));
        EHMM_TRY(first.reserve(sizeof(unsigned long long) * cap));
        EHMM_CK(cudaMemsetAsync(tags.p, 0, sizeof(unsigned long long) * cap, s->stream));
        EHMM_CK(cudaMemsetAsync(first.p, 0xff, sizeof(unsigned long long) * cap, s->stream));
        EHMM_CK(cudaMemsetAsync(misc2.p, 0, 64, s->stream));
        int *fail = misc2.as<int>();
        unsigned long long *count = misc2.as<unsigned long long>() + 1;
        PROF(s, KID_MISC), grp_insert_kernel<<<148 * 8, 256, 0, s->stream>>>(
            M, s->N, s->d_counts, s->d_offsets, s->d_controls, D, tags.as<unsigned long long>(),
            first.as<unsigned long long>(), cap - 1, cellslot, fail);
        EHMM_CK(cudaGetLastError());
        const unsigned long long maxlist = cap / 2;
        // sort buffers: two key arrays, two value arrays, the radix sort's counters (one 256-bin histogram per 4096 pairs)
        const size_t kb = sizeof(unsigned long long) * maxlist, vb = sizeof(uint32_t) * maxlist, hb = sizeof(unsigned) * 256 * (maxlist / 4096 + 1);
        EHMM_TRY(sortb.reserve(2 * kb + 2 * vb + hb));
        keys0 = sortb.as<unsigned long long>();
        keys1 = keys0 + maxlist;
        v0 = reinterpret_cast<uint32_t *>(keys1 + maxlist);
        v1 = v0 + maxlist;
        hist = reinterpret_cast<unsigned *>(v1 + maxlist);
        PROF(s, KID_MISC), grp_compact_kernel<<<148 * 8, 256, 0, s->stream>>>(
            cap, tags.as<unsigned long long>(), first.as<unsigned long long>(), keys0, v1, maxlist, count);
        EHMM_CK(cudaGetLastError());
        unsigned long long *hc = reinterpret_cast<unsigned long long *>(s->h_pinned + 420);
        EHMM_CK(cudaMemcpyAsync(hc, misc2.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
        EHMM_CK(cudaStreamSynchronize(s->stream));
        const int failed = (int)(hc[0] & 0xffffffffULL);
        G = hc[1];
        if (!failed && G <= maxlist) break;
        EHMM_REQUIRE(attempt == 0, EHMM_ERR_NOMEM, "build_groups: more than %llu distinct groups", maxlist);
        cap <<= 3;  // one retry with an 8x larger table
    }
    // .GRP: groups numbered by first appearance -- the (first cell, slot) pairs sorted by first cell, on the device
    uint64_t *ksorted = nullptr;
    uint32_t *vsorted = nullptr;
    EHMM_TRY(radix_sort_pairs(s, reinterpret_cast<uint64_t *>(keys0), reinterpret_cast<uint64_t *>(keys1), v0, v1, hist, (int64_t)G, v1,
                              &ksorted, &vsorted));
    EHMM_TRY(rank.reserve(sizeof(int32_t) * cap));
    EHMM_CK(cudaMemsetAsync(rank.p, 0xff, sizeof(int32_t) * cap, s->stream));
    PROF(s, KID_MISC), grp_rank_kernel<<<(unsigned)((G + 255) / 256), 256, 0, s->stream>>>((int64_t)G, vsorted, rank.as<int32_t>());
    EHMM_CK(cudaGetLastError());
    EHMM_CK(cudaMemsetAsync(misc2.p, 0, 64, s->stream));
    PROF(s, KID_MISC), grp_assign_kernel<<<148 * 8, 256, 0, s->stream>>>(
        M, s->N, s->d_counts, s->d_offsets, s->d_controls, D, cellslot, rank.as<int32_t>(),
        first.as<unsigned long long>(), s->group.as<int32_t>(), misc2.as<int>());
    EHMM_CK(cudaGetLastError());
    // dtUnique
    const size_t gb = (size_t)(G + 1) * sizeof(double);
    EHMM_TRY(s->u_chip.reserve(gb));
    EHMM_TRY(s->u_off.reserve(gb));
    if (s->d_controls) EHMM_TRY(s->u_ctrl.reserve(gb)); else s->u_ctrl.release();
    EHMM_TRY(list.reserve(sizeof(unsigned long long) * (G + 1) + sizeof(int32_t) * (G + 1)));
    EHMM_CK(cudaMemcpyAsync(list.p, ksorted, sizeof(unsigned long long) * G, cudaMemcpyDeviceToDevice, s->stream));
    int32_t *usample = reinterpret_cast<int32_t *>(list.as<unsigned long long>() + (G + 1));
    PROF(s, KID_MISC), grp_unique_kernel<<<(unsigned)((G + 1 + 255) / 256), 256, 0, s->stream>>>(
        (int64_t)G, M, list.as<unsigned long long>(), s->d_counts, s->d_offsets, s->d_controls, s->u_chip.as<double>(),
        s->u_off.as<double>(), s->d_controls ? s->u_ctrl.as<double>() : nullptr, usample);
    EHMM_CK(cudaGetLastError());
    if (s->have_design) {
        EHMM_TRY(s->u_dsg.reserve((size_t)(G + 1) * s->B));
        EHMM_CK(cudaMemcpyAsync(misc2.as<uint8_t>() + 64, s->design, (size_t)s->B * s->N, cudaMemcpyHostToDevice, s->stream));
        const int64_t tot = (int64_t)(G + 1) * s->B;
        PROF(s, KID_MISC), grp_dsg_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s->stream>>>(
            (int64_t)G, s->B, s->N, usample, misc2.as<uint8_t>() + 64, s->u_dsg.as<uint8_t>());
        EHMM_CK(cudaGetLastError());
    }
    int *hf = reinterpret_cast<int *>(s->h_pinned + 424);
    EHMM_CK(cudaMemcpyAsync(hf, misc2.p, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    // first cell of every group (ehmm_groups_first_cell: bin blocks merge their dictionaries with it)
    s->group_first.resize(G);
    EHMM_CK(cudaMemcpyAsync(s->group_first.data(), list.p, sizeof(int64_t) * G, cudaMemcpyDeviceToHost, s->stream));
    EHMM_CK(cudaStreamSynchronize(s->stream));
    EHMM_REQUIRE(*hf == 0, EHMM_ERR_STATE, "build_groups: 64-bit key-hash collision detected; group the table on the host instead");
    s->G = (int64_t)G;
    posteriors_changed(s);
    s->groups_consistent = true;
    s->rc_columns = 0;
    if (G_out) *G_out = (int64_t)G;
    return EHMM_OK;
}

namespace {
__global__ void grp_remap_kernel(int64_t n, const int32_t *__restrict__ map, int32_t *__restrict__ group) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        group[i] = map[group[i]];
}
}  // namespace

int groups_remap(ehmm_session *s, const int32_t *new_id, int64_t G_new) {
    EHMM_REQUIRE(s->G > 0 && s->group.p, EHMM_ERR_STATE, "groups_remap: no groups");
    std::vector<int32_t> map((size_t)s->G + 1, 0);
    for (int64_t g = 0; g < s->G; g++) {
        EHMM_REQUIRE(new_id[g] >= 1 && new_id[g] <= G_new, EHMM_ERR_ARG, "groups_remap: new id %d of group %lld outside [1, %lld]",
                     new_id[g], (long long)(g + 1), (long long)G_new);
        map[(size_t)g + 1] = new_id[g];
    }
    EHMM_TRY(s->rc_tmp.reserve(sizeof(int32_t) * map.size()));
    EHMM_CK(cudaMemcpyAsync(s->rc_tmp.p, map.data(), sizeof(int32_t) * map.size(), cudaMemcpyHostToDevice, s->stream));
    PROF(s, KID_MISC), grp_remap_kernel<<<148 * 8, 256, 0, s->stream>>>(s->M * s->N, s->rc_tmp.as<int32_t>(), s->group.as<int32_t>());
    EHMM_CK(cudaGetLastError());
    EHMM_CK(cudaStreamSynchronize(s->stream));  // `map` is pageable host memory
    s->group_first.clear();
    return EHMM_OK;
}

int groups_fetch(ehmm_session *s, int32_t *group, double *u_chip, double *u_off, double *u_ctrl, uint8_t *u_dsg) {
    EHMM_REQUIRE(s->G > 0, EHMM_ERR_STATE, "no groups");
    const int64_t G = s->G;
    if (group)
        EHMM_CK(cudaMemcpyAsync(group, s->group.p, sizeof(int32_t) * s->M * s->N, cudaMemcpyDeviceToHost, s->stream));
    if (u_chip) EHMM_CK(cudaMemcpyAsync(u_chip, s->u_chip.as<double>() + 1, sizeof(double) * G, cudaMemcpyDeviceToHost, s->stream));
    if (u_off) EHMM_CK(cudaMemcpyAsync(u_off, s->u_off.as<double>() + 1, sizeof(double) * G, cudaMemcpyDeviceToHost, s->stream));
    if (u_ctrl) {
        EHMM_REQUIRE(s->u_ctrl.p, EHMM_ERR_STATE, "dtUnique has no controls column");
        EHMM_CK(cudaMemcpyAsync(u_ctrl, s->u_ctrl.as<double>() + 1, sizeof(double) * G, cudaMemcpyDeviceToHost, s->stream));
    }
    if (u_dsg) {
        EHMM_REQUIRE(s->u_dsg.p && s->B > 0, EHMM_ERR_STATE, "dtUnique has no Dsg.Mix columns");
        for (int b = 0; b < s->B; b++)
            EHMM_CK(cudaMemcpyAsync(u_dsg + (size_t)b * G, s->u_dsg.as<uint8_t>() + (size_t)b * (G + 1) + 1, (size_t)G,
                                    cudaMemcpyDeviceToHost, s->stream));
    }
    EHMM_CK(cudaStreamSynchronize(s->stream));
    return EHMM_OK;
}

}  // namespace ehmm
