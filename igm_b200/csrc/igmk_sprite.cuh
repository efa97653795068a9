// igmk_sprite.cuh - K4: SPRITE cluster radius of gyration with exhaustive choice of
// chromosome copies (next row f3).
//
// GPU form of the reference's one native component, get_rg2s_cpp
// (igm/cython_compiled/cpp_sprite_assignment.cpp:79-143; Cython entry get_rgs2,
// sprite.pyx:36-101): for every structure, enumerate all prod K(i) combinations of
// alternative locations (copy k of region i), compute Rg^2 of the chosen points and
// keep the first strictly smallest; then the first structure with the strictly
// smallest Rg^2.
//
// Arithmetic (gyration_radius_sq :49-61, Vec3 :9-43; the reference is built with
// plain -O2 for baseline x86-64, i.e. no FMA): float32 throughout,
//     mean = (((0 + p0) + p1) + ...) / float(n)        component-wise
//     rg   = sum_i ((dx*dx + dy*dy) + dz*dz)           sequential, d = p_i - mean
//     Rg^2 = rg / float(n)
// Combination k selects copy (k / prod_{j<i} K(j)) % K(i) of region i (:63-77).
// INF = 1e8 (:4): a structure whose every combination has Rg^2 >= 1e8 reports 1e8
// and copy indices -1, exactly like the reference.
//
// One thread per (cluster, structure); coordinates come from the population
// resident in HBM (coalesced over structures), not from a per-cluster gather.
#pragma once
#include "igmk_device.cuh"

namespace igmk {

constexpr int kSpMaxRegions = 24;
constexpr int kSpMaxCopies = 48;     // total alternative locations of one cluster
constexpr float kSpInf = 100000000.0f;

struct SpriteParams {
    const float*   coords;
    const int32_t* region_ptr;   // [n_clusters + 1] -> regions
    const int32_t* copy_ptr;     // [n_regions_total + 1] -> beads
    const int32_t* beads;        // bead id of every alternative location
    float*   rg2s;               // [n_clusters][nstruct]
    int32_t* copy_idx;           // [region_ptr[c] * nstruct + s * n_regions(c) + i]
    int32_t* min_struct;         // [n_clusters]  (-1: no structure below INF)
    int n_clusters, nstruct, npad, nbead;
};

__global__ void __launch_bounds__(128)
sprite_rg2_kernel(const SpriteParams P) {
    const int c = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.nstruct) return;
    const int r0 = __ldg(P.region_ptr + c), nreg = __ldg(P.region_ptr + c + 1) - r0;
    float px[kSpMaxCopies], py[kSpMaxCopies], pz[kSpMaxCopies];
    int first[kSpMaxRegions], ncopy[kSpMaxRegions];
    const size_t rowf = (size_t)3 * P.npad;
    const size_t so = coord_off(s);
    long long ncomb = 1;
    int tot = 0;
    for (int i = 0; i < nreg; ++i) {
        const int b0 = __ldg(P.copy_ptr + r0 + i), b1 = __ldg(P.copy_ptr + r0 + i + 1);
        first[i] = tot;
        ncopy[i] = b1 - b0;
        ncomb *= (b1 - b0);
        for (int b = b0; b < b1; ++b, ++tot) {
            const float* p = P.coords + (size_t)__ldg(P.beads + b) * rowf + so;
            px[tot] = __ldg(p); py[tot] = __ldg(p + kSeg); pz[tot] = __ldg(p + 2 * kSeg);
        }
    }
    const float fn = (float)nreg;
    float best = kSpInf;
    long long best_k = -1;
    for (long long k = 0; k < ncomb; ++k) {
        // mean of the chosen points
        float mx = 0.f, my = 0.f, mz = 0.f;
        long long kk = k;
        for (int i = 0; i < nreg; ++i) {
            const int si = (int)(kk % ncopy[i]);
            kk /= ncopy[i];
            const int t = first[i] + si;
            mx = __fadd_rn(mx, px[t]); my = __fadd_rn(my, py[t]); mz = __fadd_rn(mz, pz[t]);
        }
        mx = __fdiv_rn(mx, fn); my = __fdiv_rn(my, fn); mz = __fdiv_rn(mz, fn);
        float rg = 0.f;
        kk = k;
        for (int i = 0; i < nreg; ++i) {
            const int si = (int)(kk % ncopy[i]);
            kk /= ncopy[i];
            const int t = first[i] + si;
            const float dx = __fsub_rn(px[t], mx), dy = __fsub_rn(py[t], my), dz = __fsub_rn(pz[t], mz);
            rg = __fadd_rn(rg, __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        }
        const float rg2 = __fdiv_rn(rg, fn);
        if (rg2 < best) { best = rg2; best_k = k; }
    }
    P.rg2s[(size_t)c * P.nstruct + s] = best;
    int32_t* ci = P.copy_idx + (size_t)r0 * P.nstruct + (size_t)s * nreg;
    long long kk = best_k;
    for (int i = 0; i < nreg; ++i) {
        ci[i] = (best_k < 0) ? -1 : (int)(kk % ncopy[i]);
        if (best_k >= 0) kk /= ncopy[i];
    }
}

// first structure with the strictly smallest Rg^2 below INF (one CTA per cluster)
__global__ void __launch_bounds__(256)
sprite_argmin_kernel(const float* __restrict__ rg2s, int nstruct, int32_t* __restrict__ min_struct) {
    __shared__ float s_v[256];
    __shared__ int s_i[256];
    const float* v = rg2s + (size_t)blockIdx.x * nstruct;
    float bv = kSpInf;
    int bi = -1;
    for (int s = threadIdx.x; s < nstruct; s += blockDim.x) {
        const float x = v[s];
        if (x < bv) { bv = x; bi = s; }          // ascending s per thread: first minimum
    }
    s_v[threadIdx.x] = bv; s_i[threadIdx.x] = bi;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) {
            const float ov = s_v[threadIdx.x + w];
            const int oi = s_i[threadIdx.x + w];
            const bool take = oi >= 0 && (s_i[threadIdx.x] < 0 || ov < s_v[threadIdx.x] ||
                                          (ov == s_v[threadIdx.x] && oi < s_i[threadIdx.x]));
            if (take) { s_v[threadIdx.x] = ov; s_i[threadIdx.x] = oi; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) min_struct[blockIdx.x] = s_i[0];
}

}  // namespace igmk
