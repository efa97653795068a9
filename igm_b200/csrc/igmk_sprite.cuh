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

// K4b: Rg^2 of a WHOLE cluster under a per-structure choice of chromosome copies - the
// second half of compute_gyration_radius (igm/cython_compiled/sprite.pyx:238-283): every
// segment of the cluster follows the copy that stage one (K4 on one representative segment
// per chromosome) selected for its chromosome in that structure, and get_rgs2 is evaluated
// on the selected beads with one location per segment (:279-282).  Also serves the
// single-chromosome branch (:200-215) with a constant selection per copy.
//
// Same float32 arithmetic as above (gyration_radius_sq, cpp:49-61), segments in the order
// given (all_segments of :243); with one combination the reference reports
// min(Rg^2, INF) (cpp:104-131: `if (rg2 < best)` from best = INF).  A negative selection
// (stage one found nothing below INF) indexes from the end, as NumPy does at :270-271.
//
// One thread per (cluster, structure), two passes over the segments; the coordinate reads
// are coalesced over structures (both passes hit the same rows, the second from L1 / L2).
struct SpriteClusterParams {
    const float*   coords;
    const int32_t* seg_ptr;      // [n_clusters + 1] -> segments
    const int32_t* loc_ptr;      // [n_segments_total + 1] -> beads (the copies of each segment)
    const int32_t* beads;
    const int32_t* seg_group;    // [n_segments_total] selection column of the segment within its cluster
    const int32_t* group_ptr;    // [n_clusters + 1] -> selection columns
    const int32_t* sel;          // [group_ptr[c] * nstruct + s * ngroups(c) + g]  (K4's copy_idx layout)
    float* rg2s;                 // [n_clusters][nstruct]
    int n_clusters, nstruct, npad;
};

__global__ void __launch_bounds__(128)
sprite_cluster_rg2_kernel(const SpriteClusterParams P) {
    const int c = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.nstruct) return;
    const int s0 = __ldg(P.seg_ptr + c), nseg = __ldg(P.seg_ptr + c + 1) - s0;
    const int g0 = __ldg(P.group_ptr + c), ngrp = __ldg(P.group_ptr + c + 1) - g0;
    const int32_t* sel = P.sel + (size_t)g0 * P.nstruct + (size_t)s * ngrp;
    const size_t rowf = (size_t)3 * P.npad;
    const size_t so = coord_off(s);
    auto point = [&](int i) -> const float* {
        const int l0 = __ldg(P.loc_ptr + s0 + i), nl = __ldg(P.loc_ptr + s0 + i + 1) - l0;
        int k = __ldg(sel + __ldg(P.seg_group + s0 + i));
        if (k < 0) k += nl;
        k = (k < 0) ? 0 : (k >= nl ? nl - 1 : k);           // (never out of range for valid input)
        return P.coords + (size_t)__ldg(P.beads + l0 + k) * rowf + so;
    };
    const float fn = (float)nseg;
    float mx = 0.f, my = 0.f, mz = 0.f;
    for (int i = 0; i < nseg; ++i) {
        const float* p = point(i);
        mx = __fadd_rn(mx, __ldg(p)); my = __fadd_rn(my, __ldg(p + kSeg)); mz = __fadd_rn(mz, __ldg(p + 2 * kSeg));
    }
    mx = __fdiv_rn(mx, fn); my = __fdiv_rn(my, fn); mz = __fdiv_rn(mz, fn);
    float rg = 0.f;
    for (int i = 0; i < nseg; ++i) {
        const float* p = point(i);
        const float dx = __fsub_rn(__ldg(p), mx), dy = __fsub_rn(__ldg(p + kSeg), my), dz = __fsub_rn(__ldg(p + 2 * kSeg), mz);
        rg = __fadd_rn(rg, __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    }
    const float rg2 = __fdiv_rn(rg, fn);
    P.rg2s[(size_t)c * P.nstruct + s] = (rg2 < kSpInf) ? rg2 : kSpInf;
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
