// igmk_rank.cuh - K5: population rank matching (next row f4: the FISH and polymer
// assignment steps, which share the bead-row gather of the A-step).
//
// Both steps take, for one item (a FISH probe, a FISH probe pair, a polymer bond), one
// distance per structure of the population, rank the structures by it and hand the
// structure of rank k the k-th value of a sorted target distribution:
//   FISH   igm/steps/FishAssignmentStep.py:23-79   get_pair_dists / get_rad_dists /
//          get_min_max_and_idx (min and max over the copy combinations per structure,
//          idx = np.argsort(np.argsort(.))), task :189-193 / :214-219 target[idx]
//   polymer igm/steps/PolymerAssignmentStep.py:24-33 get_polymer_dists, task :118-125
//          sampled_distances[sorting_idx]
// Arithmetic of np.linalg.norm(x - y, axis=1) on float32 rows (NumPy 2.3, probed in this
// image, tests/test_rank_cpu.py): sqrtf_rn(fl32(fl32(dx*dx + dy*dy) + dz*dz)), dx = fl32(x - y),
// sequential and without FMA - the A-step's d2 followed by a float32 square root.  Radial
// distances (norm of the coordinates themselves) use the all-zero row behind the
// population as the partner: x - 0 is exact.
//
// One CTA per item.  The N values become 64-bit keys (float bits << 32 | structure), so a
// plain bitonic sort in shared memory is *stable by construction*: equal distances are
// ranked by structure index.  (NumPy's default argsort is not stable; the reference's order
// inside a tie group is an implementation accident - parity is exact wherever the
// distances differ, and the multiset of targets handed to a tie group is equal.)
// After the sort key k holds (value, s): rank[s] = k goes to a second shared array so the
// results leave in coalesced, structure-ordered rows.
#pragma once
#include "igmk_device.cuh"

namespace igmk {

struct RankParams {
    const float*   coords;      // [bead][segment][xyz][128]
    const int32_t* a;           // [n_items][2] bead ids of the copies of the first locus (-1: absent)
    const int32_t* b;           // [n_items][2] second locus; NULL: radial distance (partner = origin)
    const float*   target;      // sorted target distribution(s); NULL: none
    long long      target_stride;   // floats between the targets of consecutive items (0: one shared target)
    float*         matched;     // [n_items][nstruct] target[rank[s]]          (NULL: skip)
    int32_t*       rank;        // [n_items][nstruct]                           (NULL: skip)
    float*         value;       // [n_items][nstruct] the reduced distance      (NULL: skip)
    long long n_items;
    int nstruct, npad, nbead;
    int zero_bead;
    int reduce;                 // 0: min over the copy combinations, 1: max
    int m;                      // power of two >= nstruct (keys sorted)
};

__device__ __forceinline__ float norm_nofma(const float* __restrict__ pa, const float* __restrict__ pb, size_t o) {
    return __fsqrt_rn(d2_nofma(__ldg(pa + o), __ldg(pa + o + kSeg), __ldg(pa + o + 2 * kSeg),
                               __ldg(pb + o), __ldg(pb + o + kSeg), __ldg(pb + o + 2 * kSeg)));
}

__global__ void rank_match_kernel(const RankParams P) {
    extern __shared__ __align__(16) unsigned char rk_smem[];
    u64* keys = reinterpret_cast<u64*>(rk_smem);
    uint32_t* rk = reinterpret_cast<uint32_t*>(rk_smem + (size_t)P.m * sizeof(u64));
    const size_t rowf = (size_t)3 * P.npad;
    const int tid = threadIdx.x, nthr = blockDim.x;

    for (long long item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        int a0 = __ldg(P.a + 2 * item), a1 = __ldg(P.a + 2 * item + 1);
        int b0 = P.b ? __ldg(P.b + 2 * item) : P.zero_bead;
        int b1 = P.b ? __ldg(P.b + 2 * item + 1) : -1;
        const bool ok = a0 >= 0 && a0 < P.nbead && a1 < P.nbead &&
                        b0 >= 0 && b0 <= P.nbead && b1 < P.nbead && (P.b == nullptr || b0 < P.nbead);
        const float* pa0 = P.coords + (size_t)(ok ? a0 : 0) * rowf;
        const float* pa1 = P.coords + (size_t)((ok && a1 >= 0) ? a1 : 0) * rowf;
        const float* pb0 = P.coords + (size_t)(ok ? b0 : 0) * rowf;
        const float* pb1 = P.coords + (size_t)((ok && b1 >= 0) ? b1 : 0) * rowf;
        const bool max_ = P.reduce != 0;

        // 1. one reduced distance per structure -> 64-bit key
        for (int s = tid; s < P.m; s += nthr) {
            u64 key = ~0ull;
            if (s < P.nstruct) {
                const size_t o = coord_off(s);
                float v = ok ? norm_nofma(pa0, pb0, o) : __int_as_float(0x7fc00000);
                if (ok && b1 >= 0) { const float w = norm_nofma(pa0, pb1, o); v = max_ ? fmaxf(v, w) : fminf(v, w); }
                if (ok && a1 >= 0) {
                    const float w = norm_nofma(pa1, pb0, o); v = max_ ? fmaxf(v, w) : fminf(v, w);
                    if (b1 >= 0) { const float x = norm_nofma(pa1, pb1, o); v = max_ ? fmaxf(v, x) : fminf(v, x); }
                }
                if (P.value) P.value[item * P.nstruct + s] = v;
                key = ((u64)__float_as_uint(v) << 32) | (uint32_t)s;       // v >= 0: bit order = float order; NaN sorts last
            }
            keys[s] = key;
        }
        __syncthreads();

        // 2. bitonic sort, ascending
        for (int k = 2; k <= P.m; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < (P.m >> 1); t += nthr) {
                    const int lo = 2 * t - (t & (j - 1));
                    const int hi = lo + j;
                    const u64 x = keys[lo], y = keys[hi];
                    const bool up = (lo & k) == 0;
                    if ((x > y) == up) { keys[lo] = y; keys[hi] = x; }
                }
                __syncthreads();
            }
        }

        // 3. rank of every structure, then structure-ordered output rows
        for (int k = tid; k < P.nstruct; k += nthr) rk[(uint32_t)keys[k]] = (uint32_t)k;
        __syncthreads();
        const float* tgt = P.target ? P.target + item * P.target_stride : nullptr;
        for (int s = tid; s < P.nstruct; s += nthr) {
            const uint32_t r = rk[s];
            if (P.rank) P.rank[item * P.nstruct + s] = (int32_t)r;
            if (P.matched) P.matched[item * P.nstruct + s] = __ldg(tgt + r);
        }
        __syncthreads();
    }
}

}  // namespace igmk
