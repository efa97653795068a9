// igmk_contact.cuh - K2: population contact-frequency counts (sm_100a).
//
// counts[a][b] = #{ s : d2_s(a, b) <= (cr * (r_a + r_b))^2 }   ('<' if strict)
// for a tile of bead pairs - the arithmetic behind HssFile.buildContactMap as
// called from igm/steps/HicEvaluationStep.py:107-112 and
// alabtools.analysis.get_simulated_hic (igm/report/hic.py:51).  alabtools is
// not in the reference tree, so the contact definition follows the reference's
// own A-step (ActivationDistanceStep.py:396,442: float32 rcutsq, inclusive
// compare on the non-FMA float32 d2); strict=1 gives the '<' of the commented
// in-tree statement (HicEvaluationStep.py:89-92).  PARITY UNPINNED against
// alabtools.
//
// Bound: FP32 CUDA-core issue (about 10 instructions per bead pair per
// structure), not HBM: every coordinate is read once per tile row/column from
// L2 and reused 32 times out of shared memory.  Tensor cores do not apply
// (3-term non-FMA float32 sums that must match NumPy bit for bit).
//
// CTA = 256 threads = 64 (8 x 8) pair-positions x 4 structure slices; each
// thread owns a 4 x 4 block of bead pairs (beads ta + 8 aa, tb + 8 bb) and, per
// staged chunk of 32 structures, the 8 structures of its slice.
#pragma once
#include "igmk_device.cuh"

namespace igmk {

constexpr int kCtTile = 32;            // beads per tile side
constexpr int kCtThreads = 256;
constexpr int kCtStruct = 32;          // structures staged per step
constexpr int kCtRow = 3 * kCtStruct + 4;   // floats per bead in smem; /4 is odd -> conflict-free LDS.128

struct ContactParams {
    const float* coords;   // [nbead][3][npad]
    const float* radii;    // [nbead]
    uint32_t* counts;      // [nrows][ncols]
    int nstruct, npad, nbead;
    int row0, nrows, col0, ncols;
    float contact_range;
    int strict;
};

template <bool STRICT, bool TAIL>
__device__ __forceinline__ void contact_accumulate(const float* sa, const float* sb, int ta, int tb,
                                                   int slice, int nvalid,
                                                   const float (&rc)[4][4], int (&cnt)[4][4]) {
#pragma unroll
    for (int gq = 0; gq < 2; ++gq) {
        const int s4 = slice * 8 + gq * 4;       // first of 4 structures within the chunk
        float4 ax[4], ay[4], az[4];
#pragma unroll
        for (int aa = 0; aa < 4; ++aa) {
            const float* p = sa + (ta + 8 * aa) * kCtRow + s4;
            ax[aa] = *reinterpret_cast<const float4*>(p);
            ay[aa] = *reinterpret_cast<const float4*>(p + kCtStruct);
            az[aa] = *reinterpret_cast<const float4*>(p + 2 * kCtStruct);
        }
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
            const float* p = sb + (tb + 8 * bb) * kCtRow + s4;
            const float4 bx = *reinterpret_cast<const float4*>(p);
            const float4 by = *reinterpret_cast<const float4*>(p + kCtStruct);
            const float4 bz = *reinterpret_cast<const float4*>(p + 2 * kCtStruct);
            const float BX[4] = {bx.x, bx.y, bx.z, bx.w};
            const float BY[4] = {by.x, by.y, by.z, by.w};
            const float BZ[4] = {bz.x, bz.y, bz.z, bz.w};
#pragma unroll
            for (int aa = 0; aa < 4; ++aa) {
                const float AX[4] = {ax[aa].x, ax[aa].y, ax[aa].z, ax[aa].w};
                const float AY[4] = {ay[aa].x, ay[aa].y, ay[aa].z, ay[aa].w};
                const float AZ[4] = {az[aa].x, az[aa].y, az[aa].z, az[aa].w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float d2 = d2_nofma(AX[q], AY[q], AZ[q], BX[q], BY[q], BZ[q]);
                    bool hit = STRICT ? (d2 < rc[aa][bb]) : (d2 <= rc[aa][bb]);
                    if (TAIL) hit = hit && (s4 + q < nvalid);
                    cnt[aa][bb] += hit ? 1 : 0;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kCtThreads, 2)
contact_tile_kernel(const ContactParams P) {
    __shared__ __align__(16) float s_a[kCtTile * kCtRow];
    __shared__ __align__(16) float s_b[kCtTile * kCtRow];
    __shared__ uint32_t s_cnt[kCtTile * kCtTile];

    const int t = threadIdx.x;
    const int pos = t & 63, slice = t >> 6;
    const int ta = pos >> 3, tb = pos & 7;
    const int a_base = P.row0 + blockIdx.y * kCtTile;
    const int b_base = P.col0 + blockIdx.x * kCtTile;
    const int a_end = P.row0 + P.nrows, b_end = P.col0 + P.ncols;

    for (int e = t; e < kCtTile * kCtTile; e += kCtThreads) s_cnt[e] = 0u;

    float rc[4][4];
    int cnt[4][4];
#pragma unroll
    for (int aa = 0; aa < 4; ++aa) {
        const int a = a_base + ta + 8 * aa;
        const float ra = (a < a_end) ? __ldg(P.radii + a) : 0.f;
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
            const int b = b_base + tb + 8 * bb;
            const float rb = (b < b_end) ? __ldg(P.radii + b) : 0.f;
            const float r = __fmul_rn(P.contact_range, __fadd_rn(ra, rb));
            rc[aa][bb] = __fmul_rn(r, r);
            cnt[aa][bb] = 0;
        }
    }

    const size_t row = (size_t)3 * P.npad;
    for (int s0 = 0; s0 < P.nstruct; s0 += kCtStruct) {
        __syncthreads();
        // stage 32 beads x 3 components x 32 structures of each side
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int f = t + kCtThreads * k;          // 0 .. 767
            const int bead = f / 24, rem = f - bead * 24;
            const int comp = rem >> 3, v4 = rem & 7;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const int a = a_base + bead, b = b_base + bead;
            const size_t so = coord_off(s0 + 4 * v4) + (size_t)comp * kSeg;
            const float4 va = (a < a_end)
                ? __ldg(reinterpret_cast<const float4*>(P.coords + (size_t)a * row + so)) : z;
            const float4 vb = (b < b_end)
                ? __ldg(reinterpret_cast<const float4*>(P.coords + (size_t)b * row + so)) : z;
            *reinterpret_cast<float4*>(s_a + bead * kCtRow + comp * kCtStruct + 4 * v4) = va;
            *reinterpret_cast<float4*>(s_b + bead * kCtRow + comp * kCtStruct + 4 * v4) = vb;
        }
        __syncthreads();
        const int nvalid = P.nstruct - s0;      // structures of this chunk that exist
        if (nvalid >= kCtStruct) {
            if (P.strict) contact_accumulate<true, false>(s_a, s_b, ta, tb, slice, nvalid, rc, cnt);
            else          contact_accumulate<false, false>(s_a, s_b, ta, tb, slice, nvalid, rc, cnt);
        } else {
            if (P.strict) contact_accumulate<true, true>(s_a, s_b, ta, tb, slice, nvalid, rc, cnt);
            else          contact_accumulate<false, true>(s_a, s_b, ta, tb, slice, nvalid, rc, cnt);
        }
    }

    // combine the 4 structure slices
#pragma unroll
    for (int aa = 0; aa < 4; ++aa)
#pragma unroll
        for (int bb = 0; bb < 4; ++bb)
            atomicAdd(&s_cnt[(ta + 8 * aa) * kCtTile + (tb + 8 * bb)], (uint32_t)cnt[aa][bb]);
    __syncthreads();
    for (int e = t; e < kCtTile * kCtTile; e += kCtThreads) {
        const int a = a_base + (e >> 5), b = b_base + (e & 31);
        if (a < a_end && b < b_end)
            P.counts[(size_t)(a - P.row0) * P.ncols + (b - P.col0)] = s_cnt[e];
    }
}

}  // namespace igmk
