// igmk_restraint.cuh - K3: Hi-C restraint selection for the M-step (next row f2).
//
// For every actdist record (row, col, dist) and every structure s the reference's
// intraHiC / interHiC._apply (igm/restraints/intra_hic.py:39-58,
// inter_hic.py:39-58) adds a HarmonicUpperBound when
//     model.particles[row] - model.particles[col] <= dist
// with Particle.__sub__ = np.linalg.norm(pos_i - pos_j) on float32 positions
// (igm/model/particle.py:35-36), and chrom[row] == chrom[col] (intra) or != (inter).
// This kernel evaluates that test for all records x structures at once and
// returns a bitmap (bit s of record k) plus the per-record popcount.
//
// Arithmetic of np.linalg.norm on a float32 3-vector in this image (NumPy 2.3 +
// OpenBLAS 0.3.30 sdot; verified against 20 000 random vectors):
//     d   = fl32(x_i - x_j)                       (component-wise)
//     dot = fl32( f64(fl32(dx*dx)) + f64(fl32(dy*dy)) + f64(fl32(dz*dz)) )
//     norm = sqrtf_rn(dot)
// (float32 products, float64 accumulation, one rounding back to float32).  sqrtf is
// monotonic, so norm <= dist  <=>  dot <= T with T the largest float32 whose
// correctly rounded square root is <= dist; T is found once per record.
#pragma once
#include "igmk_device.cuh"

namespace igmk {

struct RestraintParams {
    const float*   coords;     // [bead][segment][xyz][128]
    const int32_t* chrom;      // [nbead] chromosome id of every bead (hss index chrom)
    const int32_t* row;
    const int32_t* col;
    const float*   dist;
    uint32_t*      bitmap;     // [n_rec][npad / 32]; bit s % 32 of word s / 32 = structure s
    int32_t*       counts;     // [n_rec]
    long long n_rec;
    int nstruct, npad, nbead;
    int kind;                  // 0: intra (same chromosome), 1: inter, 2: no chromosome test
};

// largest float32 t with sqrtf_rn(t) <= dist   (dist >= 0, finite)
__device__ __forceinline__ float sqrt_threshold(float dist) {
    float t = __fmul_rn(dist, dist);
    if (!(t < __int_as_float(0x7f800000))) t = __int_as_float(0x7f7fffff);
    for (int k = 0; k < 4 && __fsqrt_rn(t) > dist && t > 0.f; ++k)
        t = __uint_as_float(__float_as_uint(t) - 1u);
    for (int k = 0; k < 4; ++k) {
        const float u = __uint_as_float(__float_as_uint(t) + 1u);
        if (!(u < __int_as_float(0x7f800000)) || __fsqrt_rn(u) > dist) break;
        t = u;
    }
    return t;
}

__device__ __forceinline__ float dot3_blas(float dx, float dy, float dz) {
    const double s = __dadd_rn(__dadd_rn((double)__fmul_rn(dx, dx), (double)__fmul_rn(dy, dy)),
                               (double)__fmul_rn(dz, dz));
    return __double2float_rn(s);
}

constexpr int kRsWarps = 8;

// one warp per record; lane c handles the float4 chunks c, c + 32, ...
__global__ void __launch_bounds__(32 * kRsWarps)
restraint_select_kernel(const RestraintParams P) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * kRsWarps + (threadIdx.x >> 5);
    const long long stride = (long long)gridDim.x * kRsWarps;
    const int nseg = P.npad >> 7;                    // 128-structure segments
    const size_t rowf = (size_t)3 * P.npad;
    for (long long k = warp0; k < P.n_rec; k += stride) {
        const int a = __ldg(P.row + k), b = __ldg(P.col + k);
        const float dist = __ldg(P.dist + k);
        uint32_t* out = P.bitmap + (size_t)k * (P.npad >> 5);
        bool ok = a >= 0 && b >= 0 && a < P.nbead && b < P.nbead && dist >= 0.f;   // NaN: false
        if (ok && P.kind != 2) {
            const bool same = __ldg(P.chrom + a) == __ldg(P.chrom + b);
            ok = (P.kind == 0) ? same : !same;
        }
        int total = 0;
        if (!ok) {
            for (int w = lane; w < (P.npad >> 5); w += 32) out[w] = 0u;
        } else {
            const float T = isinf(dist) ? __int_as_float(0x7f800000) : sqrt_threshold(dist);
            const float* pa = P.coords + (size_t)a * rowf + (size_t)lane * 4;
            const float* pb = P.coords + (size_t)b * rowf + (size_t)lane * 4;
            for (int v = 0; v < nseg; ++v) {
                const float4 ax = __ldg(reinterpret_cast<const float4*>(pa));
                const float4 ay = __ldg(reinterpret_cast<const float4*>(pa + kSeg));
                const float4 az = __ldg(reinterpret_cast<const float4*>(pa + 2 * kSeg));
                const float4 bx = __ldg(reinterpret_cast<const float4*>(pb));
                const float4 by = __ldg(reinterpret_cast<const float4*>(pb + kSeg));
                const float4 bz = __ldg(reinterpret_cast<const float4*>(pb + 2 * kSeg));
                const int s0 = v * 128 + lane * 4;
                uint32_t nib = 0u;
                nib |= (dot3_blas(__fsub_rn(ax.x, bx.x), __fsub_rn(ay.x, by.x), __fsub_rn(az.x, bz.x)) <= T && s0 < P.nstruct) ? 1u : 0u;
                nib |= (dot3_blas(__fsub_rn(ax.y, bx.y), __fsub_rn(ay.y, by.y), __fsub_rn(az.y, bz.y)) <= T && s0 + 1 < P.nstruct) ? 2u : 0u;
                nib |= (dot3_blas(__fsub_rn(ax.z, bx.z), __fsub_rn(ay.z, by.z), __fsub_rn(az.z, bz.z)) <= T && s0 + 2 < P.nstruct) ? 4u : 0u;
                nib |= (dot3_blas(__fsub_rn(ax.w, bx.w), __fsub_rn(ay.w, by.w), __fsub_rn(az.w, bz.w)) <= T && s0 + 3 < P.nstruct) ? 8u : 0u;
                total += __popc(nib);
                // natural bit order: word w of the segment = lanes 8w .. 8w+7, nibble of lane 8w+m at bit 4m
                uint32_t word = nib << (4 * (lane & 7));
                word |= __shfl_xor_sync(0xffffffffu, word, 1);
                word |= __shfl_xor_sync(0xffffffffu, word, 2);
                word |= __shfl_xor_sync(0xffffffffu, word, 4);
                if ((lane & 7) == 0) out[v * 4 + (lane >> 3)] = word;
                pa += kSegFloats; pb += kSegFloats;
            }
        }
        total = __reduce_add_sync(0xffffffffu, total);
        if (lane == 0) P.counts[k] = total;
    }
}

}  // namespace igmk
