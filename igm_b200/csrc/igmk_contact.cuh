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
// Bound: FP32 CUDA-core issue (5.5 instructions per bead pair per structure
// with packed float32x2 arithmetic; 10 in scalar form), not HBM: every coordinate is read once per tile row/column from
// L2 and reused 32 times out of shared memory.  Tensor cores do not apply
// (3-term non-FMA float32 sums that must match NumPy bit for bit).
//
// CTA = 256 threads = 128 (8 x 16) pair-positions x 2 structure slices; each
// thread owns a 4 x 2 block of bead pairs (beads ta + 8 aa, tb + 16 bb) and, per
// staged chunk of 32 structures, the 16 structures of its slice.  (4 x 4 blocks with
// 4 slices need 128 registers and spill: 2.59 T vs 2.87 T bead-pair-structs/s.)
#pragma once
#include "igmk_device.cuh"

namespace igmk {

constexpr int kCtTile = 32;            // beads per tile side
constexpr int kCtThreads = 256;
constexpr int kCtStruct = 32;          // structures staged per step
constexpr int kCtRow = 3 * kCtStruct + 4;   // floats per bead in smem; /4 is odd -> conflict-free LDS.128
#ifndef IGMK_CT_MINB
#define IGMK_CT_MINB 2
#endif
#ifndef IGMK_CT_BB
#define IGMK_CT_BB 2
#endif
constexpr int kCtBB = IGMK_CT_BB;           // column beads per thread (row beads: 4)
constexpr int kCtTbN = kCtTile / kCtBB;     // column positions (8 or 16)
constexpr int kCtPos = 8 * kCtTbN;          // pair positions per structure slice
constexpr int kCtSlices = kCtThreads / kCtPos;       // structure slices (4 or 2)
constexpr int kCtGq = kCtStruct / kCtSlices / 4;     // float4 groups per slice and step
constexpr int kCtRcRow = kCtTile + 8;       // cut-off tile row: (ta * 40 + tb) % 32 distinct within a warp

struct ContactParams {
    const float* coords;   // [nbead][3][npad]
    const float* radii;    // [nbead]
    uint32_t* counts;      // [nrows][ncols]
    int nstruct, npad, nbead;
    int row0, nrows, col0, ncols;
    float contact_range;
    int strict;
    u64 negzero2;          // {-0.0f, -0.0f}, see f2sq()
    const HapEntry* hap;   // haploid mode: copies of every locus
    int haploid;           // 0: tile indices are beads, 1: haploid loci (copies summed)
};

// Full chunks: packed float32x2 arithmetic (two structures per FADD2 / FFMA2,
// sequentially rounded like the scalar form), contact flags as 1.0 / 0.0
// (FSET.BF) accumulated two per FADD2 - 5.5 instructions per bead pair and
// structure instead of 10.
template <bool STRICT>
__device__ __forceinline__ void contact_accumulate_packed(const float* sa, const float* sb, int ta, int tb,
                                                          int slice, u64 nz,
                                                          const float* s_rc, u64 (&cnt2)[4][kCtBB]) {
#pragma unroll
    for (int gq = 0; gq < kCtGq; ++gq) {
        const int s4 = slice * (4 * kCtGq) + gq * 4;       // first of 4 structures within the chunk
        // two row beads at a time (24 registers of row coordinates live)
#pragma unroll
        for (int ah = 0; ah < 4; ah += 2) {
            Row6 a[2];
#pragma unroll
            for (int a2 = 0; a2 < 2; ++a2) {
                const float* p = sa + (ta + 8 * (ah + a2)) * kCtRow + s4;
                lds_v2b64(p, a[a2].x01, a[a2].x23);
                lds_v2b64(p + kCtStruct, a[a2].y01, a[a2].y23);
                lds_v2b64(p + 2 * kCtStruct, a[a2].z01, a[a2].z23);
            }
#pragma unroll
            for (int bb = 0; bb < kCtBB; ++bb) {
                const float* p = sb + (tb + kCtTbN * bb) * kCtRow + s4;
                Row6 b;
                lds_v2b64(p, b.x01, b.x23);
                lds_v2b64(p + kCtStruct, b.y01, b.y23);
                lds_v2b64(p + 2 * kCtStruct, b.z01, b.z23);
#pragma unroll
                for (int a2 = 0; a2 < 2; ++a2) {
                    const int aa = ah + a2;
                    float d0, d1, d2, d3;
                    f2split(d2pair<0>(a[a2], b, nz), d0, d1);
                    f2split(d2pair<1>(a[a2], b, nz), d2, d3);
                    const float r = s_rc[(ta + 8 * aa) * kCtRcRow + tb + kCtTbN * bb];
#ifndef IGMK_CT_INTCOUNT
                    if (STRICT) {
                        cnt2[aa][bb] = f2add(cnt2[aa][bb], f2pack(f_lt_one(d0, r), f_lt_one(d1, r)));
                        cnt2[aa][bb] = f2add(cnt2[aa][bb], f2pack(f_lt_one(d2, r), f_lt_one(d3, r)));
                    } else {
                        cnt2[aa][bb] = f2add(cnt2[aa][bb], f2pack(f_le_one(d0, r), f_le_one(d1, r)));
                        cnt2[aa][bb] = f2add(cnt2[aa][bb], f2pack(f_le_one(d2, r), f_le_one(d3, r)));
                    }
#else
                    // integer flags on the ALU pipe: the FMA pipes carry only the 8 packed ops
                    int c = (int)cnt2[aa][bb];
                    if (STRICT) c += (d0 < r) + (d1 < r) + (d2 < r) + (d3 < r);
                    else        c += (d0 <= r) + (d1 <= r) + (d2 <= r) + (d3 <= r);
                    cnt2[aa][bb] = (u64)(unsigned)c;
#endif
                }
            }
        }
    }
}

// Bead-level tile: tile indices are bead ids.  The coordinate chunks are double-buffered:
// chunk c + 1 travels global -> shared memory by per-thread asynchronous 16-byte copies
// (cp.async / LDGSTS, zero fill for beads outside the tile) while chunk c is being computed on,
// so a CTA waits for L2 only once, in front of its first chunk.
#ifndef IGMK_CT_ASYNC
#define IGMK_CT_ASYNC 1
#endif
constexpr int kCtBufFloats = kCtTile * kCtRow;               // one side, one buffer
constexpr size_t kCtDynBytes = IGMK_CT_ASYNC ? (size_t)4 * kCtBufFloats * sizeof(float) : 0;   // [2 buffers][a | b]

__device__ __forceinline__ void cp_async16_zfill(float* dst_smem, const float* src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(src), "r"(n) : "memory");
}

__global__ void __launch_bounds__(kCtThreads, IGMK_CT_MINB)
contact_tile_kernel(const ContactParams P) {
#if IGMK_CT_ASYNC
    extern __shared__ __align__(16) float s_dyn_ct[];      // [2][a: kCtBufFloats | b: kCtBufFloats]
#else
    __shared__ __align__(16) float s_a[kCtTile * kCtRow];
    __shared__ __align__(16) float s_b[kCtTile * kCtRow];
#endif
    __shared__ uint32_t s_cnt[kCtTile * kCtTile];
    __shared__ float s_rc[kCtTile * kCtRcRow];

    const int t = threadIdx.x;
    const int pos = t % kCtPos, slice = t / kCtPos;
    const int ta = pos / kCtTbN, tb = pos % kCtTbN;
    const int a_base = P.row0 + blockIdx.y * kCtTile;
    const int b_base = P.col0 + blockIdx.x * kCtTile;
    const int a_end = P.row0 + P.nrows, b_end = P.col0 + P.ncols;

    for (int e = t; e < kCtTile * kCtTile; e += kCtThreads) s_cnt[e] = 0u;

    // squared cut-offs of the tile (float32, as the A-step computes rcutsq)
    for (int e = t; e < kCtTile * kCtTile; e += kCtThreads) {
        const int a = a_base + (e >> 5), b = b_base + (e & 31);
        const float ra = (a < a_end) ? __ldg(P.radii + a) : 0.f;
        const float rb = (b < b_end) ? __ldg(P.radii + b) : 0.f;
        const float r = __fmul_rn(P.contact_range, __fadd_rn(ra, rb));
        s_rc[(e >> 5) * kCtRcRow + (e & 31)] = __fmul_rn(r, r);
    }
    u64 cnt2[4][kCtBB];        // {even, odd structures} hit counts as floats (exact: < 2^24)
#pragma unroll
    for (int aa = 0; aa < 4; ++aa)
#pragma unroll
        for (int bb = 0; bb < kCtBB; ++bb) cnt2[aa][bb] = 0ull;

    const size_t row = (size_t)3 * P.npad;
#if IGMK_CT_ASYNC
    // this thread's three 16-byte pieces of a chunk: (bead, component, 4 structures)
    auto issue = [&](int s0, int buf) {
        float* sa = s_dyn_ct + (size_t)buf * 2 * kCtBufFloats;
        float* sb = sa + kCtBufFloats;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int f = t + kCtThreads * k;          // 0 .. 767
            const int bead = f / 24, rem = f - bead * 24;
            const int comp = rem >> 3, v4 = rem & 7;
            const int a = a_base + bead, b = b_base + bead;
            const size_t so = coord_off(s0 + 4 * v4) + (size_t)comp * kSeg;
            const int dst = bead * kCtRow + comp * kCtStruct + 4 * v4;
            cp_async16_zfill(sa + dst, (a < a_end) ? P.coords + (size_t)a * row + so : P.coords, a < a_end);
            cp_async16_zfill(sb + dst, (b < b_end) ? P.coords + (size_t)b * row + so : P.coords, b < b_end);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int nchunk = (P.nstruct + kCtStruct - 1) / kCtStruct;
    issue(0, 0);
    for (int c = 0; c < nchunk; ++c) {
        const int s0 = c * kCtStruct;
        float* sa = s_dyn_ct + (size_t)(c & 1) * 2 * kCtBufFloats;
        const float* sb = sa + kCtBufFloats;
        if (c + 1 < nchunk) {
            issue(s0 + kCtStruct, (c + 1) & 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            // structures past the end of the population: NaN on the row side, so their d2 is
            // NaN and never counts (the padding in HBM is zero); each thread patches the
            // pieces it copied itself
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int f = t + kCtThreads * k;
                const int bead = f / 24, rem = f - bead * 24;
                const int comp = rem >> 3, v4 = rem & 7;
                const int sv = s0 + 4 * v4;
                if (sv + 3 >= P.nstruct) {
                    const float qn = __int_as_float(0x7fffffff);
                    float* q = sa + bead * kCtRow + comp * kCtStruct + 4 * v4;
                    if (sv >= P.nstruct) q[0] = qn;
                    if (sv + 1 >= P.nstruct) q[1] = qn;
                    if (sv + 2 >= P.nstruct) q[2] = qn;
                    q[3] = qn;
                }
            }
        }
        __syncthreads();
        if (P.strict) contact_accumulate_packed<true>(sa, sb, ta, tb, slice, P.negzero2, s_rc, cnt2);
        else          contact_accumulate_packed<false>(sa, sb, ta, tb, slice, P.negzero2, s_rc, cnt2);
        __syncthreads();                                   // buffer (c & 1) is refilled by the next iteration's issue
    }
#else
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
            float4 va = (a < a_end)
                ? __ldg(reinterpret_cast<const float4*>(P.coords + (size_t)a * row + so)) : z;
            // structures past the end of the population: NaN on the row side, so
            // their d2 is NaN and never counts (the padding in HBM is zero)
            const int sv = s0 + 4 * v4;
            if (sv + 3 >= P.nstruct) {
                const float qn = __int_as_float(0x7fffffff);
                if (sv >= P.nstruct) va.x = qn;
                if (sv + 1 >= P.nstruct) va.y = qn;
                if (sv + 2 >= P.nstruct) va.z = qn;
                va.w = qn;
            }
            const float4 vb = (b < b_end)
                ? __ldg(reinterpret_cast<const float4*>(P.coords + (size_t)b * row + so)) : z;
            *reinterpret_cast<float4*>(s_a + bead * kCtRow + comp * kCtStruct + 4 * v4) = va;
            *reinterpret_cast<float4*>(s_b + bead * kCtRow + comp * kCtStruct + 4 * v4) = vb;
        }
        __syncthreads();
        if (P.strict) contact_accumulate_packed<true>(s_a, s_b, ta, tb, slice, P.negzero2, s_rc, cnt2);
        else          contact_accumulate_packed<false>(s_a, s_b, ta, tb, slice, P.negzero2, s_rc, cnt2);
    }
#endif

    // combine the 4 structure slices
#pragma unroll
    for (int aa = 0; aa < 4; ++aa)
#pragma unroll
        for (int bb = 0; bb < kCtBB; ++bb)
        {
#ifndef IGMK_CT_INTCOUNT
            float c_lo, c_hi;
            f2split(cnt2[aa][bb], c_lo, c_hi);
            atomicAdd(&s_cnt[(ta + 8 * aa) * kCtTile + (tb + kCtTbN * bb)], (uint32_t)(int)(c_lo + c_hi));
#else
            atomicAdd(&s_cnt[(ta + 8 * aa) * kCtTile + (tb + kCtTbN * bb)], (uint32_t)cnt2[aa][bb]);
#endif
        }
    __syncthreads();
    for (int e = t; e < kCtTile * kCtTile; e += kCtThreads) {
        const int a = a_base + (e >> 5), b = b_base + (e & 31);
        if (a < a_end && b < b_end)
            P.counts[(size_t)(a - P.row0) * P.ncols + (b - P.col0)] = s_cnt[e];
    }
}

// Haploid variant: tile indices are haploid loci and the counts of all copy combinations of a locus pair are summed
// (Contactmatrix.sumCopies() after buildContactMap, HicEvaluationStep.py:109-111):
//   counts[i][j] = sum over a in copies(i), b in copies(j) of #{s : d2_s(a,b) <= rc2(a,b)}
__global__ void __launch_bounds__(kCtThreads, IGMK_CT_MINB)
contact_tile_hap_kernel(const ContactParams P) {
    constexpr bool HAP = true;
#if IGMK_CT_ASYNC
    extern __shared__ __align__(16) float s_dyn_ct[];      // [2][a: kCtBufFloats | b: kCtBufFloats]
#else
    __shared__ __align__(16) float s_a[kCtTile * kCtRow];
    __shared__ __align__(16) float s_b[kCtTile * kCtRow];
#endif
    __shared__ uint32_t s_cnt[kCtTile * kCtTile];
    __shared__ float s_rc[kCtTile * kCtRcRow];
    __shared__ int s_ida[2][kCtTile], s_idb[2][kCtTile];     // bead id per copy, -1: absent

    const int t = threadIdx.x;
    const int pos = t % kCtPos, slice = t / kCtPos;
    const int ta = pos / kCtTbN, tb = pos % kCtTbN;
    const int a_base = P.row0 + blockIdx.y * kCtTile;
    const int b_base = P.col0 + blockIdx.x * kCtTile;
    const int a_end = P.row0 + P.nrows, b_end = P.col0 + P.ncols;

    for (int e = t; e < kCtTile * kCtTile; e += kCtThreads) s_cnt[e] = 0u;
    if (HAP && t < 2 * kCtTile) {
        const int side = t >> 5, k = t & 31;
        const int idx = (side ? b_base : a_base) + k;
        const bool in = idx < (side ? b_end : a_end);
        int id0 = -1, id1 = -1;
        if (in) {
            const int4 h = __ldg(reinterpret_cast<const int4*>(P.hap + idx));
            id0 = h.x; id1 = h.y;
        }
        if (side) { s_idb[0][k] = id0; s_idb[1][k] = id1; }
        else      { s_ida[0][k] = id0; s_ida[1][k] = id1; }
    }
    u64 cnt2[4][kCtBB];        // {even, odd structures} hit counts as floats (exact: < 2^24)
#pragma unroll
    for (int aa = 0; aa < 4; ++aa)
#pragma unroll
        for (int bb = 0; bb < kCtBB; ++bb) cnt2[aa][bb] = 0ull;

    const size_t row = (size_t)3 * P.npad;
    constexpr int ncopy = HAP ? 2 : 1;
    const float qn = __int_as_float(0x7fffffff);
    for (int ca = 0; ca < ncopy; ++ca)
    for (int cb = 0; cb < ncopy; ++cb) {
        __syncthreads();                                   // ids visible / previous combination done
        // squared cut-offs of the tile (float32, as the A-step computes rcutsq)
        for (int e = t; e < kCtTile * kCtTile; e += kCtThreads) {
            int a, b;
            if (HAP) { a = s_ida[ca][e >> 5]; b = s_idb[cb][e & 31]; }
            else {
                a = a_base + (e >> 5); b = b_base + (e & 31);
                a = (a < a_end) ? a : -1; b = (b < b_end) ? b : -1;
            }
            const float ra = (a >= 0) ? __ldg(P.radii + a) : 0.f;
            const float rb = (b >= 0) ? __ldg(P.radii + b) : 0.f;
            const float r = __fmul_rn(P.contact_range, __fadd_rn(ra, rb));
            s_rc[(e >> 5) * kCtRcRow + (e & 31)] = __fmul_rn(r, r);
        }
#if IGMK_CT_ASYNC
        // double-buffered staging as in contact_tile_kernel; absent beads (outside the tile /
        // haploid locus without a second copy) are zero-filled by the copy and overwritten with
        // NaN behind the wait, like the structures past the end of the population
        auto issue = [&](int s0, int buf) {
            float* sa = s_dyn_ct + (size_t)buf * 2 * kCtBufFloats;
            float* sb = sa + kCtBufFloats;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int f = t + kCtThreads * k;          // 0 .. 767
                const int bead = f / 24, rem = f - bead * 24;
                const int comp = rem >> 3, v4 = rem & 7;
                const int a = s_ida[ca][bead], b = s_idb[cb][bead];
                const size_t so = coord_off(s0 + 4 * v4) + (size_t)comp * kSeg;
                const int dst = bead * kCtRow + comp * kCtStruct + 4 * v4;
                cp_async16_zfill(sa + dst, (a >= 0) ? P.coords + (size_t)a * row + so : P.coords, a >= 0);
                cp_async16_zfill(sb + dst, (b >= 0) ? P.coords + (size_t)b * row + so : P.coords, b >= 0);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        const int nchunk = (P.nstruct + kCtStruct - 1) / kCtStruct;
        __syncthreads();                                   // (ids visible before the first issue)
        issue(0, 0);
        for (int c = 0; c < nchunk; ++c) {
            const int s0 = c * kCtStruct;
            float* sa = s_dyn_ct + (size_t)(c & 1) * 2 * kCtBufFloats;
            float* sb = sa + kCtBufFloats;
            if (c + 1 < nchunk) {
                issue(s0 + kCtStruct, (c + 1) & 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {                  // each thread patches the pieces it copied itself
                const int f = t + kCtThreads * k;
                const int bead = f / 24, rem = f - bead * 24;
                const int comp = rem >> 3, v4 = rem & 7;
                const int dst = bead * kCtRow + comp * kCtStruct + 4 * v4;
                const int sv = s0 + 4 * v4;
                if (s_ida[ca][bead] < 0) {
                    *reinterpret_cast<float4*>(sa + dst) = make_float4(qn, qn, qn, qn);
                } else if (sv + 3 >= P.nstruct) {
                    if (sv >= P.nstruct) sa[dst] = qn;
                    if (sv + 1 >= P.nstruct) sa[dst + 1] = qn;
                    if (sv + 2 >= P.nstruct) sa[dst + 2] = qn;
                    sa[dst + 3] = qn;
                }
                if (s_idb[cb][bead] < 0) *reinterpret_cast<float4*>(sb + dst) = make_float4(qn, qn, qn, qn);
            }
            __syncthreads();
            if (P.strict) contact_accumulate_packed<true>(sa, sb, ta, tb, slice, P.negzero2, s_rc, cnt2);
            else          contact_accumulate_packed<false>(sa, sb, ta, tb, slice, P.negzero2, s_rc, cnt2);
            __syncthreads();                               // buffer (c & 1) is refilled by the next issue; s_rc by the next combination
        }
    }
#else
        for (int s0 = 0; s0 < P.nstruct; s0 += kCtStruct) {
            __syncthreads();
            // stage 32 beads x 3 components x 32 structures of each side
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int f = t + kCtThreads * k;          // 0 .. 767
                const int bead = f / 24, rem = f - bead * 24;
                const int comp = rem >> 3, v4 = rem & 7;
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                int a, b;
                if (HAP) { a = s_ida[ca][bead]; b = s_idb[cb][bead]; }
                else {
                    a = a_base + bead; b = b_base + bead;
                    a = (a < a_end) ? a : -1; b = (b < b_end) ? b : -1;
                }
                const size_t so = coord_off(s0 + 4 * v4) + (size_t)comp * kSeg;
                // absent beads (outside the tile / haploid locus without a second copy)
                // and structures past the end of the population: NaN on the row side,
                // so their d2 is NaN and never counts (the padding in HBM is zero)
                float4 va = (a >= 0)
                    ? __ldg(reinterpret_cast<const float4*>(P.coords + (size_t)a * row + so))
                    : make_float4(qn, qn, qn, qn);
                const int sv = s0 + 4 * v4;
                if (sv + 3 >= P.nstruct) {
                    if (sv >= P.nstruct) va.x = qn;
                    if (sv + 1 >= P.nstruct) va.y = qn;
                    if (sv + 2 >= P.nstruct) va.z = qn;
                    va.w = qn;
                }
                const float4 vb = (b >= 0)
                    ? __ldg(reinterpret_cast<const float4*>(P.coords + (size_t)b * row + so))
                    : (HAP ? make_float4(qn, qn, qn, qn) : z);
                *reinterpret_cast<float4*>(s_a + bead * kCtRow + comp * kCtStruct + 4 * v4) = va;
                *reinterpret_cast<float4*>(s_b + bead * kCtRow + comp * kCtStruct + 4 * v4) = vb;
            }
            __syncthreads();
            if (P.strict) contact_accumulate_packed<true>(s_a, s_b, ta, tb, slice, P.negzero2, s_rc, cnt2);
            else          contact_accumulate_packed<false>(s_a, s_b, ta, tb, slice, P.negzero2, s_rc, cnt2);
        }
    }

#endif

    // combine the 4 structure slices
#pragma unroll
    for (int aa = 0; aa < 4; ++aa)
#pragma unroll
        for (int bb = 0; bb < kCtBB; ++bb)
        {
#ifndef IGMK_CT_INTCOUNT
            float c_lo, c_hi;
            f2split(cnt2[aa][bb], c_lo, c_hi);
            atomicAdd(&s_cnt[(ta + 8 * aa) * kCtTile + (tb + kCtTbN * bb)], (uint32_t)(int)(c_lo + c_hi));
#else
            atomicAdd(&s_cnt[(ta + 8 * aa) * kCtTile + (tb + kCtTbN * bb)], (uint32_t)cnt2[aa][bb]);
#endif
        }
    __syncthreads();
    for (int e = t; e < kCtTile * kCtTile; e += kCtThreads) {
        const int a = a_base + (e >> 5), b = b_base + (e & 31);
        if (a < a_end && b < b_end)
            P.counts[(size_t)(a - P.row0) * P.ncols + (b - P.col0)] = s_cnt[e];
    }
}

}  // namespace igmk
