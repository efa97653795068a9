// igmk_actdist.cuh - K1: the Hi-C activation-distance kernels (sm_100a).
//
// Replaces the per-pair NumPy body of get_actdist
// (igm/steps/ActivationDistanceStep.py:405-473; GP flavour
// igm/steps/GP_activation.py:367-424).
//
// Each candidate pair is owned by a group of G threads (G = 32: one warp per
// pair, no barriers at all, for nstruct <= 1024; G = one CTA of up to 512
// threads for larger populations, two CTAs resident per SM so that one pair's
// memory phase overlaps the other's select phase):
//   1. fill: every thread streams its float4 chunks (4 structures each) of the
//      <= 4 bead rows with 128-bit loads and computes the copy-combination d2
//      values with PACKED float32x2 arithmetic (FADD2 / FFMA2: two structures
//      per instruction, sequentially rounded - bit-identical to NumPy), counts
//      d2 <= rcutsq, and keeps only the HIGH 16 BITS of each kept d2, two per
//      word (bf16x2), parked in the group's shared-memory key array.
//      d2 >= 0, so bf16 order == float order.
//   2. p and the order-statistic index o are evaluated in float64.
//   3. select: bisection on the 16-bit key interval [kmin, kmax]; a pass streams
//      the keys back from shared memory (LDS.128 = 8 keys) with one packed
//      compare + one packed add per TWO elements and one group reduction,
//      until <= 32 candidates remain or the interval is a single key.
//   4. finish: the locations of the candidates are compacted into a short list,
//      re-materialised in full float32 from the coordinates by the whole group
//      in parallel, and ranked exactly by one warp.  The selected value is an
//      actual element, bit-exact.
//   No sort, no shared-memory histogram, no atomics in the main loops.
#pragma once
#include "igmk_device.cuh"

namespace igmk {

constexpr int kRankCap = 32;        // bisection stops at <= kRankCap candidates
constexpr int kWarpListCap = 64;    // candidate list words per warp group
constexpr int kBlockListCap = 1024; // candidate list words per CTA group
constexpr int kMaxQuads = 64;       // key quads per thread: bf16 pass counters stay exact

// ------------------------------------------------------------------ groups
// Shared scratch is addressed through 32-bit shared-window addresses.
template <bool BLOCK>
struct Group {
    int tid;            // thread index inside the group
    int nthr;           // group size (multiple of 32)
    uint32_t kscr;      // this thread's first key quad
    uint32_t kstride;   // bytes between consecutive quads of one thread (nthr * 16)
    uint32_t list;      // candidate list (codes, then values)
    uint32_t ctl;       // candidate counter
    int cap;            // list capacity
    uint32_t red;       // BLOCK: [2][3][32] words
    uint32_t list2;     // BLOCK: short list of the final <= kRankCap candidates
    int parity;

    __device__ __forceinline__ int sum(int x) {
        if (!BLOCK) return __reduce_add_sync(0xffffffffu, x);
        const int lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
        const uint32_t r = red + parity * 384;
        parity ^= 1;
        const int w = __reduce_add_sync(0xffffffffu, x);
        if (lane == 0) sts32(r + warp * 4, (uint32_t)w);
        __syncthreads();
        const int v = (lane < nw) ? (int)lds32(r + lane * 4) : 0;
        return __reduce_add_sync(0xffffffffu, v);
    }
    __device__ __forceinline__ void sum_min_max(int& s, uint32_t& mn, uint32_t& mx) {
        s = __reduce_add_sync(0xffffffffu, s);
        mn = __reduce_min_sync(0xffffffffu, mn);
        mx = __reduce_max_sync(0xffffffffu, mx);
        if (!BLOCK) return;
        const int lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
        const uint32_t r = red + parity * 384;
        parity ^= 1;
        if (lane == 0) {
            sts32(r + warp * 4, (uint32_t)s);
            sts32(r + 128 + warp * 4, mn);
            sts32(r + 256 + warp * 4, mx);
        }
        __syncthreads();
        const int vs = (lane < nw) ? (int)lds32(r + lane * 4) : 0;
        const uint32_t vmn = (lane < nw) ? lds32(r + 128 + lane * 4) : 0xffffffffu;
        const uint32_t vmx = (lane < nw) ? lds32(r + 256 + lane * 4) : 0u;
        s = __reduce_add_sync(0xffffffffu, vs);
        mn = __reduce_min_sync(0xffffffffu, vmn);
        mx = __reduce_max_sync(0xffffffffu, vmx);
    }
    // BLOCK only: exclusive prefix sum of x over the group's threads (one barrier).
    __device__ __forceinline__ int exclusive_scan(int x) {
        const int lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
        int incl = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        const uint32_t r = red + parity * 384;
        parity ^= 1;
        if (lane == 31) sts32(r + warp * 4, (uint32_t)incl);
        __syncthreads();
        int vs = (lane < nw) ? (int)lds32(r + lane * 4) : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, vs, d);
            if (lane >= d) vs += y;
        }
        const int below = __shfl_sync(0xffffffffu, vs, (warp > 0) ? warp - 1 : 0);
        return ((warp > 0) ? below : 0) + incl - x;
    }
    __device__ __forceinline__ void sync() {
        if (BLOCK) __syncthreads(); else __syncwarp();
    }
    __device__ __forceinline__ bool leader_warp() const { return !BLOCK || tid < 32; }
};

// --------------------------------------------------------- stage 1: fill keys
// Key quad q = v * NH + h of a thread (v: chunk index c = tid + v * nthr, h: slot
// pair) = { key(slot 2h, s0) | key(slot 2h, s1) << 16, (slot 2h: s2, s3),
//           (slot 2h+1: s0, s1), (slot 2h+1: s2, s3) },  s = structures 4c..4c+3;
// NaN pattern (0x7fff) for everything not kept.  NH = 2 when 4 values per
// structure are kept, else 1.
//
// Pair shapes (uniform over the group) get their own straight-line code:
//   SH_FULL4   LB, both loci diploid, inter-chromosomal: 4 combinations kept
//   SH_INTRA2  LB, both loci diploid, intra-chromosomal: (a0,b0),(a1,b1)
//   SH_GP4     GP, both loci diploid: 2 smallest of the 4 combinations
//   SH_GENERIC anything with a haploid locus (male X/Y): branch-free selects
enum : int { SH_FULL4 = 0, SH_INTRA2 = 1, SH_GP4 = 2, SH_GENERIC = 3 };

__device__ __forceinline__ int pair_shape(const PairDesc& d, int mode) {
    if (d.a1 >= 0 && d.b1 >= 0) {
        if (mode == IGMK_MODE_GP) return SH_GP4;
        return (d.cmask == 15) ? SH_FULL4 : SH_INTRA2;
    }
    return SH_GENERIC;
}

struct PairPtrs { const float *A0, *A1, *B0, *B1; };

__device__ __forceinline__ PairPtrs pair_ptrs(const ActdistParams& P, const PairDesc& d) {
    const size_t row = (size_t)3 * P.npad;
    PairPtrs pp;
    pp.A0 = P.coords + (size_t)d.a0 * row;
    pp.B0 = P.coords + (size_t)d.b0 * row;
    pp.A1 = P.coords + (size_t)(d.a1 >= 0 ? d.a1 : d.a0) * row;
    pp.B1 = P.coords + (size_t)(d.b1 >= 0 ? d.b1 : d.b0) * row;
    return pp;
}

// Four structures (4c .. 4c+3) x the kept copy-combination values of one pair shape,
// from the four coordinate row chunks: s[q][slot], NaN where nothing is kept.
template <int SH, int NS>
__device__ __forceinline__ void chunk_values(const PairDesc& d, int mode, const Row6& a0, const Row6& a1,
                                             const Row6& b0, const Row6& b1, u64 nz, float (&s)[4][NS]) {
    const float qnan = __int_as_float(0x7fffffff);
    if (SH == SH_INTRA2) {
        f2split(d2pair<0>(a0, b0, nz), s[0][0], s[1][0]);
        f2split(d2pair<1>(a0, b0, nz), s[2][0], s[3][0]);
        f2split(d2pair<0>(a1, b1, nz), s[0][1], s[1][1]);
        f2split(d2pair<1>(a1, b1, nz), s[2][1], s[3][1]);
    } else {
        float e[4][4];   // [q][combination d0..d3]
        f2split(d2pair<0>(a0, b0, nz), e[0][0], e[1][0]);
        f2split(d2pair<1>(a0, b0, nz), e[2][0], e[3][0]);
        f2split(d2pair<0>(a0, b1, nz), e[0][1], e[1][1]);
        f2split(d2pair<1>(a0, b1, nz), e[2][1], e[3][1]);
        f2split(d2pair<0>(a1, b0, nz), e[0][2], e[1][2]);
        f2split(d2pair<1>(a1, b0, nz), e[2][2], e[3][2]);
        f2split(d2pair<0>(a1, b1, nz), e[0][3], e[1][3]);
        f2split(d2pair<1>(a1, b1, nz), e[2][3], e[3][3]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (SH == SH_FULL4) {
#pragma unroll
                for (int k = 0; k < NS; ++k) s[q][k] = e[q][k];
            } else if (SH == SH_GP4) {
                const float lo1 = fminf(e[q][0], e[q][1]), hi1 = fmaxf(e[q][0], e[q][1]);
                const float lo2 = fminf(e[q][2], e[q][3]), hi2 = fmaxf(e[q][2], e[q][3]);
                s[q][0] = fminf(lo1, lo2);
                s[q][1] = fminf(fmaxf(lo1, lo2), fminf(hi1, hi2));
            } else {
                const float e1 = (d.cmask & CM_D1) ? e[q][1] : qnan;
                const float e2 = (d.cmask & CM_D2) ? e[q][2] : qnan;
                const float e3 = (d.cmask & CM_D3) ? e[q][3] : qnan;
                float t[4];
                pack_slots(d, mode, e[q][0], e1, e2, e3, t);
#pragma unroll
                for (int k = 0; k < NS; ++k) s[q][k] = t[k];
            }
        }
    }
}

// AS: the rows of locus i come from the CTA's shared-memory tile (as0 / as1 =
// shared addresses of the a0 / a1 rows, same layout as in HBM).
template <int SH, bool AS>
__device__ __forceinline__ void fill_keys(const ActdistParams& P, const PairDesc& d,
                                          const PairPtrs& pp, int tid, int nthr, int V,
                                          uint32_t kscr, uint32_t kstride, uint32_t as0, uint32_t as1,
                                          int& cnt, uint32_t& mn2, uint32_t& mx2) {
    constexpr int NS = (SH == SH_FULL4) ? 4 : (SH == SH_INTRA2 || SH == SH_GP4) ? 2 : 4;
    const float qnan = __int_as_float(0x7fffffff);
    const float rc = d.rcutsq;
    const u64 nz = P.negzero2;
    // chunk c = tid + v * nthr lives in segment c >> 5 at lane offset (c & 31) * 4;
    // nthr is a multiple of 32, so only the segment advances with v.
    const size_t off0 = (size_t)(tid >> 5) * kSegFloats + (size_t)(tid & 31) * 4;
    const size_t vstride = (size_t)(nthr >> 5) * kSegFloats;
    const float* pa0 = pp.A0 + off0;
    const float* pb0 = pp.B0 + off0;
    const float* pa1 = pp.A1 + off0;
    const float* pb1 = pp.B1 + off0;
    uint32_t sa0 = as0 + (uint32_t)off0 * 4u, sa1 = as1 + (uint32_t)off0 * 4u;
    u64 c_local = 0ull;            // {count of even, count of odd structures} as floats
    uint32_t lmn = 0x7fff7fffu, lmx = 0x7fff7fffu;
    // in generic pairs only NH is not known at compile time
    const int nh = (NS == 4) ? ((d.keep > 2) ? 2 : 1) : 1;
    uint32_t dst = kscr;

    // The chunk loop is deliberately NOT unrolled (instruction-cache footprint and
    // register pressure: 48 registers of loaded coordinates are live here).
#pragma unroll 1
    for (int v = 0; v < V; ++v) {
        const int c = tid + v * nthr;
        uint32_t nk[NS][2];
        if (c >= P.nchunks) {              // padding chunk: NaN keys, nothing to load
            sts128(dst, 0x7fff7fffu, 0x7fff7fffu, 0x7fff7fffu, 0x7fff7fffu);
            if (NS == 4) {
                if (SH == SH_FULL4 || nh == 2) sts128(dst + kstride, 0x7fff7fffu, 0x7fff7fffu, 0x7fff7fffu, 0x7fff7fffu);
            }
            dst += (uint32_t)nh * kstride;
            continue;                      // (pointers are not used again: c only grows)
        }
        float s[4][NS];   // [q][slot]
        {
            Row6 a0, a1;
            const Row6 b0 = load_row6<IGMK_JLOAD>(pb0), b1 = load_row6<IGMK_JLOAD>(pb1);
            if (AS) {
                a0 = load_row6_shared(sa0); a1 = load_row6_shared(sa1);
            } else {
                a0 = load_row6<LD_KEEP>(pa0); a1 = load_row6<LD_KEEP>(pa1);
            }
            chunk_values<SH, NS>(d, P.mode, a0, a1, b0, b1, nz, s);
        }
        if (4 * c + 4 > P.nstruct) {        // tail chunk of the population (one thread)
#pragma unroll
            for (int q = 1; q < 4; ++q)
                if (4 * c + q >= P.nstruct) {
#pragma unroll
                    for (int k = 0; k < NS; ++k) s[q][k] = qnan;
                }
        }
        // contact count: 1.0 / 0.0 flags accumulated two per FADD2 (exact: < 2^24)
#pragma unroll
        for (int q = 0; q < 4; q += 2)
#pragma unroll
            for (int k = 0; k < NS; ++k)
                c_local = f2add(c_local, f2pack(f_le_one(s[q][k], rc), f_le_one(s[q + 1][k], rc)));
#pragma unroll
        for (int k = 0; k < NS; ++k) {
#pragma unroll
            for (int qh = 0; qh < 2; ++qh) {
                // {hi16(s[2qh]), hi16(s[2qh+1])}: bytes 2,3 of each -> one PRMT
                nk[k][qh] = __byte_perm(__float_as_uint(s[2 * qh][k]),
                                        __float_as_uint(s[2 * qh + 1][k]), 0x7632);
            }
            lmn = bf2_min(lmn, bf2_min(nk[k][0], nk[k][1]));   // NaN halves are ignored
            lmx = bf2_max(lmx, bf2_max(nk[k][0], nk[k][1]));
        }
        sts128(dst, nk[0][0], nk[0][1], nk[1][0], nk[1][1]);
        if (NS == 4) {
            if (SH == SH_FULL4 || nh == 2) sts128(dst + kstride, nk[2][0], nk[2][1], nk[3][0], nk[3][1]);
        }
        dst += (uint32_t)nh * kstride;
        pa0 += vstride; pb0 += vstride; pa1 += vstride; pb1 += vstride;
        sa0 += (uint32_t)vstride * 4u; sa1 += (uint32_t)vstride * 4u;
    }
    float c_lo, c_hi;
    f2split(c_local, c_lo, c_hi);
    cnt = (int)(c_lo + c_hi);
    mn2 = lmn;
    mx2 = lmx;
}

// One bisection pass over this thread's nq key quads: #{key <= pivot}.
// set.le yields 0xffff / 0 per half; subtracting the masks as 32-bit integers
// (two per IADD3) leaves  acc = (#hi - #lo) * 65536 + #lo.
template <int NQ>
__device__ __forceinline__ int count_le_fixed(uint32_t kscr, uint32_t kstride, uint32_t piv2) {
    uint32_t a0 = 0u, a1 = 0u;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        uint32_t k0, k1, k2, k3;
        lds128(kscr + (uint32_t)q * kstride, k0, k1, k2, k3);
        a0 = a0 - bf2_le_mask(k0, piv2) - bf2_le_mask(k1, piv2);
        a1 = a1 - bf2_le_mask(k2, piv2) - bf2_le_mask(k3, piv2);
    }
    const uint32_t acc = a0 + a1;
    return ((int)acc >> 16) + 2 * (int)(acc & 0xffffu);
}
__device__ __forceinline__ int count_le(uint32_t kscr, uint32_t kstride, int nq, uint32_t piv2) {
    if (nq == 8) return count_le_fixed<8>(kscr, kstride, piv2);      // nstruct in (896, 1024], 1-2 kept
    if (nq == 16) return count_le_fixed<16>(kscr, kstride, piv2);    // same, 4 kept
    uint32_t a0 = 0u, a1 = 0u;
#pragma unroll 4
    for (int q = 0; q < nq; ++q) {
        uint32_t k0, k1, k2, k3;
        lds128(kscr + (uint32_t)q * kstride, k0, k1, k2, k3);
        a0 = a0 - bf2_le_mask(k0, piv2) - bf2_le_mask(k1, piv2);
        a1 = a1 - bf2_le_mask(k2, piv2) - bf2_le_mask(k3, piv2);
    }
    const uint32_t acc = a0 + a1;
    return ((int)acc >> 16) + 2 * (int)(acc & 0xffffu);
}

// Calls f(e) for every element of this thread whose key lies in [lo, hi];
// e = q * 8 + r * 2 + half  (quad q, word r, half-word).  Matches of four quads
// are first collected in one bitmap word (bit w / 16 + w <-> low / high half of
// word w = 4 (q - q0) + r), so the divergent part costs one trip per match, not
// one per quad.
template <bool EQ, class F>
__device__ __forceinline__ void scan_range(uint32_t kscr, uint32_t kstride, int nq,
                                           uint32_t lo2, uint32_t hi2, F&& f) {
#pragma unroll 1
    for (int q0 = 0; q0 < nq; q0 += 4) {
        uint32_t bm = 0u;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
            if (q0 + qq < nq) {            // uniform over the group
                uint32_t k[4];
                lds128(kscr + (uint32_t)(q0 + qq) * kstride, k[0], k[1], k[2], k[3]);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const uint32_t m = EQ ? bf2_eq_mask(k[r], lo2)
                                          : (bf2_ge_mask(k[r], lo2) & bf2_le_mask(k[r], hi2));
                    bm |= m & (0x00010001u << (qq * 4 + r));
                }
            }
        }
        while (bm) {
            const int bit = __ffs(bm) - 1;
            bm &= bm - 1;
            const int w = bit & 15;
            f((q0 + (w >> 2)) * 8 + (w & 3) * 2 + (bit >> 4));
        }
    }
}
template <class F>
__device__ __forceinline__ void scan_range(uint32_t kscr, uint32_t kstride, int nq,
                                           uint32_t lo2, uint32_t hi2, F&& f) {
    if (lo2 == hi2) scan_range<true>(kscr, kstride, nq, lo2, hi2, f);
    else            scan_range<false>(kscr, kstride, nq, lo2, hi2, f);
}

// Full float32 pattern of one kept value, re-materialised from the coordinates
// (deliberately out of line).  LB: slot -> one combination.
__device__ __noinline__ uint32_t recompute_lb(const float* pa, const float* pb, int st) {
    return __float_as_uint(d2_scalar(pa, pb, st));
}
// GP: all combinations of the structure, then the kept slot.
__device__ __noinline__ uint32_t recompute_gp(const float* A0, const float* A1, const float* B0,
                                              const float* B1, int cmask, int keep, int st,
                                              int slot) {
    const float qnan = __int_as_float(0x7fffffff);
    PairDesc d;
    d.cmask = cmask; d.keep = keep;
    const float d0 = d2_scalar(A0, B0, st);
    const float d1 = (cmask & CM_D1) ? d2_scalar(A0, B1, st) : qnan;
    const float d2 = (cmask & CM_D2) ? d2_scalar(A1, B0, st) : qnan;
    const float d3 = (cmask & CM_D3) ? d2_scalar(A1, B1, st) : qnan;
    float s[4];
    pack_slots(d, IGMK_MODE_GP, d0, d1, d2, d3, s);
    return __float_as_uint(slot == 0 ? s[0] : s[1]);
}

// element e (scan_range numbering) of thread `owner`
__device__ __forceinline__ uint32_t element_bits(const PairPtrs& pp, const PairDesc& d, int mode,
                                                 int owner, int nthr, int nh, int e) {
    const int q = e >> 3, r = (e >> 1) & 3, half = e & 1;
    const int v = (nh == 2) ? (q >> 1) : q, h = (nh == 2) ? (q & 1) : 0;
    const int slot = 2 * h + (r >> 1), qh = r & 1;
    const int st = 4 * (owner + v * nthr) + 2 * qh + half;
    if (mode == IGMK_MODE_GP)
        return recompute_gp(pp.A0, pp.A1, pp.B0, pp.B1, d.cmask, d.keep, st, slot);
    // LB: the slot-th existing combination (enumeration order d0..d3)
    int cm = d.cmask;
    for (int t = 0; t < slot; ++t) cm &= cm - 1;
    const int comb = __ffs(cm) - 1;
    return recompute_lb((comb & 2) ? pp.A1 : pp.A0, (comb & 1) ? pp.B1 : pp.B0, st);
}

// r-th smallest (0-based) of the n <= 32 * k list words, by one full warp.
__device__ __forceinline__ uint32_t warp_select(uint32_t list, int n, int r, int lane) {
    if (n <= 32) {
        const uint32_t x = (lane < n) ? lds32(list + lane * 4) : 0xffffffffu;
        int rank = 0;
        for (int t = 0; t < n; ++t) {
            const uint32_t y = __shfl_sync(0xffffffffu, x, t);
            rank += (y < x || (y == x && t < lane)) ? 1 : 0;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, lane < n && rank == r);
        // hit is non-empty by construction; the guard keeps a corrupted input from hanging
        const int src = hit ? (__ffs(hit) - 1) : 0;
        return __shfl_sync(0xffffffffu, x, src);
    }
    // most-significant-bit-first binary radix select over the list
    uint32_t prefix = 0u, mask = 0u;
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t b = 1u << bit;
        int c0 = 0;
        for (int t = lane; t < n; t += 32) {
            const uint32_t x = lds32(list + t * 4);
            c0 += ((x & mask) == prefix && !(x & b)) ? 1 : 0;
        }
        c0 = __reduce_add_sync(0xffffffffu, c0);
        if (r >= c0) { r -= c0; prefix |= b; }
        mask |= b;
    }
    return prefix;
}

// ------------------------------------------------------------- one pair
// ---------------------------------------------------- locus-i tile (warp kernel)
// Two shared-memory slots per CTA, each holding the coordinate rows (both copies) of
// one locus i, so the warps that work on the consecutive pairs (i, j1), (i, j2), ...
// read them from shared memory instead of L2.  No warp ever waits: a warp whose locus
// is not resident either claims a free slot and loads the rows itself (one extra
// shared-memory round trip for that one pair) or simply streams them from L2 as
// before.  Slot word = locus << 12 | state << 10 | users; every transition is a
// single-word atomic, so there is no ABA window.
// How a tile is filled (IGMK_TILE_BULK):
//   0  the claiming warp copies the rows itself (LDG.128 + STS.128), then uses the tile;
//   1  the claiming warp issues two bulk asynchronous copies and WAITS for them (measured
//      12 % slower than 0: a 24 KB bulk copy takes longer than 48 warp-wide load / store
//      rounds, and every warp of that locus streams from L2 meanwhile);
//   2  asynchronous: the claiming warp issues the bulk copies (TMA engine, completion on the
//      slot's mbarrier) and goes on at once, streaming its own pair from L2; the first warp
//      that wants the locus after the bytes have landed (non-blocking mbarrier probe) flips
//      the slot to READY.  Slot word = locus << 12 | state << 10 | generation << 6 | users;
//      generation g (mod 16) completes phase g & 1 of the mbarrier.
#ifndef IGMK_TILE_BULK
#define IGMK_TILE_BULK 0
#endif
enum : uint32_t { TS_EMPTY = 0u, TS_LOADING = 1u, TS_READY = 2u };
struct TileCtl {
    uint32_t base;        // shared address of slot 0 (0: tiles disabled)
    uint32_t slot_bytes;  // 2 rows * 12 * npad
    uint32_t words;       // shared address of the slot words
    int nslots;           // 1 or 2
    uint32_t bars;        // shared address of one mbarrier per slot (bulk copies of the rows complete on it)
    uint32_t pars;        // shared address of one word per slot: phase parity of the next load
};
// Shared control block of the tiles: slot words, phase parities, mbarriers.
struct __align__(8) TileShared {
    unsigned long long bar[2];
    uint32_t slot[2];
    uint32_t par[2];
};
__device__ __forceinline__ void tile_init(TileShared* ts) {   // one thread, before a CTA barrier
    ts->slot[0] = (0xfffffu << 12) | (0u << 10) | ((IGMK_TILE_BULK == 2) ? (15u << 6) : 0u);
    ts->slot[1] = (0xfffffu << 12) | (0u << 10) | ((IGMK_TILE_BULK == 2) ? (15u << 6) : 0u);
    ts->par[0] = 0u; ts->par[1] = 0u;
    mbar_init((uint32_t)__cvta_generic_to_shared(&ts->bar[0]), 1u);
    mbar_init((uint32_t)__cvta_generic_to_shared(&ts->bar[1]), 1u);
    mbar_fence_init();
}
__device__ __forceinline__ void tile_bind(TileShared* ts, TileCtl& tile) {
    tile.words = (uint32_t)__cvta_generic_to_shared(&ts->slot[0]);
    tile.pars = (uint32_t)__cvta_generic_to_shared(&ts->par[0]);
    tile.bars = (uint32_t)__cvta_generic_to_shared(&ts->bar[0]);
}
__device__ __forceinline__ uint32_t atoms_cas(uint32_t addr, uint32_t cmp, uint32_t val) {
    uint32_t old;
    asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(addr), "r"(cmp), "r"(val) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t lds32_volatile(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
#if IGMK_TILE_BULK == 2
__device__ __forceinline__ int tile_acquire(const ActdistParams& P, const TileCtl& tile, int i,
                                            const PairDesc& d, const PairPtrs& pp, int lane) {
    int found = -1;
    if (lane == 0) {
        const uint32_t want = (uint32_t)i;
        bool same_busy = false;
        for (int s = 0; s < tile.nslots && found < 0; ++s) {
            const uint32_t addr = tile.words + 4u * (uint32_t)s;
            uint32_t w = lds32_volatile(addr);
            for (int tries = 0; tries < 4; ++tries) {
                if ((w >> 12) != want) break;
                const uint32_t st = (w >> 10) & 3u;
                if (st == TS_READY) {
                    const uint32_t old = atoms_cas(addr, w, w + 1u);
                    if (old == w) { found = s; break; }
                    w = old;
                    continue;
                }
                if (st == TS_LOADING) {
                    // have the bytes of this generation landed?  (non-blocking probe)
                    if (mbar_test(tile.bars + 8u * (uint32_t)s, (w >> 6) & 1u)) {
                        const uint32_t nw = (w & ~(3u << 10)) | (TS_READY << 10) | 1u;   // first user
                        const uint32_t old = atoms_cas(addr, w, nw);
                        if (old == w) { found = s; break; }
                        w = old;
                        continue;
                    }
                    same_busy = true;
                }
                break;
            }
        }
        if (found < 0 && !same_busy) {
            // claim an idle slot (empty, ready without users, or loaded and never used) and
            // start the copy; this warp streams its own pair from L2
            for (int s = 0; s < tile.nslots; ++s) {
                const uint32_t addr = tile.words + 4u * (uint32_t)s;
                const uint32_t w = lds32_volatile(addr);
                const uint32_t st = (w >> 10) & 3u;
                if ((w & 0x3fu) != 0u) continue;
                if ((w >> 12) == want && st != TS_EMPTY) continue;
                const uint32_t bar = tile.bars + 8u * (uint32_t)s;
                if (st == TS_LOADING && !mbar_test(bar, (w >> 6) & 1u)) continue;
                const uint32_t gen = (((w >> 6) & 15u) + 1u) & 15u;
                const uint32_t nw = (want << 12) | (TS_LOADING << 10) | (gen << 6);
                if (atoms_cas(addr, w, nw) != w) continue;
                const uint32_t dst0 = tile.base + (uint32_t)s * tile.slot_bytes;
                const uint32_t rowb = tile.slot_bytes >> 1;
                mbar_expect_tx(bar, (d.a1 >= 0) ? 2u * rowb : rowb);
                bulk_g2s(dst0, pp.A0, rowb, bar);
                if (d.a1 >= 0) bulk_g2s(dst0 + rowb, pp.A1, rowb, bar);
                break;
            }
        }
    }
    found = __shfl_sync(0xffffffffu, found, 0);
    if (found >= 0) __threadfence_block();
    return found;
}
#else
__device__ __forceinline__ int tile_acquire(const ActdistParams& P, const TileCtl& tile, int i,
                                            const PairDesc& d, const PairPtrs& pp, int lane) {
    int found = -1, claimed = -1;
    if (lane == 0) {
        const uint32_t want = (uint32_t)i;
        uint32_t w[2];
        w[1] = (0xfffffu << 12) | (TS_LOADING << 10);      // a missing second slot never matches
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (s >= tile.nslots) break;
            w[s] = lds32_volatile(tile.words + 4u * s);
            for (int tries = 0; tries < 4 && found < 0; ++tries) {
                if ((w[s] >> 12) != want || ((w[s] >> 10) & 3u) != TS_READY) break;
                const uint32_t old = atoms_cas(tile.words + 4u * s, w[s], w[s] + 1u);
                if (old == w[s]) found = s; else w[s] = old;
            }
        }
        if (found < 0) {
            const bool busy_same = ((w[0] >> 12) == want && ((w[0] >> 10) & 3u) == TS_LOADING) ||
                                   ((w[1] >> 12) == want && ((w[1] >> 10) & 3u) == TS_LOADING);
            if (!busy_same) {
                // victim: an idle slot (no users, not loading); empty first, then the
                // smaller locus (CSR order: the older one)
                int order0 = 0;
                if (((w[0] >> 10) & 3u) != TS_EMPTY &&
                    (((w[1] >> 10) & 3u) == TS_EMPTY || (w[1] >> 12) < (w[0] >> 12))) order0 = 1;
#pragma unroll
                for (int t = 0; t < 2 && claimed < 0; ++t) {
                    const int s = order0 ^ t;
                    if (s >= tile.nslots) continue;
                    const uint32_t st = (w[s] >> 10) & 3u;
                    if ((w[s] & 0x3ffu) != 0u || st == TS_LOADING) continue;
                    if (st == TS_READY && (w[s] >> 12) == want) continue;
                    const uint32_t nw = (want << 12) | (TS_LOADING << 10);
                    if (atoms_cas(tile.words + 4u * s, w[s], nw) == w[s]) claimed = s;
                }
            }
        }
    }
    found = __shfl_sync(0xffffffffu, found, 0);
    claimed = __shfl_sync(0xffffffffu, claimed, 0);
    if (found >= 0) {
        __threadfence_block();               // acquire: the loader's stores are visible
        return found;
    }
    if (claimed < 0) return -1;
    const uint32_t dst0 = tile.base + (uint32_t)claimed * tile.slot_bytes;
    const uint32_t rowb = tile.slot_bytes >> 1;
#if IGMK_TILE_BULK
    // load both rows of locus i: two bulk asynchronous copies (TMA engine, 12 * npad bytes
    // each) issued by one lane, completing on the slot's mbarrier; the warp waits for the
    // bytes, then publishes the slot.  Nobody else touches a LOADING slot.
    const uint32_t bar = tile.bars + 8u * (uint32_t)claimed;
    const uint32_t par = lds32_volatile(tile.pars + 4u * (uint32_t)claimed);
    __syncwarp();
    if (lane == 0) {
        mbar_expect_tx(bar, (d.a1 >= 0) ? 2u * rowb : rowb);
        bulk_g2s(dst0, pp.A0, rowb, bar);
        if (d.a1 >= 0) bulk_g2s(dst0 + rowb, pp.A1, rowb, bar);
        sts32(tile.pars + 4u * (uint32_t)claimed, par ^ 1u);
    }
    mbar_wait(bar, par);
#else
    {   // global -> shared, 128-bit, whole warp
        const int n16 = (int)(rowb >> 4);
        const float4* r0 = reinterpret_cast<const float4*>(pp.A0);
        const float4* r1 = reinterpret_cast<const float4*>(pp.A1);
        for (int k = lane; k < n16; k += 32) {
            const float4 x0 = __ldg(r0 + k);
            sts128(dst0 + (uint32_t)k * 16u, __float_as_uint(x0.x), __float_as_uint(x0.y),
                   __float_as_uint(x0.z), __float_as_uint(x0.w));
        }
        if (d.a1 >= 0) {
            for (int k = lane; k < n16; k += 32) {
                const float4 x1 = __ldg(r1 + k);
                sts128(dst0 + rowb + (uint32_t)k * 16u, __float_as_uint(x1.x), __float_as_uint(x1.y),
                       __float_as_uint(x1.z), __float_as_uint(x1.w));
            }
        }
    }
#endif
    __threadfence_block();
    __syncwarp();
    if (lane == 0) {
        // publish with one user (this warp); nobody touches a LOADING word
        const uint32_t ready = ((uint32_t)i << 12) | (TS_READY << 10) | 1u;
        sts32(tile.words + 4u * claimed, ready);   // after the block-scope fence above
    }
    __syncwarp();
    return claimed;
}
#endif
__device__ __forceinline__ void tile_release(const TileCtl& tile, int slot, int lane) {
    __syncwarp();                            // every lane's tile reads are done
    if (lane == 0)
        asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(tile.words + 4u * slot), "r"(0xffffffffu) : "memory");
}

// Second half of a pair: p and o, bisection on the parked keys, candidate list, exact
// rank, result.  `cnt`, `kmin`, `kmax` are the group-wide values of the fill; the group's
// candidate counter (g.ctl) must be zero.
template <bool BLOCK, bool DAMID>
__device__ __forceinline__ void select_part(const ActdistParams& P, Group<BLOCK>& g, int V,
                                            long long pair, const PairDesc& d, const PairPtrs& pp,
                                            double extra, int cnt, uint32_t kmin, uint32_t kmax) {
    const int nh = (d.keep > 2) ? 2 : 1;
    const int nq = V * nh;
    double p;
    int o;                                 // ascending index of the element to select
    int o_rep;                             // index reported to the caller
    if (DAMID) {
        const int M = d.keep * P.nstruct;
        float pf;
        cnt = M - cnt;                     // #{d_sq >= 1}: fill counted s <= largest float32 below (R-r)^2
        compute_p_o_damid(cnt, M, __ldg(P.pexp32 + pair), __ldg(P.plast32 + pair), P.it_corr, pf, o_rep);
        p = (double)pf;
        o = (o_rep >= 0) ? M - 1 - o_rep : -1;     // d_sq[::-1].sort(): descending order statistic
    } else {
        compute_p_o(cnt, d.keep, P.nstruct, __ldg(P.pwish + pair), __ldg(P.plast + pair), P.it_corr, p, o);
        o_rep = o;
    }
    if (o < 0) {
        emit_result(P, g.tid, pair, d, 0u, cnt, -1, 0.0, extra);
        return;
    }

    // ---- bisection on the 16-bit keys (each thread reads only its own quads)
    // Warp groups narrow the key bracket down to <= kRankCap candidates.  CTA groups stop
    // at <= P.block_stop (a key pass over the whole population costs far more there than a
    // pass over a short list of re-materialised values) and finish the bisection on the
    // full float32 patterns of the list (below).
    uint32_t lo = kmin, hi = (kmax < kmin) ? kmin : kmax;
    int cb = 0, ch = d.keep * P.nstruct;
    const int stop = BLOCK ? P.block_stop : kRankCap;
    int cl_lo = 0, cl_hi = -1;             // BLOCK: this thread's own #{key < lo}, #{key <= hi}
    while (lo < hi && (ch - cb) > stop) {
        const uint32_t mid = (lo + hi) >> 1;
        const int cl = count_le(g.kscr, g.kstride, nq, mid | (mid << 16));
        const int c = g.sum(cl);
        if (c > o) { hi = mid; ch = c; cl_hi = cl; } else { lo = mid + 1; cb = c; cl_lo = cl; }
    }
    const uint32_t lo2 = lo | (lo << 16), hi2 = hi | (hi << 16);
    int n_list;

    if ((ch - cb) <= g.cap) {
        // ---- compact the candidates' locations, then re-materialise them in
        // full precision with the whole group in parallel
        if (BLOCK) {
            // no atomics: every thread knows how many candidates it owns from the passes
            // that set the bracket, so an exclusive scan gives it a private list range
            if (cl_hi < 0) cl_hi = count_le(g.kscr, g.kstride, nq, hi2);     // hi never moved (uniform)
            uint32_t pos = (uint32_t)g.exclusive_scan(cl_hi - cl_lo);
            scan_range(g.kscr, g.kstride, nq, lo2, hi2, [&](int e) {
                if (pos < (uint32_t)g.cap) sts32(g.list + pos * 4, ((uint32_t)g.tid << 16) | (uint32_t)e);
                ++pos;
            });
        } else {
            g.sync();                               // counter = 0 visible
            scan_range(g.kscr, g.kstride, nq, lo2, hi2, [&](int e) {
                const uint32_t slot = atoms_inc(g.ctl);
                if (slot < (uint32_t)g.cap) sts32(g.list + slot * 4, ((uint32_t)g.tid << 16) | (uint32_t)e);
            });
        }
        g.sync();
        n_list = ch - cb;
        for (int t = g.tid; t < n_list; t += g.nthr) {
            const uint32_t code = lds32(g.list + t * 4);
            sts32(g.list + t * 4, element_bits(pp, d, P.mode, (int)(code >> 16), g.nthr, nh, (int)(code & 0xffffu)));
        }
        g.sync();
    } else {
        // Single fat key h = lo == hi with more elements than the list holds:
        // bisect the low 16 bits among the elements of that key, re-materialising
        // them on every pass (degenerate inputs with very many equal distances).
        const int cb0 = cb;                 // elements with key < h
        uint32_t l2 = 0u, h2 = 0xffffu;
        while (l2 < h2 && (ch - cb) > g.cap) {
            const uint32_t m2 = (l2 + h2) >> 1;
            int c_loc = 0;
            scan_range(g.kscr, g.kstride, nq, lo2, hi2, [&](int e) {
                const uint32_t x = element_bits(pp, d, P.mode, g.tid, g.nthr, nh, e);
                c_loc += ((x & 0xffffu) <= m2) ? 1 : 0;
            });
            const int c = cb0 + g.sum(c_loc);
            if (c > o) { h2 = m2; ch = c; } else { l2 = m2 + 1; cb = c; }
        }
        const uint32_t vlo = (lo << 16) | l2, vhi = (lo << 16) | h2;
        if ((ch - cb) > g.cap) {
            // l2 == h2: every remaining candidate has the same bit pattern.
            emit_result(P, g.tid, pair, d, vlo, cnt, o_rep, p, extra);
            return;
        }
        g.sync();
        scan_range(g.kscr, g.kstride, nq, lo2, hi2, [&](int e) {
            const uint32_t x = element_bits(pp, d, P.mode, g.tid, g.nthr, nh, e);
            if (x >= vlo && x <= vhi) {
                const uint32_t slot = atoms_inc(g.ctl);
                if (slot < (uint32_t)g.cap) sts32(g.list + slot * 4, x);
            }
        });
        g.sync();
        n_list = ch - cb;
    }

    int r = o - cb;                        // rank of the answer inside the list
    uint32_t sel = g.list;
    if (BLOCK && n_list > kRankCap) {
        // ---- CTA groups: bisection on the full 32-bit patterns of the list (a pass reads
        // n_list / nthr words per thread) until <= kRankCap values remain
        uint32_t vlo = 0u, vhi = 0x7f800000u;
        int nb = 0, nh = n_list;
        // all list values lie in [lo << 16, hi << 16 | 0xffff] except in the fat-key branch,
        // where the bracket is narrower still; the key bracket is a valid start for both
        vlo = lo << 16; vhi = (hi << 16) | 0xffffu;
        while (vlo < vhi && (nh - nb) > kRankCap) {
            const uint32_t mid = vlo + ((vhi - vlo) >> 1);
            int c = 0;
            for (int t = g.tid; t < n_list; t += g.nthr) c += (lds32(g.list + t * 4) <= mid) ? 1 : 0;
            c = g.sum(c);
            if (c > r) { vhi = mid; nh = c; } else { vlo = mid + 1; nb = c; }
        }
        if ((nh - nb) > kRankCap) {
            // vlo == vhi: every remaining value has the same bit pattern
            emit_result(P, g.tid, pair, d, vlo, cnt, o_rep, p, extra);
            return;
        }
        // survivors -> short list (<= kRankCap shared atomics)
        if (g.tid == 0) sts32(g.ctl, 0u);
        g.sync();
        for (int t = g.tid; t < n_list; t += g.nthr) {
            const uint32_t x = lds32(g.list + t * 4);
            if (x >= vlo && x <= vhi) {
                const uint32_t slot = atoms_inc(g.ctl);
                if (slot < (uint32_t)kRankCap) sts32(g.list2 + slot * 4, x);
            }
        }
        g.sync();
        sel = g.list2;
        r -= nb;
        n_list = nh - nb;
    }

    // ---- exact rank inside one warp: the r-th smallest candidate
    if (g.leader_warp()) {
        const uint32_t ans = warp_select(sel, n_list, r, g.tid & 31);
        emit_result(P, g.tid, pair, d, ans, cnt, o_rep, p, extra);
    }
}

// DAMID: the same machinery computes the lamina-DamID activation distance of one
// locus (igm/steps/DamidActivationDistanceStep.py:375-470, spherical envelope): the
// "pair" is (locus, origin) - the partner row is the all-zero bead kept behind the
// population, so d2 is the float32 sum of squares of the coordinates, exactly
// np.sum(np.square(x), axis=1) - the order statistic is taken in descending order and
// the probability arithmetic is float32 (see compute_p_o_damid).
template <bool BLOCK, bool DAMID>
__device__ __forceinline__ void process_pair(const ActdistParams& P, Group<BLOCK>& g, int V,
                                             long long slot, const TileCtl& tile) {
    // slot = position in processing order; perm maps it to the pair's index in the
    // caller's list (results stay in input order)
    const long long pair = P.perm ? (long long)__ldg(P.perm + slot) : slot;
    const int i = __ldg(P.pi + pair);
    double extra = 0.0;                    // DAMID: (R - r)^2, handed to the finish pass
    const PairDesc d = DAMID ? make_damid_desc(P, i, extra) : make_pair_desc(P, i, __ldg(P.pj + pair));
    if (!d.valid) {                        // uniform over the group
        emit_empty(P, g.tid, pair);
        return;
    }
    const PairPtrs pp = pair_ptrs(P, d);

    int cnt;
    uint32_t mn2, mx2;
    const int tslot = (!BLOCK && !DAMID && tile.base) ? tile_acquire(P, tile, i, d, pp, g.tid) : -1;
    if (tslot >= 0) {
        const uint32_t as0 = tile.base + (uint32_t)tslot * tile.slot_bytes;
        const uint32_t as1 = (d.a1 >= 0) ? as0 + (tile.slot_bytes >> 1) : as0;
        switch (pair_shape(d, P.mode)) {       // uniform over the group
            case SH_FULL4:  fill_keys<SH_FULL4, true>(P, d, pp, g.tid, g.nthr, V, g.kscr, g.kstride, as0, as1, cnt, mn2, mx2); break;
            case SH_INTRA2: fill_keys<SH_INTRA2, true>(P, d, pp, g.tid, g.nthr, V, g.kscr, g.kstride, as0, as1, cnt, mn2, mx2); break;
            case SH_GP4:    fill_keys<SH_GP4, true>(P, d, pp, g.tid, g.nthr, V, g.kscr, g.kstride, as0, as1, cnt, mn2, mx2); break;
            default:        fill_keys<SH_GENERIC, true>(P, d, pp, g.tid, g.nthr, V, g.kscr, g.kstride, as0, as1, cnt, mn2, mx2); break;
        }
        tile_release(tile, tslot, g.tid);
    } else {
        switch (pair_shape(d, P.mode)) {       // uniform over the group
            case SH_FULL4:  fill_keys<SH_FULL4, false>(P, d, pp, g.tid, g.nthr, V, g.kscr, g.kstride, 0u, 0u, cnt, mn2, mx2); break;
            case SH_INTRA2: fill_keys<SH_INTRA2, false>(P, d, pp, g.tid, g.nthr, V, g.kscr, g.kstride, 0u, 0u, cnt, mn2, mx2); break;
            case SH_GP4:    fill_keys<SH_GP4, false>(P, d, pp, g.tid, g.nthr, V, g.kscr, g.kstride, 0u, 0u, cnt, mn2, mx2); break;
            default:        fill_keys<SH_GENERIC, false>(P, d, pp, g.tid, g.nthr, V, g.kscr, g.kstride, 0u, 0u, cnt, mn2, mx2); break;
        }
    }
    uint32_t kmin = min(mn2 & 0xffffu, mn2 >> 16);
    uint32_t mxl = mx2 & 0xffffu, mxh = mx2 >> 16;
    mxl = (mxl > 0x7f80u) ? 0u : mxl;
    mxh = (mxh > 0x7f80u) ? 0u : mxh;
    uint32_t kmax = max(mxl, mxh);
    if (g.tid == 0) sts32(g.ctl, 0u);
    g.sum_min_max(cnt, kmin, kmax);
    select_part<BLOCK, DAMID>(P, g, V, pair, d, pp, extra, cnt, kmin, kmax);
}

// ---------------------------------------------------------------- kernels
// G = 32: one pair per warp, kWarpsPerBlock independent warps per CTA, no CTA
// barrier anywhere.  Consecutive pairs (CSR order: same locus i) go to the warps
// of one CTA, so the rows of locus i are shared through L1.
#ifndef IGMK_WPB
#define IGMK_WPB 20
#endif
#ifndef IGMK_MINB
#define IGMK_MINB 1
#endif
constexpr int kWarpsPerBlock = IGMK_WPB;

template <bool DAMID>
__global__ void __launch_bounds__(32 * kWarpsPerBlock, IGMK_MINB)
actdist_warp_kernel(const ActdistParams P, const int V) {
    extern __shared__ uint4 s_keys[];             // [warp][2 V][32] key quads, then 2 locus-i tiles
    __shared__ uint32_t s_list[kWarpsPerBlock][kWarpListCap];
    __shared__ uint32_t s_cnt[kWarpsPerBlock];
    __shared__ TileShared s_tile;
    __shared__ unsigned int s_ticket;
    __shared__ unsigned long long s_gblock[8];
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;           // <= kWarpsPerBlock (fewer when V is large)
    const long long n_pairs = P.n_pairs_dev ? (long long)__ldg(P.n_pairs_dev) : P.n_pairs;   // redo launch: count on the device
    Group<false> g;
    g.tid = threadIdx.x & 31;
    g.nthr = 32;
    g.list = smem_addr(&s_list[warp][0]);
    g.ctl = smem_addr(&s_cnt[warp]);
    g.cap = kWarpListCap;
    g.kscr = smem_addr(s_keys) + (uint32_t)(warp * 2 * V * 32 + g.tid) * 16u;
    g.kstride = 32u * 16u;
    g.red = 0u;
    g.list2 = 0u;
    g.parity = 0;
    TileCtl tile;
    tile.base = 0u; tile.slot_bytes = 0u; tile.nslots = P.tile_slots;
    tile_bind(&s_tile, tile);
    if (P.tile_block > 0) {
        tile.base = smem_addr(s_keys) + (uint32_t)nwarps * 2u * (uint32_t)V * 512u;
        tile.slot_bytes = 24u * (uint32_t)P.npad;
        if (threadIdx.x == 0) {
            tile_init(&s_tile);
            s_ticket = 0u;
        }
        if (threadIdx.x < 8) s_gblock[threadIdx.x] = 0ull;
        __syncthreads();
        // CTA-contiguous blocks of tile_block pairs (block k of this CTA = list block
        // blockIdx + k * gridDim), handed to the warps one pair at a time by a
        // shared ticket counter: consecutive pairs share locus i, and fast / slow
        // pairs balance out across the warps.
        const unsigned int B = (unsigned int)P.tile_block;
        if (P.block_counter) {
            // dynamic variant: the CTA's k-th block is whatever list block the device-wide
            // counter hands out next (fetched by the warp that draws the block's first
            // ticket, published to the other warps through one 64-bit word per slot:
            // (k + 1) << 32 | block), so CTAs that drew cheap blocks simply take more of them
            for (;;) {
                unsigned int t = 0u;
                if (g.tid == 0) t = atomicAdd(&s_ticket, 1u);
                t = __shfl_sync(0xffffffffu, t, 0);
                const unsigned int k = t / B, r = t - k * B;
                unsigned int gb = 0u;
                if (g.tid == 0) {
                    volatile unsigned long long* w = &s_gblock[k & 7u];
                    if (r == 0u) {
                        gb = atomicAdd(P.block_counter, 1u);
                        *w = ((unsigned long long)(k + 1u) << 32) | gb;
                    } else {
                        unsigned long long x;
                        int spins = 0;
                        while ((unsigned int)((x = *w) >> 32) != k + 1u) {
                            __nanosleep(20);
                            if (++spins > (1 << 24)) __trap();      // protocol error: never hang the GPU
                        }
                        gb = (unsigned int)x;
                    }
                }
                gb = __shfl_sync(0xffffffffu, gb, 0);
                const long long base = (long long)gb * B;
                if (base >= n_pairs) break;
                const long long pair = base + r;
                if (pair < n_pairs) process_pair<false, DAMID>(P, g, V, pair, tile);
                __syncwarp();
            }
            return;
        }
        for (;;) {
            unsigned int t = 0u;
            if (g.tid == 0) t = atomicAdd(&s_ticket, 1u);
            t = __shfl_sync(0xffffffffu, t, 0);
            const unsigned int k = t / B, r = t - k * B;
            const long long base = ((long long)blockIdx.x + (long long)k * gridDim.x) * B;
            if (base >= n_pairs) break;
            const long long pair = base + r;
            if (pair < n_pairs) process_pair<false, DAMID>(P, g, V, pair, tile);
            __syncwarp();
        }
        return;
    }
    const long long stride = (long long)gridDim.x * nwarps;
    for (long long pair = (long long)blockIdx.x * nwarps + warp; pair < n_pairs;
         pair += stride) {
        process_pair<false, DAMID>(P, g, V, pair, tile);
        __syncwarp();
    }
}

// G = blockDim.x (multiple of 32, <= MAXT): one pair per CTA, two CTAs per SM.
// MAXT = 320 leaves 96 registers per thread (all 12 row loads of a chunk in
// flight at once); MAXT = 512 (64 registers) is for very large populations.
template <int MAXT, bool DAMID>
__global__ void __launch_bounds__(MAXT, 2)
actdist_block_kernel(const ActdistParams P, const int V) {
    extern __shared__ uint4 s_keys[];             // [2 V][blockDim] key quads
    __shared__ uint32_t s_list[kBlockListCap];
    __shared__ uint32_t s_cnt;
    __shared__ uint32_t s_red[2 * 96];
    __shared__ uint32_t s_list2[kRankCap];
    Group<true> g;
    g.tid = threadIdx.x;
    g.nthr = blockDim.x;
    g.list = smem_addr(s_list);
    g.ctl = smem_addr(&s_cnt);
    g.cap = kBlockListCap;
    g.kscr = smem_addr(s_keys) + (uint32_t)threadIdx.x * 16u;
    g.kstride = (uint32_t)blockDim.x * 16u;
    g.red = smem_addr(s_red);
    g.list2 = smem_addr(s_list2);
    g.parity = 0;
    const long long n_pairs = P.n_pairs_dev ? (long long)__ldg(P.n_pairs_dev) : P.n_pairs;
    TileCtl tile;
    tile.base = 0u; tile.slot_bytes = 0u; tile.words = 0u; tile.nslots = 0; tile.bars = 0u; tile.pars = 0u;
    if (P.block_counter) {
        // pairs handed out by a device-wide counter (list order): the CTAs stay together in
        // the list and share the J-block's L2 residency instead of drifting apart
        __shared__ unsigned int s_next;
        for (;;) {
            if (threadIdx.x == 0) s_next = atomicAdd(P.block_counter, 1u);
            __syncthreads();
            const long long pair = (long long)s_next;
            if (pair >= n_pairs) break;
            process_pair<true, DAMID>(P, g, V, pair, tile);
            __syncthreads();
        }
        return;
    }
    for (long long pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        process_pair<true, DAMID>(P, g, V, pair, tile);
        __syncthreads();
    }
}

// ------------------------------------------------------- cross-check kernel
// Straightforward version kept as an on-device cross-check (IGMK_ALGO_SIMPLE):
// one CTA per pair, all kept d2 values in shared memory (scalar non-FMA
// arithmetic), 32-pass most-significant-bit-first binary radix select on the
// raw float32 patterns.
__global__ void __launch_bounds__(256)
actdist_simple_kernel(const ActdistParams P) {
    extern __shared__ uint32_t s_val[];          // keep * nstruct values
    __shared__ int s_red[2][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int parity = 0;
    auto block_sum = [&](int x) -> int {
        const int w = __reduce_add_sync(0xffffffffu, x);
        if (lane == 0) s_red[parity][warp] = w;
        __syncthreads();
        const int v = (lane < 8) ? s_red[parity][lane] : 0;
        parity ^= 1;
        return __reduce_add_sync(0xffffffffu, v);
    };
    for (long long pair = blockIdx.x; pair < P.n_pairs; pair += gridDim.x) {
        __syncthreads();
        const int i = __ldg(P.pi + pair), j = __ldg(P.pj + pair);
        const PairDesc d = make_pair_desc(P, i, j);
        igmk_pair_result* out = P.out + pair;
        if (!d.valid) {
            if (tid == 0) write_empty(out);
            continue;
        }
        const double pwish = __ldg(P.pwish + pair), plast = __ldg(P.plast + pair);
        const int N = P.nstruct;
        int c_loc = 0;
        for (int st = tid; st < N; st += blockDim.x) {
            float s[4];
            struct_slots(P, d, st, s);
            for (int k = 0; k < d.keep; ++k) {
                const float val = (k == 0) ? s[0] : (k == 1) ? s[1] : (k == 2) ? s[2] : s[3];
                s_val[k * N + st] = __float_as_uint(val);
                c_loc += (val <= d.rcutsq) ? 1 : 0;
            }
        }
        const int cnt = block_sum(c_loc);
        double p;
        int o;
        compute_p_o(cnt, d.keep, N, pwish, plast, P.it_corr, p, o);
        if (o < 0) {
            if (tid == 0) write_result(out, d, 0u, cnt, -1, 0.0);
            continue;
        }
        const int M = d.keep * N;
        uint32_t prefix = 0u, mask = 0u;
        int r = o;
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t b = 1u << bit;
            int c0 = 0;
            for (int e = tid; e < M; e += blockDim.x) {
                const uint32_t x = s_val[e];
                c0 += ((x & mask) == prefix && !(x & b)) ? 1 : 0;
            }
            c0 = block_sum(c0);
            if (r >= c0) { r -= c0; prefix |= b; }
            mask |= b;
        }
        if (tid == 0) write_result(out, d, prefix, cnt, o, p);
    }
}

// ------------------------------------------------------ selected-element index
// sel_flat_idx of SURVEY.md 8b/8d: the index, in d_sq[0:npc].ravel() (row * N + structure),
// of the element whose value is the selected order statistic - the LOWEST such index when
// several elements tie (NumPy's sort is not stable, so the reference defines no particular
// one).  d_sq is column-sorted at that point in BOTH modes (d_sq.sort(axis=0), :439), so the
// row of a value is its rank among the copy-combination values of its structure.  One warp per
// pair re-computes the pair's values and keeps the smallest matching index; -1 when the
// pair has no record.  An optional second pass: the A-step itself never needs the index.
template <int SH>
__device__ __forceinline__ int sel_index_scan(const ActdistParams& P, const PairDesc& d, const PairPtrs& pp,
                                              int lane, uint32_t want) {
    constexpr int NS = (SH == SH_FULL4) ? 4 : (SH == SH_INTRA2 || SH == SH_GP4) ? 2 : 4;
    int best = 0x7fffffff;
    for (int c = lane; c < P.nchunks; c += 32) {
        const size_t off = (size_t)(c >> 5) * kSegFloats + (size_t)(c & 31) * 4;
        float s[4][NS];
        {
            const Row6 a0 = load_row6<LD_PLAIN>(pp.A0 + off), a1 = load_row6<LD_PLAIN>(pp.A1 + off);
            const Row6 b0 = load_row6<LD_PLAIN>(pp.B0 + off), b1 = load_row6<LD_PLAIN>(pp.B1 + off);
            chunk_values<SH, NS>(d, P.mode, a0, a1, b0, b1, P.negzero2, s);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int st = 4 * c + q;
            if (st >= P.nstruct) continue;
            int rank = 0;
            bool hit = false;
#pragma unroll
            for (int k = 0; k < NS; ++k)
                if (k < d.keep) {
                    hit = hit || (__float_as_uint(s[q][k]) == want);
                    rank += (s[q][k] < __uint_as_float(want)) ? 1 : 0;      // NaN (absent) never counts
                }
            if (hit) best = min(best, rank * P.nstruct + st);
        }
    }
    return __reduce_min_sync(0xffffffffu, best);
}

__global__ void __launch_bounds__(256)
sel_index_kernel(const ActdistParams P, const igmk_pair_result* __restrict__ res, int32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (long long pair = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); pair < P.n_pairs;
         pair += (long long)gridDim.x * wpb) {
        const int o = res[pair].o;
        int idx = -1;
        if (o >= 0) {
            const PairDesc d = make_pair_desc(P, __ldg(P.pi + pair), __ldg(P.pj + pair));
            if (d.valid) {
                const PairPtrs pp = pair_ptrs(P, d);
                const uint32_t want = res[pair].d2_sel_bits;
                switch (pair_shape(d, P.mode)) {
                    case SH_FULL4:  idx = sel_index_scan<SH_FULL4>(P, d, pp, lane, want); break;
                    case SH_INTRA2: idx = sel_index_scan<SH_INTRA2>(P, d, pp, lane, want); break;
                    case SH_GP4:    idx = sel_index_scan<SH_GP4>(P, d, pp, lane, want); break;
                    default:        idx = sel_index_scan<SH_GENERIC>(P, d, pp, lane, want); break;
                }
                if (idx == 0x7fffffff) idx = -1;
            }
        }
        if (lane == 0) out[pair] = idx;
    }
}

}  // namespace igmk
