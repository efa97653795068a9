// igmk_actdist.cuh - K1: the Hi-C activation-distance kernels (sm_100a).
//
// Replaces the per-pair NumPy body of get_actdist
// (igm/steps/ActivationDistanceStep.py:405-473; GP flavour
// igm/steps/GP_activation.py:367-424).
//
// Fast kernel, per candidate pair handled by a group of G threads
// (G = 32: one warp per pair for nstruct <= 512; G = blockDim for larger
// populations, 64 threads at nstruct = 1000, 640 at 10 000):
//   1. every thread streams V <= 4 float4 chunks (4 structures each) of the
//      <= 4 bead rows with 128-bit loads, computes the copy-combination d2
//      values in registers (non-FMA float32, bit-identical to NumPy), counts
//      d2 <= rcutsq, and keeps only the HIGH 16 BITS of each kept d2, two per
//      register (bf16x2).  d2 >= 0, so bf16 order == float order.
//   2. p and the order-statistic index o are evaluated in float64.
//   3. the o-th smallest value is located by bisection on the 16-bit key
//      interval [kmin, kmax]: one packed compare + one packed add per TWO
//      elements per pass, one group reduction per pass, until <= 32 candidates
//      remain (or the interval is a single key).
//   4. the few candidates are re-materialised in full float32 precision from
//      the coordinates and ranked exactly inside one warp.
//   No sort, no shared-memory histogram, no atomics in the main loop.
#pragma once
#include "igmk_device.cuh"

namespace igmk {

constexpr int kCandCap = 32;

// ------------------------------------------------------------------ groups
// Shared scratch is addressed through 32-bit shared-window addresses.
struct WarpGroup {
    int tid;            // lane
    int nthr;           // 32
    uint32_t cand;      // shared address of kCandCap words (per warp)
    uint32_t cand_cnt;  // shared address of the candidate counter (per warp)
    uint32_t kscr;      // shared address of this thread's key scratch column
    uint32_t kstride;   // bytes between consecutive key quads of one thread

    __device__ __forceinline__ int sum(int x) { return __reduce_add_sync(0xffffffffu, x); }
    __device__ __forceinline__ void sum_min_max(int& s, uint32_t& mn, uint32_t& mx) {
        s = __reduce_add_sync(0xffffffffu, s);
        mn = __reduce_min_sync(0xffffffffu, mn);
        mx = __reduce_max_sync(0xffffffffu, mx);
    }
    __device__ __forceinline__ void sync() { __syncwarp(); }
    __device__ __forceinline__ bool leader_warp() const { return true; }
};

struct BlockGroup {
    int tid;
    int nthr;
    uint32_t cand;
    uint32_t cand_cnt;
    uint32_t kscr;
    uint32_t kstride;
    uint32_t red;       // shared address of [2][3][32] words
    int parity;

    __device__ __forceinline__ int sum(int x) {
        const int lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
        const uint32_t r = red + parity * 384;
        parity ^= 1;
        const int w = __reduce_add_sync(0xffffffffu, x);
        if (lane == 0) sts32(r + warp * 4, (uint32_t)w);
        __syncthreads();
        if (nw == 2) return w + (int)lds32(r + (warp ^ 1) * 4);
        const int v = (lane < nw) ? (int)lds32(r + lane * 4) : 0;
        return __reduce_add_sync(0xffffffffu, v);
    }
    __device__ __forceinline__ void sum_min_max(int& s, uint32_t& mn, uint32_t& mx) {
        const int lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
        const uint32_t r = red + parity * 384;
        parity ^= 1;
        const int ws = __reduce_add_sync(0xffffffffu, s);
        const uint32_t wmn = __reduce_min_sync(0xffffffffu, mn);
        const uint32_t wmx = __reduce_max_sync(0xffffffffu, mx);
        if (lane == 0) {
            sts32(r + warp * 4, (uint32_t)ws);
            sts32(r + 128 + warp * 4, wmn);
            sts32(r + 256 + warp * 4, wmx);
        }
        __syncthreads();
        const int vs = (lane < nw) ? (int)lds32(r + lane * 4) : 0;
        const uint32_t vmn = (lane < nw) ? lds32(r + 128 + lane * 4) : 0xffffffffu;
        const uint32_t vmx = (lane < nw) ? lds32(r + 256 + lane * 4) : 0u;
        s = __reduce_add_sync(0xffffffffu, vs);
        mn = __reduce_min_sync(0xffffffffu, vmn);
        mx = __reduce_max_sync(0xffffffffu, vmx);
    }
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ bool leader_warp() const { return tid < 32; }
};

// --------------------------------------------------------- stage 1: fill keys
// keys[v][slot][qh]: low half = structure 4c + 2qh, high half = 4c + 2qh + 1 of
// chunk c = tid + v * nthr; NaN pattern (0x7fff) for everything not kept.
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

struct Row3 { float4 x, y, z; };

// x / y / z of 4 consecutive structures: three 128-bit loads at constant offsets
__device__ __forceinline__ Row3 load_row3(const float* p) {
    Row3 r;
    r.x = __ldg(reinterpret_cast<const float4*>(p));
    r.y = __ldg(reinterpret_cast<const float4*>(p + kSeg));
    r.z = __ldg(reinterpret_cast<const float4*>(p + 2 * kSeg));
    return r;
}

__device__ __forceinline__ float f4get(const float4& v, int q) {
    return q == 0 ? v.x : q == 1 ? v.y : q == 2 ? v.z : v.w;
}

__device__ __forceinline__ float d2q(const Row3& a, const Row3& b, int q) {
    return d2_nofma(f4get(a.x, q), f4get(a.y, q), f4get(a.z, q),
                    f4get(b.x, q), f4get(b.y, q), f4get(b.z, q));
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

template <int V, int SH>
__device__ __forceinline__ void fill_keys(const ActdistParams& P, const PairDesc& d,
                                          const PairPtrs& pp, int tid, int nthr,
                                          uint32_t kscr, uint32_t kstride, int& cnt) {
    constexpr int NS = (SH == SH_FULL4) ? 4 : (SH == SH_INTRA2 || SH == SH_GP4) ? 2 : 4;
    const float qnan = __int_as_float(0x7fffffff);
    const float rc = d.rcutsq;
    // chunk c = tid + v * nthr lives in segment c >> 5 at lane offset (c & 31) * 4;
    // nthr is a multiple of 32, so only the segment advances with v.
    const size_t off0 = (size_t)(tid >> 5) * kSegFloats + (size_t)(tid & 31) * 4;
    const size_t vstride = (size_t)(nthr >> 5) * kSegFloats;
    const float* pa0 = pp.A0 + off0;
    const float* pb0 = pp.B0 + off0;
    const float* pa1 = pp.A1 + off0;
    const float* pb1 = pp.B1 + off0;
    int c_local = 0;

    // The chunk loop is deliberately NOT unrolled (instruction-cache footprint and
    // register pressure: 48 registers of loaded coordinates are live here).  The
    // packed keys of each chunk are parked in this thread's private column of
    // shared memory (indexable, unlike registers) and pulled back into registers
    // once after the loop, for the bisection passes.
#pragma unroll 1
    for (int v = 0; v < V; ++v) {
        const int c = tid + v * nthr;
        uint32_t nk[NS][2];
        if (c < P.nchunks) {
            const Row3 a0 = load_row3(pa0);
            const Row3 b0 = load_row3(pb0);
            const Row3 a1 = load_row3(pa1);
            const Row3 b1 = load_row3(pb1);
            float s[4][NS];   // [q][slot]
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (SH == SH_FULL4) {
                    s[q][0] = d2q(a0, b0, q);
                    s[q][1] = d2q(a0, b1, q);
                    s[q][2] = d2q(a1, b0, q);
                    s[q][3] = d2q(a1, b1, q);
                } else if (SH == SH_INTRA2) {
                    s[q][0] = d2q(a0, b0, q);
                    s[q][1] = d2q(a1, b1, q);
                } else if (SH == SH_GP4) {
                    const float e0 = d2q(a0, b0, q), e1 = d2q(a0, b1, q);
                    const float e2 = d2q(a1, b0, q), e3 = d2q(a1, b1, q);
                    const float lo1 = fminf(e0, e1), hi1 = fmaxf(e0, e1);
                    const float lo2 = fminf(e2, e3), hi2 = fmaxf(e2, e3);
                    s[q][0] = fminf(lo1, lo2);
                    s[q][1] = fminf(fmaxf(lo1, lo2), fminf(hi1, hi2));
                } else {
                    const float e0 = d2q(a0, b0, q);
                    const float e1 = (d.cmask & CM_D1) ? d2q(a0, b1, q) : qnan;
                    const float e2 = (d.cmask & CM_D2) ? d2q(a1, b0, q) : qnan;
                    const float e3 = (d.cmask & CM_D3) ? d2q(a1, b1, q) : qnan;
                    float t[4];
                    pack_slots(d, P.mode, e0, e1, e2, e3, t);
#pragma unroll
                    for (int k = 0; k < NS; ++k) s[q][k] = t[k];
                }
            }
            if (4 * c + 4 > P.nstruct) {        // tail chunk of the population (one thread)
#pragma unroll
                for (int q = 1; q < 4; ++q)
                    if (4 * c + q >= P.nstruct) {
#pragma unroll
                        for (int k = 0; k < NS; ++k) s[q][k] = qnan;
                    }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int k = 0; k < NS; ++k) c_local += (s[q][k] <= rc) ? 1 : 0;   // NaN: false
#pragma unroll
            for (int k = 0; k < NS; ++k) {
#pragma unroll
                for (int qh = 0; qh < 2; ++qh) {
                    // {hi16(s[2qh]), hi16(s[2qh+1])}: bytes 2,3 of each -> one PRMT
                    nk[k][qh] = __byte_perm(__float_as_uint(s[2 * qh][k]),
                                            __float_as_uint(s[2 * qh + 1][k]), 0x7632);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < NS; ++k) { nk[k][0] = 0x7fff7fffu; nk[k][1] = 0x7fff7fffu; }
        }
        {
            const uint32_t dst = kscr + (uint32_t)(2 * v) * kstride;
            sts128(dst, nk[0][0], nk[0][1], nk[1][0], nk[1][1]);
            if (NS == 4) sts128(dst + kstride, nk[2][0], nk[2][1], nk[3][0], nk[3][1]);
        }
        pa0 += vstride; pb0 += vstride; pa1 += vstride; pb1 += vstride;
    }
    cnt = c_local;
}

template <int V>
__device__ __forceinline__ int count_le(const uint32_t (&keys)[V][4][2], int keep, uint32_t piv2) {
    uint32_t acc0 = 0u, acc1 = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < keep) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                acc0 = bf2_add(acc0, bf2_le(keys[v][k][0], piv2));
                acc1 = bf2_add(acc1, bf2_le(keys[v][k][1], piv2));
            }
        }
    }
    return bf2_count_sum(acc0) + bf2_count_sum(acc1);
}

// Bit (r) / (16 + r) of word w <-> low / high half of register R = 16 w + r,
// R = (v * 4 + slot) * 2 + qh.
template <int V>
__device__ __forceinline__ void scan_range(const uint32_t (&keys)[V][4][2], int keep,
                                           uint32_t lo2, uint32_t hi2,
                                           uint32_t (&bm)[(V + 1) / 2]) {
#pragma unroll
    for (int w = 0; w < (V + 1) / 2; ++w) bm[w] = 0u;
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < keep) {
#pragma unroll
                for (int qh = 0; qh < 2; ++qh) {
                    const int R = (v * 4 + k) * 2 + qh;
                    const uint32_t m = bf2_ge_mask(keys[v][k][qh], lo2) & bf2_le_mask(keys[v][k][qh], hi2);
                    bm[R >> 4] |= m & (0x00010001u << (R & 15));
                }
            }
        }
}

// Full float32 pattern of one kept value, re-materialised from the coordinates
// (candidate gathering; deliberately out of line).  LB: slot -> one combination.
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

// element behind bit `bit` of word `w` of a scan_range bitmap
__device__ __forceinline__ uint32_t element_bits(const PairPtrs& pp, const PairDesc& d, int mode,
                                                 int tid, int nthr, int w, int bit) {
    const int R = (w << 4) | (bit & 15);
    const int half = bit >> 4;
    const int v = R >> 3, slot = (R >> 1) & 3, qh = R & 1;
    const int st = 4 * (tid + v * nthr) + 2 * qh + half;
    if (mode == IGMK_MODE_GP)
        return recompute_gp(pp.A0, pp.A1, pp.B0, pp.B1, d.cmask, d.keep, st, slot);
    // LB: the slot-th existing combination (enumeration order d0..d3)
    int cm = d.cmask;
    for (int t = 0; t < slot; ++t) cm &= cm - 1;
    const int comb = __ffs(cm) - 1;
    return recompute_lb((comb & 2) ? pp.A1 : pp.A0, (comb & 1) ? pp.B1 : pp.B0, st);
}

// ------------------------------------------------------------- one pair
template <int V, class G>
__device__ __forceinline__ void process_pair(const ActdistParams& P, G& g, long long pair) {
    const int i = __ldg(P.pi + pair), j = __ldg(P.pj + pair);
    const PairDesc d = make_pair_desc(P, i, j);
    igmk_pair_result* out = P.out + pair;
    if (!d.valid) {                        // uniform over the group
        if (g.tid == 0) write_empty(out);
        return;
    }
    const double pwish = __ldg(P.pwish + pair), plast = __ldg(P.plast + pair);
    const PairPtrs pp = pair_ptrs(P, d);

    int cnt;
    switch (pair_shape(d, P.mode)) {       // uniform over the group
        case SH_FULL4:  fill_keys<V, SH_FULL4>(P, d, pp, g.tid, g.nthr, g.kscr, g.kstride, cnt); break;
        case SH_INTRA2: fill_keys<V, SH_INTRA2>(P, d, pp, g.tid, g.nthr, g.kscr, g.kstride, cnt); break;
        case SH_GP4:    fill_keys<V, SH_GP4>(P, d, pp, g.tid, g.nthr, g.kscr, g.kstride, cnt); break;
        default:        fill_keys<V, SH_GENERIC>(P, d, pp, g.tid, g.nthr, g.kscr, g.kstride, cnt); break;
    }
    // keys back into registers (each thread reads only what it wrote itself)
    uint32_t keys[V][4][2];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        lds128(g.kscr + (uint32_t)(2 * v) * g.kstride, keys[v][0][0], keys[v][0][1], keys[v][1][0], keys[v][1][1]);
        if (d.keep > 2)
            lds128(g.kscr + (uint32_t)(2 * v + 1) * g.kstride, keys[v][2][0], keys[v][2][1], keys[v][3][0], keys[v][3][1]);
        else { keys[v][2][0] = keys[v][2][1] = keys[v][3][0] = keys[v][3][1] = 0x7fff7fffu; }
    }

    // per-thread key range (NaN halves are ignored by min/max.bf16x2)
    uint32_t mn2 = 0x7fff7fffu, mx2 = 0x7fff7fffu;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < d.keep) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                mn2 = bf2_min(mn2, bf2_min(keys[v][k][0], keys[v][k][1]));
                mx2 = bf2_max(mx2, bf2_max(keys[v][k][0], keys[v][k][1]));
            }
        }
    }
    uint32_t kmin = min(mn2 & 0xffffu, mn2 >> 16);
    uint32_t mxl = mx2 & 0xffffu, mxh = mx2 >> 16;
    mxl = (mxl > 0x7f80u) ? 0u : mxl;
    mxh = (mxh > 0x7f80u) ? 0u : mxh;
    uint32_t kmax = max(mxl, mxh);
    if (g.tid == 0) sts32(g.cand_cnt, 0u);
    g.sum_min_max(cnt, kmin, kmax);

    double p;
    int o;
    compute_p_o(cnt, d.keep, P.nstruct, pwish, plast, P.it_corr, p, o);
    if (o < 0) {
        if (g.tid == 0) write_result(out, d, 0u, cnt, -1, 0.0);
        return;
    }

    // ---- bisection on the 16-bit keys
    uint32_t lo = kmin, hi = (kmax < kmin) ? kmin : kmax;
    int cb = 0, ch = d.keep * P.nstruct;
    while (lo < hi && (ch - cb) > kCandCap) {
        const uint32_t mid = (lo + hi) >> 1;
        const int c = g.sum(count_le<V>(keys, d.keep, mid | (mid << 16)));
        if (c > o) { hi = mid; ch = c; } else { lo = mid + 1; cb = c; }
    }

    uint32_t bm[(V + 1) / 2];
    scan_range<V>(keys, d.keep, lo | (lo << 16), hi | (hi << 16), bm);
    uint32_t vlo = lo << 16, vhi = (hi << 16) | 0xffffu;

    if ((ch - cb) > kCandCap) {
        // Single fat key h = lo == hi with more than kCandCap elements: bisect the
        // low 16 bits among the elements of that key, re-materialising them from
        // the coordinates on every pass (1-2 passes for nstruct ~ 10^4; more only
        // for degenerate inputs with many identical distances).
        const int cb0 = cb;                 // elements with key < h
        uint32_t l2 = 0u, h2 = 0xffffu;
        while (l2 < h2 && (ch - cb) > kCandCap) {
            const uint32_t m2 = (l2 + h2) >> 1;
            int c_loc = 0;
#pragma unroll
            for (int w = 0; w < (V + 1) / 2; ++w) {
                uint32_t b = bm[w];
                while (b) {
                    const int bit = __ffs(b) - 1;
                    b &= b - 1;
                    const uint32_t x = element_bits(pp, d, P.mode, g.tid, g.nthr, w, bit);
                    c_loc += ((x & 0xffffu) <= m2) ? 1 : 0;
                }
            }
            const int c = cb0 + g.sum(c_loc);
            if (c > o) { h2 = m2; ch = c; } else { l2 = m2 + 1; cb = c; }
        }
        vlo = (lo << 16) | l2;
        vhi = (lo << 16) | h2;
        if ((ch - cb) > kCandCap) {
            // l2 == h2: every remaining candidate has the same bit pattern.
            if (g.tid == 0) write_result(out, d, vlo, cnt, o, p);
            return;
        }
    }

    // ---- gather the <= kCandCap candidates in full precision
    g.sync();                               // cand_cnt = 0 visible
#pragma unroll
    for (int w = 0; w < (V + 1) / 2; ++w) {
        uint32_t b = bm[w];
        while (b) {
            const int bit = __ffs(b) - 1;
            b &= b - 1;
            const uint32_t x = element_bits(pp, d, P.mode, g.tid, g.nthr, w, bit);
            if (x >= vlo && x <= vhi) {
                const uint32_t slot = atoms_inc(g.cand_cnt);
                if (slot < (uint32_t)kCandCap) sts32(g.cand + slot * 4, x);
            }
        }
    }
    g.sync();

    // ---- exact rank inside one warp: the (o - cb)-th smallest candidate
    if (g.leader_warp()) {
        const int lane = g.tid & 31;
        const int n = min((int)lds32(g.cand_cnt), kCandCap);
        const int r = o - cb;
        const uint32_t x = (lane < n) ? lds32(g.cand + lane * 4) : 0xffffffffu;
        int rank = 0;
        for (int t = 0; t < n; ++t) {
            const uint32_t y = __shfl_sync(0xffffffffu, x, t);
            rank += (y < x || (y == x && t < lane)) ? 1 : 0;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, lane < n && rank == r);
        // hit is non-empty by construction; guard keeps a corrupted input from hanging
        const int src = hit ? (__ffs(hit) - 1) : 0;
        const uint32_t ans = __shfl_sync(0xffffffffu, x, src);
        if (lane == 0) write_result(out, d, ans, cnt, o, p);
    }
}

// ---------------------------------------------------------------- kernels
// G = 32: one pair per warp, kWarpsPerBlock independent warps per CTA.
constexpr int kWarpsPerBlock = 8;

template <int V>
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
actdist_warp_kernel(const ActdistParams P) {
    extern __shared__ uint4 s_keys[];             // [warp][2 V][32] key quads
    __shared__ uint32_t s_cand[kWarpsPerBlock][kCandCap];
    __shared__ uint32_t s_cnt[kWarpsPerBlock];
    const int warp = threadIdx.x >> 5;
    WarpGroup g;
    g.tid = threadIdx.x & 31;
    g.nthr = 32;
    g.cand = smem_addr(&s_cand[warp][0]);
    g.cand_cnt = smem_addr(&s_cnt[warp]);
    g.kscr = smem_addr(s_keys) + (uint32_t)(warp * 2 * V * 32 + g.tid) * 16u;
    g.kstride = 32u * 16u;
    const long long stride = (long long)gridDim.x * kWarpsPerBlock;
    for (long long pair = (long long)blockIdx.x * kWarpsPerBlock + warp; pair < P.n_pairs;
         pair += stride) {
        process_pair<V, WarpGroup>(P, g, pair);
        __syncwarp();
    }
}

// G = blockDim.x (multiple of 32, <= 768): one pair per CTA.
template <int V, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
actdist_block_kernel(const ActdistParams P) {
    extern __shared__ uint4 s_keys[];             // [2 V][blockDim] key quads
    __shared__ uint32_t s_cand[kCandCap];
    __shared__ uint32_t s_cnt;
    __shared__ uint32_t s_red[2 * 96];
    BlockGroup g;
    g.tid = threadIdx.x;
    g.nthr = blockDim.x;
    g.cand = smem_addr(s_cand);
    g.cand_cnt = smem_addr(&s_cnt);
    g.kscr = smem_addr(s_keys) + (uint32_t)threadIdx.x * 16u;
    g.kstride = (uint32_t)blockDim.x * 16u;
    g.red = smem_addr(s_red);
    g.parity = 0;
    for (long long pair = blockIdx.x; pair < P.n_pairs; pair += gridDim.x) {
        process_pair<V, BlockGroup>(P, g, pair);
        __syncthreads();
    }
}

// ------------------------------------------------------- cross-check kernel
// Straightforward version kept as an on-device cross-check (IGMK_ALGO_SIMPLE):
// one CTA per pair, all kept d2 values in shared memory, 32-pass most-
// significant-bit-first binary radix select on the raw float32 patterns.
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

}  // namespace igmk
