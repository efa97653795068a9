// igmk_actdist.cuh - K1: the Hi-C activation-distance kernels (sm_100a).
//
// Replaces the per-pair NumPy body of get_actdist
// (igm/steps/ActivationDistanceStep.py:405-473; GP flavour
// igm/steps/GP_activation.py:367-424).
//
// Fast kernel, per candidate pair handled by a group of G threads
// (G = 32: one warp per pair for nstruct <= 1024; G = blockDim for larger
// populations):
//   1. every thread streams V float4 chunks (4 structures each) of the <= 4
//      bead rows with 128-bit loads, computes the copy-combination d2 values
//      in registers (non-FMA float32, bit-identical to NumPy), counts
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
struct WarpGroup {
    int tid;            // lane
    int nthr;           // 32
    uint32_t* cand;     // shared, kCandCap entries (per warp)
    int* cand_cnt;      // shared (per warp)

    __device__ __forceinline__ int sum(int x) { return __reduce_add_sync(0xffffffffu, x); }
    __device__ __forceinline__ void sum_min_max(int& s, uint32_t& mn, uint32_t& mx) {
        s = __reduce_add_sync(0xffffffffu, s);
        mn = __reduce_min_sync(0xffffffffu, mn);
        mx = __reduce_max_sync(0xffffffffu, mx);
    }
    __device__ __forceinline__ void sync() { __syncwarp(); }
    __device__ __forceinline__ bool leader_warp() const { return true; }
    __device__ __forceinline__ uint32_t bcast(uint32_t v) { return v; }  // already warp-uniform
};

struct BlockGroup {
    int tid;
    int nthr;
    uint32_t* cand;
    int* cand_cnt;
    int* red;           // shared [2][3][32]
    uint32_t* bc;       // shared broadcast slot
    int parity;

    __device__ __forceinline__ int sum(int x) {
        const int lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
        int* r = red + parity * 96;
        const int w = __reduce_add_sync(0xffffffffu, x);
        if (lane == 0) r[warp] = w;
        __syncthreads();
        const int v = (lane < nw) ? r[lane] : 0;
        parity ^= 1;
        return __reduce_add_sync(0xffffffffu, v);
    }
    __device__ __forceinline__ void sum_min_max(int& s, uint32_t& mn, uint32_t& mx) {
        const int lane = tid & 31, warp = tid >> 5, nw = nthr >> 5;
        int* r = red + parity * 96;
        const int ws = __reduce_add_sync(0xffffffffu, s);
        const uint32_t wmn = __reduce_min_sync(0xffffffffu, mn);
        const uint32_t wmx = __reduce_max_sync(0xffffffffu, mx);
        if (lane == 0) { r[warp] = ws; r[32 + warp] = (int)wmn; r[64 + warp] = (int)wmx; }
        __syncthreads();
        const int vs = (lane < nw) ? r[lane] : 0;
        const uint32_t vmn = (lane < nw) ? (uint32_t)r[32 + lane] : 0xffffffffu;
        const uint32_t vmx = (lane < nw) ? (uint32_t)r[64 + lane] : 0u;
        parity ^= 1;
        s = __reduce_add_sync(0xffffffffu, vs);
        mn = __reduce_min_sync(0xffffffffu, vmn);
        mx = __reduce_max_sync(0xffffffffu, vmx);
    }
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ bool leader_warp() const { return tid < 32; }
    // value known to warp 0 only -> everybody
    __device__ __forceinline__ uint32_t bcast(uint32_t v) {
        if (tid == 0) *bc = v;
        __syncthreads();
        const uint32_t r = *bc;
        __syncthreads();
        return r;
    }
};

// --------------------------------------------------------- stage 1: fill keys
// keys[v][slot][qh]: low half = structure 4c + 2qh, high half = 4c + 2qh + 1 of
// chunk c = tid + v * nthr; NaN pattern (0x7fff) for everything not kept.
template <int V>
__device__ __forceinline__ void fill_keys(const ActdistParams& P, const PairDesc& d,
                                          int tid, int nthr, uint32_t (&keys)[V][4][2],
                                          int& cnt) {
    const size_t row = (size_t)3 * P.npad;
    const float* A0 = P.coords + (size_t)d.a0 * row;
    const float* B0 = P.coords + (size_t)d.b0 * row;
    const float* A1 = P.coords + (size_t)(d.a1 >= 0 ? d.a1 : d.a0) * row;
    const float* B1 = P.coords + (size_t)(d.b1 >= 0 ? d.b1 : d.b0) * row;
    const bool need_a1 = (d.cmask & (CM_D2 | CM_D3)) != 0;
    const bool need_b1 = (d.cmask & (CM_D1 | CM_D3)) != 0;
    const float qnan = __int_as_float(0x7fffffff);
    const float rc = d.rcutsq;
    const int npad = P.npad;
    int c_local = 0;

#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int c = tid + v * nthr;
        float s[4][4];   // [q][slot]
        if (c < P.nchunks) {
            const int off = 4 * c;
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 ax0 = __ldg(reinterpret_cast<const float4*>(A0 + off));
            const float4 ay0 = __ldg(reinterpret_cast<const float4*>(A0 + npad + off));
            const float4 az0 = __ldg(reinterpret_cast<const float4*>(A0 + 2 * npad + off));
            const float4 bx0 = __ldg(reinterpret_cast<const float4*>(B0 + off));
            const float4 by0 = __ldg(reinterpret_cast<const float4*>(B0 + npad + off));
            const float4 bz0 = __ldg(reinterpret_cast<const float4*>(B0 + 2 * npad + off));
            float4 ax1 = zero4, ay1 = zero4, az1 = zero4, bx1 = zero4, by1 = zero4, bz1 = zero4;
            if (need_a1) {
                ax1 = __ldg(reinterpret_cast<const float4*>(A1 + off));
                ay1 = __ldg(reinterpret_cast<const float4*>(A1 + npad + off));
                az1 = __ldg(reinterpret_cast<const float4*>(A1 + 2 * npad + off));
            }
            if (need_b1) {
                bx1 = __ldg(reinterpret_cast<const float4*>(B1 + off));
                by1 = __ldg(reinterpret_cast<const float4*>(B1 + npad + off));
                bz1 = __ldg(reinterpret_cast<const float4*>(B1 + 2 * npad + off));
            }
            const float AX0[4] = {ax0.x, ax0.y, ax0.z, ax0.w}, AY0[4] = {ay0.x, ay0.y, ay0.z, ay0.w},
                        AZ0[4] = {az0.x, az0.y, az0.z, az0.w};
            const float AX1[4] = {ax1.x, ax1.y, ax1.z, ax1.w}, AY1[4] = {ay1.x, ay1.y, ay1.z, ay1.w},
                        AZ1[4] = {az1.x, az1.y, az1.z, az1.w};
            const float BX0[4] = {bx0.x, bx0.y, bx0.z, bx0.w}, BY0[4] = {by0.x, by0.y, by0.z, by0.w},
                        BZ0[4] = {bz0.x, bz0.y, bz0.z, bz0.w};
            const float BX1[4] = {bx1.x, bx1.y, bx1.z, bx1.w}, BY1[4] = {by1.x, by1.y, by1.z, by1.w},
                        BZ1[4] = {bz1.x, bz1.y, bz1.z, bz1.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float d0, d1 = qnan, d2 = qnan, d3 = qnan;
                d0 = d2_nofma(AX0[q], AY0[q], AZ0[q], BX0[q], BY0[q], BZ0[q]);
                if (d.cmask & CM_D1) d1 = d2_nofma(AX0[q], AY0[q], AZ0[q], BX1[q], BY1[q], BZ1[q]);
                if (d.cmask & CM_D2) d2 = d2_nofma(AX1[q], AY1[q], AZ1[q], BX0[q], BY0[q], BZ0[q]);
                if (d.cmask & CM_D3) d3 = d2_nofma(AX1[q], AY1[q], AZ1[q], BX1[q], BY1[q], BZ1[q]);
                pack_slots(d, P.mode, d0, d1, d2, d3, s[q]);
                if (off + q >= P.nstruct) { s[q][0] = s[q][1] = s[q][2] = s[q][3] = qnan; }
#pragma unroll
                for (int k = 0; k < 4; ++k) c_local += (s[q][k] <= rc) ? 1 : 0;   // NaN: false
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int k = 0; k < 4; ++k) s[q][k] = qnan;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int qh = 0; qh < 2; ++qh) {
                // {hi16(s[2qh]), hi16(s[2qh+1])}: bytes 2,3 of each -> PRMT
                keys[v][k][qh] = __byte_perm(__float_as_uint(s[2 * qh][k]),
                                             __float_as_uint(s[2 * qh + 1][k]), 0x7632);
            }
        }
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

// full float32 bit pattern of the element behind bit `bit` of word `w`
__device__ __forceinline__ uint32_t element_bits(const ActdistParams& P, const PairDesc& d,
                                                 int tid, int nthr, int w, int bit) {
    const int R = (w << 4) | (bit & 15);
    const int half = bit >> 4;
    const int v = R >> 3, slot = (R >> 1) & 3, qh = R & 1;
    const int st = 4 * (tid + v * nthr) + 2 * qh + half;
    float s[4];
    struct_slots(P, d, st, s);
    const float val = (slot == 0) ? s[0] : (slot == 1) ? s[1] : (slot == 2) ? s[2] : s[3];
    return __float_as_uint(val);
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

    uint32_t keys[V][4][2];
    int cnt;
    fill_keys<V>(P, d, g.tid, g.nthr, keys, cnt);

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
    if (g.tid == 0) *g.cand_cnt = 0;
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
                    const uint32_t x = element_bits(P, d, g.tid, g.nthr, w, bit);
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
            const uint32_t x = element_bits(P, d, g.tid, g.nthr, w, bit);
            if (x >= vlo && x <= vhi) {
                const int slot = atomicAdd(g.cand_cnt, 1);
                if (slot < kCandCap) g.cand[slot] = x;
            }
        }
    }
    g.sync();

    // ---- exact rank inside one warp: the (o - cb)-th smallest candidate
    if (g.leader_warp()) {
        const int lane = g.tid & 31;
        const int n = min(*g.cand_cnt, kCandCap);
        const int r = o - cb;
        const uint32_t x = (lane < n) ? g.cand[lane] : 0xffffffffu;
        int rank = 0;
#pragma unroll 8
        for (int t = 0; t < kCandCap; ++t) {
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
__global__ void __launch_bounds__(32 * kWarpsPerBlock, 2)
actdist_warp_kernel(const ActdistParams P) {
    __shared__ uint32_t s_cand[kWarpsPerBlock][kCandCap];
    __shared__ int s_cnt[kWarpsPerBlock];
    const int warp = threadIdx.x >> 5;
    WarpGroup g;
    g.tid = threadIdx.x & 31;
    g.nthr = 32;
    g.cand = s_cand[warp];
    g.cand_cnt = &s_cnt[warp];
    const long long stride = (long long)gridDim.x * kWarpsPerBlock;
    for (long long pair = (long long)blockIdx.x * kWarpsPerBlock + warp; pair < P.n_pairs;
         pair += stride) {
        process_pair<V, WarpGroup>(P, g, pair);
        __syncwarp();
    }
}

// G = blockDim.x (multiple of 32, <= 1024): one pair per CTA.
template <int V, int MAXT>
__global__ void __launch_bounds__(MAXT)
actdist_block_kernel(const ActdistParams P) {
    __shared__ uint32_t s_cand[kCandCap];
    __shared__ int s_cnt;
    __shared__ int s_red[2 * 96];
    __shared__ uint32_t s_bc;
    BlockGroup g;
    g.tid = threadIdx.x;
    g.nthr = blockDim.x;
    g.cand = s_cand;
    g.cand_cnt = &s_cnt;
    g.red = s_red;
    g.bc = &s_bc;
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
