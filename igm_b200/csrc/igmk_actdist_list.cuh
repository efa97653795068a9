// igmk_actdist_list.cuh - K1, list form (sm_100a): the fast path of the Hi-C activation
// distance (igm/steps/ActivationDistanceStep.py:405-473; GP flavour GP_activation.py:367-424).
//
// The order-statistic index is bounded BEFORE any distance is computed: p <= pwish
// (cleanProbability, :314-332, :452-470), so  o <= o_max = round(npc * pwish * N).  At the
// sigma values that make the long candidate lists o_max is 1 - 5 % of the npc * N values of
// a pair, i.e. ~95 % of the values can never be the answer.  So instead of parking every
// value (as a 16-bit key) in shared memory and bisecting over all of them
// (igmk_actdist.cuh), this kernel
//   1. sample: takes the first chunk of every thread (the first 4 * G structures) and picks
//      a threshold T from the sample's order statistics with a safety margin (bisection on
//      the 16-bit keys held in registers);  T_eff = max(T, rcutsq);
//   2. fill: streams all chunks and appends only the values <= T_eff, in FULL float32, to a
//      short thread-private list in shared memory (one compare, one predicated store, one
//      predicated pointer bump per value; no contact counting, no key packing);
//   3. select: contact count = #{list <= rcutsq} (T_eff >= rcutsq), p and o in float64, then
//      the o-th smallest of the list: bisection in value space, every thread reading its own
//      list with 128-bit loads, until <= 32 candidates remain; exact rank by a bitonic
//      sort across the lanes of one warp.
// Exactness does not depend on the sample: the list is complete for values <= T_eff, so if
// it holds at least o + 1 values the o-th smallest of the list IS the o-th smallest of the
// population.  Pairs for which that fails (sample not representative, thread list full,
// large o, many contacts) are appended to a redo list and processed by the key-array
// kernels of igmk_actdist.cuh in a second launch.
#pragma once
#include "igmk_actdist.cuh"

namespace igmk {

#ifndef IGMK_LSLOTS
#define IGMK_LSLOTS 44            // list words per thread; = 4 (mod 8): conflict-free 128-bit reads
#endif
#ifndef IGMK_LWPB
#define IGMK_LWPB 24              // warps per CTA of the warp-group kernel: 6 per scheduler at 80 registers; with ONE
                                  // locus-i tile slot 24 warps' lists (132 KB) + tile (24 KB) stay below the 164 KB
                                  // shared-memory carve-out step (20 warps: 297, 24 warps: 312 M pairs/s on config 2)
#endif
#ifndef IGMK_LBT
#define IGMK_LBT 320              // threads per CTA of the CTA-group kernel
#endif
constexpr int kListSlots = IGMK_LSLOTS;
constexpr int kListSpare = 16 + 3;                  // one chunk's appends + the sentinel padding
constexpr int kListCap = kListSlots - kListSpare;   // entries a thread may hold before a chunk
constexpr uint32_t kListBytes = (uint32_t)kListSlots * 4u;
constexpr uint32_t kListSentinel = 0x7fffffffu;     // compares above every pivot
constexpr int kListWarps = IGMK_LWPB;
constexpr int kListBlockThreads = IGMK_LBT;
static_assert(kListSlots % 8 == 4, "list stride must be 4 mod 8 words");

// ---------------------------------------------------------------- pair descriptors
__global__ void __launch_bounds__(256)
build_pairrec_kernel(const ActdistParams P, PairRec* __restrict__ rec) {
    const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= P.n_pairs) return;
    const long long pair = P.perm ? (long long)__ldg(P.perm + slot) : slot;
    const PairDesc d = make_pair_desc(P, __ldg(P.pi + pair), __ldg(P.pj + pair));
    PairRec r;
    r.pair = (int32_t)pair;
    r.a0 = d.a0; r.a1 = d.a1; r.b0 = d.b0; r.b1 = d.b1;
    r.rcutsq = d.rcutsq;
    r.bits = (uint32_t)d.keep | ((uint32_t)d.cmask << 4) | ((uint32_t)d.nrec << 8) | ((uint32_t)d.valid << 12);
    r.omax = 0;
    if (d.valid) {
        // o <= o_max: p <= pwish for pwish <= 1 (the post-check of the select does not rely on it)
        const int total = d.keep * P.nstruct;
        const double x = __dmul_rn(__dmul_rn((double)d.keep, __ldg(P.pwish + pair)), (double)P.nstruct);
        const double q = rint(x);
        r.omax = (q >= (double)(total - 1)) ? (total - 1) : ((q > 0.0) ? (int)q : 0);
    }
    reinterpret_cast<uint4*>(rec + slot)[0] = make_uint4((uint32_t)r.pair, (uint32_t)r.a0, (uint32_t)r.a1, (uint32_t)r.b0);
    reinterpret_cast<uint4*>(rec + slot)[1] = make_uint4((uint32_t)r.b1, __float_as_uint(r.rcutsq), r.bits, (uint32_t)r.omax);
}

__device__ __forceinline__ PairDesc load_pairrec(const PairRec* rec, long long slot, long long& pair, int& omax) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(rec + slot));
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(rec + slot) + 1);
    PairDesc d;
    pair = (long long)(int32_t)u.x;
    d.a0 = (int)u.y; d.a1 = (int)u.z; d.b0 = (int)u.w; d.b1 = (int)v.x;
    d.rcutsq = __uint_as_float(v.y);
    d.keep = (int)(v.z & 15u);
    d.cmask = (int)((v.z >> 4) & 15u);
    d.nrec = (int)((v.z >> 8) & 15u);
    d.valid = (int)((v.z >> 12) & 1u);
    d.always_rec = 0;
    omax = (int)v.w;
    return d;
}

struct SampleOut { uint32_t T_bits; int ok; };

// Threshold from the group's first chunks (the first `ns` structures): kw = the 16-bit keys
// (high halves of the float32 values, two per word, NaN = not a value) of this thread's
// first chunk.  ok = 0 when the list such a threshold produces is expected to exceed the
// budget (large o or many contacts): the pair then goes to the key-array kernel.
// Deliberately out of line: one copy for all pair shapes.
template <bool BLOCK>
__device__ __noinline__ SampleOut sample_threshold(uint4 kwa, uint4 kwb, int tid, int nthr, uint32_t red,
                                                   int keep, int nstruct, int omax, uint32_t rcb,
                                                   float z, float budget) {
    Group<BLOCK> g;
    g.tid = tid; g.nthr = nthr; g.red = red; g.parity = 0;
    g.kscr = 0u; g.kstride = 0u; g.list = 0u; g.ctl = 0u; g.cap = 0; g.list2 = 0u;
    const uint32_t kw[8] = {kwa.x, kwa.y, kwa.z, kwa.w, kwb.x, kwb.y, kwb.z, kwb.w};
    uint32_t lmn = 0x7fff7fffu, lmx = 0x7fff7fffu;
#pragma unroll
    for (int w = 0; w < 8; w += 2) {
        lmn = bf2_min(lmn, bf2_min(kw[w], kw[w + 1]));         // NaN halves are ignored
        lmx = bf2_max(lmx, bf2_max(kw[w], kw[w + 1]));
    }
    uint32_t kmin = min(lmn & 0xffffu, lmn >> 16);
    uint32_t mxl = lmx & 0xffffu, mxh = lmx >> 16;
    mxl = (mxl > 0x7f80u) ? 0u : mxl;
    mxh = (mxh > 0x7f80u) ? 0u : mxh;
    uint32_t kmax = max(mxl, mxh);
    int zero = 0;
    g.sum_min_max(zero, kmin, kmax);
    if (kmax < kmin) kmax = kmin;

    SampleOut out;
    out.T_bits = 0u; out.ok = 0;
    const int ns = min(nstruct, 4 * nthr);
    const float Sv = (float)(keep * ns), M = (float)(keep * nstruct);
    int k_t;
    if (ns == nstruct) {
        k_t = omax + 1;                                // the sample is the population
    } else {
        // smallest k with  k - z sqrt(k) >= r0  (r0: expected sample rank of element o_max)
        const float r0 = (float)(omax + 1) * Sv / M;
        k_t = (int)(r0 + 0.5f * z * z + 0.5f * z * sqrtf(z * z + 4.f * r0)) + 1;
    }
    if ((float)k_t * M > budget * Sv) return out;
    const int k_hi = k_t + max(2, k_t >> 2);
    uint32_t lo = kmin, hi = kmax;
    int ch = keep * ns;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const uint32_t m2 = mid | (mid << 16);
        uint32_t a0 = 0u, a1 = 0u;
#pragma unroll
        for (int w = 0; w < 8; w += 2) {
            a0 -= bf2_le_mask(kw[w], m2);
            a1 -= bf2_le_mask(kw[w + 1], m2);
        }
        const uint32_t acc = a0 + a1;
        const int c = g.sum(((int)acc >> 16) + 2 * (int)(acc & 0xffffu));
        if (c >= k_t) {
            hi = mid; ch = c;
            if (c <= k_hi) break;
        } else {
            lo = mid + 1;
        }
    }
    uint32_t T = (hi << 16) | 0xffffu;
    if (rcb > T) {                                     // the list must hold every contact
        T = rcb;
        const uint32_t rk = rcb >> 16, r2 = rk | (rk << 16);
        uint32_t a0 = 0u, a1 = 0u;
#pragma unroll
        for (int w = 0; w < 8; w += 2) {
            a0 -= bf2_le_mask(kw[w], r2);
            a1 -= bf2_le_mask(kw[w + 1], r2);
        }
        const uint32_t acc = a0 + a1;
        ch = g.sum(((int)acc >> 16) + 2 * (int)(acc & 0xffffu));
    }
    if ((float)ch * M > budget * Sv) return out;
    out.T_bits = T; out.ok = 1;
    return out;
}

// Fill of one pair.  Returns 0 when the thread lists are complete for values <= T_bits,
// 1 when the pair has to be redone by the key-array kernel.
// SAMPLE = false (slab pipeline, igmk_actdist_slab.cuh): T_bits is given, the row pointers
// start at chunk c0 of the population and V counts the chunks of this slab only.
template <int SH, bool AS, bool BLOCK, bool SAMPLE = true>
__device__ __forceinline__ int fill_list(const ActdistParams& P, const Group<BLOCK>& g, const PairDesc& d,
                                         const PairPtrs& pp, int V, uint32_t as0, uint32_t as1,
                                         uint32_t lbase, uint32_t sred, int omax,
                                         uint32_t& T_bits, int& mycnt, bool& ovf, int c0 = 0) {
    constexpr int NS = (SH == SH_FULL4) ? 4 : (SH == SH_INTRA2 || SH == SH_GP4) ? 2 : 4;
    const float qnan = __int_as_float(0x7fffffff);
    const u64 nz = P.negzero2;
    const int tid = g.tid, nthr = g.nthr;
    const size_t off0 = (size_t)(tid >> 5) * kSegFloats + (size_t)(tid & 31) * 4;
    const size_t vstride = (size_t)(nthr >> 5) * kSegFloats;
    const float* pa0 = pp.A0 + off0;
    const float* pb0 = pp.B0 + off0;
    const float* pa1 = pp.A1 + off0;
    const float* pb1 = pp.B1 + off0;
    uint32_t sa0 = as0 + (uint32_t)off0 * 4u, sa1 = as1 + (uint32_t)off0 * 4u;
    uint32_t lptr = lbase;
    const uint32_t llimit = lbase + (uint32_t)kListCap * 4u;
    float T = SAMPLE ? 0.f : __uint_as_float(T_bits);
    ovf = false;
    int status = 0;
    const u64 pol = (IGMK_JLOAD == LD_STREAM_L2KEEP || IGMK_JLOAD == LD_PLAIN_L2KEEP) ? l2_policy_evict_last() : 0ull;
#pragma unroll 1
    for (int v = 0; v < V; ++v) {
        const int c = c0 + tid + v * nthr;
        if ((v > 0 || !SAMPLE) && c >= P.nchunks) break;   // padding chunks (c only grows)
        float s[4][NS];
        {
            // a padding chunk at v = 0 (tiny populations) still takes part in the sample;
            // its loads fall inside the zero-padded rows (npad >= 128)
            Row6 a0, a1;
            const Row6 b0 = load_row6<IGMK_JLOAD>(pb0, pol), b1 = load_row6<IGMK_JLOAD>(pb1, pol);
            if (AS) {
                a0 = load_row6_shared(sa0); a1 = load_row6_shared(sa1);
            } else {
                a0 = load_row6<LD_KEEP>(pa0); a1 = load_row6<LD_KEEP>(pa1);
            }
            chunk_values<SH, NS>(d, P.mode, a0, a1, b0, b1, nz, s);
        }
        if (4 * c + 4 > P.nstruct) {                   // tail / padding chunk of the population
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (4 * c + q >= P.nstruct) {
#pragma unroll
                    for (int k = 0; k < NS; ++k) s[q][k] = qnan;
                }
        }
        if (SAMPLE && v == 0) {                        // uniform over the group
            uint32_t kw[8];
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int qh = 0; qh < 2; ++qh)
                    kw[2 * k + qh] = (k < NS) ? __byte_perm(__float_as_uint(s[2 * qh][k < NS ? k : 0]),
                                                             __float_as_uint(s[2 * qh + 1][k < NS ? k : 0]), 0x7632)
                                              : 0x7fff7fffu;
            const SampleOut so = sample_threshold<BLOCK>(make_uint4(kw[0], kw[1], kw[2], kw[3]),
                                                         make_uint4(kw[4], kw[5], kw[6], kw[7]), tid, nthr, sred,
                                                         d.keep, P.nstruct, omax, __float_as_uint(d.rcutsq),
                                                         P.list_z, P.list_budget);
            if (!so.ok) { status = 1; break; }
            T_bits = so.T_bits;
            T = __uint_as_float(so.T_bits);
        }
        if (lptr > llimit) {                           // thread list full: stop appending, redo the pair
            ovf = true;
            T = -1.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int k = 0; k < NS; ++k)
                if (s[q][k] <= T) {
                    sts32(lptr, __float_as_uint(s[q][k]));
                    lptr += 4u;
                }
        pa0 += vstride; pb0 += vstride; pa1 += vstride; pb1 += vstride;
        sa0 += (uint32_t)vstride * 4u; sa1 += (uint32_t)vstride * 4u;
    }
    // pad to a whole quad: the select reads the list with 128-bit loads
    sts32(lptr, kListSentinel); sts32(lptr + 4u, kListSentinel); sts32(lptr + 8u, kListSentinel);
    mycnt = (int)((lptr - lbase) >> 2);
    return status;
}

// #{own list values <= piv}: 1.0 / 0.0 flags (FSET.BF) accumulated two per FADD2, exact below
// 2^24; the sentinel padding is a NaN pattern and never counts.
__device__ __forceinline__ int list_count_le(uint32_t lbase, int nq, float piv) {
    u64 acc = 0ull;
#pragma unroll 2
    for (int q = 0; q < nq; ++q) {
        uint32_t x0, x1, x2, x3;
        lds128(lbase + (uint32_t)q * 16u, x0, x1, x2, x3);
        acc = f2add(acc, f2pack(f_le_one(__uint_as_float(x0), piv), f_le_one(__uint_as_float(x1), piv)));
        acc = f2add(acc, f2pack(f_le_one(__uint_as_float(x2), piv), f_le_one(__uint_as_float(x3), piv)));
    }
    float lo, hi;
    f2split(acc, lo, hi);
    return (int)(lo + hi);
}

// r-th smallest (0-based) of the n <= 32 words at `list`: bitonic sort across the lanes.
__device__ __forceinline__ uint32_t warp_rank(uint32_t list, int n, int r, int lane) {
    uint32_t x = (lane < n) ? lds32(list + (uint32_t)lane * 4u) : 0xffffffffu;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
            const bool keep_min = (((lane & k) == 0) == ((lane & j) == 0));
            x = keep_min ? min(x, y) : max(x, y);
        }
    }
    return __shfl_sync(0xffffffffu, x, r);
}

// p, o and the o-th smallest value from the thread lists.  Returns false when the lists
// cannot answer (overflow, or fewer than o + 1 values: the sample misjudged the pair).
template <bool BLOCK>
__device__ __forceinline__ bool select_list(const ActdistParams& P, Group<BLOCK>& g, long long pair,
                                            const PairDesc& d, uint32_t lbase,
                                            int mycnt, bool ovf, uint32_t T_bits) {
    const int n = g.sum(mycnt + (ovf ? (1 << 20) : 0));       // lists hold < 2^20 values
    if ((n >> 20) || T_bits >= 0x7f800000u) return false;     // (overflowed distances: not for this path)
    const int nq = (mycnt + 3) >> 2;
    const int rcb = (int)__float_as_uint(d.rcutsq), tb = (int)T_bits;
    // every value <= rcutsq is in the list (T >= rcutsq)
    const int cnt = (rcb >= tb) ? n : g.sum(list_count_le(lbase, nq, d.rcutsq));
    double p;
    int o;
    compute_p_o(cnt, d.keep, P.nstruct, __ldg(P.pwish + pair), __ldg(P.plast + pair), P.it_corr, p, o);
    if (o < 0) {
        emit_result(P, g.tid, pair, d, 0u, cnt, -1, 0.0, 0.0);
        return true;
    }
    if (n < o + 1) return false;
    // bracket (lo, hi] in bit-pattern order
    int lo = -1, hi = tb, cb = 0, ch = n;
    if (rcb < tb) {
        if (cnt > o) { hi = rcb; ch = cnt; } else { lo = rcb; cb = cnt; }
    }
    int pass = 0;
    while (ch - cb > kRankCap && hi - lo > 1) {
        int mid;
        if (pass < 8) {                                // midpoint in value space
            const float fl = (lo < 0) ? 0.f : __int_as_float(lo), fh = __int_as_float(hi);
            mid = __float_as_int(fl + (fh - fl) * 0.5f);
        } else {                                       // ... then in pattern space (bounded)
            mid = lo + ((hi - lo) >> 1);
        }
        mid = max(lo + 1, min(mid, hi - 1));
        const int c = g.sum(list_count_le(lbase, nq, __int_as_float(mid)));
        if (c > o) { hi = mid; ch = c; } else { lo = mid; cb = c; }
        ++pass;
    }
    if (ch - cb > kRankCap) {                          // hi == lo + 1: all candidates are `hi`
        emit_result(P, g.tid, pair, d, (uint32_t)hi, cnt, o, p, 0.0);
        return true;
    }
    g.sync();                                          // candidate counter = 0 is visible
    for (int q = 0; q < nq; ++q) {
        uint32_t x[4];
        lds128(lbase + (uint32_t)q * 16u, x[0], x[1], x[2], x[3]);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if ((int)x[t] > lo && (int)x[t] <= hi) {   // (the sentinel is above every hi)
                const uint32_t slot = atoms_inc(g.ctl);
                if (slot < (uint32_t)kRankCap) sts32(g.list + slot * 4, x[t]);
            }
        }
    }
    g.sync();
    if (g.leader_warp()) {
        const uint32_t ans = warp_rank(g.list, ch - cb, o - cb, g.tid & 31);
        emit_result(P, g.tid, pair, d, ans, cnt, o, p, 0.0);
    }
    return true;
}

__device__ __forceinline__ void push_redo(const ActdistParams& P, int tid, long long pair) {
    if (tid == 0) {
        const unsigned int k = atomicAdd(P.redo_count, 1u);
        P.redo[k] = (int32_t)pair;
    }
}

template <bool BLOCK>
__device__ __forceinline__ void process_pair_list(const ActdistParams& P, Group<BLOCK>& g, int V,
                                                  long long slot, const TileCtl& tile,
                                                  uint32_t lbase, uint32_t sred) {
    long long pair;
    int omax;
    const PairDesc d = load_pairrec(P.rec, slot, pair, omax);
    if (!d.valid) {                        // uniform over the group
        emit_empty(P, g.tid, pair);
        return;
    }
    const PairPtrs pp = pair_ptrs(P, d);
    const int i = d.a0;                    // key of the locus-i tile: the bead of copy 0
    if (g.tid == 0) sts32(g.ctl, 0u);
    uint32_t T_bits = 0u;
    int mycnt = 0, status;
    bool ovf = false;
    const int tslot = (!BLOCK && tile.base) ? tile_acquire(P, tile, i, d, pp, g.tid) : -1;
    if (tslot >= 0) {
        const uint32_t as0 = tile.base + (uint32_t)tslot * tile.slot_bytes;
        const uint32_t as1 = (d.a1 >= 0) ? as0 + (tile.slot_bytes >> 1) : as0;
        switch (pair_shape(d, P.mode)) {       // uniform over the group
            case SH_FULL4:  status = fill_list<SH_FULL4, true, BLOCK>(P, g, d, pp, V, as0, as1, lbase, sred, omax, T_bits, mycnt, ovf); break;
            case SH_INTRA2: status = fill_list<SH_INTRA2, true, BLOCK>(P, g, d, pp, V, as0, as1, lbase, sred, omax, T_bits, mycnt, ovf); break;
            case SH_GP4:    status = fill_list<SH_GP4, true, BLOCK>(P, g, d, pp, V, as0, as1, lbase, sred, omax, T_bits, mycnt, ovf); break;
            default:        status = fill_list<SH_GENERIC, true, BLOCK>(P, g, d, pp, V, as0, as1, lbase, sred, omax, T_bits, mycnt, ovf); break;
        }
        tile_release(tile, tslot, g.tid);
    } else {
        switch (pair_shape(d, P.mode)) {
            case SH_FULL4:  status = fill_list<SH_FULL4, false, BLOCK>(P, g, d, pp, V, 0u, 0u, lbase, sred, omax, T_bits, mycnt, ovf); break;
            case SH_INTRA2: status = fill_list<SH_INTRA2, false, BLOCK>(P, g, d, pp, V, 0u, 0u, lbase, sred, omax, T_bits, mycnt, ovf); break;
            case SH_GP4:    status = fill_list<SH_GP4, false, BLOCK>(P, g, d, pp, V, 0u, 0u, lbase, sred, omax, T_bits, mycnt, ovf); break;
            default:        status = fill_list<SH_GENERIC, false, BLOCK>(P, g, d, pp, V, 0u, 0u, lbase, sred, omax, T_bits, mycnt, ovf); break;
        }
    }
    if (status != 0 || !select_list<BLOCK>(P, g, pair, d, lbase, mycnt, ovf, T_bits))
        push_redo(P, g.tid, pair);
}

// ---------------------------------------------------------------- kernels
// G = 32: one pair per warp, no CTA barrier after the set-up.  Shared memory: kListSlots
// words of list per thread, then the locus-i tiles (igmk_actdist.cuh; their rows arrive by
// bulk asynchronous copies).
__global__ void __launch_bounds__(32 * kListWarps, 1)
actdist_list_warp_kernel(const ActdistParams P, const int V) {
    extern __shared__ uint4 s_dyn[];              // [thread] lists | locus-i tiles
    __shared__ uint32_t s_list[kListWarps][kRankCap];
    __shared__ uint32_t s_cnt[kListWarps];
    __shared__ TileShared s_tile;
    __shared__ unsigned int s_ticket;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Group<false> g;
    g.tid = lane;
    g.nthr = 32;
    g.list = smem_addr(&s_list[warp][0]);
    g.ctl = smem_addr(&s_cnt[warp]);
    g.cap = kRankCap;
    g.kscr = 0u; g.kstride = 0u; g.red = 0u; g.list2 = 0u; g.parity = 0;
    const uint32_t dyn0 = smem_addr(s_dyn);
    const uint32_t lbase = dyn0 + (uint32_t)threadIdx.x * kListBytes;
    TileCtl tile;
    tile.base = 0u; tile.slot_bytes = 24u * (uint32_t)P.npad; tile.nslots = P.tile_slots;
    tile_bind(&s_tile, tile);
    if (P.tile_slots > 0) tile.base = dyn0 + (uint32_t)blockDim.x * kListBytes;
    if (threadIdx.x == 0) {
        tile_init(&s_tile);
        s_ticket = 0u;
    }
    __syncthreads();
    // CTA-contiguous blocks of 2^bshift pairs (block k of this CTA = list block blockIdx +
    // k * gridDim), handed to the warps one pair at a time by a shared ticket counter:
    // consecutive pairs share locus i, fast and slow pairs balance out across the warps.
    const unsigned int bshift = (unsigned int)P.tile_block;      // log2 of the block size
    for (;;) {
        unsigned int t = 0u;
        if (lane == 0) t = atomicAdd(&s_ticket, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        const unsigned int k = t >> bshift, r = t & ((1u << bshift) - 1u);
        const long long base = ((long long)blockIdx.x + (long long)k * gridDim.x) << bshift;
        if (base >= P.n_pairs) break;
        if (base + r < P.n_pairs) process_pair_list<false>(P, g, V, base + r, tile, lbase, 0u);
        __syncwarp();
    }
}

// G = blockDim.x: one pair per CTA, two CTAs per SM (N > 1024).
__global__ void __launch_bounds__(kListBlockThreads, 2)
actdist_list_block_kernel(const ActdistParams P, const int V) {
    extern __shared__ uint4 s_dyn[];              // [thread] lists
    __shared__ uint32_t s_list[kRankCap];
    __shared__ uint32_t s_cnt;
    __shared__ uint32_t s_red[2 * 96];
    __shared__ uint32_t s_red2[2 * 96];           // reductions of the sample phase
    Group<true> g;
    g.tid = threadIdx.x;
    g.nthr = blockDim.x;
    g.list = smem_addr(s_list);
    g.ctl = smem_addr(&s_cnt);
    g.cap = kRankCap;
    g.kscr = 0u; g.kstride = 0u;
    g.red = smem_addr(s_red);
    g.list2 = 0u;
    g.parity = 0;
    const uint32_t lbase = smem_addr(s_dyn) + (uint32_t)threadIdx.x * kListBytes;
    TileCtl tile;
    tile.base = 0u; tile.slot_bytes = 0u; tile.words = 0u; tile.nslots = 0; tile.bars = 0u; tile.pars = 0u;
    for (long long slot = blockIdx.x; slot < P.n_pairs; slot += gridDim.x) {
        process_pair_list<true>(P, g, V, slot, tile, lbase, smem_addr(s_red2));
        __syncthreads();
    }
}

}  // namespace igmk
