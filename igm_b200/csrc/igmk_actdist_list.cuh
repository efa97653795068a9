// igmk_actdist_list.cuh - K1, list form (sm_100a): the fast path of the Hi-C activation
// distance (igm/steps/ActivationDistanceStep.py:405-473; GP flavour GP_activation.py:367-424).
//
// The order-statistic index is bounded BEFORE any distance is computed: p <= pwish
// (cleanProbability, :314-332, :452-470), so  o <= o_max = round(npc * pwish * N).  At the
// sigma values that make the long candidate lists o_max is 1 - 5 % of the npc * N values of
// a pair, i.e. ~95 % of the values can never be the answer.  So instead of parking every
// value (as a 16-bit key) in shared memory and bisecting over all of them
// (igmk_actdist.cuh), this kernel
//   1. sample: computes the first chunk of every thread (the first 4 * G structures), and
//      picks a threshold T from the sample's order statistics with a safety margin
//      (bisection on the 16-bit keys held in registers);  T_eff = max(T, rcutsq);
//   2. fill: streams all chunks and appends only the values <= T_eff, in FULL float32, to a
//      short thread-private list in shared memory (one predicated store per hit).  The rows
//      of locus j arrive through a per-warp ring of shared-memory stages filled by bulk
//      asynchronous copies (cp.async.bulk + mbarrier: one elected lane issues two 1536-byte
//      copies per stage; the stages of the NEXT pair are requested before the current pair's
//      select phase, so their L2 latency is hidden behind it);
//   3. select: contact count = #{list <= rcutsq} (T_eff >= rcutsq), p and o in float64, then
//      the o-th smallest of the list: bisection in value space over the thread-private
//      lists until <= 32 candidates remain, exact rank by one warp.
// Exactness does not depend on the sample: the list is complete for values <= T_eff, so if
// it holds at least o + 1 values the o-th smallest of the list IS the o-th smallest of the
// population.  Pairs for which that fails (sample not representative, thread list full,
// large o, many contacts) are appended to a redo list and processed by the key-array
// kernels of igmk_actdist.cuh in a second launch.
#pragma once
#include "igmk_actdist.cuh"

namespace igmk {

#ifndef IGMK_LCAP
#define IGMK_LCAP 24              // guaranteed list entries per thread
#endif
#ifndef IGMK_RING
#define IGMK_RING 2               // stages of the locus-j ring per warp (power of two)
#endif
#ifndef IGMK_LWPB
#define IGMK_LWPB 20              // warps per CTA of the warp-group kernel
#endif
#ifndef IGMK_LBT
#define IGMK_LBT 320              // threads per CTA of the CTA-group kernel
#endif
constexpr int kListCap = IGMK_LCAP;
constexpr int kListSpare = 4;                       // the bound is checked every 4 values
constexpr int kListSlots = kListCap + kListSpare;
constexpr int kRing = IGMK_RING;
constexpr uint32_t kRowSegBytes = kSegFloats * 4u;  // x | y | z of 128 structures of one bead
constexpr uint32_t kStageBytes = 2u * kRowSegBytes; // both copies of locus j
constexpr int kListWarps = IGMK_LWPB;
constexpr int kListBlockThreads = IGMK_LBT;
static_assert((kRing & (kRing - 1)) == 0 && kRing >= 1, "ring depth must be a power of two");

// ------------------------------------------------------------ mbarrier / bulk copies
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0u;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1 << 22)) __trap();          // protocol error: never hang the GPU
    }
}
// global -> shared bulk copy (TMA engine, 1-D), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Per-warp ring of locus-j stages.  Stage k of the warp's issue sequence lives in slot
// k % kRing and completes phase (k / kRing) & 1 of that slot's mbarrier; stages are
// consumed in issue order, so two counters describe the whole state.
struct Ring {
    uint32_t base;        // shared address of slot 0
    uint32_t bar;         // shared address of mbarrier 0
    uint32_t issued, consumed;
    __device__ __forceinline__ void issue(const float* b0seg, const float* b1seg, int lane) {
        const uint32_t s = issued & (uint32_t)(kRing - 1);
        if (lane == 0) {
            const uint32_t bar_s = bar + 8u * s, dst = base + s * kStageBytes;
            mbar_expect_tx(bar_s, kStageBytes);
            bulk_g2s(dst, b0seg, kRowSegBytes, bar_s);
            bulk_g2s(dst + kRowSegBytes, b1seg, kRowSegBytes, bar_s);
        }
        ++issued;
    }
    __device__ __forceinline__ uint32_t acquire() {
        const uint32_t s = consumed & (uint32_t)(kRing - 1);
        mbar_wait(bar + 8u * s, (consumed / (uint32_t)kRing) & 1u);
        ++consumed;
        return base + s * kStageBytes;
    }
};

// The pair a group will work on next: enough to request its first locus-j stages.
struct NextPair {
    long long slot;       // position in processing order, -1: none
    int b0, b1;           // beads of locus j (-1, -1: the pair needs no fill)
};

__device__ __forceinline__ NextPair peek_pair(const ActdistParams& P, long long slot) {
    NextPair n;
    n.slot = slot; n.b0 = -1; n.b1 = -1;
    if (slot < 0) return n;
    const long long pair = P.perm ? (long long)__ldg(P.perm + slot) : slot;
    const int i = __ldg(P.pi + pair), j = __ldg(P.pj + pair);
    if (i != j && i >= 0 && j >= 0 && i < P.n_hap && j < P.n_hap) {
        const int4 hb = __ldg(reinterpret_cast<const int4*>(P.hap + j));
        n.b0 = hb.x;
        n.b1 = (hb.y >= 0) ? hb.y : hb.x;
    }
    return n;
}

template <int NW>
__device__ __forceinline__ int count_le_regs(const uint32_t (&kw)[NW], uint32_t piv2) {
    uint32_t a0 = 0u, a1 = 0u;
#pragma unroll
    for (int w = 0; w < NW; w += 2) {
        a0 -= bf2_le_mask(kw[w], piv2);
        a1 -= bf2_le_mask(kw[w + 1], piv2);
    }
    const uint32_t acc = a0 + a1;
    return ((int)acc >> 16) + 2 * (int)(acc & 0xffffu);
}

// Threshold from the group's first chunks (the first `ns` structures).  Returns false when
// the list such a threshold produces is expected to exceed the budget (large o or many
// contacts): the pair then goes to the key-array kernel.
template <bool BLOCK, int NS>
__device__ __forceinline__ bool sample_threshold(const ActdistParams& P, Group<BLOCK>& g, const PairDesc& d,
                                                 const float (&s)[4][NS], int omax, uint32_t& T_bits) {
    uint32_t kw[2 * NS];
    uint32_t lmn = 0x7fff7fffu, lmx = 0x7fff7fffu;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
#pragma unroll
        for (int qh = 0; qh < 2; ++qh)
            kw[2 * k + qh] = __byte_perm(__float_as_uint(s[2 * qh][k]), __float_as_uint(s[2 * qh + 1][k]), 0x7632);
        lmn = bf2_min(lmn, bf2_min(kw[2 * k], kw[2 * k + 1]));     // NaN halves are ignored
        lmx = bf2_max(lmx, bf2_max(kw[2 * k], kw[2 * k + 1]));
    }
    uint32_t kmin = min(lmn & 0xffffu, lmn >> 16);
    uint32_t mxl = lmx & 0xffffu, mxh = lmx >> 16;
    mxl = (mxl > 0x7f80u) ? 0u : mxl;
    mxh = (mxh > 0x7f80u) ? 0u : mxh;
    uint32_t kmax = max(mxl, mxh);
    int zero = 0;
    g.sum_min_max(zero, kmin, kmax);
    if (kmax < kmin) kmax = kmin;

    const int ns = min(P.nstruct, 4 * g.nthr);
    const float Sv = (float)(d.keep * ns), M = (float)(d.keep * P.nstruct);
    int k_t;
    if (ns == P.nstruct) {
        k_t = omax + 1;                                // the sample is the population
    } else {
        // smallest k with  k - z sqrt(k) >= r0  (r0: expected sample rank of element o_max)
        const float r0 = (float)(omax + 1) * Sv / M, z = P.list_z;
        k_t = (int)(r0 + 0.5f * z * z + 0.5f * z * sqrtf(z * z + 4.f * r0)) + 1;
    }
    if ((float)k_t * M > P.list_budget * Sv) return false;
    const int k_hi = k_t + max(1, k_t >> 3);
    uint32_t lo = kmin, hi = kmax;
    int ch = d.keep * ns;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        const int c = g.sum(count_le_regs(kw, mid | (mid << 16)));
        if (c >= k_t) {
            hi = mid; ch = c;
            if (c <= k_hi) break;
        } else {
            lo = mid + 1;
        }
    }
    uint32_t T = (hi << 16) | 0xffffu;
    const uint32_t rcb = __float_as_uint(d.rcutsq);
    if (rcb > T) {                                     // the list must hold every contact
        T = rcb;
        const uint32_t rk = rcb >> 16;
        ch = g.sum(count_le_regs(kw, rk | (rk << 16)));
    }
    if ((float)ch * M > P.list_budget * Sv) return false;
    T_bits = T;
    return true;
}

// Fill of one pair.  Returns 0 when the thread lists are complete for values <= T_bits,
// 1 when the pair has to be redone by the key-array kernel.  `pre`: stages of this pair
// already requested; on return `pre` = stages of `nxt` requested.
template <int SH, bool AS, bool BLOCK>
__device__ __forceinline__ int fill_list(const ActdistParams& P, Group<BLOCK>& g, const PairDesc& d,
                                         const PairPtrs& pp, int V, uint32_t as0, uint32_t as1,
                                         Ring& ring, int& pre, const NextPair& nxt,
                                         uint32_t lbase, uint32_t lstride, int omax,
                                         uint32_t& T_bits, int& mycnt, bool& ovf) {
    constexpr int NS = (SH == SH_FULL4) ? 4 : (SH == SH_INTRA2 || SH == SH_GP4) ? 2 : 4;
    const float qnan = __int_as_float(0x7fffffff);
    const u64 nz = P.negzero2;
    const int tid = g.tid, lane = tid & 31, warp = tid >> 5, nw = g.nthr >> 5;
    const int nseg = P.npad / kSeg;
    const int Vw = (nseg > warp) ? (nseg - warp + nw - 1) / nw : 0;   // this warp's segments
    const size_t seg0 = (size_t)warp * kSegFloats, vstride = (size_t)nw * kSegFloats;
    const float* pa0 = pp.A0 + seg0 + (size_t)lane * 4;
    const float* pa1 = pp.A1 + seg0 + (size_t)lane * 4;
    uint32_t sa0 = as0 + (uint32_t)(seg0 + (size_t)lane * 4) * 4u;
    uint32_t sa1 = as1 + (uint32_t)(seg0 + (size_t)lane * 4) * 4u;
    const float* jb0 = pp.B0 + seg0;                   // segment bases of locus j (warp-uniform)
    const float* jb1 = pp.B1 + seg0;
    const size_t row = (size_t)3 * P.npad;
    const float* nb0 = (nxt.b0 >= 0) ? P.coords + (size_t)nxt.b0 * row + seg0 : nullptr;
    const float* nb1 = (nxt.b0 >= 0) ? P.coords + (size_t)nxt.b1 * row + seg0 : nullptr;

    for (int k = pre; k < kRing && k < Vw; ++k) ring.issue(jb0 + (size_t)k * vstride, jb1 + (size_t)k * vstride, lane);
    pre = 0;

    uint32_t lptr = lbase;
    const uint32_t llimit = lbase + (uint32_t)kListCap * lstride;
    float T = 0.f;
    ovf = false;
    int status = 0;
#pragma unroll 1
    for (int v = 0; v < V; ++v) {
        const bool live = v < Vw;                      // warp-uniform
        const int c = tid + v * g.nthr;
        float s[4][NS];
        if (live) {
            const uint32_t st = ring.acquire() + (uint32_t)lane * 16u;
            Row6 a0, a1;
            const Row6 b0 = load_row6_shared(st), b1 = load_row6_shared(st + kRowSegBytes);
            if (AS) {
                a0 = load_row6_shared(sa0); a1 = load_row6_shared(sa1);
            } else {
                a0 = load_row6<LD_KEEP>(pa0); a1 = load_row6<LD_KEEP>(pa1);
            }
            chunk_values<SH, NS>(d, P.mode, a0, a1, b0, b1, nz, s);
            if (4 * c + 4 > P.nstruct) {               // tail / padding chunk of the population
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (4 * c + q >= P.nstruct) {
#pragma unroll
                        for (int k = 0; k < NS; ++k) s[q][k] = qnan;
                    }
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int k = 0; k < NS; ++k) s[q][k] = qnan;
        }
        if (v == 0) {                                  // uniform over the group
            if (!sample_threshold<BLOCK, NS>(P, g, d, s, omax, T_bits)) {
                // hand the pair over; the stages already requested must still land
                const int out = min(kRing, Vw) - (live ? 1 : 0);
                for (int k = 0; k < out; ++k) ring.acquire();
                status = 1;
                break;
            }
            T = __uint_as_float(T_bits);
        }
        if (live) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int k = 0; k < NS; ++k) {
                    if (s[q][k] <= T) {
                        sts32(lptr, __float_as_uint(s[q][k]));
                        lptr += lstride;
                    }
                }
                if (NS == 4 || (q & 1)) {              // every 4 values: keep inside the spare slots
                    if (lptr > llimit) { ovf = true; lptr = llimit; }
                }
            }
            __syncwarp();                              // every lane has consumed the stage
            const int kk = v + kRing;
            if (kk < Vw) {
                ring.issue(jb0 + (size_t)kk * vstride, jb1 + (size_t)kk * vstride, lane);
            } else if (nb0 != nullptr && pre < Vw) {   // the freed slot takes the next pair's next stage
                ring.issue(nb0 + (size_t)pre * vstride, nb1 + (size_t)pre * vstride, lane);
                ++pre;
            }
        }
        pa0 += vstride; pa1 += vstride;
        sa0 += (uint32_t)vstride * 4u; sa1 += (uint32_t)vstride * 4u;
    }
    mycnt = (int)((lptr - lbase) / lstride);
    return status;
}

__device__ __forceinline__ int list_count_le(uint32_t lbase, uint32_t lstride, int n, int piv) {
    int c = 0;
#pragma unroll 4
    for (int k = 0; k < n; ++k) c += ((int)lds32(lbase + (uint32_t)k * lstride) <= piv) ? 1 : 0;
    return c;
}

// p, o and the o-th smallest value from the thread lists.  Returns false when the lists
// cannot answer (overflow, or fewer than o + 1 values: the sample misjudged the pair).
template <bool BLOCK>
__device__ __forceinline__ bool select_list(const ActdistParams& P, Group<BLOCK>& g, long long pair,
                                            const PairDesc& d, uint32_t lbase, uint32_t lstride,
                                            int mycnt, bool ovf, uint32_t T_bits) {
    const int n = g.sum(mycnt + (ovf ? (1 << 20) : 0));       // lists hold < 2^20 values
    if (n >> 20) return false;
    const int rcb = (int)__float_as_uint(d.rcutsq), tb = (int)T_bits;
    // every value <= rcutsq is in the list (T >= rcutsq)
    const int cnt = (rcb >= tb) ? n : g.sum(list_count_le(lbase, lstride, mycnt, rcb));
    double p;
    int o;
    compute_p_o(cnt, d.keep, P.nstruct, __ldg(P.pwish + pair), __ldg(P.plast + pair), P.it_corr, p, o);
    if (o < 0) {
        emit_result(P, g.tid, pair, d, 0u, cnt, -1, 0.0, 0.0);
        return true;
    }
    if (n < o + 1) return false;
    // bracket (lo, hi] in bit-pattern order (non-negative floats: same as value order)
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
        const int c = g.sum(list_count_le(lbase, lstride, mycnt, mid));
        if (c > o) { hi = mid; ch = c; } else { lo = mid; cb = c; }
        ++pass;
    }
    if (ch - cb > kRankCap) {                          // hi == lo + 1: all candidates are `hi`
        emit_result(P, g.tid, pair, d, (uint32_t)hi, cnt, o, p, 0.0);
        return true;
    }
    g.sync();                                          // candidate counter = 0 is visible
    for (int k = 0; k < mycnt; ++k) {
        const int x = (int)lds32(lbase + (uint32_t)k * lstride);
        if (x > lo && x <= hi) {
            const uint32_t slot = atoms_inc(g.ctl);
            if (slot < (uint32_t)kRankCap) sts32(g.list + slot * 4, (uint32_t)x);
        }
    }
    g.sync();
    if (g.leader_warp()) {
        const uint32_t ans = warp_select(g.list, ch - cb, o - cb, g.tid & 31);
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

// One pair in list form.  `cur` was peeked before (its first stages may be in flight:
// `pre`); `nxt` is the group's next pair.
template <bool BLOCK>
__device__ __forceinline__ void process_pair_list(const ActdistParams& P, Group<BLOCK>& g, int V,
                                                  const NextPair& cur, const NextPair& nxt, Ring& ring,
                                                  int& pre, const TileCtl& tile,
                                                  uint32_t lbase, uint32_t lstride) {
    const long long pair = P.perm ? (long long)__ldg(P.perm + cur.slot) : cur.slot;
    const int i = __ldg(P.pi + pair);
    const PairDesc d = make_pair_desc(P, i, __ldg(P.pj + pair));
    if (!d.valid) {                        // uniform over the group; nothing was requested
        emit_empty(P, g.tid, pair);
        pre = 0;
        return;
    }
    const PairPtrs pp = pair_ptrs(P, d);
    // o <= o_max: p <= pwish for pwish <= 1 (the post-check in select_list does not rely on it)
    int omax;
    {
        const int total = d.keep * P.nstruct;
        const double x = __dmul_rn(__dmul_rn((double)d.keep, __ldg(P.pwish + pair)), (double)P.nstruct);
        const double r = rint(x);
        omax = (r >= (double)(total - 1)) ? (total - 1) : ((r > 0.0) ? (int)r : 0);
    }
    if (g.tid == 0) sts32(g.ctl, 0u);
    uint32_t T_bits = 0u;
    int mycnt = 0, status;
    bool ovf = false;
    const int tslot = (!BLOCK && tile.base) ? tile_acquire(P, tile, i, d, pp, g.tid) : -1;
    if (tslot >= 0) {
        const uint32_t as0 = tile.base + (uint32_t)tslot * tile.slot_bytes;
        const uint32_t as1 = (d.a1 >= 0) ? as0 + (tile.slot_bytes >> 1) : as0;
        switch (pair_shape(d, P.mode)) {       // uniform over the group
            case SH_FULL4:  status = fill_list<SH_FULL4, true, BLOCK>(P, g, d, pp, V, as0, as1, ring, pre, nxt, lbase, lstride, omax, T_bits, mycnt, ovf); break;
            case SH_INTRA2: status = fill_list<SH_INTRA2, true, BLOCK>(P, g, d, pp, V, as0, as1, ring, pre, nxt, lbase, lstride, omax, T_bits, mycnt, ovf); break;
            case SH_GP4:    status = fill_list<SH_GP4, true, BLOCK>(P, g, d, pp, V, as0, as1, ring, pre, nxt, lbase, lstride, omax, T_bits, mycnt, ovf); break;
            default:        status = fill_list<SH_GENERIC, true, BLOCK>(P, g, d, pp, V, as0, as1, ring, pre, nxt, lbase, lstride, omax, T_bits, mycnt, ovf); break;
        }
        tile_release(tile, tslot, g.tid);
    } else {
        switch (pair_shape(d, P.mode)) {
            case SH_FULL4:  status = fill_list<SH_FULL4, false, BLOCK>(P, g, d, pp, V, 0u, 0u, ring, pre, nxt, lbase, lstride, omax, T_bits, mycnt, ovf); break;
            case SH_INTRA2: status = fill_list<SH_INTRA2, false, BLOCK>(P, g, d, pp, V, 0u, 0u, ring, pre, nxt, lbase, lstride, omax, T_bits, mycnt, ovf); break;
            case SH_GP4:    status = fill_list<SH_GP4, false, BLOCK>(P, g, d, pp, V, 0u, 0u, ring, pre, nxt, lbase, lstride, omax, T_bits, mycnt, ovf); break;
            default:        status = fill_list<SH_GENERIC, false, BLOCK>(P, g, d, pp, V, 0u, 0u, ring, pre, nxt, lbase, lstride, omax, T_bits, mycnt, ovf); break;
        }
    }
    if (status != 0 || !select_list<BLOCK>(P, g, pair, d, lbase, lstride, mycnt, ovf, T_bits))
        push_redo(P, g.tid, pair);
}

// ---------------------------------------------------------------- kernels
// G = 32: one pair per warp, no CTA barrier after the set-up.  Shared memory per warp:
// kListSlots x 128 bytes of thread lists and kRing x 3072 bytes of locus-j stages; one
// locus-i tile per CTA (igmk_actdist.cuh).
__global__ void __launch_bounds__(32 * kListWarps, 1)
actdist_list_warp_kernel(const ActdistParams P, const int V) {
    extern __shared__ uint4 s_dyn[];              // [warp] rings | [warp] lists | locus-i tile
    __shared__ __align__(8) unsigned long long s_bar[kListWarps][kRing];
    __shared__ uint32_t s_list[kListWarps][kWarpListCap];
    __shared__ uint32_t s_cnt[kListWarps];
    __shared__ uint32_t s_slot[2];
    __shared__ unsigned int s_ticket;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    Group<false> g;
    g.tid = lane;
    g.nthr = 32;
    g.list = smem_addr(&s_list[warp][0]);
    g.ctl = smem_addr(&s_cnt[warp]);
    g.cap = kWarpListCap;
    g.kscr = 0u; g.kstride = 0u; g.red = 0u; g.list2 = 0u; g.parity = 0;
    const uint32_t dyn0 = smem_addr(s_dyn);
    Ring ring;
    ring.base = dyn0 + (uint32_t)warp * (uint32_t)kRing * kStageBytes;
    ring.bar = smem_addr(&s_bar[warp][0]);
    ring.issued = 0u; ring.consumed = 0u;
    const uint32_t lists0 = dyn0 + (uint32_t)nwarps * (uint32_t)kRing * kStageBytes;
    const uint32_t lbase = lists0 + (uint32_t)warp * (uint32_t)kListSlots * 128u + (uint32_t)lane * 4u;
    TileCtl tile;
    tile.base = 0u; tile.slot_bytes = 24u * (uint32_t)P.npad; tile.words = smem_addr(s_slot); tile.nslots = 1;
    if (P.tile_block > 0 && P.tile_slots > 0)
        tile.base = lists0 + (uint32_t)nwarps * (uint32_t)kListSlots * 128u;
    if (threadIdx.x == 0) {
        s_slot[0] = (0xfffffu << 12) | (TS_EMPTY << 10);
        s_slot[1] = (0xfffffu << 12) | (TS_EMPTY << 10);
        s_ticket = 0u;
    }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kRing; ++s) mbar_init(ring.bar + 8u * s, 1u);
        mbar_fence_init();
    }
    __syncthreads();
    // CTA-contiguous blocks of B pairs (block k of this CTA = list block blockIdx + k *
    // gridDim), handed to the warps one pair at a time by a shared ticket counter:
    // consecutive pairs share locus i, fast and slow pairs balance out across the warps.
    const unsigned int B = (unsigned int)max(P.tile_block, 1);
    auto draw = [&]() -> long long {
        for (;;) {
            unsigned int t = 0u;
            if (lane == 0) t = atomicAdd(&s_ticket, 1u);
            t = __shfl_sync(0xffffffffu, t, 0);
            const unsigned int k = t / B, r = t - k * B;
            const long long base = ((long long)blockIdx.x + (long long)k * gridDim.x) * B;
            if (base >= P.n_pairs) return -1;
            if (base + r < P.n_pairs) return base + r;
        }
    };
    int pre = 0;
    NextPair nxt = peek_pair(P, draw());
    while (nxt.slot >= 0) {
        const NextPair cur = nxt;
        nxt = peek_pair(P, draw());
        process_pair_list<false>(P, g, V, cur, nxt, ring, pre, tile, lbase, 128u);
        __syncwarp();
    }
}

// G = blockDim.x: one pair per CTA, two CTAs per SM (N > 1024).  Every warp runs its own
// ring over its own segments of the locus-j rows; the rows of locus i come through L1.
__global__ void __launch_bounds__(kListBlockThreads, 2)
actdist_list_block_kernel(const ActdistParams P, const int V) {
    extern __shared__ uint4 s_dyn[];              // [warp] rings | [slot][thread] lists
    __shared__ __align__(8) unsigned long long s_bar[kListBlockThreads / 32][kRing];
    __shared__ uint32_t s_list[kRankCap];
    __shared__ uint32_t s_cnt;
    __shared__ uint32_t s_red[2 * 96];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
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
    const uint32_t dyn0 = smem_addr(s_dyn);
    Ring ring;
    ring.base = dyn0 + (uint32_t)warp * (uint32_t)kRing * kStageBytes;
    ring.bar = smem_addr(&s_bar[warp][0]);
    ring.issued = 0u; ring.consumed = 0u;
    const uint32_t lbase = dyn0 + (uint32_t)nwarps * (uint32_t)kRing * kStageBytes + (uint32_t)threadIdx.x * 4u;
    const uint32_t lstride = (uint32_t)blockDim.x * 4u;
    TileCtl tile;
    tile.base = 0u; tile.slot_bytes = 0u; tile.words = 0u; tile.nslots = 0;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kRing; ++s) mbar_init(ring.bar + 8u * s, 1u);
        mbar_fence_init();
    }
    __syncthreads();
    int pre = 0;
    long long slot = blockIdx.x;
    NextPair nxt = peek_pair(P, (slot < P.n_pairs) ? slot : -1);
    while (nxt.slot >= 0) {
        const NextPair cur = nxt;
        slot += gridDim.x;
        nxt = peek_pair(P, (slot < P.n_pairs) ? slot : -1);
        process_pair_list<true>(P, g, V, cur, nxt, ring, pre, tile, lbase, lstride);
        __syncthreads();
    }
}

}  // namespace igmk
