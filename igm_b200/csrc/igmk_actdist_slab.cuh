// igmk_actdist_slab.cuh - K1 for populations larger than 1024 structures (sm_100a):
// the list form of igmk_actdist_list.cuh, run slab by slab.
//
// For N <= 1024 one warp holds a whole pair (igmk_actdist_list.cuh).  For larger N the
// one-CTA-per-pair kernels re-read both loci of every pair from L2 (480 KB per pair at
// N = 10 000: the L2 -> SM path is their limiter) and pay a CTA barrier per reduction.
// Here the population is cut into slabs of 1024 structures and the SAME warp-per-pair fill
// runs once per (pair, slab): inside a slab the rows of locus i of consecutive pairs come
// from the CTA's shared-memory tile again (24 KB per locus and slab), the J-block order
// keeps a slab's locus-j rows in L2, and nothing ever waits on a CTA barrier.
//   A. sample kernel: threshold T of every pair from its first 128 structures
//      (sample_threshold, same margin rule), or "not for this path";
//   B. fill kernel: tasks (slab, block of pairs); a warp appends the values <= T of its
//      (pair, slab) to thread lists in shared memory, then flushes them to the pair's list in
//      global memory (one atomic reservation per warp);
//   C. select kernel: one warp per pair - contact count, p, o, bisection in value space over
//      the pair's global list (narrowed into shared memory once few candidates remain),
//      exact rank across the lanes.
// The exactness argument is that of the list form: a pair's list holds every value <= T;
// pairs whose list overflows or holds fewer than o + 1 values go to the key-array kernels.
// The host runs A - C over batches of kSlabBatch pairs (igmk.cu).
#pragma once
#include "igmk_actdist_list.cuh"

namespace igmk {

#ifndef IGMK_SWPB
#define IGMK_SWPB 24              // warps per CTA of the slab fill kernel (78 registers, no spills; 20 / 22 / 24 warps: 23.0 / 23.7 / 24.1 M pairs/s at N = 10 000)
#endif
constexpr int kSlabWarps = IGMK_SWPB;
constexpr int kSlabSegs = 8;                  // segments of 128 structures per slab
constexpr int kSlabChunks = kSlabSegs * 32;   // float4 chunks per slab and bead row
constexpr int kSlabCap = 8192;                // list words per pair in global memory
constexpr int kSlabBatch = 131072;            // pairs per batch (lists: 4 GiB; 32 k / 64 k / 128 k: 21.0 / 26.5 / 27.2 M pairs/s)
constexpr int kSelBuf = 256;                  // select: candidates narrowed into shared memory, then registers
constexpr uint32_t kNoList = 0xffffffffu;     // T value of a pair that does not take this path

struct SlabParams {
    uint32_t* T;              // [batch] threshold (float32 pattern) or kNoList
    unsigned int* cnt;        // [batch] values appended so far; > kSlabCap: overflow
    uint32_t* lists;          // [batch][kSlabCap]
    long long slot0;          // first slot (processing order) of the batch
    int nslots;               // pairs in the batch
    int nslab;                // slabs of the population
    int nblk;                 // blocks of 2^bshift pairs in the batch
    int bshift;
};

// ------------------------------------------------------------------ A: sample
template <int SH>
__device__ __forceinline__ void first_chunk_keys(const ActdistParams& P, const PairDesc& d, const PairPtrs& pp,
                                                 int lane, uint32_t (&kw)[8]) {
    constexpr int NS = (SH == SH_FULL4) ? 4 : (SH == SH_INTRA2 || SH == SH_GP4) ? 2 : 4;
    const float qnan = __int_as_float(0x7fffffff);
    float s[4][NS];
    {
        const size_t off = (size_t)lane * 4;
        const Row6 a0 = load_row6<LD_PLAIN>(pp.A0 + off), a1 = load_row6<LD_PLAIN>(pp.A1 + off);
        const Row6 b0 = load_row6<LD_PLAIN>(pp.B0 + off), b1 = load_row6<LD_PLAIN>(pp.B1 + off);
        chunk_values<SH, NS>(d, P.mode, a0, a1, b0, b1, P.negzero2, s);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (4 * lane + q >= P.nstruct) {
#pragma unroll
            for (int k = 0; k < NS; ++k) s[q][k] = qnan;
        }
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int qh = 0; qh < 2; ++qh)
            kw[2 * k + qh] = (k < NS) ? __byte_perm(__float_as_uint(s[2 * qh][k < NS ? k : 0]),
                                                     __float_as_uint(s[2 * qh + 1][k < NS ? k : 0]), 0x7632)
                                      : 0x7fff7fffu;
}

__global__ void __launch_bounds__(256)
slab_sample_kernel(const ActdistParams P, const SlabParams S) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int w = blockIdx.x * wpb + (threadIdx.x >> 5); w < S.nslots; w += gridDim.x * wpb) {
        long long pair;
        int omax;
        const PairDesc d = load_pairrec(P.rec, S.slot0 + w, pair, omax);
        uint32_t T = kNoList;
        if (!d.valid) {
            emit_empty(P, lane, pair);
        } else {
            const PairPtrs pp = pair_ptrs(P, d);
            uint32_t kw[8];
            switch (pair_shape(d, P.mode)) {
                case SH_FULL4:  first_chunk_keys<SH_FULL4>(P, d, pp, lane, kw); break;
                case SH_INTRA2: first_chunk_keys<SH_INTRA2>(P, d, pp, lane, kw); break;
                case SH_GP4:    first_chunk_keys<SH_GP4>(P, d, pp, lane, kw); break;
                default:        first_chunk_keys<SH_GENERIC>(P, d, pp, lane, kw); break;
            }
            const SampleOut so = sample_threshold<false>(make_uint4(kw[0], kw[1], kw[2], kw[3]),
                                                         make_uint4(kw[4], kw[5], kw[6], kw[7]), lane, 32, 0u,
                                                         d.keep, P.nstruct, omax,
                                                         __float_as_uint(d.rcutsq), P.list_z, P.list_budget);
            if (so.ok && so.T_bits < 0x7f800000u) T = so.T_bits;
            else push_redo(P, lane, pair);
        }
        if (lane == 0) {
            S.T[w] = T;
            S.cnt[w] = 0u;
        }
        __syncwarp();
    }
}

// -------------------------------------------------------------------- B: fill
__global__ void __launch_bounds__(32 * kSlabWarps, 1)
slab_fill_kernel(const ActdistParams P, const SlabParams S) {
    extern __shared__ uint4 s_dyn[];              // [thread] lists | locus-i tiles
    __shared__ TileShared s_tile;
    __shared__ unsigned int s_ticket;
    const int lane = threadIdx.x & 31;
    Group<false> g;
    g.tid = lane; g.nthr = 32;
    g.list = 0u; g.ctl = 0u; g.cap = 0; g.kscr = 0u; g.kstride = 0u; g.red = 0u; g.list2 = 0u; g.parity = 0;
    const uint32_t dyn0 = smem_addr(s_dyn);
    const uint32_t lbase = dyn0 + (uint32_t)threadIdx.x * kListBytes;
    TileCtl tile;
    tile.base = 0u; tile.slot_bytes = 2u * (uint32_t)kSlabSegs * kSegFloats * 4u; tile.nslots = P.tile_slots;
    tile_bind(&s_tile, tile);
    if (P.tile_slots > 0) tile.base = dyn0 + (uint32_t)blockDim.x * kListBytes;
    if (threadIdx.x == 0) {
        tile_init(&s_tile);
        s_ticket = 0u;
    }
    __syncthreads();
    const int nseg = P.npad / kSeg;
    const size_t row = (size_t)3 * P.npad;
    // tasks (slab, pair block), slab-major: task k of this CTA = blockIdx + k * gridDim; the
    // pairs of a task are handed to the warps by one ticket counter (consecutive pairs share
    // locus i: the tile is keyed by (locus, slab))
    const unsigned int bshift = (unsigned int)S.bshift;
    const long long ntask = (long long)S.nslab * S.nblk;
    for (;;) {
        unsigned int t = 0u;
        if (lane == 0) t = atomicAdd(&s_ticket, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        const long long task = (long long)blockIdx.x + (long long)(t >> bshift) * gridDim.x;
        if (task >= ntask) break;
        const int slab = (int)(task / S.nblk), blk = (int)(task - (long long)slab * S.nblk);
        const int w = (blk << bshift) + (int)(t & ((1u << bshift) - 1u));
        if (w >= S.nslots) continue;
        uint32_t T_bits = __ldg(S.T + w);
        if (T_bits == kNoList) continue;
        long long pair;
        int omax_unused;
        const PairDesc d = load_pairrec(P.rec, S.slot0 + w, pair, omax_unused);
        const int i = d.a0;                              // tile key: (bead of copy 0, slab)
        const int Vs = min(kSlabSegs, nseg - slab * kSlabSegs);
        const size_t soff = (size_t)slab * kSlabSegs * kSegFloats;
        PairPtrs pp;
        pp.A0 = P.coords + (size_t)d.a0 * row + soff;
        pp.B0 = P.coords + (size_t)d.b0 * row + soff;
        pp.A1 = P.coords + (size_t)(d.a1 >= 0 ? d.a1 : d.a0) * row + soff;
        pp.B1 = P.coords + (size_t)(d.b1 >= 0 ? d.b1 : d.b0) * row + soff;
        int mycnt = 0;
        bool ovf = false;
        const int c0 = slab * kSlabChunks;
        const int tslot = tile.base ? tile_acquire(P, tile, i * S.nslab + slab, d, pp, lane) : -1;
        if (tslot >= 0) {
            const uint32_t as0 = tile.base + (uint32_t)tslot * tile.slot_bytes;
            const uint32_t as1 = (d.a1 >= 0) ? as0 + (tile.slot_bytes >> 1) : as0;
            switch (pair_shape(d, P.mode)) {
                case SH_FULL4:  fill_list<SH_FULL4, true, false, false>(P, g, d, pp, Vs, as0, as1, lbase, 0u, 0, T_bits, mycnt, ovf, c0); break;
                case SH_INTRA2: fill_list<SH_INTRA2, true, false, false>(P, g, d, pp, Vs, as0, as1, lbase, 0u, 0, T_bits, mycnt, ovf, c0); break;
                case SH_GP4:    fill_list<SH_GP4, true, false, false>(P, g, d, pp, Vs, as0, as1, lbase, 0u, 0, T_bits, mycnt, ovf, c0); break;
                default:        fill_list<SH_GENERIC, true, false, false>(P, g, d, pp, Vs, as0, as1, lbase, 0u, 0, T_bits, mycnt, ovf, c0); break;
            }
            tile_release(tile, tslot, lane);
        } else {
            switch (pair_shape(d, P.mode)) {
                case SH_FULL4:  fill_list<SH_FULL4, false, false, false>(P, g, d, pp, Vs, 0u, 0u, lbase, 0u, 0, T_bits, mycnt, ovf, c0); break;
                case SH_INTRA2: fill_list<SH_INTRA2, false, false, false>(P, g, d, pp, Vs, 0u, 0u, lbase, 0u, 0, T_bits, mycnt, ovf, c0); break;
                case SH_GP4:    fill_list<SH_GP4, false, false, false>(P, g, d, pp, Vs, 0u, 0u, lbase, 0u, 0, T_bits, mycnt, ovf, c0); break;
                default:        fill_list<SH_GENERIC, false, false, false>(P, g, d, pp, Vs, 0u, 0u, lbase, 0u, 0, T_bits, mycnt, ovf, c0); break;
            }
        }
        // flush the thread lists to the pair's global list: one reservation per warp
        const int any_ovf = __any_sync(0xffffffffu, ovf);
        int incl = mycnt;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, s);
            if (lane >= s) incl += y;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned int base = 0u;
        if (lane == 0 && (total > 0 || any_ovf))
            base = atomicAdd(S.cnt + w, any_ovf ? (unsigned int)(kSlabCap + 1) : (unsigned int)total);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (!any_ovf && base + (unsigned int)total <= (unsigned int)kSlabCap) {
            uint32_t* dst = S.lists + (size_t)w * kSlabCap + base + (unsigned int)(incl - mycnt);
            for (int k = 0; k < mycnt; ++k) dst[k] = lds32(lbase + (uint32_t)k * 4u);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ C: select
// #{list <= piv} over a pair's list in global memory (L2); four independent loads per lane
// and iteration keep enough requests in flight (the select is latency-bound)
__device__ __forceinline__ int global_count_le(const uint32_t* list, int n, int lane, float piv) {
    u64 acc0 = 0ull, acc1 = 0ull;
    int k = lane;
    for (; k + 96 < n; k += 128) {
        const float x0 = __uint_as_float(__ldg(list + k)), x1 = __uint_as_float(__ldg(list + k + 32));
        const float x2 = __uint_as_float(__ldg(list + k + 64)), x3 = __uint_as_float(__ldg(list + k + 96));
        acc0 = f2add(acc0, f2pack(f_le_one(x0, piv), f_le_one(x1, piv)));
        acc1 = f2add(acc1, f2pack(f_le_one(x2, piv), f_le_one(x3, piv)));
    }
    for (; k < n; k += 32) acc0 = f2add(acc0, f2pack(f_le_one(__uint_as_float(__ldg(list + k)), piv), 0.f));
    float a, b, c, d;
    f2split(acc0, a, b);
    f2split(acc1, c, d);
    return __reduce_add_sync(0xffffffffu, (int)((a + b) + (c + d)));
}

__global__ void __launch_bounds__(256)
slab_select_kernel(const ActdistParams P, const SlabParams S) {
    __shared__ uint32_t s_buf[8][kSelBuf];
    __shared__ uint32_t s_cnt[8];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const uint32_t buf = smem_addr(&s_buf[wib][0]), ctl = smem_addr(&s_cnt[wib]);
    for (int w = blockIdx.x * wpb + wib; w < S.nslots; w += gridDim.x * wpb) {
        const uint32_t T_bits = __ldg(S.T + w);
        if (T_bits == kNoList) continue;                 // uniform over the warp
        long long pair;
        int omax_unused;
        const PairDesc d = load_pairrec(P.rec, S.slot0 + w, pair, omax_unused);
        const unsigned int nraw = __ldg(S.cnt + w);
        if (nraw > (unsigned int)kSlabCap) {             // a list overflowed
            push_redo(P, lane, pair);
            continue;
        }
        const int n = (int)nraw;
        const uint32_t* list = S.lists + (size_t)w * kSlabCap;
        const int rcb = (int)__float_as_uint(d.rcutsq), tb = (int)T_bits;
        const int cnt = (rcb >= tb) ? n : global_count_le(list, n, lane, d.rcutsq);
        double p;
        int o;
        compute_p_o(cnt, d.keep, P.nstruct, __ldg(P.pwish + pair), __ldg(P.plast + pair), P.it_corr, p, o);
        if (o < 0) {
            emit_result(P, lane, pair, d, 0u, cnt, -1, 0.0, 0.0);
            continue;
        }
        if (n < o + 1) {                                 // the sample misjudged the pair
            push_redo(P, lane, pair);
            continue;
        }
        int lo = -1, hi = tb, cb = 0, ch = n;
        if (rcb < tb) {
            if (cnt > o) { hi = rcb; ch = cnt; } else { lo = rcb; cb = cnt; }
        }
        // passes over the global list (L2) until the bracket fits the warp's shared buffer
        int pass = 0;
        while (ch - cb > kSelBuf && hi - lo > 1) {
            int mid;
            if (pass < 8) {
                const float fl = (lo < 0) ? 0.f : __int_as_float(lo), fh = __int_as_float(hi);
                mid = __float_as_int(fl + (fh - fl) * 0.5f);
            } else {
                mid = lo + ((hi - lo) >> 1);
            }
            mid = max(lo + 1, min(mid, hi - 1));
            const int c = global_count_le(list, n, lane, __int_as_float(mid));
            if (c > o) { hi = mid; ch = c; } else { lo = mid; cb = c; }
            ++pass;
        }
        if (ch - cb > kSelBuf) {                         // hi == lo + 1: all candidates are `hi`
            emit_result(P, lane, pair, d, (uint32_t)hi, cnt, o, p, 0.0);
            continue;
        }
        // bracket -> shared buffer (order does not matter)
        if (lane == 0) sts32(ctl, 0u);
        __syncwarp();
        for (int k = lane; k < n; k += 32) {
            const int x = (int)__ldg(list + k);
            if (x > lo && x <= hi) {
                const uint32_t pos = atoms_inc(ctl);
                if (pos < (uint32_t)kSelBuf) sts32(buf + pos * 4u, (uint32_t)x);
            }
        }
        __syncwarp();
        const int nb = ch - cb;                          // entries in the buffer
        const int cb0 = cb;                              // values below the buffer's range
        uint32_t e[kSelBuf / 32];                        // this lane's share of the buffer
#pragma unroll
        for (int t = 0; t < kSelBuf / 32; ++t) {
            const int k = lane + 32 * t;
            e[t] = (k < nb) ? lds32(buf + (uint32_t)k * 4u) : kListSentinel;
        }
        // bisection inside the buffer (registers) until <= 32 candidates remain
        while (ch - cb > kRankCap && hi - lo > 1) {
            int mid;
            if (pass < 12) {
                const float fl = (lo < 0) ? 0.f : __int_as_float(lo), fh = __int_as_float(hi);
                mid = __float_as_int(fl + (fh - fl) * 0.5f);
            } else {
                mid = lo + ((hi - lo) >> 1);
            }
            mid = max(lo + 1, min(mid, hi - 1));
            int c = 0;
#pragma unroll
            for (int t = 0; t < kSelBuf / 32; ++t) c += ((int)e[t] <= mid) ? 1 : 0;
            c = cb0 + __reduce_add_sync(0xffffffffu, c);
            if (c > o) { hi = mid; ch = c; } else { lo = mid; cb = c; }
            ++pass;
        }
        if (ch - cb > kRankCap) {                        // hi == lo + 1: all candidates are `hi`
            emit_result(P, lane, pair, d, (uint32_t)hi, cnt, o, p, 0.0);
            continue;
        }
        // the <= 32 candidates (lo, hi] -> head of the buffer (every lane holds its share in
        // registers by now), one per lane, exact rank across the lanes
        if (lane == 0) sts32(ctl, 0u);
        __syncwarp();
#pragma unroll
        for (int t = 0; t < kSelBuf / 32; ++t) {
            const int x = (int)e[t];
            if (x > lo && x <= hi) {                     // (the sentinel is above every hi)
                const uint32_t pos = atoms_inc(ctl);
                if (pos < (uint32_t)kRankCap) sts32(buf + pos * 4u, e[t]);
            }
        }
        __syncwarp();
        const uint32_t ans = warp_rank(buf, ch - cb, o - cb, lane);
        emit_result(P, lane, pair, d, ans, cnt, o, p, 0.0);
        __syncwarp();
    }
}

}  // namespace igmk
