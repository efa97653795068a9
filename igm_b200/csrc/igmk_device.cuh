// igmk_device.cuh - device-side building blocks shared by the A-step kernels.
//
// Arithmetic contract (pinned by tests against the reference's
// own get_actdist):
//   d2 = fl32(fl32(fl32(dx*dx) + fl32(dy*dy)) + fl32(dz*dz)), dx = fl32(xi - xj)
//   strictly sequential, no FMA  (np.sum(np.square(x - y), axis=1),
//   igm/steps/ActivationDistanceStep.py:418,435)
//   p / o arithmetic in float64 exactly as CPython evaluates :445-470.
//
// Coordinate layout in HBM (one block of 3 * npad floats per bead, npad a
// multiple of 128):
//   [bead][segment of 128 structures][x | y | z][128]
// so one bead's coordinates over all structures are contiguous (12 * npad
// bytes), a warp's 128-bit loads of one component are one 512-byte run, and
// the x / y / z loads of a chunk differ by the constants 512 / 1024 bytes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/igmk.h"

namespace igmk {

typedef unsigned long long u64;

constexpr int kSeg = 128;              // structures per row segment
constexpr int kSegFloats = 3 * kSeg;   // floats per segment (x, y, z rows)

// float offset of (component 0, structure s) inside one bead's block
__host__ __device__ __forceinline__ size_t coord_off(int s) {
    return (size_t)(s >> 7) * kSegFloats + (size_t)(s & (kSeg - 1));
}

// One haploid bin: bead ids of its (<= 2) copies, chromosome, radius of copy 0.
struct __align__(16) HapEntry {
    int   b0;
    int   b1;      // -1 when the bin has a single copy (male X / Y)
    int   chrom;
    float radius;  // radii[ii[0]]  (ActivationDistanceStep.py:393)
};

struct ActdistParams {
    const float*    coords;   // see layout above
    const HapEntry* hap;      // [n_hap]
    const int32_t*  pi;
    const int32_t*  pj;
    const double*   pwish;
    const double*   plast;
    igmk_pair_result* out;
    long long n_pairs;
    int   nstruct;
    int   npad;               // padded structure count, multiple of 128
    int   nchunks;            // ceil(nstruct / 4)
    int   n_hap;
    float contact_range;      // already float32 (NEP-50: python float * f32 -> f32)
    int   it_corr;
    int   mode;
    const float* pexp32;      // DamID: p_exp / plast of the float32 batch file
    const float* plast32;
    double damid_R;           // DamID: nucleus_radius * (1 - contact_range), float64
    int   zero_bead;          // DamID: index of the all-zero row behind the population
    const u64* peers;         // multi-GPU: where this rank's result slice starts in each GPU's gather buffer
    int   n_peers;            // 0: results go to `out` only
    const int32_t* perm;      // processing order (NULL: input order), see igmk.cu order_pairs()
    int   tile_slots;         // 1 or 2 locus-i tiles per CTA
    int   tile_block;         // warp kernel: pairs per CTA-contiguous block; 0 = no locus-i tile in shared memory
    unsigned int* block_counter;  // warp kernel: device-wide counter handing out pair blocks (NULL: static round-robin)
    int   block_stop;         // CTA groups: the key bisection stops at <= block_stop candidates (<= kBlockListCap)
    u64   negzero2;           // {-0.0f, -0.0f}: opaque addend of the packed squares (igmk_actdist.cuh)
    // list form (igmk_actdist_list.cuh): pairs it cannot answer are appended to `redo`; the key-array
    // kernels then take their pair count from the device (n_pairs_dev) and `redo` as processing order
    int32_t* redo;
    unsigned int* redo_count;
    const unsigned int* n_pairs_dev;
    const struct PairRec* rec;    // list form: one 32-byte descriptor per pair in processing order (build_pairrec_kernel)
    float list_z;             // safety margin of the sample threshold, in standard deviations
    float list_budget;        // expected list length beyond which a pair goes to the key-array kernel
};

// Everything the list-form kernels need to know about one pair before they touch a
// coordinate, gathered once per launch in processing order (one coalesced 32-byte read per
// pair instead of the chain  perm -> i, j, pwish -> two index entries).
struct __align__(16) PairRec {
    int32_t  pair;     // index in the caller's list
    int32_t  a0, a1;   // beads of locus i (-1: absent)
    int32_t  b0, b1;   // beads of locus j
    float    rcutsq;
    uint32_t bits;     // keep | cmask << 4 | nrec << 8 | valid << 12
    int32_t  omax;     // upper bound of the order-statistic index (p <= pwish)
};

// Combination shapes of one pair (which of the four copy combinations
// d0=(a0,b0) d1=(a0,b1) d2=(a1,b0) d3=(a1,b1) exist, in reference order).
enum : int { CM_D0 = 1, CM_D1 = 2, CM_D2 = 4, CM_D3 = 8 };

struct PairDesc {
    int   a0, a1, b0, b1;   // bead ids (-1: absent)
    int   cmask;            // combinations that are computed
    int   keep;             // n_possible_contacts: values kept per structure
    int   nrec;             // records the pair expands to when p > 0
    int   valid;            // 0: i == j / out of range -> empty result
    int   always_rec;       // DamID: records are written even when p <= 0
    float rcutsq;
};

__device__ __forceinline__ float d2_nofma(float xi, float yi, float zi,
                                          float xj, float yj, float zj) {
    const float dx = __fsub_rn(xi, xj);
    const float dy = __fsub_rn(yi, yj);
    const float dz = __fsub_rn(zi, zj);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Pair metadata, uniform over the group (ActivationDistanceStep.py:379-424,
// GP_activation.py:355-366).
__device__ __forceinline__ PairDesc make_pair_desc(const ActdistParams& P, int i, int j) {
    PairDesc d;
    d.valid = (i != j) && i >= 0 && j >= 0 && i < P.n_hap && j < P.n_hap;
    d.a0 = d.a1 = d.b0 = d.b1 = -1;
    d.cmask = 0; d.keep = 0; d.nrec = 0; d.rcutsq = 0.f; d.always_rec = 0;
    if (!d.valid) return d;
    const int4 ha = __ldg(reinterpret_cast<const int4*>(P.hap + i));
    const int4 hb = __ldg(reinterpret_cast<const int4*>(P.hap + j));
    d.a0 = ha.x; d.a1 = ha.y; d.b0 = hb.x; d.b1 = hb.y;
    const int na = (d.a1 >= 0) ? 2 : 1;
    const int nb = (d.b1 >= 0) ? 2 : 1;
    const bool intra = (ha.z == hb.z);
    const float ri = __int_as_float(ha.w), rj = __int_as_float(hb.w);
    const float rc = __fmul_rn(P.contact_range, __fadd_rn(ri, rj));
    d.rcutsq = __fmul_rn(rc, rc);
    const int all = CM_D0 | (nb == 2 ? CM_D1 : 0) | (na == 2 ? CM_D2 : 0) |
                    ((na == 2 && nb == 2) ? CM_D3 : 0);
    if (P.mode == IGMK_MODE_LB && intra) {
        // zip(ii, jj): (a0,b0),(a1,b1); the reference needs len(ii) == len(jj) here
        // (quirk q5) - the host rejects other inputs before launch.
        d.cmask = CM_D0 | ((na == 2 && nb == 2) ? CM_D3 : 0);
        d.keep = (na < nb) ? na : nb;
    } else if (P.mode == IGMK_MODE_LB) {
        d.cmask = all;
        d.keep = na * nb;
    } else {
        d.cmask = all;
        d.keep = (na < nb) ? na : nb;
    }
    d.nrec = intra ? ((na < nb) ? na : nb) : na * nb;
    return d;
}

// DamID "pair" (locus I, origin): copies of I against the all-zero bead
// (DamidActivationDistanceStep.py:424-438).  D = (R - r)^2 in float64; a value is "in
// contact with the lamina" when float64(s) / D >= 1.0 (:441,:449), which for a
// correctly rounded division is exactly s >= D - so the fill counts s <= T with T the
// largest float32 below D and the caller takes the complement.
__device__ __forceinline__ PairDesc make_damid_desc(const ActdistParams& P, int I, double& D) {
    PairDesc d;
    d.valid = I >= 0 && I < P.n_hap;
    d.a0 = d.a1 = d.b0 = d.b1 = -1;
    d.cmask = 0; d.keep = 0; d.nrec = 0; d.rcutsq = 0.f; d.always_rec = 1;
    D = 0.0;
    if (!d.valid) return d;
    const int4 ha = __ldg(reinterpret_cast<const int4*>(P.hap + I));
    d.a0 = ha.x; d.a1 = ha.y; d.b0 = P.zero_bead;
    const int na = (d.a1 >= 0) ? 2 : 1;
    d.cmask = CM_D0 | (na == 2 ? CM_D2 : 0);
    d.keep = na;
    d.nrec = na;
    const double x = __dsub_rn(P.damid_R, (double)__int_as_float(ha.w));
    D = __dmul_rn(x, x);
    float t = __double2float_rd(D);                     // largest float32 <= D
    if ((double)t >= D) t = __uint_as_float(__float_as_uint(t) - 1u);   // strictly below (D > 0)
    d.rcutsq = (D > 0.0) ? t : __int_as_float(0xff800000);             // D == 0: everything is in contact
    return d;
}

// cleanProbability + order index of get_damid_actdist_I (:445-466) with the types
// NumPy >= 2 gives them: p_exp / plast are np.float32 (float32 batch file), pnow is a
// Python float, so every operation that involves p_exp or plast is float32; only
// `1.0 - t` with t = pnow (plast >= 1 branch) is float64, rounded to float32 when it
// meets p_exp.  o_desc = -1 when p <= 0.
__device__ __forceinline__ void compute_p_o_damid(int cnt_ge, int M, float p_exp, float plast,
                                                  int it_corr, float& p, int& o_desc) {
    if (it_corr == 1) {
        const double pnow = __ddiv_rn((double)cnt_ge, (double)M);
        float num, den;
        if (plast < 1.0f) {
            float t = __fdiv_rn(__fsub_rn((float)pnow, plast), __fsub_rn(1.0f, plast));
            t = (t > 0.0f) ? t : 0.0f;                                   // max(0, t)
            if (t < 1.0f) { num = __fsub_rn(p_exp, t); den = __fsub_rn(1.0f, t); p = __fdiv_rn(num, den); }
            else p = p_exp;
        } else {
            const double t = (pnow > 0.0) ? pnow : 0.0;                  // Python float
            if (t < 1.0) { num = __fsub_rn(p_exp, (float)t); den = (float)__dsub_rn(1.0, t); p = __fdiv_rn(num, den); }
            else p = p_exp;
        }
    } else {
        p = p_exp;
    }
    o_desc = -1;
    if (p > 0.0f) {
        const float x = __fmul_rn((float)M, p);                          // n_copies * n_struct * p in float32
        const float r = rintf(x);                                        // round(): half to even
        o_desc = (r >= (float)(M - 1)) ? (M - 1) : (int)r;
        if (o_desc < 0) o_desc = 0;
    } else {
        p = 0.0f;
    }
}

// Four copy-combination values of one structure -> the `keep` kept values in
// slots s[0..3]; unused slots are NaN (never counted, never selected).
__device__ __forceinline__ void pack_slots(const PairDesc& d, int mode, float d0, float d1,
                                           float d2, float d3, float (&s)[4]) {
    const float qnan = __int_as_float(0x7fffffff);
    if (mode == IGMK_MODE_GP) {
        // keep the `keep` smallest of the existing combinations (d_sq.sort(axis=0)
        // then rows 0:npc, GP_activation.py:389-395).  fminf/fmaxf ignore NaN.
        if (d.keep == 2) {
            const float lo1 = fminf(d0, d1), hi1 = fmaxf(d0, d1);
            const float lo2 = fminf(d2, d3), hi2 = fmaxf(d2, d3);
            s[0] = fminf(lo1, lo2);
            s[1] = fminf(fmaxf(lo1, lo2), fminf(hi1, hi2));
        } else {
            s[0] = fminf(fminf(d0, d1), fminf(d2, d3));
            s[1] = qnan;
        }
        s[2] = qnan; s[3] = qnan;
    } else {
        // LB: every computed combination is kept, in enumeration order.
        const int cm = d.cmask;
        s[0] = d0;
        s[1] = (cm == (CM_D0 | CM_D3)) ? d3 : ((cm == (CM_D0 | CM_D2)) ? d2 : d1);
        s[2] = (cm == 15) ? d2 : qnan;
        s[3] = (cm == 15) ? d3 : qnan;
    }
}

// d2 of structure st between the beads whose blocks start at pa / pb (scalar loads)
__device__ __forceinline__ float d2_scalar(const float* pa, const float* pb, int st) {
    const size_t o = coord_off(st);
    return d2_nofma(__ldg(pa + o), __ldg(pa + o + kSeg), __ldg(pa + o + 2 * kSeg),
                    __ldg(pb + o), __ldg(pb + o + kSeg), __ldg(pb + o + 2 * kSeg));
}

// All kept values of structure `st` from scalar loads (slow path: candidate
// re-materialisation in GP mode, cross-check kernel).
__device__ __forceinline__ void struct_slots(const ActdistParams& P, const PairDesc& d, int st,
                                             float (&s)[4]) {
    const float qnan = __int_as_float(0x7fffffff);
    const size_t row = (size_t)3 * P.npad;
    const float* A0 = P.coords + (size_t)d.a0 * row;
    const float* B0 = P.coords + (size_t)d.b0 * row;
    const float* A1 = P.coords + (size_t)(d.a1 >= 0 ? d.a1 : d.a0) * row;
    const float* B1 = P.coords + (size_t)(d.b1 >= 0 ? d.b1 : d.b0) * row;
    const float d0 = d2_scalar(A0, B0, st);
    const float d1 = (d.cmask & CM_D1) ? d2_scalar(A0, B1, st) : qnan;
    const float d2 = (d.cmask & CM_D2) ? d2_scalar(A1, B0, st) : qnan;
    const float d3 = (d.cmask & CM_D3) ? d2_scalar(A1, B1, st) : qnan;
    pack_slots(d, P.mode, d0, d1, d2, d3, s);
}

// cleanProbability, igm/steps/ActivationDistanceStep.py:314-332 (float64).
__device__ __forceinline__ double clean_probability(double pij, double pexist) {
    double pclean = pij;
    if (pexist < 1.0) pclean = __ddiv_rn(__dsub_rn(pij, pexist), __dsub_rn(1.0, pexist));
    return (pclean > 0.0) ? pclean : 0.0;      // python max(0, x): NaN and -0.0 -> 0
}

// p and the order-statistic index o, :445-470.  o = -1 when p <= 0.
__device__ __forceinline__ void compute_p_o(int contact_count, int npc, int nstruct,
                                            double pwish, double plast, int it_corr,
                                            double& p, int& o) {
    const int total = npc * nstruct;
    if (it_corr == 1) {
        const double pnow = __ddiv_rn((double)contact_count, (double)total);
        const double t = clean_probability(pnow, plast);
        p = clean_probability(pwish, t);
    } else {
        p = pwish;
    }
    o = -1;
    if (p > 0.0) {
        // int(round(npc * p * N)): left to right, round-half-even (np.float64.__round__)
        const double x = __dmul_rn(__dmul_rn((double)npc, p), (double)nstruct);
        const double r = rint(x);
        o = (r >= (double)(total - 1)) ? (total - 1) : (int)r;
        if (o < 0) o = 0;
    } else {
        p = 0.0;
    }
}

// float32(float("%.4f" % x)) for x >= 0: the text round trip of task()/reduce()
// (:38, :230, :249).  Exact: x * 1e4 is split into hi + lo with an FMA, so the
// decimal rounding (half-even on exact ties, as CPython's dtoa does) is decided
// on the true product.
__device__ __forceinline__ float round4_to_f32(double x) {
    const double hi = __dmul_rn(x, 1e4);
    if (!(hi < 4.0e15)) return __double2float_rn(x);   // already integral at 1e-4
    const double lo = __fma_rn(x, 1e4, -hi);
    double r = rint(hi);
    const double e = __dsub_rn(hi, r);                 // exact, |e| <= 0.5
    if (e == 0.5 && lo > 0.0) r += 1.0;
    else if (e == -0.5 && lo < 0.0) r -= 1.0;
    return __double2float_rn(__ddiv_rn(r, 1e4));
}

// float32(float("%.5f" % x)) for x >= 0 (DamID text format, :35,:270,:289)
__device__ __forceinline__ float round5_to_f32(double x) {
    const double hi = __dmul_rn(x, 1e5);
    if (!(hi < 4.0e15)) return __double2float_rn(x);
    const double lo = __fma_rn(x, 1e5, -hi);
    double r = rint(hi);
    const double e = __dsub_rn(hi, r);
    if (e == 0.5 && lo > 0.0) r += 1.0;
    else if (e == -0.5 && lo < 0.0) r -= 1.0;
    return __double2float_rn(__ddiv_rn(r, 1e5));
}

// Raw per-pair result; dist / prob (the 4-decimal text round trip) are filled in
// by finish_results_kernel, one thread per pair.
__device__ __forceinline__ void st_global_256(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d,
                                              uint32_t e, uint32_t f, uint32_t g, uint32_t h) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h) : "memory");
}
// One 256-bit store per record (sm_100 STG.256): a single NVLink packet when the
// destination is a peer GPU's gather buffer.
__device__ __forceinline__ void write_result(igmk_pair_result* out, const PairDesc& d,
                                             uint32_t d2_bits, int count, int o, double p,
                                             double extra = 0.0) {
    // Hi-C: no record when p <= 0 (:464,:485); DamID: every copy gets a record (:468)
    const int nrec = (o >= 0 || d.always_rec) ? d.nrec : 0;
    const long long pb = __double_as_longlong(p), eb = __double_as_longlong(extra);
    st_global_256(out, d2_bits, (uint32_t)count, (uint32_t)o, (uint32_t)nrec,
                  (uint32_t)(pb & 0xffffffffll), (uint32_t)((unsigned long long)pb >> 32),
                  (uint32_t)(eb & 0xffffffffll), (uint32_t)((unsigned long long)eb >> 32));
}

// Group-level emit: one thread writes the local result, or - multi-GPU - thread t
// writes the same 32 bytes straight into GPU t's gather buffer over NVLink (peer
// stores issued from inside K1: the "all-gather" overlaps the compute and needs no
// separate collective).  All arguments are uniform over the group.
__device__ __forceinline__ void emit_result(const ActdistParams& P, int tid, long long pair,
                                            const PairDesc& d, uint32_t d2_bits, int count, int o, double p,
                                            double extra = 0.0) {
    if (P.n_peers == 0) {
        if (tid == 0) write_result(P.out + pair, d, d2_bits, count, o, p, extra);
    } else if (tid < P.n_peers) {
        // records that travel to the peers are finished here (dist / prob: the 4-decimal text
        // round trip), so no GPU has to re-finish the other ranks' slices after the gather
        float dist = 0.f, prob = 0.f;
        if (o >= 0) {
            dist = round4_to_f32(sqrt((double)__uint_as_float(d2_bits)));   // float64 sqrt (:473)
            prob = round4_to_f32(p);
        }
        const int nrec = (o >= 0 || d.always_rec) ? d.nrec : 0;
        const long long pb = __double_as_longlong(p);
        st_global_256(reinterpret_cast<igmk_pair_result*>(__ldg(P.peers + tid)) + pair,
                      d2_bits, (uint32_t)count, (uint32_t)o, (uint32_t)nrec,
                      (uint32_t)(pb & 0xffffffffll), (uint32_t)((unsigned long long)pb >> 32),
                      __float_as_uint(dist), __float_as_uint(prob));
    }
}

__device__ __forceinline__ void write_empty(igmk_pair_result* out) {
    st_global_256(out, 0u, 0u, 0xffffffffu, 0u, 0u, 0u, 0u, 0u);   // o = -1
}

__device__ __forceinline__ void emit_empty(const ActdistParams& P, int tid, long long pair) {
    if (P.n_peers == 0) {
        if (tid == 0) write_empty(P.out + pair);
    } else if (tid < P.n_peers) {
        write_empty(reinterpret_cast<igmk_pair_result*>(__ldg(P.peers + tid)) + pair);
    }
}

__global__ void __launch_bounds__(256)
finish_results_kernel(igmk_pair_result* out, long long n) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint4 a = reinterpret_cast<const uint4*>(out + t)[0];
    const int o = (int)a.z;
    float dist = 0.f, prob = 0.f;
    if (o >= 0) {
        const double p = out[t].p;
        dist = round4_to_f32(sqrt((double)__uint_as_float(a.x)));   // float64 sqrt (:473)
        prob = round4_to_f32(p);
    }
    reinterpret_cast<float2*>(out + t)[3] = make_float2(dist, prob);
}

// ----------------------------------------------------------- packed float32x2
__device__ __forceinline__ u64 f2sub(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 f2add(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// rn(a * a) per half.  Written as fma(a, a, -0.0) with the -0.0 pair coming from
// a kernel parameter: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into
// FFMA2 even under --fmad=false, which would break the sequential rounding
// NumPy performs; an FMA whose addend is opaque cannot be contracted further,
// and x*x + (-0.0) rounds exactly like x*x.
__device__ __forceinline__ u64 f2sq(u64 a, u64 negzero2) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(r) : "l"(a), "l"(negzero2));
    return r;
}
// a <= b ? 1.0f : 0.0f  (FSET.BF; NaN compares false)
__device__ __forceinline__ float f_le_one(float a, float b) {
    float r;
    asm("set.le.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float f_lt_one(float a, float b) {
    float r;
    asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ u64 f2pack(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2split(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// x / y / z of 4 consecutive structures as packed pairs (s0,s1) (s2,s3)
struct Row6 { u64 x01, x23, y01, y23, z01, z23; };

// Row loads with an L1 policy: the rows of locus i are shared by every warp of the
// CTA (consecutive pairs of the CSR-ordered list) and should stay in L1; the rows
// of locus j are streamed once and must not evict them.
enum : int { LD_KEEP = 0, LD_STREAM = 1, LD_PLAIN = 2, LD_STREAM_L2KEEP = 3, LD_PLAIN_L2KEEP = 4 };
// How K1 loads the rows of locus j.  Plain loads: the L1::no_allocate hint (LDG.NA) also
// demotes the lines in L2 - with it the rows of the current J-block are evicted before
// their next use (L2 hit rate 39 % against 86 %, DRAM traffic 43 GB against 6.6 GB per
// launch on config 2; ncu, profiles/r02_*).
#ifndef IGMK_JLOAD
#define IGMK_JLOAD LD_PLAIN
#endif
// L2 eviction policy "keep" (evict_last) for the rows of the current J-block
__device__ __forceinline__ u64 l2_policy_evict_last() {
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
template <int HINT>
__device__ __forceinline__ void ldg_v2b64(const float* p, u64& a, u64& b, u64 pol = 0ull) {
#ifdef IGMK_NO_L1_HINTS
    asm("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
#else
    if (HINT == LD_KEEP)
        asm("ld.global.nc.L1::evict_last.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
    else if (HINT == LD_STREAM)
        asm("ld.global.nc.L1::no_allocate.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
    else if (HINT == LD_STREAM_L2KEEP)
        asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.b64 {%0, %1}, [%2], %3;" : "=l"(a), "=l"(b) : "l"(p), "l"(pol));
    else if (HINT == LD_PLAIN_L2KEEP)
        asm("ld.global.nc.L2::cache_hint.v2.b64 {%0, %1}, [%2], %3;" : "=l"(a), "=l"(b) : "l"(p), "l"(pol));
    else
        asm("ld.global.nc.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
#endif
}
__device__ __forceinline__ void lds_v2b64(const float* p, u64& a, u64& b) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];"
                 : "=l"(a), "=l"(b) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
}
__device__ __forceinline__ Row6 load_row6_shared(uint32_t addr) {   // same layout, shared window address
    Row6 r;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(r.x01), "=l"(r.x23) : "r"(addr) : "memory");
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2+512];" : "=l"(r.y01), "=l"(r.y23) : "r"(addr) : "memory");
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2+1024];" : "=l"(r.z01), "=l"(r.z23) : "r"(addr) : "memory");
    return r;
}
template <int HINT>
__device__ __forceinline__ Row6 load_row6(const float* p, u64 pol = 0ull) {
    Row6 r;
    ldg_v2b64<HINT>(p, r.x01, r.x23, pol);
    ldg_v2b64<HINT>(p + kSeg, r.y01, r.y23, pol);
    ldg_v2b64<HINT>(p + 2 * kSeg, r.z01, r.z23, pol);
    return r;
}
// d2 of structures (4c+2h, 4c+2h+1), h = 0 / 1
template <int H>
__device__ __forceinline__ u64 d2pair(const Row6& a, const Row6& b, u64 nz) {
    const u64 dx = f2sub(H ? a.x23 : a.x01, H ? b.x23 : b.x01);
    const u64 dy = f2sub(H ? a.y23 : a.y01, H ? b.y23 : b.y01);
    const u64 dz = f2sub(H ? a.z23 : a.z01, H ? b.z23 : b.z01);
    return f2add(f2add(f2sq(dx, nz), f2sq(dy, nz)), f2sq(dz, nz));
}

// DamID finish: activation distance sqrt(float64(s) / D) (:441,:466), 2 when p <= 0
// (:459); 5-decimal text round trip.  D travels in the record's last 8 bytes.
__global__ void __launch_bounds__(256)
finish_damid_kernel(igmk_pair_result* out, long long n) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint4 a = reinterpret_cast<const uint4*>(out + t)[0];
    const uint4 b = reinterpret_cast<const uint4*>(out + t)[1];
    const int o = (int)a.z;
    float dist = 0.f, prob = 0.f;
    if ((int)a.w > 0) {                              // a record exists
        const double D = __longlong_as_double(((long long)b.w << 32) | (long long)b.z);
        const double p = __longlong_as_double(((long long)b.y << 32) | (long long)b.x);
        dist = (o >= 0) ? round5_to_f32(sqrt(__ddiv_rn((double)__uint_as_float(a.x), D))) : 2.0f;
        prob = round5_to_f32(p);
    }
    reinterpret_cast<float2*>(out + t)[3] = make_float2(dist, prob);
}

// ---- packed bf16x2 primitives (sm_90+ PTX; keys are the high 16 bits of the
// non-negative float32 d2, so bf16 order == unsigned order == float order) ----
__device__ __forceinline__ uint32_t bf2_le(uint32_t a, uint32_t b) {   // 1.0 / 0.0 per half
    uint32_t r;
    asm("set.le.bf16x2.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t bf2_le_mask(uint32_t a, uint32_t b) {  // 0xffff / 0 per half
    uint32_t r;
    asm("set.le.u32.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t bf2_ge_mask(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("set.ge.u32.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t bf2_eq_mask(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("set.eq.u32.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t bf2_add(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t bf2_min(uint32_t a, uint32_t b) {  // NaN operand ignored
    uint32_t r;
    asm("min.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t bf2_max(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
// small non-negative integer counts held as bf16 (exact up to 256) -> int
__device__ __forceinline__ int bf2_count_sum(uint32_t acc) {
    return (int)(__uint_as_float(acc << 16) + __uint_as_float(acc & 0xffff0000u));
}

// ---- shared memory through 32-bit window addresses (no generic-pointer math)
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void lds128(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}
__device__ __forceinline__ uint32_t atoms_inc(uint32_t addr) {
    uint32_t v;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

// ---- mbarrier + bulk asynchronous copies (TMA engine, 1-D): global -> shared, completion
// counted in bytes on an mbarrier.  Used for the locus-i tiles of K1 (igmk_actdist.cuh).
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
// non-blocking probe: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

}  // namespace igmk
