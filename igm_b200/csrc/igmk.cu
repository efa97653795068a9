// igmk.cu - libigmk.so: C ABI (include/igmk.h) + host-side launch logic.
//
// Build (see __graft_entry__.build()):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false
//        -shared -Xcompiler -fPIC -o igm_b200/libigmk.so igm_b200/csrc/igmk.cu
// -fmad=false: float32 d2 and float64 p/o arithmetic must not be contracted
// (bit-exact parity with NumPy / CPython).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/igmk.h"
#include "igmk_actdist.cuh"
#include "igmk_actdist_list.cuh"
#include "igmk_actdist_slab.cuh"
#include "igmk_contact.cuh"
#include "igmk_restraint.cuh"
#include "igmk_sprite.cuh"
#include "igmk_rank.cuh"

#include <cub/device/device_radix_sort.cuh>

using namespace igmk;

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess)                                                      \
            return fail(IGMK_ECUDA, "%s failed: %s (%s:%d)", #expr,                 \
                        cudaGetErrorString(_e), __FILE__, __LINE__);                \
    } while (0)

// Per-launch scratch of the A-step.  Two sets: igmk_actdist_host* alternates them (and two
// compute streams) between consecutive slices of the list, so the short kernels that end a
// slice (redo, finish) run beside the next slice's list kernel instead of after a drained GPU.
struct LaunchScratch {
    void* d_order = nullptr; size_t order_bytes = 0;    // keys / values / cub temp of order_pairs()
    void* d_redo = nullptr; size_t redo_bytes = 0;      // [256 B counter][n_pairs int32]
    void* d_rec = nullptr; size_t rec_bytes = 0;        // PairRec per pair in processing order
    void* d_slab = nullptr; size_t slab_bytes = 0;      // T | cnt | lists of one batch (slab pipeline)
};

constexpr int kPadRows = 64;

struct igmk_ctx {
    LaunchScratch scr[2];
    int cur = 0;                                        // scratch set of the launch being issued
    int device = 0;
    int nbead = 0, nstruct = 0, npad = 0, nchunks = 0, n_hap = 0;
    int sm_count = 0;
    float* d_coords = nullptr;       // [nbead][3][npad]
    float* d_radii = nullptr;        // [nbead]
    int32_t* d_chrom = nullptr;      // [nbead] (igmk_set_bead_chrom)
    HapEntry* d_hap = nullptr;       // [n_hap]
    std::vector<HapEntry> h_hap;
    bool have_coords = false, have_index = false;
    // staging for the *_host entry points
    void* d_stage = nullptr; size_t stage_bytes = 0;
    void* d_pairs = nullptr; size_t pairs_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;                   // second compute stream (odd slices of igmk_actdist_host*)
    cudaEvent_t ev_join = nullptr;
    int host_overlap = 1;                             // IGMK_HOST_OVERLAP: consecutive slices on alternating streams (N <= 1024)
    cudaStream_t s_in = nullptr, s_out = nullptr;     // copy streams of igmk_actdist_host
    cudaStream_t s_up = nullptr;                      // re-layout kernels of the pipelined population upload (highest priority)
    // igmk_actdist_host_population: beads the loci >= l need, as suffix minima per bead region
    // (region 0: first copies, region 1: second copies; merged into one when they interleave)
    int up_nreg = 0, up_lo_reg[2] = {0, 0}, up_hi_reg[2] = {0, 0};
    std::vector<int> up_lo[2];
    int up_piece = 0;                                 // pieces uploaded by the operation in flight (staging buffer parity)
    std::vector<cudaEvent_t> ev_in, ev_k;
    long long host_slice_pairs = 1 << 19;             // IGMK_HOST_SLICE
    int host_slice_ramp = 3;                          // IGMK_HOST_RAMP: slices ramp up from (and down to) slice >> ramp at the ends of a long list
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_up[4] = {nullptr, nullptr, nullptr, nullptr};   // coordinate upload: copy done x2, kernel done x2
    float last_kernel_ms = 0.f;
    int group_threads = 0;       // IGMK_GROUP_THREADS
    int warps_per_cta = 0;       // IGMK_WARPS_PER_CTA
    long long l2_budget = 80ll << 20;   // IGMK_L2_BUDGET: bytes of locus-j rows one J-block may hold
    long long order_min_pairs = 65536;  // IGMK_ORDER_MIN: shorter lists keep the input order
    int tile_block = 512;        // IGMK_TILE_BLOCK (0: no shared-memory locus-i tile)
    int tile_slots = 1;          // IGMK_TILE_SLOTS (2 slots shrink L1 to 15 KB at N = 1000: slower)
    unsigned int* d_blockctr = nullptr;   // ring of 64 block counters (one per launch in flight)
    unsigned int ctr_next = 0;
    int dynamic_blocks = -1;     // IGMK_DYNAMIC_BLOCKS: pair blocks handed out by a device-wide counter; -1 = for lists of
                                 // more than 8 M pairs (CTAs drift apart over a long list and lose the J-block's L2
                                 // residency: config 5 +4 %; config 2 -1 %)
    int list_form = 1;           // IGMK_LIST: 1 = list form first, key-array kernels for what it hands back; 0 = key arrays only
    float list_z = 1.5f;         // IGMK_LIST_Z: margin of the sample threshold (standard deviations)
    int jblock_slab = 1;         // IGMK_JBLOCK_SLAB: slab pipeline sizes J-blocks by one slab's rows (0: whole rows; +2 %)
    int slab_batch = kSlabBatch; // IGMK_SLAB_BATCH: pairs per batch of the slab pipeline (lists of a batch should stay in L2)
    int slab_form = 1;           // IGMK_SLAB: populations > 1024 structures run the list form slab by slab (0: one CTA per pair)
    int list_tile_slots = 1;     // IGMK_LIST_TILE_SLOTS: locus-i tiles per CTA of the list-form warp kernel (2: -1.5 % on config 2,
                                 // no change on config 5; the second slot's 24 KB are worth more as L1)
    float list_budget = 20.f;    // IGMK_LIST_BUDGET: expected list entries per thread beyond which a pair goes to the key arrays
    unsigned int last_redo = 0;  // pairs the list form handed back in the most recent launch (igmk_last_redo_count)
    bool redo_pending = false;
    int block_stop = 32;         // IGMK_BLOCK_STOP: CTA groups leave the key bisection at <= this many candidates (larger: measured slower)
};

static int ensure(void** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes && *p) return IGMK_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    CUDA_TRY(cudaMalloc(p, bytes));
    *cap = bytes;
    return IGMK_OK;
}

extern "C" int igmk_version(void) { return IGMK_VERSION; }
extern "C" const char* igmk_last_error(void) { return g_err.c_str(); }
extern "C" int64_t igmk_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int igmk_destroy(igmk_ctx* c);

extern "C" int igmk_create(int device, int nbead, int nstruct, igmk_ctx** out) {
    if (!out) return fail(IGMK_EINVAL, "igmk_create: out is NULL");
    *out = nullptr;
    if (nbead <= 0 || nstruct <= 0) return fail(IGMK_EINVAL, "igmk_create: nbead and nstruct must be positive");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(IGMK_ECUDA, "igmk_create: no CUDA device (%s); there is no CPU fallback",
                    cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(IGMK_EINVAL, "igmk_create: device %d out of range", device);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    igmk_ctx* c = new igmk_ctx();
    c->device = device;
    c->nbead = nbead;
    c->nstruct = nstruct;
    c->npad = (nstruct + kSeg - 1) / kSeg * kSeg;
    c->nchunks = (nstruct + 3) / 4;
    c->sm_count = prop.multiProcessorCount;
    // one extra all-zero row behind the population: the "origin bead" of the DamID path; kPadRows
    // more so that an in-place all-gather of equal bead shares (igmk_coords_device) fits
    const size_t bytes = (size_t)(nbead + 1 + kPadRows) * 3 * c->npad * sizeof(float);
    e = cudaMalloc(&c->d_coords, bytes);
    if (e != cudaSuccess) { delete c; return fail(IGMK_ECUDA, "igmk_create: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
    // every further resource is checked; a failure releases what exists so far
    e = cudaMemset(c->d_coords, 0, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_radii, (size_t)nbead * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&c->d_blockctr, 64 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking);
    if (e == cudaSuccess) {
        int pr_lo = 0, pr_hi = 0;
        e = cudaDeviceGetStreamPriorityRange(&pr_lo, &pr_hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->s_up, cudaStreamNonBlocking, pr_hi);
    }
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e != cudaSuccess) {
        const int rc = fail(IGMK_ECUDA, "igmk_create: device resources: %s", cudaGetErrorString(e));
        igmk_destroy(c);
        return rc;
    }
    const char* ov = getenv("IGMK_GROUP_THREADS");
    if (ov) c->group_threads = atoi(ov);
    ov = getenv("IGMK_L2_BUDGET");
    if (ov) c->l2_budget = atoll(ov);
    ov = getenv("IGMK_ORDER_MIN");
    if (ov) c->order_min_pairs = atoll(ov);
    ov = getenv("IGMK_TILE_SLOTS");
    if (ov) c->tile_slots = atoi(ov);
    ov = getenv("IGMK_TILE_BLOCK");
    if (ov) c->tile_block = atoi(ov);
    ov = getenv("IGMK_HOST_SLICE");
    if (ov && atoll(ov) > 0) c->host_slice_pairs = atoll(ov);
    ov = getenv("IGMK_HOST_OVERLAP");
    if (ov) c->host_overlap = atoi(ov);
    ov = getenv("IGMK_HOST_RAMP");
    if (ov) c->host_slice_ramp = atoi(ov);
    ov = getenv("IGMK_WARPS_PER_CTA");
    if (ov) c->warps_per_cta = atoi(ov);
    ov = getenv("IGMK_DYNAMIC_BLOCKS");
    if (ov) c->dynamic_blocks = atoi(ov);
    ov = getenv("IGMK_LIST");
    if (ov) c->list_form = atoi(ov);
    ov = getenv("IGMK_LIST_Z");
    if (ov) c->list_z = (float)atof(ov);
    ov = getenv("IGMK_SLAB");
    if (ov) c->slab_form = atoi(ov);
    ov = getenv("IGMK_JBLOCK_SLAB");
    if (ov) c->jblock_slab = atoi(ov);
    ov = getenv("IGMK_SLAB_BATCH");
    if (ov && atoi(ov) >= 1024) c->slab_batch = atoi(ov);
    ov = getenv("IGMK_LIST_TILE_SLOTS");
    if (ov) c->list_tile_slots = atoi(ov);
    ov = getenv("IGMK_LIST_BUDGET");
    if (ov) c->list_budget = (float)atof(ov);
    ov = getenv("IGMK_BLOCK_STOP");
    if (ov) c->block_stop = atoi(ov);
    if (c->block_stop < kRankCap) c->block_stop = kRankCap;
    if (c->block_stop > kBlockListCap) c->block_stop = kBlockListCap;
    *out = c;
    return IGMK_OK;
}

extern "C" int igmk_destroy(igmk_ctx* c) {
    if (!c) return IGMK_OK;
    cudaSetDevice(c->device);
    cudaFree(c->d_coords);
    cudaFree(c->d_radii);
    cudaFree(c->d_blockctr);
    cudaFree(c->d_chrom);
    cudaFree(c->d_hap);
    cudaFree(c->d_stage);
    cudaFree(c->d_pairs);
    for (LaunchScratch& sc : c->scr) {
        cudaFree(sc.d_order);
        cudaFree(sc.d_redo);
        cudaFree(sc.d_rec);
        cudaFree(sc.d_slab);
    }
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (cudaEvent_t e : c->ev_up) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_k) cudaEventDestroy(e);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    if (c->s_up) cudaStreamDestroy(c->s_up);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    delete c;
    return IGMK_OK;
}

// (nb, nstruct, 3) bead-major AoS  ->  [bead][segment of 128 structures][xyz][128]
// 128 threads at <= 32 registers: a block fits into the 4096 registers the persistent list
// kernel (640 threads x 96) leaves free on an SM, so the pipelined upload
// (igmk_actdist_host_population) re-lays out pieces while the pair kernel runs.
__global__ void __launch_bounds__(128, 16)
stage_coords_kernel(const float* __restrict__ src, float* __restrict__ dst, int nstruct, int npad) {
    const size_t bead = blockIdx.x;
    const float* s = src + bead * (size_t)nstruct * 3;
    float* d = dst + bead * (size_t)npad * 3;
    for (int t = threadIdx.x; t < 3 * nstruct; t += blockDim.x) {
        const int st = t / 3, c = t - 3 * st;
        d[coord_off(st) + (size_t)c * kSeg] = s[t];
    }
}

// Beads per piece of a host upload: two staging buffers of <= 64 MiB each.
static int upload_step(const igmk_ctx* c) {
    const size_t per_bead = (size_t)c->nstruct * 3 * sizeof(float);
    long long step = (long long)((64ull << 20) / per_bead);
    if (step < 1) step = 1;
    if (step > c->nbead) step = c->nbead;
    return (int)step;
}

static int upload_begin(igmk_ctx* c) {
    const size_t per_bead = (size_t)c->nstruct * 3 * sizeof(float);
    int rc = ensure(&c->d_stage, &c->stage_bytes, 2 * (size_t)upload_step(c) * per_bead);
    if (rc) return rc;
    if (!c->ev_up[0]) {
        for (int k = 0; k < 4; ++k) CUDA_TRY(cudaEventCreateWithFlags(&c->ev_up[k], cudaEventDisableTiming));
    }
    c->up_piece = 0;
    return IGMK_OK;
}

// Beads [bead0, bead0 + nb) -> HBM from the host array `xyz` (.hss layout, its first row is
// bead0), asynchronously: the copy of piece k + 1 (copy-in stream) overlaps the re-layout kernel of
// piece k (stream sk).  Pinned host memory (igmk_host_alloc, torch pin_memory) makes the
// copies asynchronous.  After the call ev_up[2 + ((up_piece - 1) & 1)] marks "everything
// uploaded so far is in place" (the re-layout kernels are ordered on sk).
static int upload_pieces(igmk_ctx* c, const float* xyz, int bead0, int nb, cudaStream_t sk) {
    const size_t per_bead = (size_t)c->nstruct * 3 * sizeof(float);
    const int step = upload_step(c);
    for (int b = 0; b < nb; b += step) {
        const int n = (nb - b < step) ? nb - b : step, k = c->up_piece++;
        float* buf = (float*)((char*)c->d_stage + (size_t)(k & 1) * step * per_bead);
        if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(c->s_in, c->ev_up[2 + (k & 1)], 0));     // the kernel that read this buffer
        CUDA_TRY(cudaMemcpyAsync(buf, xyz + (size_t)b * c->nstruct * 3, (size_t)n * per_bead, cudaMemcpyHostToDevice, c->s_in));
        CUDA_TRY(cudaEventRecord(c->ev_up[k & 1], c->s_in));
        CUDA_TRY(cudaStreamWaitEvent(sk, c->ev_up[k & 1], 0));
        stage_coords_kernel<<<n, 128, 0, sk>>>(buf, c->d_coords + (size_t)(bead0 + b) * 3 * c->npad, c->nstruct, c->npad);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(c->ev_up[2 + (k & 1)], sk));
    }
    return IGMK_OK;
}

extern "C" int igmk_upload_coords_range(igmk_ctx* c, const float* xyz, int bead0, int nb, int on_device) {
    if (!c || !xyz) return fail(IGMK_EINVAL, "igmk_upload_coords: NULL argument");
    if (bead0 < 0 || nb < 0 || bead0 + nb > c->nbead) return fail(IGMK_EINVAL, "igmk_upload_coords: bead range out of bounds");
    CUDA_TRY(cudaSetDevice(c->device));
    if (on_device) {
        // already in HBM: one re-layout launch
        if (nb > 0) {
            stage_coords_kernel<<<nb, 128, 0, c->stream>>>(xyz, c->d_coords + (size_t)bead0 * 3 * c->npad, c->nstruct, c->npad);
            g_launches++;
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaStreamSynchronize(c->stream));
        }
        c->have_coords = true;
        return IGMK_OK;
    }
    if (nb == 0) { c->have_coords = true; return IGMK_OK; }
    int rc = upload_begin(c);
    if (rc) return rc;
    rc = upload_pieces(c, xyz, bead0, nb, c->stream);
    if (rc) { cudaDeviceSynchronize(); return rc; }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->have_coords = true;
    return IGMK_OK;
}

// The staged population as device memory: rows of 3 * npad floats, nbead rows of beads, one
// all-zero row, kPadRows spare rows.  For NVLink replication: every rank uploads its share of
// the beads (igmk_upload_coords_range) and one all-gather over whole rows completes the copies.
extern "C" int igmk_coords_device(igmk_ctx* c, float** d_coords, int64_t* n_rows, int64_t* row_floats) {
    if (!c || !d_coords || !n_rows || !row_floats) return fail(IGMK_EINVAL, "igmk_coords_device: NULL argument");
    *d_coords = c->d_coords;
    *n_rows = (int64_t)c->nbead + 1 + kPadRows;
    *row_floats = 3ll * c->npad;
    return IGMK_OK;
}

// Beads [bead0, bead0 + nb) of `src` (another device of this process) -> the same rows of
// `dst`, over NVLink when the devices are peers.  Synchronises dst's stream.
extern "C" int igmk_copy_coords_peer(igmk_ctx* dst, igmk_ctx* src, int bead0, int nb) {
    if (!dst || !src) return fail(IGMK_EINVAL, "igmk_copy_coords_peer: NULL context");
    if (dst->nbead != src->nbead || dst->nstruct != src->nstruct) return fail(IGMK_EINVAL, "igmk_copy_coords_peer: populations differ in shape");
    if (bead0 < 0 || nb < 0 || bead0 + nb > dst->nbead) return fail(IGMK_EINVAL, "igmk_copy_coords_peer: bead range out of bounds");
    if (!src->have_coords) return fail(IGMK_ESTATE, "igmk_copy_coords_peer: the source holds no coordinates");
    CUDA_TRY(cudaSetDevice(dst->device));
    if (nb > 0) {
        if (dst->device != src->device) {
            int can = 0;
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, dst->device, src->device));
            if (can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess) return fail(IGMK_ECUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
            }
        }
        const size_t row = (size_t)3 * dst->npad, off = (size_t)bead0 * row;
        CUDA_TRY(cudaMemcpyPeerAsync(dst->d_coords + off, dst->device, src->d_coords + off, src->device,
                                     (size_t)nb * row * sizeof(float), dst->stream));
        CUDA_TRY(cudaStreamSynchronize(dst->stream));
    }
    dst->have_coords = true;
    return IGMK_OK;
}

extern "C" int igmk_upload_coords(igmk_ctx* c, const float* xyz, int on_device) {
    if (!c) return fail(IGMK_EINVAL, "igmk_upload_coords: NULL context");
    return igmk_upload_coords_range(c, xyz, 0, c->nbead, on_device);
}

extern "C" int igmk_set_index(igmk_ctx* c, int n_hap, const int32_t* copy_ptr,
                              const int32_t* copy_beads, const int32_t* chrom_hap,
                              const float* radii) {
    if (!c || !copy_ptr || !copy_beads || !chrom_hap || !radii) return fail(IGMK_EINVAL, "igmk_set_index: NULL argument");
    if (n_hap <= 0) return fail(IGMK_EINVAL, "igmk_set_index: n_hap must be positive");
    CUDA_TRY(cudaSetDevice(c->device));
    std::vector<HapEntry> h(n_hap);
    for (int i = 0; i < n_hap; ++i) {
        const int nc = copy_ptr[i + 1] - copy_ptr[i];
        if (nc < 1) return fail(IGMK_EINVAL, "igmk_set_index: bin %d has no copy", i);
        if (nc > 2) return fail(IGMK_ELIMIT, "igmk_set_index: bin %d has %d copies; at most 2 are supported", i, nc);
        const int b0 = copy_beads[copy_ptr[i]];
        const int b1 = (nc == 2) ? copy_beads[copy_ptr[i] + 1] : -1;
        if (b0 < 0 || b0 >= c->nbead || b1 >= c->nbead) return fail(IGMK_EINVAL, "igmk_set_index: bead id out of range in bin %d", i);
        h[i].b0 = b0; h[i].b1 = b1; h[i].chrom = chrom_hap[i]; h[i].radius = radii[b0];
    }
    if (c->d_hap) { cudaFree(c->d_hap); c->d_hap = nullptr; }
    CUDA_TRY(cudaMalloc(&c->d_hap, (size_t)n_hap * sizeof(HapEntry)));
    CUDA_TRY(cudaMemcpy(c->d_hap, h.data(), (size_t)n_hap * sizeof(HapEntry), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->d_radii, radii, (size_t)c->nbead * sizeof(float), cudaMemcpyHostToDevice));
    c->h_hap.swap(h);
    c->n_hap = n_hap;
    c->have_index = true;
    // Pipelined population upload (igmk_actdist_host_population): the beads every locus >= l
    // needs form, per copy, the range [suffix minimum, region end) - whatever the index looks
    // like; a monotone index (the usual one) makes consecutive ranges disjoint.
    {
        const int kNone = 0x7fffffff;
        int mn[2] = {kNone, kNone}, mx[2] = {-1, -1};
        for (int r = 0; r < 2; ++r) c->up_lo[r].assign((size_t)n_hap + 1, kNone);
        for (int l = n_hap - 1; l >= 0; --l) {
            const int b[2] = {c->h_hap[l].b0, c->h_hap[l].b1};
            for (int r = 0; r < 2; ++r) {
                int v = c->up_lo[r][l + 1];
                if (b[r] >= 0) {
                    v = (b[r] < v) ? b[r] : v;
                    mn[r] = (b[r] < mn[r]) ? b[r] : mn[r];
                    mx[r] = (b[r] > mx[r]) ? b[r] : mx[r];
                }
                c->up_lo[r][l] = v;
            }
        }
        const bool two = mx[1] >= 0 && (mx[0] < mn[1] || mx[1] < mn[0]);     // disjoint bead regions
        if (two) {
            c->up_nreg = 2;
            for (int r = 0; r < 2; ++r) { c->up_lo_reg[r] = mn[r]; c->up_hi_reg[r] = mx[r] + 1; }
        } else {
            c->up_nreg = 1;
            for (int l = 0; l <= n_hap; ++l) c->up_lo[0][l] = (c->up_lo[1][l] < c->up_lo[0][l]) ? c->up_lo[1][l] : c->up_lo[0][l];
            c->up_lo_reg[0] = (mn[1] < mn[0]) ? mn[1] : mn[0];
            c->up_hi_reg[0] = ((mx[1] > mx[0]) ? mx[1] : mx[0]) + 1;
        }
    }
    return IGMK_OK;
}

// ------------------------------------------------------------- K1 launches
static int launch_finish(const ActdistParams& P, cudaStream_t st) {
    if (P.n_peers > 0) return IGMK_OK;      // raw results went to the peers; each GPU finishes its gather buffer
    const long long blocks = (P.n_pairs + 255) / 256;
    if (P.pexp32) finish_damid_kernel<<<(unsigned)blocks, 256, 0, st>>>(P.out, P.n_pairs);
    else          finish_results_kernel<<<(unsigned)blocks, 256, 0, st>>>(P.out, P.n_pairs);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return IGMK_OK;
}

template <bool DAMID>
static int launch_warp(const igmk_ctx* c, ActdistParams P, cudaStream_t st, bool redo = false) {
    const int V = (c->nchunks + 31) / 32;
    // one CTA per SM: two locus-i tiles (2 rows of 12 * npad bytes each) plus as many
    // warps as the key arrays (V KiB per warp) leave room for
    const size_t budget = 208 * 1024;
    const int slots = (c->tile_slots == 1) ? 1 : 2;
    // (a redo launch works through an unordered list of unknown length: no locus-i tile)
    size_t tile_bytes = (!DAMID && !redo && c->tile_block > 0 && c->n_hap < (1 << 20)) ? (size_t)slots * 24 * c->npad : 0;
    if (tile_bytes + (size_t)8 * 2 * V * 32 * 16 > budget) tile_bytes = 0;     // keep >= 8 warps
    int warps = (int)((budget - tile_bytes) / ((size_t)2 * V * 32 * 16));
    if (warps > kWarpsPerBlock) warps = kWarpsPerBlock;
    if (c->warps_per_cta > 0 && warps > c->warps_per_cta) warps = c->warps_per_cta;
    if (warps < 1) return fail(IGMK_ELIMIT, "actdist_warp_kernel: nstruct = %d is too large for one warp per pair", c->nstruct);
    P.tile_block = tile_bytes ? c->tile_block : 0;
    P.tile_slots = slots;
    if (P.tile_block > 0) {
        // short lists: blocks are dealt to the CTAs round-robin, so keep about eight blocks
        // per CTA (a list of 229 k pairs in blocks of 512 gives most CTAs 3 blocks and some 4)
        const long long per_cta = (P.n_pairs + c->sm_count - 1) / c->sm_count;
        if (per_cta < 8LL * P.tile_block) {
            long long b = ((per_cta + 7) / 8 + 31) / 32 * 32;
            if (b < 32) b = 32;
            if (b < P.tile_block) P.tile_block = (int)b;
        }
    }
    P.block_counter = nullptr;
    const bool dyn = (c->dynamic_blocks < 0) ? (P.n_pairs > (8LL << 20)) : (c->dynamic_blocks != 0);
    if (P.tile_block > 0 && dyn && c->d_blockctr) {
        igmk_ctx* cm = const_cast<igmk_ctx*>(c);
        P.block_counter = c->d_blockctr + (cm->ctr_next++ & 63u);
        CUDA_TRY(cudaMemsetAsync(P.block_counter, 0, sizeof(unsigned int), st));
    }
    int per_sm = 0;
    const size_t smem = (size_t)warps * 2 * V * 32 * 16 + tile_bytes;
    CUDA_TRY(cudaFuncSetAttribute(actdist_warp_kernel<DAMID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, actdist_warp_kernel<DAMID>, 32 * warps, smem));
    if (per_sm < 1) return fail(IGMK_ECUDA, "actdist_warp_kernel cannot run with V = %d", V);
    long long want = P.tile_block ? (P.n_pairs + P.tile_block - 1) / P.tile_block : (P.n_pairs + warps - 1) / warps;
    long long cap = (long long)c->sm_count * per_sm;
    const int grid = (int)((want < cap) ? want : cap);
    actdist_warp_kernel<DAMID><<<grid, 32 * warps, smem, st>>>(P, V);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return launch_finish(P, st);
}

template <int MAXT, bool DAMID>
static int launch_block(const igmk_ctx* c, const ActdistParams& Pin, int threads, cudaStream_t st) {
    ActdistParams P = Pin;
    P.block_counter = nullptr;
    if (c->dynamic_blocks > 0 && c->d_blockctr && P.n_pairs < 0xffffffffLL) {      // IGMK_DYNAMIC_BLOCKS=1
        igmk_ctx* cm = const_cast<igmk_ctx*>(c);
        P.block_counter = c->d_blockctr + (cm->ctr_next++ & 63u);
        CUDA_TRY(cudaMemsetAsync(P.block_counter, 0, sizeof(unsigned int), st));
    }
    const int V = (c->nchunks + threads - 1) / threads;
    int per_sm = 0;
    const size_t smem = (size_t)2 * V * threads * 16;
    if (2 * V > kMaxQuads || smem > 200 * 1024)
        return fail(IGMK_ELIMIT, "igmk_actdist: nstruct = %d exceeds the supported 25600", c->nstruct);
    CUDA_TRY(cudaFuncSetAttribute(actdist_block_kernel<MAXT, DAMID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, actdist_block_kernel<MAXT, DAMID>, threads, smem));
    if (per_sm < 1) return fail(IGMK_ECUDA, "actdist_block_kernel cannot run with %d threads, V = %d", threads, V);
    long long cap = (long long)c->sm_count * per_sm;
    const int grid = (int)((P.n_pairs < cap) ? P.n_pairs : cap);
    actdist_block_kernel<MAXT, DAMID><<<grid, threads, smem, st>>>(P, V);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return launch_finish(P, st);
}


// ------------------------------------------------------------- K1, list form
// First launch: igmk_actdist_list.cuh over the whole list; second launch: the key-array
// kernels over the pairs the first one handed back (count and order live on the device).
static int list_prepare(igmk_ctx* c, ActdistParams& P, cudaStream_t st, int group_threads) {
    int rc = ensure(&c->scr[c->cur].d_redo, &c->scr[c->cur].redo_bytes, 256 + (size_t)P.n_pairs * sizeof(int32_t));
    if (rc) return rc;
    P.redo_count = (unsigned int*)c->scr[c->cur].d_redo;
    P.redo = (int32_t*)((char*)c->scr[c->cur].d_redo + 256);
    P.n_pairs_dev = nullptr;
    P.list_z = c->list_z;
    P.list_budget = c->list_budget * (float)group_threads;
    CUDA_TRY(cudaMemsetAsync(P.redo_count, 0, sizeof(unsigned int), st));
    // pair descriptors in processing order
    rc = ensure(&c->scr[c->cur].d_rec, &c->scr[c->cur].rec_bytes, (size_t)P.n_pairs * sizeof(PairRec));
    if (rc) return rc;
    P.rec = (const PairRec*)c->scr[c->cur].d_rec;
    build_pairrec_kernel<<<(unsigned)((P.n_pairs + 255) / 256), 256, 0, st>>>(P, (PairRec*)c->scr[c->cur].d_rec);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return IGMK_OK;
}

static int launch_list_warp(igmk_ctx* c, ActdistParams P, cudaStream_t st) {
    const int V = (c->nchunks + 31) / 32;
    int rc = list_prepare(c, P, st, 32);
    if (rc) return rc;
    const size_t budget = 220 * 1024;
    int warps = kListWarps;
    if (c->warps_per_cta > 0 && warps > c->warps_per_cta) warps = c->warps_per_cta;
    const size_t list_bytes = (size_t)warps * 32 * kListBytes;
    const size_t one_tile = (size_t)24 * c->npad;
    int slots = (c->tile_block > 0 && c->nbead < (1 << 20)) ? c->list_tile_slots : 0;     // tile key: bead id
    if (slots > 2) slots = 2;
    while (slots > 0 && list_bytes + slots * one_tile > budget) --slots;
    P.tile_slots = slots;
    // CTA-contiguous blocks of 2^shift pairs; short lists: blocks are dealt to the CTAs
    // round-robin, so keep about eight per CTA
    int shift = 9;
    if (c->tile_block > 0) { shift = 0; while ((2 << shift) <= c->tile_block) ++shift; }
    const long long per_cta = (P.n_pairs + c->sm_count - 1) / c->sm_count;
    while (shift > 5 && per_cta < (8LL << shift)) --shift;
    P.tile_block = shift;
    const size_t smem = list_bytes + (size_t)slots * one_tile;
    CUDA_TRY(cudaFuncSetAttribute(actdist_list_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long want = (P.n_pairs + (1LL << shift) - 1) >> shift;
    const int grid = (int)((want < c->sm_count) ? want : c->sm_count);
    actdist_list_warp_kernel<<<grid, 32 * warps, smem, st>>>(P, V);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return IGMK_OK;
}

static int launch_list_block(igmk_ctx* c, ActdistParams P, int threads, cudaStream_t st) {
    if (threads > kListBlockThreads) threads = kListBlockThreads;
    const int V = (c->nchunks + threads - 1) / threads;
    int rc = list_prepare(c, P, st, threads);
    if (rc) return rc;
    const size_t smem = (size_t)threads * kListBytes;
    int per_sm = 0;
    CUDA_TRY(cudaFuncSetAttribute(actdist_list_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, actdist_list_block_kernel, threads, smem));
    if (per_sm < 1) return fail(IGMK_ECUDA, "actdist_list_block_kernel cannot run (shared memory %zu)", smem);
    long long cap = (long long)c->sm_count * per_sm;
    const int grid = (int)((P.n_pairs < cap) ? P.n_pairs : cap);
    actdist_list_block_kernel<<<grid, threads, smem, st>>>(P, V);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return IGMK_OK;
}

// Populations of more than 1024 structures: sample / fill / select kernels of
// igmk_actdist_slab.cuh over batches of kSlabBatch pairs (processing order).
static int launch_slab(igmk_ctx* c, ActdistParams P, cudaStream_t st) {
    int rc = list_prepare(c, P, st, 32);
    if (rc) return rc;
    P.list_budget = (float)kSlabCap / 1.35f;          // expected list length a pair may have
    const int nseg = c->npad / kSeg;
    SlabParams S;
    S.nslab = (nseg + kSlabSegs - 1) / kSlabSegs;
    const long long batch = c->slab_batch;
    const size_t off_cnt = (size_t)batch * 4, off_lists = 2 * (size_t)batch * 4;
    rc = ensure(&c->scr[c->cur].d_slab, &c->scr[c->cur].slab_bytes, off_lists + (size_t)batch * kSlabCap * 4);
    if (rc) return rc;
    S.T = (uint32_t*)c->scr[c->cur].d_slab;
    S.cnt = (unsigned int*)((char*)c->scr[c->cur].d_slab + off_cnt);
    S.lists = (uint32_t*)((char*)c->scr[c->cur].d_slab + off_lists);
    int warps = kSlabWarps;
    if (c->warps_per_cta > 0 && warps > c->warps_per_cta) warps = c->warps_per_cta;
    const size_t list_bytes = (size_t)warps * 32 * kListBytes;
    const size_t one_tile = (size_t)2 * kSlabSegs * kSegFloats * 4;
    int slots = (c->tile_block > 0 && (long long)c->nbead * S.nslab < (1 << 20)) ? c->list_tile_slots : 0;   // key: (bead, slab)
    if (slots > 2) slots = 2;
    P.tile_slots = slots;
    const size_t smem = list_bytes + (size_t)slots * one_tile;
    CUDA_TRY(cudaFuncSetAttribute(slab_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (long long b0 = 0; b0 < P.n_pairs; b0 += batch) {
        S.slot0 = b0;
        S.nslots = (int)((P.n_pairs - b0 < batch) ? P.n_pairs - b0 : batch);
        int shift = 9;
        if (c->tile_block > 0) { shift = 0; while ((2 << shift) <= c->tile_block) ++shift; }
        // about eight tasks per CTA at least
        while (shift > 5 && (long long)S.nslab * ((S.nslots + (1 << shift) - 1) >> shift) < 8LL * c->sm_count) --shift;
        S.bshift = shift;
        S.nblk = (S.nslots + (1 << shift) - 1) >> shift;
        const int wgrid = (S.nslots + 7) / 8;
        const int grid_w = (wgrid < 8 * c->sm_count) ? wgrid : 8 * c->sm_count;
        slab_sample_kernel<<<grid_w, 256, 0, st>>>(P, S);
        const long long ntask = (long long)S.nslab * S.nblk;
        const int grid_f = (int)((ntask < c->sm_count) ? ntask : c->sm_count);
        slab_fill_kernel<<<grid_f, 32 * warps, smem, st>>>(P, S);
        slab_select_kernel<<<grid_w, 256, 0, st>>>(P, S);
        g_launches += 3;
        CUDA_TRY(cudaGetLastError());
    }
    return IGMK_OK;
}

static int launch_simple(const igmk_ctx* c, const ActdistParams& P, cudaStream_t st) {
    const size_t smem = (size_t)4 * P.nstruct * sizeof(uint32_t);
    if (smem > 227 * 1024) return fail(IGMK_ELIMIT, "IGMK_ALGO_SIMPLE supports nstruct <= %d", 227 * 1024 / 16);
    CUDA_TRY(cudaFuncSetAttribute(actdist_simple_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, actdist_simple_kernel, 256, smem));
    if (per_sm < 1) per_sm = 1;
    long long cap = (long long)c->sm_count * per_sm;
    const int grid = (int)((P.n_pairs < cap) ? P.n_pairs : cap);
    actdist_simple_kernel<<<grid, 256, smem, st>>>(P);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return launch_finish(P, st);
}

// ------------------------------------------------------- processing order
// The rows of locus j are streamed once per pair; when the whole population does not
// fit in L2 they come from HBM again for every row i of the list.  Processing the
// list one J-block at a time (loci whose rows fit the L2 budget together), in the
// caller's order inside a block, turns those into L2 hits: DRAM traffic drops from
// ~11 KB to < 1 KB per pair at N = 1000.  Stable 1-pass radix sort on the block id.
__global__ void jblock_keys_kernel(const int32_t* __restrict__ pj, long long n, int loci_per_block,
                                   int n_hap, uint32_t* __restrict__ keys, int32_t* __restrict__ vals) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int j = pj[t];
    j = (j < 0) ? 0 : (j >= n_hap ? n_hap - 1 : j);
    keys[t] = (uint32_t)(j / loci_per_block);
    vals[t] = (int32_t)t;
}

static int order_pairs(igmk_ctx* c, int64_t n_pairs, const int32_t* d_j, cudaStream_t st,
                       const int32_t** perm_out) {
    *perm_out = nullptr;
    if (c->l2_budget <= 0 || n_pairs < c->order_min_pairs || n_pairs > 0x7fffffffLL) return IGMK_OK;
    // rows one locus contributes to the working set: whole rows, or - slab pipeline
    // (IGMK_JBLOCK_SLAB=1) - the rows of one slab only: a slab's J-block then fills the L2 budget
    long long per_locus = 2ll * 12 * c->npad;
    if (c->jblock_slab && c->list_form && c->slab_form && c->group_threads == 0 && c->nchunks > 256)
        per_locus = 2ll * 12 * 1024;
    long long lpb = c->l2_budget / per_locus;
    if (lpb < 1) lpb = 1;
    const int nblk = (int)((c->n_hap + lpb - 1) / lpb);
    if (nblk <= 1) return IGMK_OK;
    int bits = 1;
    while ((1 << bits) < nblk) ++bits;
    const size_t n = (size_t)n_pairs;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    size_t temp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, (int)n, 0, bits, st);
    const size_t total = 4 * up(n * 4) + up(temp);
    int rc = ensure(&c->scr[c->cur].d_order, &c->scr[c->cur].order_bytes, total);
    if (rc) return rc;
    char* base = (char*)c->scr[c->cur].d_order;
    uint32_t* k_in = (uint32_t*)base;
    uint32_t* k_out = (uint32_t*)(base + up(n * 4));
    int32_t* v_in = (int32_t*)(base + 2 * up(n * 4));
    int32_t* v_out = (int32_t*)(base + 3 * up(n * 4));
    void* d_temp = base + 4 * up(n * 4);
    jblock_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_j, (long long)n, (int)lpb, c->n_hap, k_in, v_in);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(d_temp, temp, k_in, k_out, v_in, v_out, (int)n, 0, bits, st));
    *perm_out = v_out;
    return IGMK_OK;
}

static int group_threads_for(const igmk_ctx* c) {
    int T = c->group_threads;
    if (T == 0) {
        if (c->nchunks <= 32 * 8) T = 32;
        else {
            T = ((c->nchunks + 7) / 8 + 31) / 32 * 32;          // about 8 chunks per thread
            if (T < 64) T = 64;
            if (T > 320) T = ((c->nchunks + 319) / 320 <= 12) ? 320 : 512;
        }
    }
    return T;
}

template <bool DAMID>
static int dispatch_groups(const igmk_ctx* c, const ActdistParams& P, cudaStream_t st, bool redo = false) {
    int T = group_threads_for(c);
    if (T == 32 && c->nchunks <= 32 * 32) return launch_warp<DAMID>(c, P, st, redo);
    if (T < 64) T = 64;
    if (T > 512) T = 512;
    T = (T + 31) / 32 * 32;
    if (T <= 320) return launch_block<320, DAMID>(c, P, T, st);
    return launch_block<512, DAMID>(c, P, T, st);
}

static int actdist_launch(igmk_ctx* c, int64_t n_pairs,
                          const int32_t* d_i, const int32_t* d_j,
                          const double* d_pwish, const double* d_plast,
                          float contact_range, int it_corr, int mode, int algo,
                          igmk_pair_result* d_out, const uint64_t* d_peers, int n_peers, void* stream) {
    if (!c) return fail(IGMK_EINVAL, "igmk_actdist: NULL context");
    if (!c->have_coords || !c->have_index) return fail(IGMK_ESTATE, "igmk_actdist: upload coordinates and index first");
    if (n_pairs < 0) return fail(IGMK_EINVAL, "igmk_actdist: negative n_pairs");
    if (mode != IGMK_MODE_LB && mode != IGMK_MODE_GP) return fail(IGMK_EINVAL, "igmk_actdist: bad mode %d", mode);
    if (n_pairs == 0) return IGMK_OK;
    if (!d_i || !d_j || !d_pwish || !d_plast || (!d_out && n_peers == 0)) return fail(IGMK_EINVAL, "igmk_actdist: NULL buffer");
    if (n_peers < 0 || n_peers > 32 || (n_peers > 0 && (!d_peers || algo != IGMK_ALGO_FAST)))
        return fail(IGMK_EINVAL, "igmk_actdist: bad peer table");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    ActdistParams P;
    P.coords = c->d_coords; P.hap = c->d_hap;
    P.pi = d_i; P.pj = d_j; P.pwish = d_pwish; P.plast = d_plast; P.out = d_out;
    P.n_pairs = n_pairs; P.nstruct = c->nstruct; P.npad = c->npad; P.nchunks = c->nchunks;
    P.n_hap = c->n_hap; P.contact_range = contact_range; P.it_corr = it_corr; P.mode = mode;
    P.negzero2 = 0x8000000080000000ull;
    P.block_stop = c->block_stop;
    P.perm = nullptr;
    P.peers = (const u64*)d_peers; P.n_peers = n_peers;
    P.pexp32 = nullptr; P.plast32 = nullptr; P.damid_R = 0.0; P.zero_bead = c->nbead;
    P.tile_block = 0;
    P.tile_slots = 0;
    P.redo = nullptr; P.redo_count = nullptr; P.n_pairs_dev = nullptr; P.list_z = 0.f; P.list_budget = 0.f;
    P.rec = nullptr;
    c->redo_pending = false;

    if (algo == IGMK_ALGO_SIMPLE) return launch_simple(c, P, st);
    if (algo != IGMK_ALGO_FAST) return fail(IGMK_EINVAL, "igmk_actdist: bad algo %d", algo);
    {
        int rc = order_pairs(c, n_pairs, d_j, st, &P.perm);
        if (rc) return rc;
    }

    // Thread-group shape (igmk_actdist.cuh): V float4 chunks (4 structures x <= 4
    // combinations each) per thread.
    //   nstruct <= 1024: one warp per pair, V = ceil(nchunks / 32) <= 8
    //   larger:          one CTA per pair, T <= 320 threads with V = ceil(nchunks / T) about 8
    //                    (T = 512 beyond nstruct = 15360)
    // IGMK_GROUP_THREADS (tuning knob): 0 = default, 32 = force one warp per pair
    // (nstruct <= 4096), otherwise the CTA size to use.
    // List form first (igmk_actdist_list.cuh); the key-array kernels then redo what it
    // handed back, in the order it was handed back.  Forced group sizes and IGMK_LIST=0
    // keep the key-array kernels alone (cross-check).
    if (c->list_form && c->group_threads == 0 && n_pairs <= 0x7fffffffLL) {
        const int T = group_threads_for(c);
        int rc = (T == 32) ? launch_list_warp(c, P, st)
                           : (c->slab_form ? launch_slab(c, P, st) : launch_list_block(c, P, (T + 31) / 32 * 32, st));
        if (rc) return rc;
        ActdistParams R = P;
        R.redo_count = (unsigned int*)c->scr[c->cur].d_redo;
        R.redo = (int32_t*)((char*)c->scr[c->cur].d_redo + 256);
        R.perm = R.redo;
        R.n_pairs_dev = R.redo_count;
        c->redo_pending = true;
        return dispatch_groups<false>(c, R, st, true);
    }
    return dispatch_groups<false>(c, P, st);
}

// Pairs the list form handed to the key-array kernels in the most recent launch on this
// context (diagnostic; synchronises the device).
extern "C" int64_t igmk_last_redo_count(igmk_ctx* c) {
    if (!c) return -1;
    if (!c->redo_pending || !c->scr[c->cur].d_redo) return 0;
    if (cudaSetDevice(c->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return -1;
    unsigned int v = 0;
    if (cudaMemcpy(&v, c->scr[c->cur].d_redo, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    c->last_redo = v;
    return (int64_t)v;
}

extern "C" int igmk_actdist_device(igmk_ctx* c, int64_t n_pairs,
                                   const int32_t* d_i, const int32_t* d_j,
                                   const double* d_pwish, const double* d_plast,
                                   float contact_range, int it_corr, int mode, int algo,
                                   igmk_pair_result* d_out, void* stream) {
    return actdist_launch(c, n_pairs, d_i, d_j, d_pwish, d_plast, contact_range, it_corr, mode, algo,
                          d_out, nullptr, 0, stream);
}

extern "C" int igmk_actdist_device_peers(igmk_ctx* c, int64_t n_pairs,
                                         const int32_t* d_i, const int32_t* d_j,
                                         const double* d_pwish, const double* d_plast,
                                         float contact_range, int it_corr, int mode,
                                         const uint64_t* d_peer_slices, int n_peers, void* stream) {
    return actdist_launch(c, n_pairs, d_i, d_j, d_pwish, d_plast, contact_range, it_corr, mode,
                          IGMK_ALGO_FAST, nullptr, d_peer_slices, n_peers, stream);
}

// sel_flat_idx of every pair: see sel_index_kernel.  d_results: the records
// igmk_actdist_* produced for the same pair list.
extern "C" int igmk_actdist_sel_index_device(igmk_ctx* c, int64_t n_pairs, const int32_t* d_i, const int32_t* d_j,
                                             const igmk_pair_result* d_results, int mode, int32_t* d_sel_idx,
                                             void* stream) {
    if (!c) return fail(IGMK_EINVAL, "igmk_actdist_sel_index: NULL context");
    if (!c->have_coords || !c->have_index) return fail(IGMK_ESTATE, "igmk_actdist_sel_index: upload coordinates and index first");
    if (mode != IGMK_MODE_LB && mode != IGMK_MODE_GP) return fail(IGMK_EINVAL, "igmk_actdist_sel_index: bad mode %d", mode);
    if (n_pairs < 0) return fail(IGMK_EINVAL, "igmk_actdist_sel_index: negative n_pairs");
    if (n_pairs == 0) return IGMK_OK;
    if (!d_i || !d_j || !d_results || !d_sel_idx) return fail(IGMK_EINVAL, "igmk_actdist_sel_index: NULL buffer");
    CUDA_TRY(cudaSetDevice(c->device));
    ActdistParams P;
    memset(&P, 0, sizeof P);
    P.coords = c->d_coords; P.hap = c->d_hap; P.pi = d_i; P.pj = d_j;
    P.n_pairs = n_pairs; P.nstruct = c->nstruct; P.npad = c->npad; P.nchunks = c->nchunks;
    P.n_hap = c->n_hap; P.mode = mode; P.contact_range = 2.0f;
    P.negzero2 = 0x8000000080000000ull;
    const long long want = (n_pairs + 7) / 8, cap = 16LL * c->sm_count;
    sel_index_kernel<<<(unsigned)((want < cap) ? want : cap), 256, 0, (cudaStream_t)stream>>>(P, d_results, d_sel_idx);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return IGMK_OK;
}

extern "C" int igmk_actdist_sel_index_host(igmk_ctx* c, int64_t n_pairs, const int32_t* i, const int32_t* j,
                                           const igmk_pair_result* results, int mode, int32_t* sel_idx) {
    if (!c) return fail(IGMK_EINVAL, "igmk_actdist_sel_index_host: NULL context");
    if (n_pairs == 0) return IGMK_OK;
    if (n_pairs < 0 || !i || !j || !results || !sel_idx) return fail(IGMK_EINVAL, "igmk_actdist_sel_index_host: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t n = (size_t)n_pairs;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t off_j = up(n * 4), off_res = off_j + up(n * 4), off_out = off_res + up(n * sizeof(igmk_pair_result));
    int rc = ensure(&c->d_pairs, &c->pairs_bytes, off_out + n * 4);
    if (rc) return rc;
    char* base = (char*)c->d_pairs;
    CUDA_TRY(cudaMemcpyAsync(base, i, n * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(base + off_j, j, n * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(base + off_res, results, n * sizeof(igmk_pair_result), cudaMemcpyHostToDevice, c->stream));
    rc = igmk_actdist_sel_index_device(c, n_pairs, (const int32_t*)base, (const int32_t*)(base + off_j),
                                       (const igmk_pair_result*)(base + off_res), mode, (int32_t*)(base + off_out), c->stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(sel_idx, base + off_out, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return IGMK_OK;
}

extern "C" int igmk_finish_results_device(igmk_ctx* c, igmk_pair_result* d_results, int64_t n, void* stream) {
    if (!c) return fail(IGMK_EINVAL, "igmk_finish_results: NULL context");
    if (n < 0 || (n > 0 && !d_results)) return fail(IGMK_EINVAL, "igmk_finish_results: bad argument");
    if (n == 0) return IGMK_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    finish_results_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_results, n);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return IGMK_OK;
}

// ------------------------------------------------------------- DamID (next row f1)
extern "C" int igmk_damid_actdist_device(igmk_ctx* c, int64_t n_loci, const int32_t* d_loci,
                                         const float* d_pexp, const float* d_plast,
                                         double nucleus_radius, double contact_range, int it_corr,
                                         igmk_pair_result* d_out, void* stream) {
    if (!c) return fail(IGMK_EINVAL, "igmk_damid_actdist: NULL context");
    if (!c->have_coords || !c->have_index) return fail(IGMK_ESTATE, "igmk_damid_actdist: upload coordinates and index first");
    if (n_loci < 0) return fail(IGMK_EINVAL, "igmk_damid_actdist: negative n_loci");
    if (n_loci == 0) return IGMK_OK;
    if (!d_loci || !d_pexp || !d_plast || !d_out) return fail(IGMK_EINVAL, "igmk_damid_actdist: NULL buffer");
    CUDA_TRY(cudaSetDevice(c->device));
    ActdistParams P;
    memset(&P, 0, sizeof P);
    P.coords = c->d_coords; P.hap = c->d_hap;
    P.pi = d_loci; P.pj = nullptr; P.pwish = nullptr; P.plast = nullptr; P.out = d_out;
    P.n_pairs = n_loci; P.nstruct = c->nstruct; P.npad = c->npad; P.nchunks = c->nchunks;
    P.n_hap = c->n_hap; P.contact_range = 0.f; P.it_corr = it_corr; P.mode = IGMK_MODE_LB;
    P.negzero2 = 0x8000000080000000ull;
    P.block_stop = c->block_stop;
    P.pexp32 = d_pexp; P.plast32 = d_plast;
    P.damid_R = nucleus_radius * (1.0 - contact_range);       // np.array(nucleus_param) * (1 - contact_range), :436
    P.zero_bead = c->nbead;
    return dispatch_groups<true>(c, P, (cudaStream_t)stream);
}

extern "C" int igmk_damid_actdist_host(igmk_ctx* c, int64_t n_loci, const int32_t* loci,
                                       const float* pexp, const float* plast,
                                       double nucleus_radius, double contact_range, int it_corr,
                                       igmk_pair_result* out) {
    if (!c) return fail(IGMK_EINVAL, "igmk_damid_actdist_host: NULL context");
    if (n_loci == 0) return IGMK_OK;
    if (n_loci < 0 || !loci || !pexp || !plast || !out) return fail(IGMK_EINVAL, "igmk_damid_actdist_host: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t n = (size_t)n_loci;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t off_pe = up(n * 4), off_pl = off_pe + up(n * 4), off_out = off_pl + up(n * 4);
    int rc = ensure(&c->d_pairs, &c->pairs_bytes, off_out + n * sizeof(igmk_pair_result));
    if (rc) return rc;
    char* base = (char*)c->d_pairs;
    CUDA_TRY(cudaMemcpyAsync(base, loci, n * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(base + off_pe, pexp, n * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(base + off_pl, plast, n * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    rc = igmk_damid_actdist_device(c, n_loci, (const int32_t*)base, (const float*)(base + off_pe),
                                   (const float*)(base + off_pl), nucleus_radius, contact_range, it_corr,
                                   (igmk_pair_result*)(base + off_out), c->stream);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaMemcpyAsync(out, base + off_out, n * sizeof(igmk_pair_result), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaEventElapsedTime(&c->last_kernel_ms, c->ev0, c->ev1));
    return IGMK_OK;
}

// Host-buffer form of the A-step.  xyz == nullptr: the population is resident.  Otherwise
// the population (.hss layout, host memory) is staged by the same call, overlapped with the
// pair kernels: the pair list is worked off from its END in slices, and before slice k is
// launched only the beads its loci (and all higher ones) need are uploaded - for a list
// sorted by (i, j > i), as setup() writes it (ActivationDistanceStep.py:166-178), the last
// slice needs the top 1 / nslices of the loci, and each earlier slice a little more.
static int actdist_host_impl(igmk_ctx* c, const float* xyz, int64_t n_pairs,
                             const int32_t* i, const int32_t* j,
                             const double* pwish, const double* plast,
                             float contact_range, int it_corr, int mode, int algo,
                             igmk_pair_result* out) {
    const size_t n = (size_t)n_pairs;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t off_j = up(n * 4), off_pw = off_j + up(n * 4), off_pl = off_pw + up(n * 8);
    const size_t off_out = off_pl + up(n * 8);
    const size_t total = off_out + n * sizeof(igmk_pair_result);
    int rc = ensure(&c->d_pairs, &c->pairs_bytes, total);
    if (rc) return rc;
    char* base = (char*)c->d_pairs;
    // Three-stage pipeline over slices of the pair list: inputs of the next slice go up
    // (copy-in stream) and results of the previous one come down (copy-out stream) while
    // the kernel works on the current slice.  Overlap needs pinned host buffers
    // (igmk_host_alloc); pageable ones still work, serialised by the driver.
    // Slice sizes ramp up from slice >> ramp and down again (long lists only): the first slice's
    // inputs (and, pipelined, the beads it needs) and the last slice's results are the only
    // copies nothing hides.
    const size_t slice = (size_t)c->host_slice_pairs;
    std::vector<size_t> cut;                 // slice k = [cut[k], cut[k + 1])
    cut.push_back(0);
    const int ramp = (c->host_slice_ramp < 0) ? 0 : (c->host_slice_ramp > 4 ? 4 : c->host_slice_ramp);
    if (ramp && n >= 6 * slice && slice >= 1024) {
        const size_t small = slice >> ramp;
        size_t pos = 0;
        for (size_t sz = small; sz < slice; sz *= 2) { pos += sz; cut.push_back(pos); }
        const size_t tail = slice - small;                           // = small + 2 small + ... + slice / 2
        while (n - pos > tail + slice + slice / 2) { pos += slice; cut.push_back(pos); }
        const size_t mid = n - pos - tail;                           // in (slice/2, 3 slice/2]
        pos += mid; cut.push_back(pos);
        for (size_t sz = slice / 2; sz >= small; sz /= 2) { pos += sz; cut.push_back(pos); }
    } else {
        for (size_t pos = slice; pos < n; pos += slice) cut.push_back(pos);
        cut.push_back(n);
    }
    const int nsl = (int)cut.size() - 1;
    if ((int)c->ev_in.size() < nsl) {
        const size_t old = c->ev_in.size();
        c->ev_in.resize(nsl); c->ev_k.resize(nsl);
        for (size_t t = old; t < (size_t)nsl; ++t) {
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_in[t], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&c->ev_k[t], cudaEventDisableTiming));
        }
    }
    int cur[2] = {0, 0}, m_run = c->n_hap;
    if (xyz) {
        rc = upload_begin(c);
        if (rc) return rc;
        for (int r = 0; r < c->up_nreg; ++r) cur[r] = c->up_hi_reg[r];
        c->have_coords = true;           // every launch below is ordered after the pieces it reads
    }
    // consecutive slices alternate between two compute streams and two scratch sets (warp
    // list kernel only: the slab pipeline's batches want the L2 for themselves)
    const bool alt = c->host_overlap && nsl > 1 && c->list_form && c->group_threads == 0 &&
                     group_threads_for(c) == 32 && algo == IGMK_ALGO_FAST;
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    if (alt) CUDA_TRY(cudaStreamWaitEvent(c->stream2, c->ev0, 0));
    for (int t = 0; t < nsl; ++t) {
        const int k = xyz ? nsl - 1 - t : t;
        c->cur = alt ? (t & 1) : 0;
        cudaStream_t cs = c->cur ? c->stream2 : c->stream;
        const size_t lo = cut[k], cnt = cut[k + 1] - cut[k];
        if (xyz) {
            int m = m_run;
            for (size_t q = lo; q < lo + cnt; ++q) {
                const int a = i[q], b = j[q];
                m = (a < m) ? a : m;
                m = (b < m) ? b : m;
            }
            m_run = (m < 0) ? 0 : m;                   // (out-of-range pairs are flagged by the kernel)
            for (int r = 0; r < c->up_nreg; ++r) {
                const int want = c->up_lo[r][m_run];
                if (want < cur[r]) {
                    rc = upload_pieces(c, xyz + (size_t)want * c->nstruct * 3, want, cur[r] - want, c->s_up);
                    if (rc) { cudaDeviceSynchronize(); c->cur = 0; c->have_coords = false; return rc; }
                    cur[r] = want;
                }
            }
            if (c->up_piece > 0) CUDA_TRY(cudaStreamWaitEvent(cs, c->ev_up[2 + ((c->up_piece - 1) & 1)], 0));
        }
        CUDA_TRY(cudaMemcpyAsync(base + lo * 4, i + lo, cnt * 4, cudaMemcpyHostToDevice, c->s_in));
        CUDA_TRY(cudaMemcpyAsync(base + off_j + lo * 4, j + lo, cnt * 4, cudaMemcpyHostToDevice, c->s_in));
        CUDA_TRY(cudaMemcpyAsync(base + off_pw + lo * 8, pwish + lo, cnt * 8, cudaMemcpyHostToDevice, c->s_in));
        CUDA_TRY(cudaMemcpyAsync(base + off_pl + lo * 8, plast + lo, cnt * 8, cudaMemcpyHostToDevice, c->s_in));
        CUDA_TRY(cudaEventRecord(c->ev_in[k], c->s_in));
        CUDA_TRY(cudaStreamWaitEvent(cs, c->ev_in[k], 0));
        igmk_pair_result* d_out = (igmk_pair_result*)(base + off_out) + lo;
        rc = igmk_actdist_device(c, (int64_t)cnt, (const int32_t*)base + lo, (const int32_t*)(base + off_j) + lo,
                                 (const double*)(base + off_pw) + lo, (const double*)(base + off_pl) + lo,
                                 contact_range, it_corr, mode, algo, d_out, cs);
        if (rc) { cudaDeviceSynchronize(); c->cur = 0; if (xyz) c->have_coords = false; return rc; }
        CUDA_TRY(cudaEventRecord(c->ev_k[k], cs));
        CUDA_TRY(cudaStreamWaitEvent(c->s_out, c->ev_k[k], 0));
        CUDA_TRY(cudaMemcpyAsync(out + lo, d_out, cnt * sizeof(igmk_pair_result), cudaMemcpyDeviceToHost, c->s_out));
    }
    c->cur = 0;
    if (alt) {
        CUDA_TRY(cudaEventRecord(c->ev_join, c->stream2));
        CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    }
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    if (xyz) {
        // the beads no pair of the list touched: the context ends up fully staged
        int iv[2][2], niv = c->up_nreg;
        for (int r = 0; r < niv; ++r) { iv[r][0] = cur[r]; iv[r][1] = c->up_hi_reg[r]; }
        if (niv == 2 && iv[1][0] < iv[0][0]) { std::swap(iv[0][0], iv[1][0]); std::swap(iv[0][1], iv[1][1]); }
        int pos = 0;
        for (int r = 0; r <= niv; ++r) {
            const int end = (r < niv) ? iv[r][0] : c->nbead;
            if (end > pos) {
                rc = upload_pieces(c, xyz + (size_t)pos * c->nstruct * 3, pos, end - pos, c->s_up);
                if (rc) { cudaDeviceSynchronize(); c->have_coords = false; return rc; }
            }
            if (r < niv) pos = (iv[r][1] > pos) ? iv[r][1] : pos;
        }
        CUDA_TRY(cudaStreamSynchronize(c->s_up));
    }
    CUDA_TRY(cudaStreamSynchronize(c->s_out));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaEventElapsedTime(&c->last_kernel_ms, c->ev0, c->ev1));
    return IGMK_OK;
}

// Device buffers of the host entry points for lists of up to n_pairs pairs, allocated now
// instead of growing (cudaFree + cudaMalloc, tens of milliseconds and a device-wide
// synchronisation each) from one A-step's list to the next larger one.
extern "C" int igmk_reserve_pairs(igmk_ctx* c, int64_t n_pairs) {
    if (!c || n_pairs < 0) return fail(IGMK_EINVAL, "igmk_reserve_pairs: bad argument");
    if (n_pairs == 0) return IGMK_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t n = (size_t)n_pairs;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    int rc = ensure(&c->d_pairs, &c->pairs_bytes, 2 * up(n * 4) + 2 * up(n * 8) + n * sizeof(igmk_pair_result));
    if (rc) return rc;
    // per-launch scratch: the host pipeline launches slices of at most 3/2 host_slice_pairs
    size_t m = (size_t)c->host_slice_pairs * 3 / 2;
    if (m > n) m = n;
    size_t temp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, (int)m, 0, 16, c->stream);
    for (int k = 0; k < 2; ++k) {
        LaunchScratch& sc = c->scr[k];
        if ((rc = ensure(&sc.d_order, &sc.order_bytes, 4 * up(m * 4) + up(temp) + 4096))) return rc;
        if ((rc = ensure(&sc.d_redo, &sc.redo_bytes, 256 + m * sizeof(int32_t)))) return rc;
        if ((rc = ensure(&sc.d_rec, &sc.rec_bytes, m * sizeof(PairRec)))) return rc;
    }
    return IGMK_OK;
}

extern "C" int igmk_actdist_host(igmk_ctx* c, int64_t n_pairs,
                                 const int32_t* i, const int32_t* j,
                                 const double* pwish, const double* plast,
                                 float contact_range, int it_corr, int mode, int algo,
                                 igmk_pair_result* out) {
    if (!c) return fail(IGMK_EINVAL, "igmk_actdist_host: NULL context");
    if (n_pairs == 0) return IGMK_OK;
    if (n_pairs < 0 || !i || !j || !pwish || !plast || !out) return fail(IGMK_EINVAL, "igmk_actdist_host: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    return actdist_host_impl(c, nullptr, n_pairs, i, j, pwish, plast, contact_range, it_corr, mode, algo, out);
}

extern "C" int igmk_actdist_host_population(igmk_ctx* c, const float* xyz, int64_t n_pairs,
                                            const int32_t* i, const int32_t* j,
                                            const double* pwish, const double* plast,
                                            float contact_range, int it_corr, int mode, int algo,
                                            igmk_pair_result* out) {
    if (!c || !xyz) return fail(IGMK_EINVAL, "igmk_actdist_host_population: NULL argument");
    if (!c->have_index) return fail(IGMK_ESTATE, "igmk_actdist_host_population: set the index first");
    if (n_pairs < 0 || (n_pairs > 0 && (!i || !j || !pwish || !plast || !out)))
        return fail(IGMK_EINVAL, "igmk_actdist_host_population: bad argument");
    if (mode != IGMK_MODE_LB && mode != IGMK_MODE_GP) return fail(IGMK_EINVAL, "igmk_actdist_host_population: bad mode %d", mode);
    if (algo != IGMK_ALGO_FAST && algo != IGMK_ALGO_SIMPLE) return fail(IGMK_EINVAL, "igmk_actdist_host_population: bad algo %d", algo);
    if (n_pairs == 0) return igmk_upload_coords(c, xyz, 0);
    CUDA_TRY(cudaSetDevice(c->device));
    c->have_coords = false;
    return actdist_host_impl(c, xyz, n_pairs, i, j, pwish, plast, contact_range, it_corr, mode, algo, out);
}

extern "C" float igmk_last_kernel_ms(igmk_ctx* c) { return c ? c->last_kernel_ms : 0.f; }

// ---------------------------------------------------------------- host phases of setup()
// Candidate filter of the reference's setup loop (ActivationDistanceStep.py:166-178) over
// a strict-upper-triangle CSR matrix: entry (r, c, p) is kept if r != c and
// p >= intra_sigma (same chromosome) / p >= inter_sigma (different chromosomes), the
// comparison in float32 as NumPy >= 2 evaluates `float32 >= python float`.  Pure host
// code (no device needed).  Returns the number of candidates, or -1 - needed when
// `capacity` is too small (nothing is written beyond capacity).
// Host threads of the setup / task phases (IGMK_HOST_THREADS, default min(8, cores)): the
// lists are tens of MB, one core leaves them at ~50 ms per pass.
static int host_threads(int64_t work) {
    static const int conf = [] {
        const char* ov = getenv("IGMK_HOST_THREADS");
        int t = ov ? atoi(ov) : (int)std::thread::hardware_concurrency();
        if (!ov && t > 8) t = 8;
        return t < 1 ? 1 : t;
    }();
    const int64_t by_work = work / 65536;
    return (int)((by_work < 1) ? 1 : (by_work < conf ? by_work : conf));
}
template <class F>
static void run_threads(int nt, F&& body) {
    if (nt <= 1) { body(0); return; }
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    for (int t = 1; t < nt; ++t) th.emplace_back([&body, t] { body(t); });
    body(0);
    for (auto& x : th) x.join();
}

extern "C" int64_t igmk_filter_candidates(int64_t n_rows, const int64_t* indptr, const int32_t* indices,
                                          const float* data, const int32_t* chrom,
                                          int use_intra, float intra_sigma, int use_inter, float inter_sigma,
                                          int32_t* out_i, int32_t* out_j, double* out_p, int64_t capacity) {
    if (n_rows < 0 || !indptr || (indptr[n_rows] > 0 && (!indices || !data || !chrom))) return -1;
    const int64_t nnz = indptr[n_rows];
    const int nt = host_threads(nnz);
    // row ranges of about equal stored entries; count, then write at the prefix offsets
    std::vector<int64_t> r0((size_t)nt + 1, n_rows), cnt((size_t)nt + 1, 0);
    r0[0] = 0;
    for (int t = 1; t < nt; ++t)
        r0[t] = std::lower_bound(indptr, indptr + n_rows, nnz / nt * t) - indptr;
    auto keep_entry = [&](int64_t r, int cr, int64_t e) {
        const int32_t c = indices[e];
        const float p = data[e];
        const bool intra = chrom[c] == cr;
        return (intra ? (use_intra && p >= intra_sigma) : (use_inter && p >= inter_sigma)) && c != (int32_t)r;
    };
    run_threads(nt, [&](int t) {
        int64_t k = 0;
        for (int64_t r = r0[t]; r < r0[t + 1]; ++r) {
            const int cr = chrom[r];
            for (int64_t e = indptr[r]; e < indptr[r + 1]; ++e) k += keep_entry(r, cr, e) ? 1 : 0;
        }
        cnt[t + 1] = k;
    });
    for (int t = 0; t < nt; ++t) cnt[t + 1] += cnt[t];
    const int64_t total = cnt[nt];
    if (total > capacity) return -1 - total;
    run_threads(nt, [&](int t) {
        int64_t k = cnt[t];
        for (int64_t r = r0[t]; r < r0[t + 1]; ++r) {
            const int cr = chrom[r];
            for (int64_t e = indptr[r]; e < indptr[r + 1]; ++e) {
                if (!keep_entry(r, cr, e)) continue;
                out_i[k] = (int32_t)r; out_j[k] = indices[e]; out_p[k] = (double)data[e];
                ++k;
            }
        }
    });
    return total;
}

// plast[i, j] of setup (:144-160, :177) as a merge join: the stored records (row, col, prob)
// restricted to row < n and col < n, against the candidate list (ii, jj), both in strictly
// increasing (row, col) order - what files written by this step and CSR-ordered candidate
// lists are.  Returns 1 when done, 0 when either side is not strictly increasing (the caller
// then takes the general path), -1 on bad arguments.
extern "C" int igmk_join_plast(int64_t n_rec, const int32_t* row, const int32_t* col, const float* prob,
                               int32_t n, int64_t n_pairs, const int32_t* ii, const int32_t* jj,
                               double* out) {
    if (n_rec < 0 || n_pairs < 0 || (n_rec > 0 && (!row || !col || !prob)) || (n_pairs > 0 && (!ii || !jj || !out)))
        return -1;
    auto key = [n](int32_t a, int32_t b) { return (int64_t)a * n + b; };
    const int nt = host_threads(n_rec + n_pairs);
    std::atomic<int> ordered{1};
    // candidate side: strictly increasing (checked range by range, one element of overlap); zero the output
    run_threads(nt, [&](int t) {
        const int64_t lo = n_pairs * t / nt, hi = n_pairs * (t + 1) / nt;
        int64_t last = (lo > 0) ? key(ii[lo - 1], jj[lo - 1]) : -1;
        bool ok = true;
        for (int64_t q = lo; q < hi; ++q) {
            const int64_t kq = key(ii[q], jj[q]);
            ok = ok && kq > last;
            last = kq;
            out[q] = 0.0;
        }
        if (!ok) ordered.store(0);
    });
    if (!ordered.load()) return 0;
    // record side: the records with row < n and col < n strictly increasing inside every range ...
    std::vector<int64_t> first((size_t)nt, -1), lastk((size_t)nt, -1);
    run_threads(nt, [&](int t) {
        const int64_t lo = n_rec * t / nt, hi = n_rec * (t + 1) / nt;
        int64_t last = -1, fst = -1;
        bool ok = true;
        for (int64_t e = lo; e < hi; ++e) {
            if (row[e] >= n || col[e] >= n) continue;
            const int64_t kr = key(row[e], col[e]);
            if (fst < 0) fst = kr;
            ok = ok && kr > last;
            last = kr;
        }
        first[t] = fst; lastk[t] = last;
        if (!ok) ordered.store(0);
    });
    if (!ordered.load()) return 0;
    {   // ... and across the ranges
        int64_t last = -1;
        for (int t = 0; t < nt; ++t) {
            if (first[t] < 0) continue;
            if (first[t] <= last) return 0;
            last = lastk[t];
        }
    }
    // merge: every record range starts at the candidate its first key points to; the keys of
    // different ranges are disjoint, so the ranges write different candidates
    run_threads(nt, [&](int t) {
        if (first[t] < 0) return;
        const int64_t lo = n_rec * t / nt, hi = n_rec * (t + 1) / nt;
        int64_t a = 0, b = n_pairs;                       // first candidate with key >= first[t]
        while (a < b) {
            const int64_t m = (a + b) >> 1;
            if (key(ii[m], jj[m]) < first[t]) a = m + 1; else b = m;
        }
        int64_t q = a;
        for (int64_t e = lo; e < hi && q < n_pairs; ++e) {
            if (row[e] >= n || col[e] >= n) continue;
            const int64_t kr = key(row[e], col[e]);
            while (q < n_pairs && key(ii[q], jj[q]) < kr) ++q;
            if (q < n_pairs && key(ii[q], jj[q]) == kr) out[q] = (double)prob[e];
        }
    });
    return 1;
}

// Record expansion, host side: task() flattening + get_actdist's result lists
// (ActivationDistanceStep.py:221-222, 476-483).
extern "C" int igmk_expand_records(igmk_ctx* c, int64_t n_pairs,
                                   const int32_t* pi, const int32_t* pj,
                                   const igmk_pair_result* res,
                                   int32_t* row, int32_t* col, float* dist, float* prob,
                                   int64_t capacity, int64_t* n_records) {
    if (!c || !c->have_index) return fail(IGMK_ESTATE, "igmk_expand_records: index not set");
    if (n_pairs < 0 || (n_pairs > 0 && (!pi || !pj || !res))) return fail(IGMK_EINVAL, "igmk_expand_records: bad argument");
    const int nt = host_threads(n_pairs);
    // records of a pair: the kernel's nrec, re-derived here from the index (and checked against
    // it while writing); offsets by a prefix sum over pair ranges
    std::vector<int64_t> off((size_t)nt + 1, 0), bad((size_t)nt, -1);
    const int n_hap = c->n_hap;
    run_threads(nt, [&](int t) {
        const int64_t lo = n_pairs * t / nt, hi = n_pairs * (t + 1) / nt;
        int64_t k = 0;
        for (int64_t q = lo; q < hi; ++q) {
            if (res[q].nrec <= 0) continue;
            if (pi[q] < 0 || pj[q] < 0 || pi[q] >= n_hap || pj[q] >= n_hap) { if (bad[t] < 0) bad[t] = q; continue; }
            k += res[q].nrec;
        }
        off[t + 1] = k;
    });
    for (int t = 0; t < nt; ++t) {
        if (bad[t] >= 0) return fail(IGMK_EINVAL, "igmk_expand_records: pair %lld out of range", (long long)bad[t]);
        off[t + 1] += off[t];
    }
    if (off[nt] > capacity) return fail(IGMK_EINVAL, "igmk_expand_records: capacity %lld too small", (long long)capacity);
    std::atomic<int> mismatch{0};
    run_threads(nt, [&](int t) {
        const int64_t lo = n_pairs * t / nt, hi = n_pairs * (t + 1) / nt;
        int64_t k = off[t];
        for (int64_t q = lo; q < hi; ++q) {
            const igmk_pair_result& r = res[q];
            if (r.nrec <= 0) continue;
            const HapEntry& a = c->h_hap[pi[q]];
            const HapEntry& b = c->h_hap[pj[q]];
            const int ab[2] = {a.b0, a.b1}, bb[2] = {b.b0, b.b1};
            const int na = a.b1 >= 0 ? 2 : 1, nb = b.b1 >= 0 ? 2 : 1;
            const int64_t k0 = k;
            if (k + (a.chrom == b.chrom ? (na < nb ? na : nb) : na * nb) > off[t + 1]) { mismatch.store(1); return; }
            if (a.chrom == b.chrom) {
                const int m = na < nb ? na : nb;                    // zip(ii, jj)
                for (int u = 0; u < m; ++u) { row[k] = ab[u]; col[k] = bb[u]; dist[k] = r.dist; prob[k] = r.prob; ++k; }
            } else {
                for (int u = 0; u < na; ++u)                        // for i0 in ii for i1 in jj
                    for (int w = 0; w < nb; ++w) { row[k] = ab[u]; col[k] = bb[w]; dist[k] = r.dist; prob[k] = r.prob; ++k; }
            }
            if (k - k0 != r.nrec) { mismatch.store(1); return; }
        }
    });
    if (mismatch.load()) return fail(IGMK_EINVAL, "igmk_expand_records: record counts do not match the index");
    if (n_records) *n_records = off[nt];
    return IGMK_OK;
}

// ------------------------------------------------------------- K2 launches
static int contact_launch(igmk_ctx* c, int haploid, int row0, int nrows, int col0, int ncols,
                          float contact_range, int strict, uint32_t* d_counts, void* stream) {
    if (!c) return fail(IGMK_EINVAL, "igmk_contact_counts: NULL context");
    if (!c->have_coords || !c->have_index) return fail(IGMK_ESTATE, "igmk_contact_counts: upload coordinates and index first");
    const int lim = haploid ? c->n_hap : c->nbead;
    if (row0 < 0 || col0 < 0 || nrows < 0 || ncols < 0 || row0 + nrows > lim || col0 + ncols > lim)
        return fail(IGMK_EINVAL, "igmk_contact_counts: tile out of range");
    if (nrows == 0 || ncols == 0) return IGMK_OK;
    if (!d_counts) return fail(IGMK_EINVAL, "igmk_contact_counts: NULL output");
    CUDA_TRY(cudaSetDevice(c->device));
    ContactParams P;
    P.coords = c->d_coords; P.radii = c->d_radii; P.counts = d_counts;
    P.nstruct = c->nstruct; P.npad = c->npad; P.nbead = c->nbead;
    P.row0 = row0; P.nrows = nrows; P.col0 = col0; P.ncols = ncols;
    P.contact_range = contact_range; P.strict = strict;
    P.negzero2 = 0x8000000080000000ull;
    P.hap = c->d_hap; P.haploid = haploid;
    dim3 grid((ncols + kCtTile - 1) / kCtTile, (nrows + kCtTile - 1) / kCtTile);
    if (haploid) {
        if (kCtDynBytes > 48 * 1024)
            CUDA_TRY(cudaFuncSetAttribute(contact_tile_hap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCtDynBytes));
        contact_tile_hap_kernel<<<grid, kCtThreads, kCtDynBytes, (cudaStream_t)stream>>>(P);
    }
    else {
        if (kCtDynBytes > 48 * 1024)
            CUDA_TRY(cudaFuncSetAttribute(contact_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCtDynBytes));
        contact_tile_kernel<<<grid, kCtThreads, kCtDynBytes, (cudaStream_t)stream>>>(P);
    }
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return IGMK_OK;
}

static int contact_host(igmk_ctx* c, int haploid, int row0, int nrows, int col0, int ncols,
                        float contact_range, int strict, uint32_t* counts) {
    if (!c) return fail(IGMK_EINVAL, "igmk_contact_counts_host: NULL context");
    if (nrows == 0 || ncols == 0) return IGMK_OK;
    if (nrows < 0 || ncols < 0 || !counts) return fail(IGMK_EINVAL, "igmk_contact_counts_host: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t bytes = (size_t)nrows * ncols * sizeof(uint32_t);
    int rc = ensure(&c->d_stage, &c->stage_bytes, bytes);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    rc = contact_launch(c, haploid, row0, nrows, col0, ncols, contact_range, strict,
                        (uint32_t*)c->d_stage, c->stream);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaMemcpyAsync(counts, c->d_stage, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaEventElapsedTime(&c->last_kernel_ms, c->ev0, c->ev1));
    return IGMK_OK;
}

extern "C" int igmk_contact_counts_device(igmk_ctx* c, int row0, int nrows, int col0, int ncols,
                                          float contact_range, int strict,
                                          uint32_t* d_counts, void* stream) {
    return contact_launch(c, 0, row0, nrows, col0, ncols, contact_range, strict, d_counts, stream);
}
extern "C" int igmk_contact_counts_host(igmk_ctx* c, int row0, int nrows, int col0, int ncols,
                                        float contact_range, int strict, uint32_t* counts) {
    return contact_host(c, 0, row0, nrows, col0, ncols, contact_range, strict, counts);
}
extern "C" int igmk_contact_counts_haploid_device(igmk_ctx* c, int row0, int nrows, int col0, int ncols,
                                                  float contact_range, int strict,
                                                  uint32_t* d_counts, void* stream) {
    return contact_launch(c, 1, row0, nrows, col0, ncols, contact_range, strict, d_counts, stream);
}
extern "C" int igmk_contact_counts_haploid_host(igmk_ctx* c, int row0, int nrows, int col0, int ncols,
                                                float contact_range, int strict, uint32_t* counts) {
    return contact_host(c, 1, row0, nrows, col0, ncols, contact_range, strict, counts);
}

// ------------------------------------------------------------- K3 (next row f2)
extern "C" int igmk_set_bead_chrom(igmk_ctx* c, const int32_t* chrom_bead) {
    if (!c || !chrom_bead) return fail(IGMK_EINVAL, "igmk_set_bead_chrom: NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    if (!c->d_chrom) CUDA_TRY(cudaMalloc(&c->d_chrom, (size_t)c->nbead * sizeof(int32_t)));
    CUDA_TRY(cudaMemcpy(c->d_chrom, chrom_bead, (size_t)c->nbead * sizeof(int32_t), cudaMemcpyHostToDevice));
    return IGMK_OK;
}

extern "C" int igmk_restraint_words(igmk_ctx* c) { return c ? c->npad / 32 : 0; }

extern "C" int igmk_restraint_select_device(igmk_ctx* c, int64_t n_rec, const int32_t* d_row,
                                            const int32_t* d_col, const float* d_dist, int kind,
                                            uint32_t* d_bitmap, int32_t* d_counts, void* stream) {
    if (!c) return fail(IGMK_EINVAL, "igmk_restraint_select: NULL context");
    if (!c->have_coords) return fail(IGMK_ESTATE, "igmk_restraint_select: upload coordinates first");
    if (kind < 0 || kind > 2) return fail(IGMK_EINVAL, "igmk_restraint_select: bad kind %d", kind);
    if (kind != 2 && !c->d_chrom) return fail(IGMK_ESTATE, "igmk_restraint_select: igmk_set_bead_chrom first");
    if (n_rec < 0) return fail(IGMK_EINVAL, "igmk_restraint_select: negative n_rec");
    if (n_rec == 0) return IGMK_OK;
    if (!d_row || !d_col || !d_dist || !d_bitmap || !d_counts) return fail(IGMK_EINVAL, "igmk_restraint_select: NULL buffer");
    CUDA_TRY(cudaSetDevice(c->device));
    RestraintParams P;
    P.coords = c->d_coords; P.chrom = c->d_chrom; P.row = d_row; P.col = d_col; P.dist = d_dist;
    P.bitmap = d_bitmap; P.counts = d_counts; P.n_rec = n_rec;
    P.nstruct = c->nstruct; P.npad = c->npad; P.nbead = c->nbead; P.kind = kind;
    long long want = (n_rec + kRsWarps - 1) / kRsWarps, cap = (long long)c->sm_count * 8;
    restraint_select_kernel<<<(unsigned)(want < cap ? want : cap), 32 * kRsWarps, 0, (cudaStream_t)stream>>>(P);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return IGMK_OK;
}

extern "C" int igmk_restraint_select_host(igmk_ctx* c, int64_t n_rec, const int32_t* row,
                                          const int32_t* col, const float* dist, int kind,
                                          uint32_t* bitmap, int32_t* counts) {
    if (!c) return fail(IGMK_EINVAL, "igmk_restraint_select_host: NULL context");
    if (n_rec == 0) return IGMK_OK;
    if (n_rec < 0 || !row || !col || !dist || !bitmap || !counts) return fail(IGMK_EINVAL, "igmk_restraint_select_host: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t n = (size_t)n_rec, words = (size_t)c->npad / 32;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t off_c = up(n * 4), off_d = off_c + up(n * 4), off_n = off_d + up(n * 4), off_b = off_n + up(n * 4);
    int rc = ensure(&c->d_pairs, &c->pairs_bytes, off_b + n * words * 4);
    if (rc) return rc;
    char* base = (char*)c->d_pairs;
    CUDA_TRY(cudaMemcpyAsync(base, row, n * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(base + off_c, col, n * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(base + off_d, dist, n * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    rc = igmk_restraint_select_device(c, n_rec, (const int32_t*)base, (const int32_t*)(base + off_c),
                                      (const float*)(base + off_d), kind, (uint32_t*)(base + off_b),
                                      (int32_t*)(base + off_n), c->stream);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaMemcpyAsync(counts, base + off_n, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(bitmap, base + off_b, n * words * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaEventElapsedTime(&c->last_kernel_ms, c->ev0, c->ev1));
    return IGMK_OK;
}

// ------------------------------------------------------------- K4 (next row f3)
extern "C" int igmk_sprite_cluster_rg2_host(igmk_ctx* c, int n_clusters, const int32_t* seg_ptr,
                                            const int32_t* loc_ptr, const int32_t* beads,
                                            const int32_t* seg_group, const int32_t* group_ptr,
                                            const int32_t* sel, float* rg2s) {
    if (!c) return fail(IGMK_EINVAL, "igmk_sprite_cluster_rg2: NULL context");
    if (!c->have_coords) return fail(IGMK_ESTATE, "igmk_sprite_cluster_rg2: upload coordinates first");
    if (n_clusters < 0) return fail(IGMK_EINVAL, "igmk_sprite_cluster_rg2: negative n_clusters");
    if (n_clusters == 0) return IGMK_OK;
    if (!seg_ptr || !loc_ptr || !beads || !seg_group || !group_ptr || !sel || !rg2s)
        return fail(IGMK_EINVAL, "igmk_sprite_cluster_rg2: NULL buffer");
    if (seg_ptr[0] != 0 || loc_ptr[0] != 0 || group_ptr[0] != 0)
        return fail(IGMK_EINVAL, "igmk_sprite_cluster_rg2: CSR arrays must start at 0");
    const int nseg_tot = seg_ptr[n_clusters], ngrp_tot = group_ptr[n_clusters];
    const int nloc_tot = loc_ptr[nseg_tot];
    for (int k = 0; k < n_clusters; ++k) {
        const int ng = group_ptr[k + 1] - group_ptr[k];
        if (seg_ptr[k + 1] <= seg_ptr[k] || ng <= 0)
            return fail(IGMK_EINVAL, "igmk_sprite_cluster_rg2: cluster %d is empty", k);
        for (int i = seg_ptr[k]; i < seg_ptr[k + 1]; ++i) {
            if (loc_ptr[i + 1] <= loc_ptr[i]) return fail(IGMK_EINVAL, "igmk_sprite_cluster_rg2: segment %d has no location", i);
            if (seg_group[i] < 0 || seg_group[i] >= ng) return fail(IGMK_EINVAL, "igmk_sprite_cluster_rg2: segment %d: selection column out of range", i);
        }
    }
    for (int b = 0; b < nloc_tot; ++b)
        if (beads[b] < 0 || beads[b] >= c->nbead) return fail(IGMK_EINVAL, "igmk_sprite_cluster_rg2: bead id out of range");
    CUDA_TRY(cudaSetDevice(c->device));
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t N = (size_t)c->nstruct;
    const size_t o_lp = up((size_t)(n_clusters + 1) * 4), o_b = o_lp + up((size_t)(nseg_tot + 1) * 4);
    const size_t o_sg = o_b + up((size_t)nloc_tot * 4), o_gp = o_sg + up((size_t)nseg_tot * 4);
    const size_t o_sel = o_gp + up((size_t)(n_clusters + 1) * 4), o_rg = o_sel + up((size_t)ngrp_tot * N * 4);
    const size_t total = o_rg + up((size_t)n_clusters * N * 4);
    int rc = ensure(&c->d_pairs, &c->pairs_bytes, total);
    if (rc) return rc;
    char* base = (char*)c->d_pairs;
    CUDA_TRY(cudaMemcpyAsync(base, seg_ptr, (size_t)(n_clusters + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(base + o_lp, loc_ptr, (size_t)(nseg_tot + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(base + o_b, beads, (size_t)nloc_tot * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(base + o_sg, seg_group, (size_t)nseg_tot * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(base + o_gp, group_ptr, (size_t)(n_clusters + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(base + o_sel, sel, (size_t)ngrp_tot * N * 4, cudaMemcpyHostToDevice, c->stream));
    SpriteClusterParams P;
    P.coords = c->d_coords; P.seg_ptr = (const int32_t*)base; P.loc_ptr = (const int32_t*)(base + o_lp);
    P.beads = (const int32_t*)(base + o_b); P.seg_group = (const int32_t*)(base + o_sg);
    P.group_ptr = (const int32_t*)(base + o_gp); P.sel = (const int32_t*)(base + o_sel);
    P.rg2s = (float*)(base + o_rg);
    P.n_clusters = n_clusters; P.nstruct = c->nstruct; P.npad = c->npad;
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    for (int k0 = 0; k0 < n_clusters; k0 += 65535) {        // gridDim.y limit
        SpriteClusterParams Q = P;
        const int nk = (n_clusters - k0 < 65535) ? n_clusters - k0 : 65535;
        Q.seg_ptr = P.seg_ptr + k0; Q.group_ptr = P.group_ptr + k0; Q.rg2s = P.rg2s + (size_t)k0 * N; Q.n_clusters = nk;
        dim3 grid((c->nstruct + 127) / 128, nk);
        sprite_cluster_rg2_kernel<<<grid, 128, 0, c->stream>>>(Q);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaMemcpyAsync(rg2s, base + o_rg, (size_t)n_clusters * N * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaEventElapsedTime(&c->last_kernel_ms, c->ev0, c->ev1));
    return IGMK_OK;
}

extern "C" int igmk_sprite_rg2_host(igmk_ctx* c, int n_clusters, const int32_t* region_ptr,
                                    const int32_t* copy_ptr, const int32_t* beads,
                                    float* rg2s, int32_t* copy_idx, int32_t* min_struct) {
    if (!c) return fail(IGMK_EINVAL, "igmk_sprite_rg2: NULL context");
    if (!c->have_coords) return fail(IGMK_ESTATE, "igmk_sprite_rg2: upload coordinates first");
    if (n_clusters < 0) return fail(IGMK_EINVAL, "igmk_sprite_rg2: negative n_clusters");
    if (n_clusters == 0) return IGMK_OK;
    if (!region_ptr || !copy_ptr || !beads || !rg2s || !copy_idx || !min_struct)
        return fail(IGMK_EINVAL, "igmk_sprite_rg2: NULL buffer");
    const int nreg_tot = region_ptr[n_clusters];
    if (region_ptr[0] != 0 || copy_ptr[0] != 0) return fail(IGMK_EINVAL, "igmk_sprite_rg2: CSR arrays must start at 0");
    const int ncopy_tot = copy_ptr[nreg_tot];
    for (int k = 0; k < n_clusters; ++k) {
        const int r0 = region_ptr[k], r1 = region_ptr[k + 1];
        if (r1 <= r0 || r1 - r0 > kSpMaxRegions)
            return fail(IGMK_ELIMIT, "igmk_sprite_rg2: cluster %d has %d regions (1..%d supported)", k, r1 - r0, kSpMaxRegions);
        double comb = 1.0;
        for (int i = r0; i < r1; ++i) {
            if (copy_ptr[i + 1] <= copy_ptr[i]) return fail(IGMK_EINVAL, "igmk_sprite_rg2: region %d has no location", i);
            comb *= copy_ptr[i + 1] - copy_ptr[i];
        }
        if (copy_ptr[r1] - copy_ptr[r0] > kSpMaxCopies || comb > 16777216.0)
            return fail(IGMK_ELIMIT, "igmk_sprite_rg2: cluster %d is too large (%d locations, %.0f combinations)", k, copy_ptr[r1] - copy_ptr[r0], comb);
    }
    for (int b = 0; b < ncopy_tot; ++b)
        if (beads[b] < 0 || beads[b] >= c->nbead) return fail(IGMK_EINVAL, "igmk_sprite_rg2: bead id out of range");
    CUDA_TRY(cudaSetDevice(c->device));
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t N = (size_t)c->nstruct;
    const size_t o_cp = up((size_t)(n_clusters + 1) * 4), o_b = o_cp + up((size_t)(nreg_tot + 1) * 4);
    const size_t o_rg = o_b + up((size_t)ncopy_tot * 4), o_ci = o_rg + up((size_t)n_clusters * N * 4);
    const size_t o_ms = o_ci + up((size_t)nreg_tot * N * 4), total = o_ms + up((size_t)n_clusters * 4);
    int rc = ensure(&c->d_pairs, &c->pairs_bytes, total);
    if (rc) return rc;
    char* base = (char*)c->d_pairs;
    CUDA_TRY(cudaMemcpyAsync(base, region_ptr, (size_t)(n_clusters + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(base + o_cp, copy_ptr, (size_t)(nreg_tot + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(base + o_b, beads, (size_t)ncopy_tot * 4, cudaMemcpyHostToDevice, c->stream));
    SpriteParams P;
    P.coords = c->d_coords; P.region_ptr = (const int32_t*)base; P.copy_ptr = (const int32_t*)(base + o_cp);
    P.beads = (const int32_t*)(base + o_b); P.rg2s = (float*)(base + o_rg); P.copy_idx = (int32_t*)(base + o_ci);
    P.min_struct = (int32_t*)(base + o_ms);
    P.n_clusters = n_clusters; P.nstruct = c->nstruct; P.npad = c->npad; P.nbead = c->nbead;
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    for (int k0 = 0; k0 < n_clusters; k0 += 65535) {        // gridDim.y limit
        SpriteParams Q = P;
        const int nk = (n_clusters - k0 < 65535) ? n_clusters - k0 : 65535;
        Q.region_ptr = P.region_ptr + k0; Q.rg2s = P.rg2s + (size_t)k0 * N; Q.n_clusters = nk;
        dim3 grid((c->nstruct + 127) / 128, nk);
        sprite_rg2_kernel<<<grid, 128, 0, c->stream>>>(Q);
        g_launches++;
        CUDA_TRY(cudaGetLastError());
    }
    sprite_argmin_kernel<<<n_clusters, 256, 0, c->stream>>>(P.rg2s, c->nstruct, P.min_struct);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaMemcpyAsync(rg2s, base + o_rg, (size_t)n_clusters * N * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(copy_idx, base + o_ci, (size_t)nreg_tot * N * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(min_struct, base + o_ms, (size_t)n_clusters * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaEventElapsedTime(&c->last_kernel_ms, c->ev0, c->ev1));
    return IGMK_OK;
}

// ------------------------------------------------------------- K5 (next row f4)
extern "C" int igmk_rank_match_device(igmk_ctx* c, int64_t n_items, const int32_t* d_a, const int32_t* d_b,
                                      int reduce, const float* d_target, int64_t target_stride,
                                      float* d_matched, int32_t* d_rank, float* d_value, void* stream) {
    if (!c) return fail(IGMK_EINVAL, "igmk_rank_match: NULL context");
    if (!c->have_coords) return fail(IGMK_ESTATE, "igmk_rank_match: upload coordinates first");
    if (n_items < 0) return fail(IGMK_EINVAL, "igmk_rank_match: negative n_items");
    if (reduce != 0 && reduce != 1) return fail(IGMK_EINVAL, "igmk_rank_match: bad reduce %d", reduce);
    if (n_items == 0) return IGMK_OK;
    if (!d_a || (d_matched && !d_target) || target_stride < 0) return fail(IGMK_EINVAL, "igmk_rank_match: bad buffer");
    if (c->nstruct > 16384) return fail(IGMK_ELIMIT, "igmk_rank_match: nstruct = %d exceeds the supported 16384", c->nstruct);
    CUDA_TRY(cudaSetDevice(c->device));
    RankParams P;
    P.coords = c->d_coords; P.a = d_a; P.b = d_b; P.target = d_target; P.target_stride = target_stride;
    P.matched = d_matched; P.rank = d_rank; P.value = d_value;
    P.n_items = n_items; P.nstruct = c->nstruct; P.npad = c->npad; P.nbead = c->nbead;
    P.zero_bead = c->nbead; P.reduce = reduce;
    int m = 2;
    while (m < c->nstruct) m <<= 1;
    P.m = m;
    const int threads = (m >= 2048) ? 512 : ((m >= 512) ? 256 : 128);
    const size_t smem = (size_t)m * 8 + (size_t)c->nstruct * 4;
    CUDA_TRY(cudaFuncSetAttribute(rank_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rank_match_kernel, threads, smem));
    if (per_sm < 1) return fail(IGMK_ECUDA, "rank_match_kernel cannot run with nstruct = %d", c->nstruct);
    const long long cap = (long long)c->sm_count * per_sm;
    rank_match_kernel<<<(unsigned)(n_items < cap ? n_items : cap), threads, smem, (cudaStream_t)stream>>>(P);
    g_launches++;
    CUDA_TRY(cudaGetLastError());
    return IGMK_OK;
}

extern "C" int igmk_rank_match_host(igmk_ctx* c, int64_t n_items, const int32_t* a, const int32_t* b,
                                    int reduce, const float* target, int64_t target_stride,
                                    float* matched, int32_t* rank, float* value) {
    if (!c) return fail(IGMK_EINVAL, "igmk_rank_match_host: NULL context");
    if (n_items == 0) return IGMK_OK;
    if (n_items < 0 || !a || (matched && !target) || target_stride < 0)
        return fail(IGMK_EINVAL, "igmk_rank_match_host: bad argument");
    for (int64_t t = 0; t < n_items; ++t) {
        const int a0 = a[2 * t], a1 = a[2 * t + 1];
        if (a0 < 0 || a0 >= c->nbead || a1 >= c->nbead) return fail(IGMK_EINVAL, "igmk_rank_match_host: bead id out of range in item %lld", (long long)t);
        if (b) {
            const int b0 = b[2 * t], b1 = b[2 * t + 1];
            if (b0 < 0 || b0 >= c->nbead || b1 >= c->nbead) return fail(IGMK_EINVAL, "igmk_rank_match_host: bead id out of range in item %lld", (long long)t);
        }
    }
    CUDA_TRY(cudaSetDevice(c->device));
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t n = (size_t)n_items, N = (size_t)c->nstruct;
    const size_t n_tgt = target ? (target_stride ? n * (size_t)target_stride : N) : 0;
    if (target && target_stride && (size_t)target_stride < N) return fail(IGMK_EINVAL, "igmk_rank_match_host: target_stride < nstruct");
    const size_t o_b = up(n * 8), o_t = o_b + up(n * 8), o_m = o_t + up(n_tgt * 4);
    const size_t o_r = o_m + (matched ? up(n * N * 4) : 0), o_v = o_r + (rank ? up(n * N * 4) : 0);
    const size_t total = o_v + (value ? up(n * N * 4) : 0);
    int rc = ensure(&c->d_pairs, &c->pairs_bytes, total);
    if (rc) return rc;
    char* base = (char*)c->d_pairs;
    CUDA_TRY(cudaMemcpyAsync(base, a, n * 8, cudaMemcpyHostToDevice, c->stream));
    if (b) CUDA_TRY(cudaMemcpyAsync(base + o_b, b, n * 8, cudaMemcpyHostToDevice, c->stream));
    if (target) CUDA_TRY(cudaMemcpyAsync(base + o_t, target, n_tgt * 4, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    rc = igmk_rank_match_device(c, n_items, (const int32_t*)base, b ? (const int32_t*)(base + o_b) : nullptr, reduce,
                                target ? (const float*)(base + o_t) : nullptr, target_stride,
                                matched ? (float*)(base + o_m) : nullptr, rank ? (int32_t*)(base + o_r) : nullptr,
                                value ? (float*)(base + o_v) : nullptr, c->stream);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    if (matched) CUDA_TRY(cudaMemcpyAsync(matched, base + o_m, n * N * 4, cudaMemcpyDeviceToHost, c->stream));
    if (rank) CUDA_TRY(cudaMemcpyAsync(rank, base + o_r, n * N * 4, cudaMemcpyDeviceToHost, c->stream));
    if (value) CUDA_TRY(cudaMemcpyAsync(value, base + o_v, n * N * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaEventElapsedTime(&c->last_kernel_ms, c->ev0, c->ev1));
    return IGMK_OK;
}

extern "C" int igmk_host_alloc(void** ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return fail(IGMK_EINVAL, "igmk_host_alloc: bad argument");
    CUDA_TRY(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault));
    return IGMK_OK;
}

extern "C" int igmk_host_free(void* ptr) {
    if (ptr) CUDA_TRY(cudaFreeHost(ptr));
    return IGMK_OK;
}
