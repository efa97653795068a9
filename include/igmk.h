/*
 * igmk.h - C ABI of libigmk.so: the B200 (sm_100a) kernels behind IGM's Hi-C
 * A-step and population contact-frequency map.
 *
 * The reference (bonimba87/igm) is pure Python on this path and has no FFI of
 * its own (its only native code is the SPRITE helper,
 * igm/cython_compiled/sprite.pyx:21-31).  The entry points below are what a
 * ctypes binding inside the reference's Step would call instead of the NumPy
 * loops they replace; each one cites the reference lines it stands in for
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes
 * stub.
 *
 * Conventions: plain pointers and sizes only; the caller allocates every
 * buffer; every function returns 0 on success and a negative IGMK_E* code on
 * failure (igmk_last_error() gives the message); no exception crosses the
 * ABI; one context per GPU; a context is not thread-safe, distinct contexts
 * are independent.  There is no CPU fallback: without a CUDA device every
 * call fails with IGMK_ECUDA.
 */
#ifndef IGMK_H_
#define IGMK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IGMK_VERSION 100

#define IGMK_OK        0
#define IGMK_EINVAL   -1   /* bad argument */
#define IGMK_ECUDA    -2   /* CUDA runtime error (no device, launch failure, OOM) */
#define IGMK_ESTATE   -3   /* call order: coordinates / index not set */
#define IGMK_ELIMIT   -4   /* outside supported range (ploidy > 2, nstruct > 25600) */

/* A-step flavours.  LB = igm/steps/ActivationDistanceStep.py:336-485 (what
 * igm-run executes); GP = igm/steps/GP_activation.py:317-445 and
 * igm/utils/actdist.py:15-96 (keep the min(len(ii),len(jj)) smallest copy
 * combinations per structure). */
#define IGMK_MODE_LB 0
#define IGMK_MODE_GP 1

/* Kernel selection for igmk_actdist_*: 0 = production kernel (register-resident
 * packed-key select), 1 = straightforward shared-memory bitwise select kept as
 * an on-device cross-check. */
#define IGMK_ALGO_FAST   0
#define IGMK_ALGO_SIMPLE 1

/* Per-pair result, 32 bytes.  One per candidate pair, in input order. */
typedef struct igmk_pair_result {
    uint32_t d2_sel_bits;   /* float32 bit pattern of sortdist_sq[o]           (:448,:473) */
    int32_t  contact_count; /* #{d_sq[0:npc] <= rcutsq}                         (:442)     */
    int32_t  o;             /* order-statistic index, -1 when p <= 0 or i == j  (:469-470) */
    int32_t  nrec;          /* records this pair expands to: 0,1,2 or 4         (:476-483) */
    double   p;             /* corrected probability (float64)                  (:452-462) */
    float    dist;          /* float32("%10.4f" % sqrt_f64(d2_sel))             (:38,:230,:249) */
    float    prob;          /* float32("%.4f" % p)                              (:38,:230,:249) */
} igmk_pair_result;

typedef struct igmk_ctx igmk_ctx;

int         igmk_version(void);
const char* igmk_last_error(void);
/* Number of kernel launches issued by this library in the calling process
 * (bench.py's gpu_launches claim). */
int64_t     igmk_launch_count(void);

/* Context for a population of nbead beads x nstruct structures on `device`. */
int igmk_create(int device, int nbead, int nstruct, igmk_ctx** out);
int igmk_destroy(igmk_ctx* ctx);

/* Stage coordinates once per A-step.  `xyz` is the .hss layout
 * (nbead, nstruct, 3) float32, bead-major (igm/core/step.py:373,
 * igm/_preprocess.py:102-105); replaces the per-pair hss.get_bead_crd(k) reads
 * (ActivationDistanceStep.py:415-416,431-432).  Stored in HBM as
 * [bead][segment of 128 structures][xyz][128] so one bead's row over all structures is
 * contiguous.  `on_device` != 0: xyz is a device pointer. */
int igmk_upload_coords(igmk_ctx* ctx, const float* xyz, int on_device);
/* Partial upload of beads [bead0, bead0 + nb): lets a loader stream HDF5
 * chunks (pack_beads x nstruct x 3, igm/steps/ModelingStep.py:753-760). */
int igmk_upload_coords_range(igmk_ctx* ctx, const float* xyz, int bead0, int nb, int on_device);
/* Replication over NVLink instead of one PCIe upload per GPU ("coordinates replicated",
 * north_star): every GPU uploads only its share of the beads, then the shares are exchanged.
 * igmk_coords_device exposes the staged rows (3 * npad floats each; nbead bead rows, one
 * all-zero row, 64 spare rows so an in-place all-gather of equal shares fits) to the
 * caller's collective (one process per GPU); igmk_copy_coords_peer copies rows between two
 * contexts of one process (one task driving all GPUs, igm/core/step.py:259-274). */
int igmk_coords_device(igmk_ctx* ctx, float** d_coords, int64_t* n_rows, int64_t* row_floats);
int igmk_copy_coords_peer(igmk_ctx* dst, igmk_ctx* src, int bead0, int nb);

/* Index tables (host pointers): copy_ptr[n_hap+1] / copy_beads = CSR of
 * hss.index.copy_index; chrom_hap[n_hap] = hss.index.chrom indexed by haploid
 * bin; radii[nbead]  (ActivationDistanceStep.py:383-393). */
int igmk_set_index(igmk_ctx* ctx, int n_hap, const int32_t* copy_ptr,
                   const int32_t* copy_beads, const int32_t* chrom_hap,
                   const float* radii);

/* get_actdist for n_pairs candidate pairs (ActivationDistanceStep.py:336-485;
 * loop at :215-219).  *_device: every pointer is device memory, asynchronous
 * on `stream` (a cudaStream_t, NULL = default stream).  *_host: host pointers;
 * copies in, runs, copies out and synchronises before returning. */
int igmk_actdist_device(igmk_ctx* ctx, int64_t n_pairs,
                        const int32_t* d_i, const int32_t* d_j,
                        const double* d_pwish, const double* d_plast,
                        float contact_range, int it_corr, int mode, int algo,
                        igmk_pair_result* d_out, void* stream);
int igmk_actdist_host(igmk_ctx* ctx, int64_t n_pairs,
                      const int32_t* i, const int32_t* j,
                      const double* pwish, const double* plast,
                      float contact_range, int it_corr, int mode, int algo,
                      igmk_pair_result* out);

/* Optional: size the device buffers of the *_host entry points for lists of up to n_pairs
 * pairs now (the sigma sweep's lists grow from step to step, :69-98; growing the buffers costs a
 * device-wide synchronisation and tens of milliseconds each time). */
int igmk_reserve_pairs(igmk_ctx* ctx, int64_t n_pairs);

/* The whole device side of ActivationDistanceStep.task in one call: stage the population
 * `xyz` (host memory, the .hss layout of igmk_upload_coords; every A-step follows an M-step
 * that rewrote it, igm/steps/ModelingStep.py:730-782, read back through
 * igm/core/step.py:386-392) AND run get_actdist over the pair list, with the upload hidden
 * behind the pair kernels: slices of the list are taken from its end, and a slice is launched
 * as soon as the beads of its loci and of all higher ones are in HBM (setup() writes the
 * list sorted by i with j > i, ActivationDistanceStep.py:166-178, so the last slice needs
 * only the top of the population).  Any list order and any copy index give the same results
 * as igmk_upload_coords + igmk_actdist_host; they only overlap less.  The index must be set;
 * afterwards the context holds the whole population.  Page-locked buffers make the copies
 * asynchronous. */
int igmk_actdist_host_population(igmk_ctx* ctx, const float* xyz, int64_t n_pairs,
                                 const int32_t* i, const int32_t* j,
                                 const double* pwish, const double* plast,
                                 float contact_range, int it_corr, int mode, int algo,
                                 igmk_pair_result* out);

/* sel_flat_idx (the "selected structure index" of the A-step): for every pair the index, in
 * d_sq[0:npc].ravel() = row * nstruct + structure (:439-448; d_sq is column-sorted there, so
 * the row is the value's rank among the copy combinations of its structure), of the element
 * that IS the selected
 * order statistic - the lowest index when several elements tie (np.sort is not stable, so
 * the reference defines no particular one); -1 for pairs without a record.  `results`: the
 * records igmk_actdist_* returned for the same pairs and mode.  A separate pass because the
 * A-step itself never needs the index. */
int igmk_actdist_sel_index_device(igmk_ctx* ctx, int64_t n_pairs, const int32_t* d_i, const int32_t* d_j,
                                  const igmk_pair_result* d_results, int mode, int32_t* d_sel_idx,
                                  void* stream);
int igmk_actdist_sel_index_host(igmk_ctx* ctx, int64_t n_pairs, const int32_t* i, const int32_t* j,
                                const igmk_pair_result* results, int mode, int32_t* sel_idx);

/* Diagnostic: how many pairs of the most recent igmk_actdist_* launch on this context the
 * list-form kernel handed back to the key-array kernels (large order index, many contacts,
 * or a population sample that misjudged the pair).  Synchronises the device. */
int64_t igmk_last_redo_count(igmk_ctx* ctx);

/* Multi-GPU form (one process per GPU, pairs sharded, coordinates replicated;
 * the reference farms 1000-pair batches to CPU workers and meets again on the
 * shared filesystem, igm/steps/ActivationDistanceStep.py:181-194,
 * igm/core/step.py:274): the kernel stores every raw pair result straight into
 * the gather buffers of all n_peers GPUs over NVLink (peer stores from inside the
 * kernel, so the all-gather overlaps the compute and no collective follows).
 * d_peer_slices: device array of n_peers addresses - where THIS rank's slice of
 * igmk_pair_result records starts in GPU p's gather buffer (peer-mapped memory,
 * e.g. torch symmetric memory).  The records arrive finished (dist / prob filled in by
 * the kernel); what remains after the launch is one barrier across the ranks. */
int igmk_actdist_device_peers(igmk_ctx* ctx, int64_t n_pairs,
                              const int32_t* d_i, const int32_t* d_j,
                              const double* d_pwish, const double* d_plast,
                              float contact_range, int it_corr, int mode,
                              const uint64_t* d_peer_slices, int n_peers, void* stream);
/* dist / prob (float64 sqrt + the reference's 4-decimal text round trip,
 * ActivationDistanceStep.py:38,230,249,473) for n raw results in device memory. */
int igmk_finish_results_device(igmk_ctx* ctx, igmk_pair_result* d_results, int64_t n, void* stream);

/* Lamina-DamID activation distance, spherical envelope
 * (get_damid_actdist_I, igm/steps/DamidActivationDistanceStep.py:375-470; loop at
 * :246-253).  One entry per locus: loci[] = haploid index I, pexp[] / plast[] = the
 * float32 columns of the '%d.damid.in.npy' batch (:178-186).  Result fields:
 * d2_sel_bits = float32 sum of squares behind d_sq[o]; contact_count =
 * #{d_sq >= 1}; o = index in DESCENDING order (-1 when p <= 0, the reference then
 * stores distance 2); nrec = number of copies (records are written even for
 * p <= 0, :468); p = corrected probability (a float32 value);
 * dist / prob = float32("%.5f" % value) as reduce() stores them (:35,:289).
 * Ellipsoidal envelopes never reach this function in the reference (:245 compares
 * two string literals) and are not supported. */
int igmk_damid_actdist_device(igmk_ctx* ctx, int64_t n_loci, const int32_t* d_loci,
                              const float* d_pexp, const float* d_plast,
                              double nucleus_radius, double contact_range, int it_corr,
                              igmk_pair_result* d_out, void* stream);
int igmk_damid_actdist_host(igmk_ctx* ctx, int64_t n_loci, const int32_t* loci,
                            const float* pexp, const float* plast,
                            double nucleus_radius, double contact_range, int it_corr,
                            igmk_pair_result* out);

/* Record expansion of task()/reduce() (ActivationDistanceStep.py:221-222,
 * 476-483, 249-257): pair results -> the four actdist.hdf5 columns, reference
 * order.  Host pointers; row/col/dist/prob must hold sum(nrec) entries;
 * *n_records receives that sum. */
int igmk_expand_records(igmk_ctx* ctx, int64_t n_pairs,
                        const int32_t* i, const int32_t* j,
                        const igmk_pair_result* res,
                        int32_t* row, int32_t* col, float* dist, float* prob,
                        int64_t capacity, int64_t* n_records);

/* Host phases of setup() (pure host code, no device needed).
 * igmk_filter_candidates: the candidate filter of the reference's setup loop
 * (igm/steps/ActivationDistanceStep.py:166-178) over the strict-upper-triangle CSR matrix of
 * the .hcs file: entry (r, c, p) is kept if r != c and p >= intra_sigma (chrom[r] == chrom[c])
 * or p >= inter_sigma (otherwise), compared in float32; use_intra / use_inter = 0 disables a
 * class (the reference's `False` sigma).  Writes (i, j, pwish as float64) in CSR order and
 * returns their number, or -1 - needed when capacity is too small.
 * igmk_join_plast: plast[i, j] of :144-160,177 - the stored `prob` of the previous
 * iteration's record whose (row, col) are the haploid indices (records with row >= n or
 * col >= n are skipped, quirk q7) - as a merge join for inputs in strictly increasing
 * (row, col) order.  Returns 1 when done, 0 when an input is not in that order (nothing
 * useful written; the caller takes its general path), -1 on bad arguments. */
int64_t igmk_filter_candidates(int64_t n_rows, const int64_t* indptr, const int32_t* indices,
                               const float* data, const int32_t* chrom,
                               int use_intra, float intra_sigma, int use_inter, float inter_sigma,
                               int32_t* out_i, int32_t* out_j, double* out_p, int64_t capacity);
int igmk_join_plast(int64_t n_rec, const int32_t* row, const int32_t* col, const float* prob,
                    int32_t n, int64_t n_pairs, const int32_t* ii, const int32_t* jj, double* out);

/* Population contact-frequency counts for a tile of bead pairs
 * (HssFile.buildContactMap as used by igm/steps/HicEvaluationStep.py:107-112
 * and igm/report/hic.py:51; in-tree statement of the formula:
 * HicEvaluationStep.py:73-93).  counts[(a - row0) * ncols + (b - col0)] =
 * #{s : d2_s(a,b) <= (cr*(r_a+r_b))^2}  (strict '<' when strict != 0) for
 * a in [row0,row0+nrows), b in [col0,col0+ncols).  d_counts is device memory. */
int igmk_contact_counts_device(igmk_ctx* ctx, int row0, int nrows, int col0, int ncols,
                               float contact_range, int strict,
                               uint32_t* d_counts, void* stream);
int igmk_contact_counts_host(igmk_ctx* ctx, int row0, int nrows, int col0, int ncols,
                             float contact_range, int strict, uint32_t* counts);

/* The same counts summed over the copies of each haploid locus - what
 * Contactmatrix.sumCopies() makes of the bead-level map
 * (igm/steps/HicEvaluationStep.py:109-111; igm/report/hic.py:51 reads the
 * haploid matrix of get_simulated_hic):
 *   counts[(i - row0) * ncols + (j - col0)] =
 *       sum over a in copies(i), b in copies(j) of #{s : d2_s(a,b) <= (cr*(r_a+r_b))^2}
 * Row / column indices are haploid loci in [0, n_hap).  alabtools is not in the
 * reference tree: the normalisation of sumCopies is UNPINNED. */
int igmk_contact_counts_haploid_device(igmk_ctx* ctx, int row0, int nrows, int col0, int ncols,
                                       float contact_range, int strict,
                                       uint32_t* d_counts, void* stream);
int igmk_contact_counts_haploid_host(igmk_ctx* ctx, int row0, int nrows, int col0, int ncols,
                                     float contact_range, int strict, uint32_t* counts);

/* Hi-C restraint selection of the M-step (intraHiC / interHiC._apply,
 * igm/restraints/intra_hic.py:39-58, inter_hic.py:39-58; one call per structure in
 * igm/steps/ModelingStep.py:376-398): for every actdist record k = (row, col, dist)
 * and every structure s,  bit s of record k =
 *     np.linalg.norm(x[row,s] - x[col,s]) <= dist      (igm/model/particle.py:35-36)
 *     and chrom[row] == chrom[col] (kind 0) / != (kind 1) / no test (kind 2).
 * bitmap: n_rec rows of igmk_restraint_words() uint32 words, structure s at bit
 * s % 32 of word s / 32; counts[k] = number of structures that get restraint k.
 * igmk_set_bead_chrom: hss index chrom, one id per bead (host pointer). */
int igmk_set_bead_chrom(igmk_ctx* ctx, const int32_t* chrom_bead);
int igmk_restraint_words(igmk_ctx* ctx);
int igmk_restraint_select_device(igmk_ctx* ctx, int64_t n_rec, const int32_t* d_row,
                                 const int32_t* d_col, const float* d_dist, int kind,
                                 uint32_t* d_bitmap, int32_t* d_counts, void* stream);
int igmk_restraint_select_host(igmk_ctx* ctx, int64_t n_rec, const int32_t* row,
                               const int32_t* col, const float* dist, int kind,
                               uint32_t* bitmap, int32_t* counts);

/* SPRITE cluster radius of gyration with exhaustive choice of copies - the GPU form
 * of the reference's one native function, get_rg2s_cpp
 * (igm/cython_compiled/cpp_sprite_assignment.cpp:79-143, bound by Cython at
 * igm/cython_compiled/sprite.pyx:21-31,98-99), batched over clusters and reading
 * the population resident in HBM instead of a per-cluster gathered array.
 * Cluster k has regions region_ptr[k] .. region_ptr[k+1]-1; region i has the
 * alternative locations beads[copy_ptr[i] .. copy_ptr[i+1]-1] (bead ids).  Out:
 * rg2s[k * nstruct + s]; copy_idx[region_ptr[k] * nstruct + s * M_k + i] (M_k = regions
 * of cluster k; the reference's copy_idxs[n_regions*s + i]); min_struct[k] (first
 * structure with the strictly smallest Rg^2, -1 if none is below 1e8).  Host pointers. */
int igmk_sprite_rg2_host(igmk_ctx* ctx, int n_clusters, const int32_t* region_ptr,
                         const int32_t* copy_ptr, const int32_t* beads,
                         float* rg2s, int32_t* copy_idx, int32_t* min_struct);

/* Rg^2 of whole SPRITE clusters under a per-structure choice of chromosome copies - the
 * second half of compute_gyration_radius (igm/cython_compiled/sprite.pyx:238-283: the copies
 * selected on one representative segment per chromosome are applied to every segment of the
 * cluster and get_rgs2 is evaluated on the selected beads) and its single-chromosome branch
 * (:200-215, one constant selection per copy).  Cluster k has the segments
 * seg_ptr[k] .. seg_ptr[k+1]-1 in the order the reference concatenates them (:243); segment
 * i has the copies beads[loc_ptr[i] .. loc_ptr[i+1]-1] and follows selection column
 * seg_group[i] (0 .. G_k-1, G_k = group_ptr[k+1] - group_ptr[k]);
 * sel[group_ptr[k] * nstruct + s * G_k + g] is the copy chosen in structure s (the layout of
 * igmk_sprite_rg2_host's copy_idx; negative values index from the end as NumPy does).
 * Out: rg2s[k * nstruct + s] = min(Rg^2, 1e8).  No limit on the segments per cluster.
 * Host pointers. */
int igmk_sprite_cluster_rg2_host(igmk_ctx* ctx, int n_clusters, const int32_t* seg_ptr,
                                 const int32_t* loc_ptr, const int32_t* beads,
                                 const int32_t* seg_group, const int32_t* group_ptr,
                                 const int32_t* sel, float* rg2s);

/* Population rank matching - the arithmetic of the FISH and polymer assignment steps
 * (igm/steps/FishAssignmentStep.py:23-79 get_pair_dists / get_rad_dists /
 * get_min_max_and_idx and task :189-193, :214-219; igm/steps/PolymerAssignmentStep.py:24-33
 * get_polymer_dists and task :118-125).  Item t has the bead ids a[2t], a[2t+1] (second
 * -1 when the locus has one copy) and, when b != NULL, b[2t], b[2t+1]; b == NULL means the
 * radial distance of a (np.linalg.norm of the coordinates).  Per structure s:
 *   value[t][s]   = min (reduce 0) or max (reduce 1) over the copy combinations of
 *                   np.linalg.norm(x_a - x_b) as float32
 *   rank[t][s]    = np.argsort(np.argsort(value[t]))[s] with ties broken by structure index
 *   matched[t][s] = target[t * target_stride + rank[t][s]]   (target_stride 0: one shared
 *                   sorted distribution; target / matched may both be NULL)
 * nstruct <= 16384.  As intended by the reference's docstring ALL len(ii)*len(jj)
 * combinations of a pair enter the min / max; the reference's own get_pair_dists never
 * advances its row counter (:33-40), so only single-copy pairs are defined there. */
int igmk_rank_match_device(igmk_ctx* ctx, int64_t n_items, const int32_t* d_a, const int32_t* d_b,
                           int reduce, const float* d_target, int64_t target_stride,
                           float* d_matched, int32_t* d_rank, float* d_value, void* stream);
int igmk_rank_match_host(igmk_ctx* ctx, int64_t n_items, const int32_t* a, const int32_t* b,
                         int reduce, const float* target, int64_t target_stride,
                         float* matched, int32_t* rank, float* value);

/* Pinned host memory for zero-staging transfers (optional). */
int igmk_host_alloc(void** ptr, int64_t bytes);
int igmk_host_free(void* ptr);

/* Device-time of the last igmk_actdist_host / igmk_contact_counts_host call:
 * kernel only, in milliseconds (CUDA events on the library's stream). */
float igmk_last_kernel_ms(igmk_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* IGMK_H_ */
