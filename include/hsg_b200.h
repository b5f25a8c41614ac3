/* hsg_b200 -- C ABI of the B200-native (sm_100a) implementation of HSG's
 * clustering + contrastive hot path.
 *
 * This header is the drop-in boundary.  The reference (twke18/HSG) is pure
 * Python/PyTorch and has no FFI of its own; each entry point below replaces a
 * Python operator of the reference (cited per function, paths relative to the
 * reference root) and is what a maintainer binds with ctypes (see
 * INTEGRATION.md; hsg_b200/_lib.py is that binding).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no torch types.
 *  - every pointer is DEVICE memory unless the name ends in _host;
 *    all tensors are contiguous, row-major; floats are fp32, labels int64
 *    (the reference's dtypes).
 *  - the caller owns every buffer, including outputs and the workspace
 *    (size it with the matching hsg_*_workspace_bytes()).  The library keeps
 *    no pointer after a call returns and allocates nothing on the device.
 *  - `stream` is a cudaStream_t passed as void*; work is enqueued on it and
 *    the call returns without synchronising.  The device is the current one.
 *  - return value: HSG_OK (0) or a negative HSG_E_* code; the message is in
 *    hsg_last_error() (thread-local).  No C++ exception crosses the boundary.
 *  - re-entrant: may be called concurrently from one host thread per GPU
 *    (the reference's DataParallel threading model).
 */
#ifndef HSG_B200_H_
#define HSG_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define HSG_API __attribute__((visibility("default")))
#else
#define HSG_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define HSG_OK 0
#define HSG_E_INVALID (-1)     /* bad argument */
#define HSG_E_CUDA (-2)        /* CUDA runtime / launch error */
#define HSG_E_WORKSPACE (-3)   /* workspace too small */
#define HSG_E_UNSUPPORTED (-4) /* shape outside what the kernels cover */
#define HSG_E_COMM (-5)        /* collective error */

/* flags for the k-means entry points */
#define HSG_KMEANS_AUTO 0        /* tensor-core E-step when the shape allows it */
#define HSG_KMEANS_FORCE_SIMT 1  /* fp32 CUDA-core E-step (any shape) */
#define HSG_KMEANS_FORCE_TC 2    /* fail with HSG_E_UNSUPPORTED instead of falling back */
#define HSG_KMEANS_FULL_MSTEP 4  /* OR-able: re-sum every row in every M-step (no incremental update) */

/* fp16 side copy of the pixel rows used by the tensor-core E-step: each row has
 * d16 + HSG_XH_TAIL fp16 columns (the tail carries the split trailing features,
 * see hsg_b200/csrc/prep.cu); at most HSG_XH_MAX_TRAILING trailing features. */
#define HSG_XH_TAIL 16
#define HSG_XH_MAX_TRAILING 5

/* modes of hsg_segment_reduce_f32 */
#define HSG_REDUCE_SUM 0
#define HSG_REDUCE_NORMALIZE 1 /* sum then L2-normalise: calculate_prototypes_from_labels */
#define HSG_REDUCE_MEAN 2      /* sum / max(count,1): segment_mean */

HSG_API const char* hsg_last_error(void);
HSG_API int hsg_version(void);
/* number of SMs of the current device, or a negative error code */
HSG_API int hsg_device_sms(void);

/* instrumentation used by bench.py: number of kernels launched by this library
 * so far (process-wide), and optional CUDA-event timing of the phases of the
 * path on the launching stream (phase ids: 0 prep, 1 M-step sort, 2 M-step
 * gather, 3 M-step combine, 4 E-step, 5 E-step float64 re-decision, 6 relabel,
 * 7 pooling, 8 NCE fwd, 9 NCE bwd, 10 centroid fp16 conversion, 11 one whole
 * hsg_kmeans_* call).  on = 1: every phase; on = 2: coarse -- phases 1-5 and 10
 * are not recorded (two events per k-means call instead of ~200: the events
 * themselves cost ~2 us each on the stream). */
HSG_API long long hsg_launch_count(void);
HSG_API int hsg_profile_enable(int on);
HSG_API int hsg_profile_collect(double* total_ms_host, long long* counts_host, int n_phases);

/* test hook: the next tensor-core E-steps also write their screening
 * similarities to sims [N,kmax] (device); NULL switches the dump off. */
HSG_API int hsg_debug_set_tc_dump(float* sims);
/* profiling hook (tools/estep_timeline.py): with HSG_TC_EXP=2 in the environment the single-pass
 * tensor-core E-step writes, per CTA and role (TMA producer, MMA issuer, epilogue of accumulator 0 / 1),
 * {cycles alive, cycles in its first wait, cycles in its second wait} to clk [grid,4,3] (device). */
HSG_API int hsg_debug_set_tc_clock(long long* clk);
/* test hook, a bit mask that pins a path to one of its two implementations (0 = the library decides):
 *   1 NCE forward on the fp32 CUDA-core kernel      4 NCE backward GEMMs on CUDA cores     8 NCE backward G chunk on CUDA cores
 *  16 attention forward on CUDA cores              32 attention forward on tcgen05 for every shape it supports
 *  64 attention backward on CUDA cores            128 attention backward on tcgen05 for every shape it supports */
HSG_API int hsg_debug_set_flags(int flags);

/* ---- a1: normalize_embedding  (hsg/utils/general/common.py:101-120) -------
 * y[r,:] = x[r,:] / max(||x[r,:]||_2, 1e-12).  x may alias y. */
HSG_API int hsg_normalize_f32(const float* x, float* y, int64_t rows, int dim, void* stream);
/* gx = (gy - y <y,gy>) / ||x||   (gy/1e-12 when ||x|| < 1e-12) */
HSG_API int hsg_normalize_bwd_f32(const float* x, const float* gy, float* gx, int64_t rows, int dim,
                          void* stream);

/* ---- K0 prep: front half of segment_by_kmeans
 *      (hsg/utils/segsort/common.py:305-365: NCHW->NHWC, normalise, concat the
 *      local features, re-normalise, drop ignore pixels; :376-381 batch index)
 * emb_nchw      [B,D,H,W]
 * loc           [.,H,W,L] local features; image b starts at loc + b*loc_image_stride
 *               (stride 0 = one map shared by all images, the reference's expand())
 * labels        [B,H,W] or NULL (-> all zero); pixels with labels == ignore_index
 *               are dropped when use_ignore != 0
 * init_clusters [.,H,W] dense initial cluster ids (already passed through
 *               unique(return_inverse), :341); image stride like loc
 * outputs, each with room for B*H*W rows; valid rows are [0, seg_offsets[B]):
 *   x_out [N,D], xloc_out [N,D+L], labels_out [N], clusters_out [N],
 *   batch_out [N] (= b + batch_index_base), seg_offsets [B+1] (device, int64).
 *   xh_out   (optional, may be NULL) [N,D+HSG_XH_TAIL] fp16 side copy of xloc_out
 *   xerr_out (optional, with xh_out) [N] ||xloc[:, :D] - fp16(xloc[:, :D])||_2
 *   pixel_out (optional) [N] flat source pixel b*H*W + y*W + x of each kept row
 */
HSG_API size_t hsg_prep_workspace_bytes(int B, int H, int W);
HSG_API int hsg_prep_f32(const float* emb_nchw, int B, int D, int H, int W,
                 const float* loc, int L, int64_t loc_image_stride,
                 const int64_t* labels, int use_ignore, int64_t ignore_index,
                 const int64_t* init_clusters, int64_t init_image_stride,
                 int64_t batch_index_base,
                 float* x_out, float* xloc_out, void* xh_out, float* xerr_out,
                 int64_t* labels_out, int64_t* clusters_out, int64_t* batch_out,
                 int64_t* pixel_out, int64_t* seg_offsets, void* workspace, size_t workspace_bytes,
                 void* stream);

/* hsg_prep_f32 that also emits the partial sums of the FIRST k-means M-step while the rows are on chip
 * (segment_by_kmeans starts k-means from the grid labels it has just written, common.py:337-372, so the
 * first calculate_prototypes_from_labels would re-read every xloc row): for each tile of 64 consecutive
 * source pixels of an image, one partial sum per run of equal initial cluster id among the kept pixels,
 * at most HSG_PREP_RUNS runs per tile.
 *   run_sums    [B*R, D+L] fp32, R = hsg_prep_runs_per_image(H,W); slot (b*tiles + t)*HSG_PREP_RUNS + run
 *   run_cluster [B*R] int32 initial cluster id of the run, -1 = unused slot (its sum row is not written)
 *   run_count   [B*R] int32 pixels in the run
 *   run_overflow[1]   int32, 1 when some tile had more runs than slots: the sums are then incomplete and
 *               hsg_kmeans_presummed_f32 falls back to the ordinary first M-step (decided on the device) */
#define HSG_PREP_RUNS 4
HSG_API int64_t hsg_prep_runs_per_image(int H, int W);
HSG_API int hsg_prep_sums_f32(const float* emb_nchw, int B, int D, int H, int W,
                 const float* loc, int L, int64_t loc_image_stride,
                 const int64_t* labels, int use_ignore, int64_t ignore_index,
                 const int64_t* init_clusters, int64_t init_image_stride,
                 int64_t batch_index_base,
                 float* x_out, float* xloc_out, void* xh_out, float* xerr_out,
                 int64_t* labels_out, int64_t* clusters_out, int64_t* batch_out,
                 int64_t* pixel_out, int64_t* seg_offsets, void* workspace, size_t workspace_bytes,
                 float* run_sums, int32_t* run_cluster, int32_t* run_count, int32_t* run_overflow,
                 void* stream);

/* Backward of the prep chain (the reference's autograd through permute / normalize / cat / normalize /
 * index_select, hsg/utils/segsort/common.py:305-365): gradients w.r.t. the two float outputs of hsg_prep_f32
 * (either may be NULL) -> gradient w.r.t. the NCHW input, zero at dropped pixels.  One pass.
 *   x            [N,D]   the forward's x_out
 *   row_of_pixel [B*H*W] output row of every source pixel, -1 = dropped; NULL = nothing was dropped */
HSG_API int hsg_prep_bwd_f32(const float* emb_nchw, int B, int D, int H, int W,
                 const float* x, const float* loc, int L, int64_t loc_image_stride,
                 const int64_t* row_of_pixel, const float* grad_x, const float* grad_xloc,
                 float* grad_emb_nchw, void* stream);

/* fp16 side copy used by the tensor-core E-step, for callers that did not go
 * through hsg_prep_f32: xh [rows, d16+HSG_XH_TAIL] fp16, xerr[r] = rounding norm
 * of the first d16 columns; dim - d16 <= HSG_XH_MAX_TRAILING */
HSG_API int hsg_make_half_copy_f32(const float* x, int64_t rows, int dim, int d16, void* xh_out,
                           float* xerr_out, void* stream);

/* ---- K1: spherical k-means
 *      kmeans_with_initial_labels (hsg/utils/segsort/common.py:67-97), batched
 *      over S independent segments (the per-image loop of segment_by_kmeans,
 *      :337-372) or S == 1 for the flat problem.
 * x            [N,dim] unit rows (embedding with location features)
 * seg_offsets  [S+1] device int64, segment s = rows [seg_offsets[s], seg_offsets[s+1])
 * max_seg_len  host upper bound on any segment length (sizes the grids; no sync)
 * seg_k        [S] device int32 clusters per segment, or NULL (= kmax everywhere)
 * init_labels  [N] int64 in [0, k_s)
 * labels_out   [N] int64
 * centroids_out optional [S,kmax,dim]: the centroids the last E-step used
 * xh, xerr     optional fp16 copy [N,d16+HSG_XH_TAIL] (see prep); enables the
 *              tcgen05 E-step when d16 in {64,128,256}, dim-d16 <= 5, kmax <= 256
 * Each iteration = M-step (deterministic segmented sum + normalise; empty
 * cluster -> zero row) then E-step (arg-max of <x,c>, ties -> lowest index).
 * The E-step result is the arg-max of the float64 dot products of the fp32
 * inputs: cheap fp32/fp16 passes only prune, every pixel whose top-2 gap is
 * inside the pass's rigorous error bound is re-decided in float64.
 */
HSG_API size_t hsg_kmeans_workspace_bytes(int64_t N, int dim, int S, int kmax, int64_t max_seg_len);
HSG_API int hsg_kmeans_f32(const float* x, int64_t N, int dim,
                   const void* xh, int d16, const float* xerr,
                   const int64_t* seg_offsets, int S, int64_t max_seg_len,
                   const int32_t* seg_k, int kmax,
                   const int64_t* init_labels, int iterations,
                   int64_t* labels_out, float* centroids_out, int flags,
                   void* workspace, size_t workspace_bytes, void* stream);

/* hsg_kmeans_f32 whose first M-step takes the run sums of hsg_prep_sums_f32 (the init labels must be the
 * clusters that call wrote, segment s = image s, runs_per_segment = hsg_prep_runs_per_image) instead of
 * re-reading every row; falls back to the ordinary first M-step on the device when *run_overflow != 0. */
HSG_API int hsg_kmeans_presummed_f32(const float* x, int64_t N, int dim,
                   const void* xh, int d16, const float* xerr,
                   const int64_t* seg_offsets, int S, int64_t max_seg_len,
                   const int32_t* seg_k, int kmax,
                   const int64_t* init_labels, int iterations,
                   int64_t* labels_out, float* centroids_out, int flags,
                   void* workspace, size_t workspace_bytes,
                   const float* run_sums, const int32_t* run_cluster, const int32_t* run_count,
                   const int32_t* run_overflow, int64_t runs_per_segment, void* stream);

/* single steps, for per-iteration (teacher-forced) parity:
 * M-step == calculate_prototypes_from_labels per segment (common.py:11-41),
 * E-step == find_nearest_prototypes per segment (common.py:44-64). */
HSG_API int hsg_kmeans_mstep_f32(const float* x, int64_t N, int dim,
                         const int64_t* seg_offsets, int S, int64_t max_seg_len,
                         const int32_t* seg_k, int kmax, const int64_t* labels,
                         float* centroids_out, void* workspace, size_t workspace_bytes,
                         void* stream);
HSG_API int hsg_kmeans_estep_f32(const float* x, int64_t N, int dim,
                         const void* xh, int d16, const float* xerr,
                         const int64_t* seg_offsets, int S, int64_t max_seg_len,
                         const int32_t* seg_k, int kmax, const float* centroids,
                         int64_t* labels_out, int64_t* num_rechecked_out /* [2]: re-decided, full scans */, int flags,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- K3: segmented reduction by label
 *      calculate_prototypes_from_labels (common.py:11-41, mode NORMALIZE),
 *      segment_mean (hsg/utils/general/common.py:123-147, mode MEAN).
 * labels [N] int64 in [0,P).  When seg_offsets != NULL the labels of segment s
 * must lie in [seg_base[s], seg_base[s] + kmax) (true for prototype ids made by
 * segment_by_kmeans: they are ranked by image first); with seg_offsets == NULL
 * the call is one segment with kmax = P.  Deterministic: rows of one label are
 * added in a fixed order that depends only on the shapes.
 * out [P,dim]; counts_out optional [P] float.
 * The backward pass is a row gather: gx[i,:] = gs[labels[i],:] with
 *   NORMALIZE: gs_k = (g_k - p_k <p_k,g_k>)/||s_k||  (g_k/1e-12 below eps)
 *   MEAN:      gs_k = g_k / max(count_k,1)           SUM: gs_k = g_k
 */
HSG_API size_t hsg_segment_reduce_workspace_bytes(int64_t N, int dim, int64_t P, int S, int kmax,
                                          int64_t max_seg_len);
HSG_API int hsg_segment_reduce_f32(const float* x, int64_t N, int dim, const int64_t* labels, int64_t P,
                           const int64_t* seg_offsets, int S, int64_t max_seg_len,
                           const int64_t* seg_base, int kmax, int mode,
                           float* out, float* sums_out, float* counts_out,
                           void* workspace, size_t workspace_bytes, void* stream);
HSG_API int hsg_segment_reduce_bwd_f32(const float* grad_out, const float* out, const float* sums,
                               const float* counts, const int64_t* labels, int64_t N, int dim,
                               int64_t P, int mode, float* grad_x, void* workspace,
                               size_t workspace_bytes, void* stream);

/* Exact, order-independent bin sums: every element is converted to 2^-36 fixed point and summed in int64
 * (rows with |x| <= 1, up to 2^26 rows per bin).  sums_out [P,dim] int64; sum * 2^-36 is the value.  Integer sums do
 * not depend on how the rows are split over launches or GPUs, so a row-sharded k-means that all-reduces these
 * (ncclSum on int64 is exact) yields bit-identical centroids and labels at every GPU count (SURVEY 8e notes). */
HSG_API size_t hsg_segment_sum_exact_workspace_bytes(int64_t N, int dim, int64_t P, int S, int kmax,
                                                     int64_t max_seg_len);
HSG_API int hsg_segment_sum_exact_i64(const float* x, int64_t N, int dim, const int64_t* labels, int64_t P,
                                      const int64_t* seg_offsets, int S, int64_t max_seg_len,
                                      const int64_t* seg_base, int kmax, long long* sums_out,
                                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- K1 over row shards: kmeans_with_initial_labels (common.py:67-97) on rows split across GPUs, as the reference's
 *      flat k-means is used by hsg/models/embeddings/clusters.py:30-42 (BASELINE configs[4]); one iteration =
 *        hsg_kmeans_dist_local_i64   this shard's exact int64 contribution to the [kmax,dim] centroid sums
 *                                    (first != 0: the sums of init_labels; afterwards: only the rows whose label changed,
 *                                    sum[new] += x, sum[old] -= x)
 *        <caller: ncclAllReduce(sum, int64) over the shards; running += contribution>
 *        hsg_kmeans_dist_assign_f32  centroids = normalise(running * 2^-36), E-step (same kernels and float64
 *                                    re-decision as hsg_kmeans_f32); the new labels stay in the workspace
 *      and hsg_kmeans_dist_labels_i64 copies the final labels out.  The workspace carries the loop's state and must be
 *      left untouched between the calls (same N, dim, d16, kmax in all of them).  Integer sums make the centroids -- and
 *      the labels -- bit-identical for every way of sharding the rows. */
HSG_API size_t hsg_kmeans_dist_workspace_bytes(int64_t N, int dim, int kmax);
HSG_API int hsg_kmeans_dist_local_i64(const float* x, int64_t N, int dim, int d16,
                                      const int64_t* seg_offsets /* device {0,N} */, int kmax, int first,
                                      const int64_t* init_labels, long long* sums_out /* [kmax,dim] */,
                                      void* workspace, size_t workspace_bytes, void* stream);
HSG_API int hsg_kmeans_dist_assign_f32(const float* x, int64_t N, int dim, const void* xh, int d16, const float* xerr,
                                       const int64_t* seg_offsets, int kmax, const long long* sums /* [kmax,dim] */,
                                       int flags, void* workspace, size_t workspace_bytes, void* stream);
HSG_API int hsg_kmeans_dist_labels_i64(int64_t N, int dim, int d16, int kmax, int64_t* labels_out,
                                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- K4: pixel-to-prototype NCE ("SegSort+") loss
 *      _calculate_log_likelihood (hsg/utils/segsort/loss.py:15-82).
 * e [N,dim], prototypes [P,dim], inst [N] (own prototype id), n_sets label
 * sets evaluated in ONE pass over e x prototypes (Hsg.losses calls the loss 3x
 * on the same (e, prototypes), hsg/models/predictions/hsg.py:105,130,149):
 *   sem  [n_sets,N]  psem [n_sets,P]   group_plus[n_sets] (1 = 'segsort+')
 * The forward runs on tensor cores (fp16 hi/lo split, fp32-grade products) when
 * dim is 64, 128 or 256 and the workspace of hsg_nce_workspace_bytes() is given;
 * otherwise on the exact fp32 CUDA-core kernel.
 * per_pixel_out [n_sets,N] = -log(num/den);  stats_out [n_sets,N,4] =
 * (num, den, own, flags) saved for the backward pass (may be NULL).
 * Backward for L = sum_s sum_i w[s,i] * l[s,i]:
 *   grad_e [N,dim], grad_p [P,dim]  (SURVEY.md A.1 closed form).
 */
HSG_API size_t hsg_nce_workspace_bytes(int64_t N, int64_t P, int dim, int n_sets);
HSG_API int hsg_nce_fwd_f32(const float* e, const float* prototypes, int64_t N, int64_t P, int dim,
                    const int64_t* inst, const int64_t* sem, const int64_t* psem, int n_sets,
                    const int32_t* group_plus_host, float concentration,
                    float* per_pixel_out, float* stats_out,
                    void* workspace, size_t workspace_bytes, void* stream);
HSG_API int hsg_nce_bwd_f32(const float* e, const float* prototypes, int64_t N, int64_t P, int dim,
                    const int64_t* inst, const int64_t* sem, const int64_t* psem, int n_sets,
                    const int32_t* group_plus_host, float concentration,
                    const float* stats, const float* w, float* grad_e, float* grad_p,
                    void* workspace, size_t workspace_bytes, void* stream);
/* The same two calls with the number of valid prototypes on the DEVICE: P is the capacity of the prototype arrays,
 * *num_prototypes_dev (<= P; NULL = P) the rows that count.  Rows and labels beyond it contribute to no sum and
 * receive a zero gradient, so a step whose prototype count is only known on the device (empty clusters dropped by
 * hsg_relabel_i64, prototypes all-gathered by hsg_exchange_*) needs no host read between k-means and the loss.
 */
HSG_API int hsg_nce_fwd_counted_f32(const float* e, const float* prototypes, int64_t N, int64_t P,
                    const int64_t* num_prototypes_dev, int dim,
                    const int64_t* inst, const int64_t* sem, const int64_t* psem, int n_sets,
                    const int32_t* group_plus_host, float concentration,
                    float* per_pixel_out, float* stats_out,
                    void* workspace, size_t workspace_bytes, void* stream);
HSG_API int hsg_nce_bwd_counted_f32(const float* e, const float* prototypes, int64_t N, int64_t P,
                    const int64_t* num_prototypes_dev, int dim,
                    const int64_t* inst, const int64_t* sem, const int64_t* psem, int n_sets,
                    const int32_t* group_plus_host, float concentration,
                    const float* stats, const float* w, float* grad_e, float* grad_p,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---- a13: prototype exchange around ONE all-gather, no host read
 *      (replaces the gather + cat + unique of hsg/models/utils.py:127-217 for one-process-per-GPU launches;
 *       the library holds no communicator: the caller all-gathers the records, e.g. ncclAllGather /
 *       torch.distributed.all_gather_into_tensor).
 * pack:   record [hsg_exchange_record_bytes(capacity, dim, dim_loc)] = this rank's count (read from the device,
 *         clamped to capacity), prototypes [capacity,dim], prototypes_with_loc [capacity,dim_loc] and the three
 *         label vectors [capacity]; rows beyond the count are written as zeros / -1.
 * unpack: gathered = the `world` records in rank order.  Outputs have world*capacity rows: the valid rows of rank 0,
 *         then rank 1, ... (the reference's global prototype order), the rest zeros / -1; *total_out = number of
 *         valid rows, *offset_out = valid rows of the ranks below `rank` (add it to this rank's pixel -> prototype ids).
 */
HSG_API size_t hsg_exchange_record_bytes(int64_t capacity, int dim, int dim_loc);
HSG_API int hsg_exchange_pack(const float* prototypes, const float* prototypes_with_loc, const int64_t* sem,
                    const int64_t* inst, const int64_t* batch, const int64_t* num_prototypes_dev,
                    int64_t capacity, int dim, int dim_loc, void* record, void* stream);
HSG_API int hsg_exchange_unpack(const void* gathered, int world, int rank, int64_t capacity, int dim, int dim_loc,
                    float* prototypes_out, float* prototypes_with_loc_out, int64_t* sem_out, int64_t* inst_out,
                    int64_t* batch_out, int64_t* total_out, int64_t* offset_out, void* stream);

/* ---- K5: fused attention core of the clustering transformer
 *      (nn.MultiheadAttention slow path used by hsg/models/heads/transformer.py:235,300,304:
 *       q/sqrt(hd), baddbmm with the -inf key-padding mask, softmax, dropout, bmm with v)
 * q [B*heads, L, hd], k/v [B*heads, S, hd] (head index fastest, as torch lays them out),
 * key_padding_mask [B,S] bytes (1 = ignore) or NULL; scale = 1/sqrt(hd).
 * out [B*heads, L, hd]; lse [B*heads, L] saved for the backward pass.  dropout_p > 0
 * drops probabilities with a counter-based generator keyed by `seed` (train-mode
 * parity with the reference is defined at p = 0).  A row whose keys are all
 * masked yields NaN, as in the reference. */
/* backward workspace: large enough for either backward -- tcgen05 (head dim 64, L, S <= 256, where the shape
 * pays for the operand preparation: dq and dk/dv kernels, scores recomputed in TMEM) or CUDA cores. */
HSG_API size_t hsg_mha_workspace_bytes(int B, int heads, int L, int S);
/* forward workspace: with it (head dim 64, S <= 256) the two contractions run on tcgen05 (fp16 hi/lo split,
 * fp32-grade); without it, or for other shapes, on CUDA cores.  0 bytes = not needed. */
HSG_API size_t hsg_mha_fwd_workspace_bytes(int B, int heads, int L, int S, int hd);
HSG_API int hsg_mha_fwd_f32(const float* q, const float* k, const float* v,
                            const unsigned char* key_padding_mask, int B, int heads, int L, int S,
                            int hd, float scale, float dropout_p, unsigned long long seed,
                            float* out, float* lse, void* workspace, size_t workspace_bytes, void* stream);
HSG_API int hsg_mha_bwd_f32(const float* q, const float* k, const float* v,
                            const unsigned char* key_padding_mask, int B, int heads, int L, int S,
                            int hd, float scale, float dropout_p, unsigned long long seed,
                            const float* out, const float* lse, const float* dout,
                            float* dq, float* dk, float* dv,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ---- k-NN affinity graph of the DMoN regulariser
 * Replaces the masking / per-segment top-k / binarise part of affinity_matrix_as_attention
 * (hsg/utils/graph/common.py:76-125) on a given kernel matrix A [B,n,n]: padded nodes
 * (padding_mask [B,n] != 0, may be NULL) and, when remove_self_loop and the graph has more
 * than one valid node, the diagonal are zeroed; with knn > 0 entry (i,j) survives iff fewer
 * than min(#valid nodes of j's segment, knn) valid columns of j's segment (segment_labels
 * [B,n], NULL = one segment) are strictly larger in row i; binarize maps positives to 1.
 * out [B,n,n].  No host synchronisation (the reference takes two per segment). */
HSG_API int hsg_knn_adjacency_f32(const float* A, const unsigned char* padding_mask,
                                  const int64_t* segment_labels, int B, int n, int knn,
                                  int remove_self_loop, int binarize, float* out, void* stream);

/* ---- K2: dense relabel  (segment_by_kmeans tail, common.py:397-405;
 *      prepare_prototype_labels :192-218)
 * ids_out[i] = rank of the triple (batch[i], cluster[i], label[i]) among the
 * distinct triples present, in lexicographic order.  label_values must be a
 * sorted list of the distinct label values (n_label_values of them, device).
 * proto_label_out/proto_batch_out/proto_cluster_out [>= n_protos] describe each
 * id; n_protos_out is a device int64.  batch values are in [batch_base,
 * batch_base + B), cluster in [0,kmax).
 */
HSG_API size_t hsg_relabel_workspace_bytes(int B, int kmax, int64_t n_label_values);
HSG_API int hsg_relabel_i64(const int64_t* batch, const int64_t* cluster, const int64_t* label, int64_t N,
                    int64_t batch_base, int B, int kmax,
                    const int64_t* label_values, int64_t n_label_values,
                    int64_t* ids_out, int64_t* proto_label_out, int64_t* proto_batch_out,
                    int64_t* proto_cluster_out, int64_t* n_protos_out,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---- f3: top_k_ranking  (hsg/utils/segsort/eval.py:9-52)
 * indices_out[r, 0..k) = the k prototypes with the largest inner product with embeddings[r,:], best first
 * (ties: lower index first), k <= 8.  fp32 products; the [N,M] affinity matrix is never written.
 * values_out optional [N,k].  The caller gathers the labels / averages the hits (tiny tensors). */
HSG_API int hsg_topk_affinity_f32(const float* embeddings, int64_t N, const float* prototypes, int64_t M,
                                  int dim, int k, int64_t* indices_out, float* values_out, void* stream);

/* ---- f2: SynchronizedBatchNorm for one process per GPU  (lib/nn/sync_batchnorm/batchnorm.py:55-118)
 * x viewed as [B,C,L].  Forward: stats_out[0..C) = sum x, [C..2C) = sum x^2 (grad NULL); the caller all-reduces
 * them over the ranks, forms mean / inv_std and calls hsg_bn_apply_f32 (grad NULL): y = (x-mean)*inv_std*w + b.
 * Backward: hsg_bn_stats_f32 with grad -> sum g and sum g*xhat; all-reduce; hsg_bn_apply_f32 with grad ->
 * gx = (g - mean_g - xhat*mean_gx) * w * inv_std, mean_* = grad_stats * inv_count. */
HSG_API int hsg_bn_stats_f32(const float* x, const float* grad_or_null, const float* mean, const float* inv_std,
                             int B, int C, int L, float* stats_out, void* stream);
HSG_API int hsg_bn_apply_f32(const float* x, const float* grad_or_null, const float* mean, const float* inv_std,
                             const float* weight, const float* bias, const float* grad_stats, float inv_count,
                             int B, int C, int L, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HSG_B200_H_ */
