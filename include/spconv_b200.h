/*
 * spconv_b200.h -- C ABI of the B200-native sparse-convolution hot path.
 *
 * This is the drop-in boundary: plain pointers, sizes and a cudaStream_t; no torch /
 * tv::Tensor types.  Every entry point replaces one pybind entry of the reference's
 * generated `core_cc` module (paths relative to the reference tree):
 *
 *   spx_subm_rulebook / spx_conv_rulebook_stage1+2 / spx_native_pairs
 *        <- SpconvOps.get_indice_pairs              spconv/csrc/sparse/all.py:2020-2218
 *        <- SpconvOps.get_indice_pairs_implicit_gemm spconv/csrc/sparse/all.py:1660-2016
 *           (kernels spconv/csrc/sparse/indices.py:292-939)
 *   spx_mask_argsort
 *        <- SpconvOps.sort_1d_by_key_allocator[_v2]  spconv/csrc/sparse/all.py:935-1134
 *   spx_rulebook_workspace_size
 *        <- SpconvOps.get_indice_gen_workspace_size  spconv/csrc/sparse/all.py:1582-1656
 *   spx_implicit_gemm_fwd
 *        <- ConvGemmOps.implicit_gemm                spconv/csrc/sparse/convops.py:2075-2243
 *   spx_implicit_gemm_dgrad / spx_implicit_gemm_wgrad
 *        <- ConvGemmOps.implicit_gemm_backward       spconv/csrc/sparse/convops.py:2247-2440
 *   spx_pairs_to_table (+ the three above)
 *        <- ConvGemmOps.indice_conv / indice_conv_backward
 *                                                    spconv/csrc/sparse/convops.py:1504-2071
 *   spx_bias_act_inplace
 *        <- InferenceOps.bias_add_act_inplace        spconv/csrc/sparse/inference.py:166-252
 *   spx_point2voxel_stage1 / _stage2
 *        <- SpconvOps.point2voxel_cuda               spconv/csrc/sparse/all.py:1349-1490
 *   spx_indice_pool_fwd / spx_indice_pool_bwd / spx_global_pool_rearrange
 *        <- SpconvOps.maxpool_forward / maxpool_backward / maxpool_implicit_gemm_forward /
 *           maxpool_implicit_gemm_backward / avgpool_implicit_gemm_forward / _backward /
 *           global_pool_rearrange                    spconv/csrc/sparse/all.py:664-905
 *           (kernels spconv/csrc/sparse/maxpool.py:41-341)
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in `_host`;
 *   - every function returns 0 on success, non-zero on failure; spx_last_error() gives the
 *     message of the last failure on the calling thread (reference: C++ exception text);
 *   - nothing allocates: outputs and scratch are caller-provided (the reference routes all
 *     allocations through ExternalAllocator callbacks, spconv/csrc/sparse/alloc.py:38-123);
 *   - all launches go to `stream`; no function synchronises the stream except
 *     spx_conv_rulebook_stage1, which must return the data-dependent output count
 *     (the reference syncs at the same point, spconv/csrc/sparse/indices.py:1454-1455).
 */
#ifndef SPCONV_B200_H_
#define SPCONV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPX_MAX_NDIM 4

/* element types of features / filters */
enum spx_dtype { SPX_F32 = 0, SPX_F16 = 1, SPX_BF16 = 2, SPX_I8 = 3 };
/* tv::gemm::Activation as used by the reference epilogues */
enum spx_act { SPX_ACT_NONE = 0, SPX_ACT_RELU = 1, SPX_ACT_SIGMOID = 2, SPX_ACT_LEAKY_RELU = 3 };
/* how fp32 features are multiplied: exact fp32 FMA, or TF32 tensor cores
 * (reference: SPCONV_ALLOW_TF32, spconv/constants.py:117) */
enum spx_f32_mode { SPX_F32_EXACT = 0, SPX_F32_TF32 = 1 };

typedef void *spx_stream_t; /* cudaStream_t */

const char *spx_last_error(void);
int spx_version(void);
/* 0 if device `dev` is usable by this library (compute capability 10.x); fills sm count */
int spx_device_check(int dev, int *sm_count, int *cc_major, int *cc_minor);

/* ------------------------------------------------------------------ rulebook */

typedef struct {
    int ndim;                       /* 1..4 */
    int batch_size;
    int in_dims[SPX_MAX_NDIM];      /* input spatial shape */
    int out_dims[SPX_MAX_NDIM];     /* output spatial shape (== in_dims for SubM) */
    int ksize[SPX_MAX_NDIM];
    int stride[SPX_MAX_NDIM];
    int padding[SPX_MAX_NDIM];
    int dilation[SPX_MAX_NDIM];
    int transposed;                 /* regular conv only */
} spx_conv_geometry;

/* scratch bytes needed by the rulebook entry points for `num_in` inputs
 * (`max_out` = upper bound on outputs; ignored for SubM) */
size_t spx_rulebook_workspace_size(const spx_conv_geometry *g, int64_t num_in, int64_t max_out,
                                   int is_subm);
/* upper bound on active outputs of a regular conv (reference: get_handcrafted_max_act_out,
 * spconv/csrc/sparse/all.py:1559-1580) */
int64_t spx_conv_max_out(const spx_conv_geometry *g, int64_t num_in);

/*
 * SubM rulebook.  indices [N, ndim+1] int32 (b, d0, d1, ..).  Writes every element of
 *   pair_fwd [kv, N]   pair_fwd[k][o] = i  (-1 = none)
 *   pair_bwd [kv, N]   pair_bwd[k][i] = o  (may be NULL)
 *   mask     [N, words] uint32, bit k%32 of word k/32 set iff pair_fwd[k][o] != -1  (may be NULL)
 * `words` = ceil(kv/32).
 */
int spx_subm_rulebook(const spx_conv_geometry *g, const int32_t *indices, int64_t N,
                      int32_t *pair_fwd, int32_t *pair_bwd, uint32_t *mask, int32_t *row_table,
                      void *workspace, size_t workspace_bytes, spx_stream_t stream);

/*
 * Optional by-product of spx_subm_rulebook: row_table [N][32] int32 (16-byte aligned), row o =
 * pair_fwd[0..kv-1][o] padded with -1 -- the forward table transposed to one 128-byte line per
 * voxel.  spx_build_tile_table re-reads the rulebook in mask_argsort order; from this copy that
 * costs 4 sectors per row instead of one per (row, offset).  Produced only for geometries where
 * this returns 1 (3-D 3x3x3 with 32-bit keys); pass NULL otherwise.
 */
int spx_subm_row_table_supported(const spx_conv_geometry *g);

/*
 * Regular / transposed conv rulebook, two-phase because the output count M is
 * data-dependent.  Stage 1 hashes every (offset, input) hit, ranks the distinct outputs in
 * the reference CPU's first-touch order and returns M (host sync).  Stage 2 fills
 *   out_inds [M, ndim+1], pair_fwd [kv, M], pair_bwd [kv, N],
 *   mask_fwd [M, words], mask_bwd [N, words]   (masks may be NULL).
 * The same workspace must be passed, untouched, to both stages.
 */
int spx_conv_rulebook_stage1(const spx_conv_geometry *g, const int32_t *indices, int64_t N,
                             int64_t *num_out_host, void *workspace, size_t workspace_bytes,
                             spx_stream_t stream);
int spx_conv_rulebook_stage2(const spx_conv_geometry *g, const int32_t *indices, int64_t N,
                             int64_t M, int32_t *out_inds, int32_t *pair_fwd, int32_t *pair_bwd,
                             uint32_t *mask_fwd, uint32_t *mask_bwd, void *workspace,
                             size_t workspace_bytes, spx_stream_t stream);

/*
 * Fused host entry points: one call = the launches of spx_subm_rulebook + spx_mask_argsort +
 * spx_build_tile_table (resp. spx_conv_rulebook_stage2 + both argsorts + both tile tables), all
 * scratch carved from ONE workspace.  They exist because an eager Python caller pays ~10 us of
 * interpreter / ctypes / allocator time per separate call; results are identical.
 * `mask` is left sorted; tile_table / tile_mask (spx_tile_table_elems, tiles * words) may be NULL to
 * skip the tile table.  argsort_bwd / table_bwd may be NULL (inference: no backward direction).
 * spx_conv_rulebook_stage2_all continues a spx_conv_rulebook_stage1 that was given a workspace of
 * spx_conv_rulebook_all_workspace_size bytes.
 */
size_t spx_subm_rulebook_all_workspace_size(const spx_conv_geometry *g, int64_t N);
int spx_subm_rulebook_all(const spx_conv_geometry *g, const int32_t *indices, int64_t N,
                          int32_t *pair_fwd, int32_t *pair_bwd, uint32_t *mask, int32_t *argsort,
                          int do_sort, int32_t *tile_table, uint32_t *tile_mask, void *workspace,
                          size_t workspace_bytes, spx_stream_t stream);
size_t spx_conv_rulebook_all_workspace_size(const spx_conv_geometry *g, int64_t N);
int spx_conv_rulebook_stage2_all(const spx_conv_geometry *g, const int32_t *indices, int64_t N,
                                 int64_t M, int32_t *out_inds, int32_t *pair_fwd, int32_t *pair_bwd,
                                 uint32_t *mask_fwd, uint32_t *mask_bwd, int32_t *argsort_fwd,
                                 int32_t *argsort_bwd, int do_sort, int32_t *table_fwd,
                                 uint32_t *tmask_fwd, int32_t *table_bwd, uint32_t *tmask_bwd,
                                 void *workspace, size_t workspace_bytes, spx_stream_t stream);

/*
 * Compact "Native" rulebook  pairs [2, kv, N] (-1 padded) + indice_pair_num [kv]  in the
 * reference CPU order (ascending input index per offset), derived from pair_bwd [kv, N] by a
 * stable scan.  For SubM only offsets < kv/2 are counted and their mirrors written, the centre
 * row is the identity (spconv/csrc/sparse/indices.py:1670-1703).
 */
int spx_native_pairs(const int32_t *pair_bwd, int64_t N, int kv, int is_subm, int32_t *pairs,
                     int32_t *indice_pair_num, void *workspace, size_t workspace_bytes,
                     spx_stream_t stream);
size_t spx_native_pairs_workspace_size(int64_t N, int kv);

/*
 * Inverse of spx_native_pairs for the ConvAlgo.Native operator path: scatter a compact
 * rulebook into dense tables  table_fwd [kv, n_out] / table_bwd [kv, n_in]  and row masks.
 * `inverse` swaps the roles of pairs[0] / pairs[1] (SparseInverseConv).  Any output may be NULL.
 */
int spx_pairs_to_table(const int32_t *pairs, const int32_t *indice_pair_num, int kv,
                       int64_t pair_stride, int64_t n_in, int64_t n_out, int is_subm, int inverse,
                       int32_t *table_fwd, int32_t *table_bwd, uint32_t *mask_fwd,
                       uint32_t *mask_bwd, spx_stream_t stream);

/*
 * argsort[N] <- stable ascending argsort of mask[N, words] (word 0 most significant) and mask
 * is left SORTED, as thrust::sort_by_key leaves it in the reference.  do_sort == 0: iota only.
 */
size_t spx_mask_argsort_workspace_size(int64_t N, int words);
int spx_mask_argsort(uint32_t *mask, int32_t *argsort, int64_t N, int words, int kv, int do_sort,
                     void *workspace, size_t workspace_bytes, spx_stream_t stream);

/* ------------------------------------------------------------------ conv arithmetic */

typedef struct {
    int dtype;                  /* spx_dtype of features, filters, outputs */
    int f32_mode;               /* spx_f32_mode, only read when dtype == SPX_F32 */
    int kv;                     /* kernel volume */
    int c_in, c_out;            /* C, K of the KRSC filter [K, kv, C] */
    int64_t n_in, n_out;        /* rows of the input / output feature matrices */
    const int32_t *pair;        /* [kv, rows] gather table of THIS pass (see each function) */
    int64_t pair_stride;        /* elements between consecutive offsets of `pair` */
    const uint32_t *mask;       /* [rows, words] in argsort order, or NULL = all offsets */
    const int32_t *argsort;     /* [rows] row visiting order, or NULL = identity */
    int reverse_offsets;        /* 1: offset k of `pair`/`mask` multiplies filter kv-1-k
                                   (SubM dgrad through the forward table; reference
                                   reverse_mask, spconv/csrc/sparse/convops.py:2412) */
    const int32_t *tile_table;  /* optional: spx_build_tile_table output for (pair, argsort);  */
    const uint32_t *tile_mask;  /* both or neither.  Required by the tcgen05 kernels: without   */
                                /* them the call runs on the generic FMA kernels.               */
} spx_gemm_desc;

/*
 * Tile-blocked gather table: the (pair, argsort, mask) triple re-laid so that one 128-row tile is
 * one contiguous block the kernels fetch with a single bulk async copy:
 *   table     [tiles][kv + 1][128] int32:  table[t][k][r] = pair[k][row(t*128 + r)]  (k < kv),
 *                                          table[t][kv][r] = row(t*128 + r)   (-1 past the end)
 *             with row(j) = argsort ? argsort[j] : j
 *   tile_mask [tiles][words] uint32: OR of mask[t*128 .. t*128+127] (mask in visiting order;
 *             NULL mask = all kv offsets) == the reference's mask_output_fwd with mask_width 128
 *             (spconv/csrc/sparse/convops.py:2180-2189)
 * tiles = ceil(rows / 128).  Built once per rulebook, shared by fwd / dgrad / wgrad of every
 * layer that shares the indice_key.
 *
 * The `table` buffer (spx_tile_table_elems int32 elements, 16-byte aligned) continues behind the
 * blocks with what the dynamically scheduled kernels need:
 *   records [tiles][8] int32: {tile, mask words[4], 0, 0, 0}, tiles in order of decreasing
 *             offset count (ties: ascending tile) -- fwd / dgrad CTAs draw tickets from an atomic
 *             counter and take tiles in this order (longest-processing-time-first scheduling);
 *   scratch [64] int32: 32 {ticket counter, finished-CTA counter} pairs.  Zero after the build; every
 *             launch takes the next pair (round-robin per process) and its last CTA zeroes it again,
 *             so up to 32 launches may be in flight on one table (other streams, graph branches).
 */
size_t spx_tile_table_elems(int64_t rows, int kv);
/*
 * row_table (optional, kv <= 32): the [rows][32] by-product of spx_subm_rulebook for the same
 * `pair`; when given, `pair` is not read.
 */
int spx_build_tile_table(const int32_t *pair, int64_t pair_stride, int kv, const int32_t *argsort,
                         const uint32_t *mask, int64_t rows, const int32_t *row_table,
                         int32_t *table, uint32_t *tile_mask, spx_stream_t stream);

/*
 * out[o, :] = act( sum_k x[pair[k][o], :] @ W[:, k, :]^T  + bias )      rows = n_out
 * filters: KRSC [c_out, kv, c_in].  bias (same dtype as features) may be NULL.
 * mask_out [ceil(n_out/128), words] (may be NULL) receives the per-128-row-tile OR of `mask`
 * (the reference's mask_output_fwd, mask_width = 128).
 */
int spx_implicit_gemm_fwd(const spx_gemm_desc *d, const void *features, const void *filters,
                          void *out, const void *bias, int act, float act_alpha,
                          uint32_t *mask_out, spx_stream_t stream);

/*
 * din[i, :] = sum_k dout[pair[k][i], :] @ W[:, k', :]        rows = n_in, k' = k or kv-1-k
 * `pair` is the backward table [kv, n_in] (in -> out).
 */
int spx_implicit_gemm_dgrad(const spx_gemm_desc *d, const void *out_bp, const void *filters,
                            void *din, spx_stream_t stream);

/*
 * dW[:, k, :] = sum_o dout[o, :]^T  x[pair[k][o], :]          rows = n_out, pair = forward table
 * dfilters: KRSC, same dtype as features.  workspace holds fp32 partial sums.
 */
size_t spx_implicit_gemm_wgrad_workspace_size(const spx_gemm_desc *d);
int spx_implicit_gemm_wgrad(const spx_gemm_desc *d, const void *features, const void *out_bp,
                            void *dfilters, void *workspace, size_t workspace_bytes,
                            spx_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Data-parallel weight-gradient exchange over NVLink peer memory.  The reference has no
 * distributed code: users wrap it in torch DDP, i.e. an NCCL all-reduce of dW after the backward
 * pass.  Here the SEND side is the tail of the weight-gradient kernel itself (csrc/peer.cu): the kernel
 * that reduces the split-K partials writes this rank's fp32 slice sums into its own exchange buffer and
 * its last CTA publishes the epoch to every rank (one system fence, `world` flag stores over NVLink).
 * The RECEIVE side (finish) reads every rank's slices through the peer mapping and sums them in rank
 * order, so all replicas end with bit-identical gradients after one rounding.  Put independent work (the
 * input gradient of the same layer) between push and finish and the NVLink latency is hidden.
 *
 * Set-up (once per process group; the host side passes the 64-byte handles around, e.g. with
 * torch.distributed.all_gather_object): every rank creates its buffer, opens the others', and fills
 * a spx_peer_group with the addresses AS MAPPED IN ITS OWN PROCESS (buffers[rank] = its own).
 * Contract: on each rank, push and finish of one group alternate in stream order (one exchange in
 * flight), and all ranks issue the same sequence of exchanges.
 */
#define SPX_MAX_PEERS 16
typedef struct spx_peer_group {
    int world, rank;
    int timeout_ms;                 /* a peer that does not arrive in time: NaN result + spx_peer_error (0 = 20 s) */
    int colocated;                  /* ranks of this group that share ONE device (tests); 0 or 1 = one rank per GPU */
    uint64_t capacity_bytes;        /* largest exchanged tensor, as fp32 (what spx_peer_buffer_create got) */
    void *buffers[SPX_MAX_PEERS];   /* exchange buffer of every rank */
} spx_peer_group;

size_t spx_peer_buffer_bytes(size_t capacity_bytes, int world);
int spx_peer_buffer_create(size_t capacity_bytes, int world, void **buffer, unsigned char handle[64]);
int spx_peer_buffer_open(const unsigned char handle[64], void **mapped);
int spx_peer_buffer_close(void *mapped);
int spx_peer_buffer_destroy(void *buffer);
/* sticky error word of this rank's buffer (1 = a peer timed out); synchronous copy */
int spx_peer_error(const spx_peer_group *pg, int *error);

/* Weight gradient of this rank (see spx_implicit_gemm_wgrad), published to the group in fp32 by the kernel
 * that reduces the split-K partials.  dfilters is NOT valid afterwards (scratch for shapes
 * the tcgen05 kernel does not tile): spx_peer_finish(pg, dfilters, kv*C*K, dtype, scale) writes it. */
int spx_implicit_gemm_wgrad_push(const spx_gemm_desc *d, const void *features, const void *out_bp,
                                 void *dfilters, void *workspace, size_t workspace_bytes,
                                 const spx_peer_group *pg, spx_stream_t stream);
/* push + finish back to back: dfilters = scale * sum over ranks */
int spx_implicit_gemm_wgrad_allreduce(const spx_gemm_desc *d, const void *features, const void *out_bp,
                                      void *dfilters, void *workspace, size_t workspace_bytes,
                                      const spx_peer_group *pg, float scale, spx_stream_t stream);
/* the same exchange for an existing small tensor (bias gradients ...): push sends `data` (dtype SPX_F32 /
 * SPX_F16 / SPX_BF16), finish writes out = scale * sum over ranks (out may be data); allreduce = both */
int spx_peer_push(const spx_peer_group *pg, const void *data, int64_t count, int dtype, spx_stream_t stream);
int spx_peer_finish(const spx_peer_group *pg, void *out, int64_t count, int dtype, float scale, spx_stream_t stream);
int spx_peer_allreduce(const spx_peer_group *pg, void *data, int64_t count, int dtype, float scale,
                       spx_stream_t stream);

/* x[r, j] = act(x[r, j] + bias[j])   in place; bias may be NULL */
int spx_bias_act_inplace(void *x, const void *bias, int64_t rows, int cols, int dtype, int act,
                         float act_alpha, spx_stream_t stream);

/* ------------------------------------------------------------------ point cloud -> voxels */

/*
 * Replaces SpconvOps.point2voxel_cuda / Point2Voxel (spconv/csrc/sparse/all.py:1349-1490,
 * spconv/csrc/sparse/pointops.py:120-490) with the deterministic semantics of the reference's CPU
 * implementation (Point2VoxelCPU, pointops.py:589-695): voxel id = rank of the voxel's first point
 * in input order; voxels beyond max_voxels are dropped; a voxel keeps its first
 * max_points_per_voxel points in input order.
 *   points [N, num_features] fp32 (device), the first ndim features are coordinates (x, y, z, ..);
 *   vsize / grid_size / coors_range are HOST arrays in the internal axis order that
 *   calc_meta_data produces (zyx != 0: axis j reads point feature ndim-1-j);  coors_range holds the
 *   ndim lower bounds first.
 * Stage 1 hashes and ranks the voxels and returns their number (host sync, as the reference's
 * sliced return tensors require): *num_voxels = min(total, max_voxels).  Stage 2 (same untouched
 * workspace) fills
 *   voxels [>= num_voxels, max_points, num_features] (unused slots are left as the caller set them,
 *          or receive the voxel mean when empty_mean != 0),  indices [>= num_voxels, ndim],
 *   num_per_voxel [>= num_voxels],  pc_voxel_id [N] int64 (-1 = no voxel).
 */
size_t spx_point2voxel_workspace_size(int64_t num_points, int ndim);
int spx_point2voxel_stage1(const float *points, int64_t N, int num_features, int ndim, int zyx,
                           const float *vsize_host, const int *grid_size_host,
                           const float *coors_range_host, int64_t max_voxels, int64_t *num_voxels_host,
                           int64_t *total_voxels_host, void *workspace, size_t workspace_bytes,
                           spx_stream_t stream);
int spx_point2voxel_stage2(const float *points, int64_t N, int num_features, int ndim, int zyx,
                           const float *vsize_host, const int *grid_size_host,
                           const float *coors_range_host, int64_t num_voxels, int64_t total_voxels,
                           int max_points_per_voxel, int empty_mean, float *voxels, int32_t *indices,
                           int32_t *num_per_voxel, int64_t *pc_voxel_id, void *workspace,
                           size_t workspace_bytes, spx_stream_t stream);

/* ------------------------------------------------------------------ pooling on the rulebook */

/*
 * out[o, :] = reduce over the offsets k with pair_fwd[k][o] >= 0 of x[pair_fwd[k][o], :]
 *   mode 0  max, rows without any input get the dtype's lowest value
 *           (SparseMaxPool, ConvAlgo.MaskImplicitGemm: maxpool.py:76-117)
 *   mode 1  max with a floor of 0 (ConvAlgo.Native: the reference raises a zero-initialised buffer,
 *           spconv/pytorch/ops.py:1910-1936 + maxpool.py:41-73); the caller passes the dense table
 *           made by spx_pairs_to_table from the compact pairs
 *   mode 2  mean over the valid inputs; count_out [n_out] int32 (may be NULL) receives their number
 *           (SparseAvgPool: maxpool.py:211-259)
 * channels * element size must be a multiple of 16 bytes.  dtype: f32 / f16 / bf16, int8 for max.
 */
int spx_indice_pool_fwd(int mode, const void *features, void *out, const int32_t *pair_fwd,
                        int64_t pair_stride, int kv, int64_t n_out, int channels, int dtype,
                        int32_t *count_out, spx_stream_t stream);
/*
 * Input gradient through pair_bwd [kv, n_in] (in -> out):
 *   modes 0, 1  din[i] = sum_k (x[i] == y[o_k]) ? dy[o_k] : 0        (maxpool.py:120-208)
 *   mode 2      din[i] = sum_k dy[o_k] * count_out[o_k]              (maxpool.py:262-300, as is)
 */
int spx_indice_pool_bwd(int mode, const void *features, const void *out_features, const void *out_bp,
                        void *din, const int32_t *pair_bwd, int64_t pair_stride, int kv, int64_t n_in,
                        int channels, int dtype, const int32_t *count_out, spx_stream_t stream);
/*
 * Rows of every sample in input order: out_indices [batch_size, n] (only the first counts[b]
 * entries of row b are written), counts [batch_size]; coords [n, row_ints] with the batch index
 * first.  Deterministic (the reference appends with atomics, maxpool.py:303-341; the CPU version
 * keeps input order, :599-620).
 */
int spx_global_pool_rearrange(const int32_t *coords, int64_t n, int row_ints, int batch_size,
                              int32_t *out_indices, int32_t *counts, spx_stream_t stream);

/*
 * int8 inference forward (reference formula: test/test_all_algo.py:272-287,
 * spconv/pytorch/quantization/quantized/conv.py:368-377):
 *   acc_i32 = sum_k x_i8[pair[k][o]] @ W_i8[:, k, :]^T
 *   y = acc * scale[j] + bias[j] (+ add_i8[o, j] * add_scale);  act;  q = clip(rint(y), -128, 127)
 * out_dtype: SPX_I8 (quantised) or SPX_F32 / SPX_F16 (y stored directly).
 */
int spx_implicit_gemm_fwd_int8(const spx_gemm_desc *d, const int8_t *features,
                               const int8_t *filters, void *out, int out_dtype,
                               const float *scale, const float *bias, const int8_t *output_add,
                               float output_add_scale, int act, float act_alpha,
                               spx_stream_t stream);

/* which kernel family served the last call of each kind on this thread: 0 none, 1 SIMT,
 * 2 tcgen05.  Used by tests/bench to prove the tensor-core path ran. */
int spx_last_kernel_family(void);
/* number of kernel launches issued by this library (all host threads) since the last reset */
int64_t spx_launch_count(int reset);
/*
 * Test / perf-triage switches (never needed for correct operation; the reference's counterpart is
 * the SPCONV_DEBUG_* environment, spconv/constants.py:100-125).
 *   force_family: -1 keep, 0 automatic, 1 generic FMA kernels, 2 tcgen05 kernels (error if the
 *                 shape does not tile) -- the start-up value comes from SPX_FORCE_SIMT / SPX_FORCE_TC,
 *                 read once when the library is loaded;
 *   tc_ctas:      0 keep, 1 or 2 resident CTAs per SM for the forward / dgrad kernel;
 *   debug_bits:   A/B and ablation mask (bits 1..32: ablations of the tcgen05 kernels, results are wrong by
 *                 construction; 64 / 512: alternative sorts; 128: legacy regular-conv rulebook; 256: fp32+TF32
 *                 input gradient on the FMA kernel instead of tcgen05, 4096: the same for the weight gradient;
 *                 1024: weight-gradient pass split; 2048: regular-conv mask sorts / tile tables one job per launch);
 *   trace_buf:    NULL or a DEVICE buffer of at least 8*2048 int64 that receives clock stamps.
 */
int spx_debug_configure(int force_family, int tc_ctas, int debug_bits, void *trace_buf,
                        size_t trace_bytes);

#ifdef __cplusplus
}
#endif
#endif /* SPCONV_B200_H_ */
