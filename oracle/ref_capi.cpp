// C entry points over the reference's own CPU code (oracle/_ref/spconv_ref_gen.h, extracted from
// /root/reference by oracle/make_ref.py).  Same signatures as the restatement in spconv_oracle.c so
// tests can run both on identical inputs.  TEST INFRASTRUCTURE ONLY.
#include "_ref/spconv_ref_gen.h"

namespace {
template <int ND> tv::array<int, ND> arr(const int *p) {
    tv::array<int, ND> a;
    for (int i = 0; i < ND; ++i) a[i] = p[i];
    return a;
}
tv::Tensor t_i32(const void *p, std::vector<int64_t> shape) { return tv::Tensor(const_cast<void *>(p), std::move(shape), tv::int32); }
tv::Tensor t_f32(const void *p, std::vector<int64_t> shape) { return tv::Tensor(const_cast<void *>(p), std::move(shape), tv::float32); }
}  // namespace

#define REF_DISPATCH_NDIM(ndim, EXPR)                                                     \
    switch (ndim) {                                                                       \
        case 1: { namespace R = ref_nd1; constexpr int ND = 1; EXPR; } break;             \
        case 2: { namespace R = ref_nd2; constexpr int ND = 2; EXPR; } break;             \
        case 3: { namespace R = ref_nd3; constexpr int ND = 3; EXPR; } break;             \
        case 4: { namespace R = ref_nd4; constexpr int ND = 4; EXPR; } break;             \
        default: return -1;                                                               \
    }

extern "C" {

// pairs [2, kv, N] pre-filled with -1, num [kv] zeroed by the caller (spconv/csrc/sparse/all.py:2071-2075)
int ref_subm_rulebook(const int32_t *indices, int N, int ndim, int batch_size, const int *dims, const int *ksize,
                      const int *dilation, int32_t *pairs, int32_t *num) {
    int kv = 1;
    for (int a = 0; a < ndim; ++a) kv *= ksize[a];
    try {
        REF_DISPATCH_NDIM(ndim, return R::SparseConvIndicesCPU::generate_subm_conv_inds(
            t_i32(indices, {N, ndim + 1}), t_i32(pairs, {2, kv, N}), tv::Tensor(), t_i32(num, {kv}), batch_size,
            arr<ND>(dims), arr<ND>(ksize), arr<ND>(dilation)));
    } catch (const std::exception &e) {
        return std::strstr(e.what(), "odd ksize") ? -2 : -3;
    }
    return -1;
}

// out_inds must hold kv * N rows (all.py:2121); returns the number of active outputs
int ref_conv_rulebook(const int32_t *indices, int N, int ndim, int batch_size, const int *out_dims,
                      const int *in_dims, const int *ksize, const int *stride, const int *padding,
                      const int *dilation, int transposed, int32_t *pairs, int32_t *out_inds, int32_t *num) {
    int kv = 1;
    for (int a = 0; a < ndim; ++a) kv *= ksize[a];
    try {
        REF_DISPATCH_NDIM(ndim, return R::SparseConvIndicesCPU::generate_conv_inds(
            t_i32(indices, {N, ndim + 1}), t_i32(pairs, {2, kv, N}), t_i32(out_inds, {(int64_t)kv * N, ndim + 1}),
            t_i32(num, {kv}), batch_size, arr<ND>(out_dims), arr<ND>(in_dims), arr<ND>(ksize), arr<ND>(stride),
            arr<ND>(padding), arr<ND>(dilation), transposed != 0));
    } catch (const std::exception &) {
        return -3;
    }
    return -1;
}

void ref_gather_f32(float *buf, const float *src, const int32_t *inds, int n, int channels, int src_rows) {
    GatherCPU::gather(t_f32(buf, {n, channels}), t_f32(src, {src_rows, channels}), t_i32(inds, {n}));
}

void ref_scatter_add_f32(float *dst, const float *buf, const int32_t *inds, int n, int channels, int dst_rows) {
    GatherCPU::scatter_add(t_f32(dst, {dst_rows, channels}), t_f32(buf, {n, channels}), t_i32(inds, {n}));
}

// one offset of the Native max pool: out[out_inds[i]] = max(out[..], in[in_inds[i]])  (maxpool.py:623-658)
void ref_maxpool_fwd_f32(float *out, const float *in, const int32_t *out_inds, const int32_t *in_inds, int nhot,
                         int channels, int out_rows, int in_rows) {
    IndiceMaxPoolCPU::forward(t_f32(out, {out_rows, channels}), t_f32(in, {in_rows, channels}), t_i32(out_inds, {nhot}),
                              t_i32(in_inds, {nhot}));
}

// din[in_inds[i]] += dout[out_inds[i]] where in == out   (maxpool.py:661-700)
void ref_maxpool_bwd_f32(const float *out, const float *in, const float *dout, float *din, const int32_t *out_inds,
                         const int32_t *in_inds, int nhot, int channels, int out_rows, int in_rows) {
    IndiceMaxPoolCPU::backward(t_f32(out, {out_rows, channels}), t_f32(in, {in_rows, channels}),
                               t_f32(dout, {out_rows, channels}), t_f32(din, {in_rows, channels}),
                               t_i32(out_inds, {nhot}), t_i32(in_inds, {nhot}));
}

// out_indices [batch, n], counts [batch] zeroed by the caller   (maxpool.py:599-620)
void ref_global_pool_rearrange(int32_t *out_indices, const int32_t *coords, int32_t *counts, int n, int row_ints,
                               int batch) {
    IndiceMaxPoolCPU::global_pool_rearrange(t_i32(out_indices, {batch, n}), t_i32(coords, {n, row_ints}),
                                            t_i32(counts, {batch}));
}

// Point2VoxelCPU::point_to_voxel_static / _empty_mean_static (pointops.py:589-695), 3-D zyx.
// densehash: int32 [grid_size...] filled with -1 by the caller (spconv/pytorch/utils.py:56-60).
// vsize / grid / range are in INTERNAL (zyx) order as calc_meta_data returns them.  Returns the voxel count.
int ref_point2voxel_3d(const float *points, int n, int nf, float *voxels, int32_t *indices, int32_t *num_per_voxel,
                       int32_t *densehash, int64_t *pc_voxel_id, const float *vsize, const int *grid,
                       const float *range, int max_voxels, int max_points, int empty_mean, int clear_voxels) {
    std::array<float, 3> vs{vsize[0], vsize[1], vsize[2]};
    std::array<int, 3> gs{grid[0], grid[1], grid[2]}, gstride{grid[1] * grid[2], grid[2], 1};
    std::array<float, 6> cr{range[0], range[1], range[2], range[3], range[4], range[5]};
    tv::Tensor pts = t_f32(points, {n, nf}), vox = t_f32(voxels, {max_voxels, max_points, nf});
    tv::Tensor ind = t_i32(indices, {max_voxels, 3}), num = t_i32(num_per_voxel, {max_voxels});
    tv::Tensor dh = t_i32(densehash, {grid[0], grid[1], grid[2]});
    tv::Tensor ids(pc_voxel_id, {n}, tv::int64);
    auto res = empty_mean
        ? ref_p2v3::Point2VoxelCPU::point_to_voxel_empty_mean_static(pts, vox, ind, num, dh, ids, vs, gs, gstride, cr, clear_voxels != 0)
        : ref_p2v3::Point2VoxelCPU::point_to_voxel_static(pts, vox, ind, num, dh, ids, vs, gs, gstride, cr, clear_voxels != 0);
    return (int)std::get<0>(res).dim(0);
}

// Point2VoxelCommon::calc_meta_data (pointops.py:42-88): xyz inputs -> internal-order vsize[3], grid[3], range[6]
void ref_point2voxel_meta_3d(const float *vsize_xyz, const float *range_xyz, float *vsize, int *grid, int64_t *stride,
                             float *range) {
    auto r = ref_p2v3::Point2VoxelCommon::calc_meta_data({vsize_xyz[0], vsize_xyz[1], vsize_xyz[2]},
                                                         {range_xyz[0], range_xyz[1], range_xyz[2], range_xyz[3], range_xyz[4], range_xyz[5]});
    for (int i = 0; i < 3; ++i) { vsize[i] = std::get<0>(r)[i]; grid[i] = std::get<1>(r)[i]; stride[i] = std::get<2>(r)[i]; }
    for (int i = 0; i < 6; ++i) range[i] = std::get<3>(r)[i];
}

int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
