// Internal interface between the C-ABI dispatcher (api_gemm.cu) and the two kernel families
// (gemm_simt.cu: generic fp32-FMA kernels; gemm_tc.cu: tcgen05 / TMEM / TMA kernels).
#pragma once
#include "common.cuh"

namespace spx {

// y[r, :] = act( sum_k  X[pair[k][row(r)], :] * W_k'  + bias )        r in [0, rows)
//   row(r) = argsort ? argsort[r] : r ;  y row written = row(r)
//   W is the KRSC filter [c_out, kv, c_in]:
//     transpose_w == 0 (forward): X has c_in channels, y has c_out:  W_k'[x][y] = W[y][k'][x]
//     transpose_w == 1 (dgrad)  : X has c_out channels, y has c_in:  W_k'[x][y] = W[x][k'][y]
//   k' = reverse ? kv-1-k : k
struct GatherGemmArgs {
    int dtype, f32_mode, kv, c_in, c_out, transpose_w, reverse, act;
    float alpha;
    int64_t rows, x_rows;
    const void *x, *w, *bias;
    void *y;
    const int32_t *pair;
    int64_t pair_stride;
    const uint32_t *mask;      // [rows, words] in visiting order, or NULL
    const int32_t *argsort;    // [rows] or NULL
    uint32_t *mask_out;        // [ceil(rows/128), words] or NULL
    const int32_t *tile_table; // [tiles][kv+1][128] or NULL (spx_build_tile_table)
    const uint32_t *tile_mask; // [tiles][words] or NULL
    __host__ __device__ int cx() const { return transpose_w ? c_out : c_in; }
    __host__ __device__ int cy() const { return transpose_w ? c_in : c_out; }
};

// dW[:, k, :] = sum_o dout[o, :]^T x[pair[k][o], :]
struct WgradArgs {
    int dtype, f32_mode, kv, c_in, c_out;
    int64_t n_in, n_out;
    const void *x, *dout;
    void *dw;
    const int32_t *pair;       // forward table [kv, n_out]
    int64_t pair_stride;
    const uint32_t *mask;      // [n_out, words] in visiting order or NULL
    const int32_t *argsort;    // [n_out] or NULL
    const int32_t *tile_table; // [tiles][kv+1][128] or NULL
    const uint32_t *tile_mask; // [tiles][words] or NULL
    void *workspace;
    size_t workspace_bytes;
    const spx_peer_group *peers;   // NULL, or: push this rank's fp32 dW to these ranks instead of writing dw
};

// Layout of the buffer spx_build_tile_table fills (int32 elements):
//   [tiles][kv+1][128]              gather blocks
//   [tiles][TT_REC_INTS]            schedule records {tile, mask[4], 0, 0, 0}, heaviest tile first
//   [TT_STATE_INTS]                 scheduler scratch: [0] ticket counter, [1] finished CTAs
//                                   (zero between launches; a launch leaves it zero again)
constexpr int TT_REC_INTS = 8;
constexpr int TT_STATE_INTS = 64;
__host__ __device__ inline int64_t tt_blocks_elems(int64_t tiles, int kv) { return tiles * (int64_t)(kv + 1) * 128; }
__host__ __device__ inline int64_t tt_total_elems(int64_t tiles, int kv) {
    return tt_blocks_elems(tiles, kv) + tiles * TT_REC_INTS + TT_STATE_INTS;
}

int simt_gather_gemm(const GatherGemmArgs &a, cudaStream_t stream);
int simt_wgrad(const WgradArgs &a, cudaStream_t stream);

bool tc_gather_gemm_supported(const GatherGemmArgs &a);
int tc_gather_gemm(const GatherGemmArgs &a, cudaStream_t stream);
bool tc_wgrad_supported(const WgradArgs &a);
size_t tc_wgrad_workspace_size(const WgradArgs &a);
int tc_wgrad(const WgradArgs &a, cudaStream_t stream);

struct Int8Args {
    GatherGemmArgs g;          // dtype = SPX_I8; bias / act fields unused
    int out_dtype;
    const float *scale, *bias_f32;
    const int8_t *output_add;
    float output_add_scale;
};
int simt_gather_gemm_int8(const Int8Args &a, cudaStream_t stream);
bool tc_gather_gemm_int8_supported(const Int8Args &a);
int tc_gather_gemm_int8(const Int8Args &a, cudaStream_t stream);

}  // namespace spx
