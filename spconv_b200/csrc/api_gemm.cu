// C-ABI entry points of the conv arithmetic: validation + dispatch between the tcgen05
// kernels (gemm_tc.cu) and the generic fp32-FMA kernels (gemm_simt.cu).
#include "gemm.cuh"
#include "peer.cuh"
#include <stdlib.h>

using namespace spx;

namespace spx {
int write_tile_masks(const uint32_t *mask, int64_t rows, int kv, uint32_t *out, cudaStream_t stream);
}

// SPX_FORCE_SIMT=1 pins the generic kernels (debug / A-B comparisons in tests); read once at load
static bool force_simt() { return runtime_cfg().force_simt != 0; }
static bool force_tc() { return runtime_cfg().force_tc != 0; }

static int check_desc(const spx_gemm_desc *d, const char *who) {
    SPX_REQUIRE(d != nullptr, "%s: descriptor is NULL", who);
    SPX_REQUIRE(d->kv >= 1 && d->kv <= 128, "%s: kernel volume %d not in [1,128]", who, d->kv);
    SPX_REQUIRE(d->c_in >= 1 && d->c_out >= 1, "%s: bad channel counts %d -> %d", who, d->c_in, d->c_out);
    SPX_REQUIRE(d->n_in >= 0 && d->n_out >= 0, "%s: negative row counts", who);
    SPX_REQUIRE(d->n_in < 2147483647ll && d->n_out < 2147483647ll, "%s: row counts must fit in int32", who);
    SPX_REQUIRE(d->pair != nullptr || (d->n_in == 0 || d->n_out == 0), "%s: pair table is NULL", who);
    SPX_REQUIRE((d->tile_table == nullptr) == (d->tile_mask == nullptr), "%s: tile_table and tile_mask go together", who);
    return 0;
}

static GatherGemmArgs make_args(const spx_gemm_desc *d, bool dgrad) {
    GatherGemmArgs a;
    memset(&a, 0, sizeof(a));
    a.dtype = d->dtype; a.f32_mode = d->f32_mode; a.kv = d->kv; a.c_in = d->c_in; a.c_out = d->c_out;
    a.transpose_w = dgrad ? 1 : 0;
    a.reverse = d->reverse_offsets;
    a.rows = dgrad ? d->n_in : d->n_out;
    a.x_rows = dgrad ? d->n_out : d->n_in;
    a.pair = d->pair; a.pair_stride = d->pair_stride;
    a.mask = d->mask; a.argsort = d->argsort;
    a.tile_table = d->tile_table; a.tile_mask = d->tile_mask;
    return a;
}

static int run_gather_gemm(const GatherGemmArgs &a, cudaStream_t stream) {
    if (a.rows == 0) return 0;
    bool exact_f32 = a.dtype == SPX_F32 && a.f32_mode == SPX_F32_EXACT;
    bool tc_ok = !force_simt() && !exact_f32 && tc_gather_gemm_supported(a);
    if (force_tc() && !tc_ok) {
        set_error("SPX_FORCE_TC=1 but the tcgen05 path does not support this call (dtype %d, C %d, K %d)",
                  a.dtype, a.c_in, a.c_out);
        return 3;
    }
    if (tc_ok) { set_family(2); return tc_gather_gemm(a, stream); }
    set_family(1);
    return simt_gather_gemm(a, stream);
}

extern "C" int spx_implicit_gemm_fwd(const spx_gemm_desc *d, const void *features, const void *filters, void *out,
                                     const void *bias, int act, float act_alpha, uint32_t *mask_out,
                                     spx_stream_t stream) {
    if (check_desc(d, "implicit_gemm_fwd")) return 2;
    SPX_REQUIRE(d->dtype == SPX_F32 || d->dtype == SPX_F16 || d->dtype == SPX_BF16,
                "implicit_gemm_fwd: dtype %d not supported (int8 has its own entry point)", d->dtype);
    if (d->n_out == 0) return 0;
    SPX_REQUIRE(features && filters && out, "implicit_gemm_fwd: NULL tensor");
    GatherGemmArgs a = make_args(d, false);
    a.x = features; a.w = filters; a.y = out; a.bias = bias; a.act = act; a.alpha = act_alpha;
    a.mask_out = mask_out;
    return run_gather_gemm(a, (cudaStream_t)stream);
}

extern "C" int spx_implicit_gemm_dgrad(const spx_gemm_desc *d, const void *out_bp, const void *filters, void *din,
                                       spx_stream_t stream) {
    if (check_desc(d, "implicit_gemm_dgrad")) return 2;
    SPX_REQUIRE(d->dtype == SPX_F32 || d->dtype == SPX_F16 || d->dtype == SPX_BF16,
                "implicit_gemm_dgrad: dtype %d not supported", d->dtype);
    if (d->n_in == 0) return 0;
    SPX_REQUIRE(out_bp && filters && din, "implicit_gemm_dgrad: NULL tensor");
    GatherGemmArgs a = make_args(d, true);
    a.x = out_bp; a.w = filters; a.y = din; a.bias = nullptr; a.act = SPX_ACT_NONE;
    return run_gather_gemm(a, (cudaStream_t)stream);
}

static WgradArgs make_wgrad(const spx_gemm_desc *d) {
    WgradArgs w;
    memset(&w, 0, sizeof(w));
    w.dtype = d->dtype; w.f32_mode = d->f32_mode; w.kv = d->kv; w.c_in = d->c_in; w.c_out = d->c_out;
    w.n_in = d->n_in; w.n_out = d->n_out;
    w.pair = d->pair; w.pair_stride = d->pair_stride; w.mask = d->mask; w.argsort = d->argsort;
    w.tile_table = d->tile_table; w.tile_mask = d->tile_mask;
    return w;
}

extern "C" size_t spx_implicit_gemm_wgrad_workspace_size(const spx_gemm_desc *d) {
    if (!d) return 0;
    WgradArgs w = make_wgrad(d);
    bool exact_f32 = d->dtype == SPX_F32 && d->f32_mode == SPX_F32_EXACT;
    if (!force_simt() && !exact_f32 && tc_wgrad_supported(w)) return tc_wgrad_workspace_size(w);
    return 256;
}

// pg == NULL: plain weight gradient.  pg != NULL: push this rank's fp32 gradient to the group; `finish` also
// runs the receive side right away (otherwise the caller does, after the work it wants to overlap).
static int wgrad_entry(const spx_gemm_desc *d, const void *features, const void *out_bp, void *dfilters, void *workspace,
                       size_t workspace_bytes, const spx_peer_group *pg, bool finish, float scale, spx_stream_t stream);

extern "C" int spx_implicit_gemm_wgrad(const spx_gemm_desc *d, const void *features, const void *out_bp,
                                       void *dfilters, void *workspace, size_t workspace_bytes,
                                       spx_stream_t stream) {
    return wgrad_entry(d, features, out_bp, dfilters, workspace, workspace_bytes, nullptr, false, 1.f, stream);
}

extern "C" int spx_implicit_gemm_wgrad_push(const spx_gemm_desc *d, const void *features, const void *out_bp,
                                            void *dfilters, void *workspace, size_t workspace_bytes,
                                            const spx_peer_group *pg, spx_stream_t stream) {
    SPX_REQUIRE(pg != nullptr, "implicit_gemm_wgrad_push: peer group is NULL");
    return wgrad_entry(d, features, out_bp, dfilters, workspace, workspace_bytes, pg, false, 1.f, stream);
}

extern "C" int spx_implicit_gemm_wgrad_allreduce(const spx_gemm_desc *d, const void *features, const void *out_bp,
                                                 void *dfilters, void *workspace, size_t workspace_bytes,
                                                 const spx_peer_group *pg, float scale, spx_stream_t stream) {
    SPX_REQUIRE(pg != nullptr, "implicit_gemm_wgrad_allreduce: peer group is NULL");
    return wgrad_entry(d, features, out_bp, dfilters, workspace, workspace_bytes, pg, true, scale, stream);
}

static int wgrad_entry(const spx_gemm_desc *d, const void *features, const void *out_bp, void *dfilters, void *workspace,
                       size_t workspace_bytes, const spx_peer_group *pg, bool finish, float scale, spx_stream_t stream) {
    if (check_desc(d, "implicit_gemm_wgrad")) return 2;
    SPX_REQUIRE(d->dtype == SPX_F32 || d->dtype == SPX_F16 || d->dtype == SPX_BF16,
                "implicit_gemm_wgrad: dtype %d not supported", d->dtype);
    SPX_REQUIRE(dfilters != nullptr, "implicit_gemm_wgrad: dfilters is NULL");
    const int64_t dw_count = (int64_t)d->kv * d->c_in * d->c_out;
    if (d->n_out == 0 || d->n_in == 0) {       // an empty shard still takes part in the exchange
        SPX_CHECK_CUDA(cudaMemsetAsync(dfilters, 0, (size_t)dw_count * dtype_bytes(d->dtype), (cudaStream_t)stream));
        if (pg) {
            if (int rc = peer_push(nullptr, 0, 0, dfilters, dw_count, d->dtype, pg, (cudaStream_t)stream)) return rc;
            if (finish) {
                return peer_finish(dfilters, dw_count, d->dtype, pg, scale, (cudaStream_t)stream);
            }
        }
        return 0;
    }
    SPX_REQUIRE(features && out_bp, "implicit_gemm_wgrad: NULL tensor");
    WgradArgs w = make_wgrad(d);
    w.peers = pg;
    w.x = features; w.dout = out_bp; w.dw = dfilters; w.workspace = workspace; w.workspace_bytes = workspace_bytes;
    bool exact_f32 = d->dtype == SPX_F32 && d->f32_mode == SPX_F32_EXACT;
    bool tc_ok = !force_simt() && !exact_f32 && tc_wgrad_supported(w);
    if (force_tc() && !tc_ok) {
        set_error("SPX_FORCE_TC=1 but the tcgen05 wgrad does not support this call (dtype %d, C %d, K %d)", d->dtype,
                  d->c_in, d->c_out);
        return 3;
    }
    if (tc_ok) {
        SPX_REQUIRE(workspace && workspace_bytes >= tc_wgrad_workspace_size(w),
                    "implicit_gemm_wgrad: workspace too small (%zu < %zu)", workspace_bytes, tc_wgrad_workspace_size(w));
        set_family(2);
        if (int rc = tc_wgrad(w, (cudaStream_t)stream)) return rc;     // with peers: partial sums pushed, dW not written yet
        if (pg && finish) {
            return peer_finish(dfilters, dw_count, d->dtype, pg, scale, (cudaStream_t)stream);
        }
        return 0;
    }
    set_family(1);
    if (int rc = simt_wgrad(w, (cudaStream_t)stream)) return rc;
    if (pg) {
        if (int rc = peer_push(nullptr, 0, 0, dfilters, dw_count, d->dtype, pg, (cudaStream_t)stream)) return rc;
        if (finish) {
            return peer_finish(dfilters, dw_count, d->dtype, pg, scale, (cudaStream_t)stream);
        }
    }
    return 0;
}

extern "C" int spx_implicit_gemm_fwd_int8(const spx_gemm_desc *d, const int8_t *features, const int8_t *filters,
                                          void *out, int out_dtype, const float *scale, const float *bias,
                                          const int8_t *output_add, float output_add_scale, int act,
                                          float act_alpha, spx_stream_t stream) {
    if (check_desc(d, "implicit_gemm_fwd_int8")) return 2;
    SPX_REQUIRE(d->dtype == SPX_I8, "implicit_gemm_fwd_int8: descriptor dtype must be SPX_I8");
    SPX_REQUIRE(out_dtype == SPX_I8 || out_dtype == SPX_F32 || out_dtype == SPX_F16,
                "implicit_gemm_fwd_int8: out dtype %d not supported", out_dtype);
    if (d->n_out == 0) return 0;
    SPX_REQUIRE(features && filters && out && scale, "implicit_gemm_fwd_int8: NULL tensor");
    Int8Args q;
    memset(&q, 0, sizeof(q));
    q.g = make_args(d, false);
    q.g.x = features; q.g.w = filters; q.g.y = out; q.g.act = act; q.g.alpha = act_alpha;
    q.out_dtype = out_dtype; q.scale = scale; q.bias_f32 = bias; q.output_add = output_add;
    q.output_add_scale = output_add_scale;
    bool tc_ok = !force_simt() && tc_gather_gemm_int8_supported(q);
    if (force_tc() && !tc_ok) {
        set_error("SPX_FORCE_TC=1 but the tcgen05 int8 path does not support C %d, K %d", d->c_in, d->c_out);
        return 3;
    }
    if (tc_ok) { set_family(2); return tc_gather_gemm_int8(q, (cudaStream_t)stream); }
    set_family(1);
    return simt_gather_gemm_int8(q, (cudaStream_t)stream);
}
