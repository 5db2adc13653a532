// tcgen05 / TMEM / TMA masked implicit GEMM: forward and input-gradient (and int8 forward).
//
// One persistent CTA per SM walks 128-row output tiles (rows in mask_argsort order).  For each
// kernel offset whose bit is set in the OR of the tile's neighbour masks:
//   * 4 producer warps gather the 128 input rows named by the pair table into a swizzled
//     K-major shared-memory tile with 16-byte cp.async (zero-fill for "-1 = no neighbour"),
//   * one lane TMA-loads that offset's KRSC weight slice W[:, k, :] (a strided 2-D box),
//   * one lane issues tcgen05.mma (M = 128, N = out channels, K = 32 bytes per instruction)
//     accumulating the whole offset sum in TMEM,
// and 4 epilogue warps drain the previous tile's accumulator (tcgen05.ld -> bias/act ->
// vectorised row stores) while the next tile is being multiplied (two TMEM accumulators).
//
// Reference being replaced: ConvMain::implicit_gemm2 as called from
// spconv/csrc/sparse/convops.py:2196-2235 (fwd) and :2394-2419 (dgrad); the kernel itself lives
// in the un-vendored cumm package (Ampere mma.sync at best, spconv/core.py:502-830).
//
// The input-gradient pass is the same kernel: A = gathered dout rows, contraction over the
// out channels, and the SAME weight box is consumed as an MN-major B operand (UMMA
// descriptor b_major = MN), so no transposed copy of the filter is ever made.
#include "gemm.cuh"
#include <cuda.h>
#include <mutex>

namespace spx {

constexpr int TC_TILE_M = 128;
constexpr int TC_THREADS = 288;          // warps 0-3 epilogue | 4-7 gather producers | 8 MMA issuer
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_SMEM_BUDGET = 200 * 1024;

struct TcParams {
    // gathered operand A
    const uint8_t *x;
    int xb;                 // bytes per gathered row (contraction length in bytes)
    int span_a;             // swizzle span of A sub-tiles: min(128, xb)
    int ksteps;             // xb / 32
    // weight operand B (TMA box = [b_rows x span_b bytes] per sub-tile)
    int span_b, b_subtiles, b_rows, b_sub_bytes, b_bytes, b_mn_major, umma_k_rows;
    int w_inner_elems;      // c_in (elements between consecutive offsets along the TMA inner dim)
    int span_b_elems;       // span_b / elem bytes
    int n;                  // UMMA N (output channels of this pass)
    uint32_t idesc;
    int stages, a_stage_bytes, stage_bytes;
    uint32_t tmem_cols;
    // rows
    int64_t rows;
    const int32_t *pair;
    int64_t pair_stride;
    const uint32_t *mask;
    const int32_t *argsort;
    int kv, words, reverse;
    // epilogue
    void *y;
    int out_dtype;          // spx_dtype of y
    int epi_mode;           // 0: float bias+act, 1: int8 quantised inference
    const void *bias;       // dtype of y (mode 0)
    int act; float alpha;
    const float *scale, *bias_f32;
    const int8_t *output_add;
    float output_add_scale;
    uint32_t *mask_out;
};

// iterate set bits of a <=128-bit tile mask in ascending order
struct BitIter {
    uint32_t m[4];
    int w;
    __device__ __forceinline__ int next() {
        while (w < 4) {
            if (m[w]) {
                int b = __ffs(m[w]) - 1;
                m[w] &= m[w] - 1;
                return w * 32 + b;
            }
            ++w;
        }
        return -1;
    }
};

// OR of the visiting-order masks of rows [base, base+128); all lanes get the result
__device__ __forceinline__ void tile_mask_or(const uint32_t *__restrict__ mask, int64_t base, int64_t rows, int words,
                                             int kv, int lane, uint32_t (&out)[4]) {
    bool any = false;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        uint32_t m = 0;
        if (w < words) {
            if (mask) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    int64_t row = base + r * 32 + lane;
                    if (row < rows) m |= __ldg(mask + row * words + w);
                }
                m = __reduce_or_sync(0xffffffffu, m);
            } else {
                int hi = kv - 32 * w;
                m = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
            }
        }
        out[w] = m;
        any = any || m != 0;
    }
    if (!any) out[0] = 1u;   // keep the pipeline uniform: offset 0 with all rows "-1" -> zeros
}

template <int KIND>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gather_gemm_kernel(const __grid_constant__ CUtensorMap tmap_w, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // dynamic smem base is only guaranteed 16-byte aligned: align manually to 1024
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
    uint8_t *smem = smem_raw + pad;
    const uint32_t smem_base = raw_addr + pad;

    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)p.stages * p.stage_bytes);
    uint64_t *full = bars;                         // [stages]  producers + TMA -> MMA
    uint64_t *empty = bars + TC_MAX_STAGES;        // [stages]  MMA -> producers
    uint64_t *tmem_full = bars + 2 * TC_MAX_STAGES;      // [2] MMA -> epilogue
    uint64_t *tmem_empty = bars + 2 * TC_MAX_STAGES + 2; // [2] epilogue -> MMA
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(bars + 2 * TC_MAX_STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t num_tiles = (p.rows + TC_TILE_M - 1) / TC_TILE_M;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 128 + 1);    // 128 cp.async arrivals + 1 expect_tx arrival
            mbar_init(&empty[s], 1);         // one tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 4);    // one arrival per epilogue warp
        }
        mbar_fence_init();
        tma_prefetch_desc(&tmap_w);
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_ptr_smem)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp >= 4 && warp < 8) {
        // ================================================= gather producers
        const int pw = warp - 4;
        const int cpr = p.xb >> 4;                 // 16-byte chunks per row
        int stage = 0; uint32_t phase = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int64_t base = tile * TC_TILE_M;
            uint32_t tm[4];
            tile_mask_or(p.mask, base, p.rows, p.words, p.kv, lane, tm);
            const int64_t my_row = base + pw * 32 + lane;
            int32_t src_row = -1;
            if (my_row < p.rows) src_row = p.argsort ? __ldg(p.argsort + my_row) : (int32_t)my_row;
            BitIter it{{tm[0], tm[1], tm[2], tm[3]}, 0};
            int k = it.next();
            int32_t idx = (src_row >= 0) ? __ldg(p.pair + (int64_t)k * p.pair_stride + src_row) : -1;
            while (k >= 0) {
                const int kn = it.next();
                int32_t idx_n = -1;
                if (kn >= 0 && src_row >= 0) idx_n = __ldg(p.pair + (int64_t)kn * p.pair_stride + src_row);
                mbar_wait(&empty[stage], phase ^ 1u);
                const uint32_t a_stage = smem_base + (uint32_t)stage * p.stage_bytes;
                // 32 rows x cpr chunks per warp; consecutive lanes take consecutive 16-byte chunks
                for (int itc = 0; itc < cpr; ++itc) {
                    const int flat = itc * 32 + lane;
                    const int r = flat / cpr;
                    const int ch = flat - r * cpr;
                    const int32_t ridx = __shfl_sync(0xffffffffu, idx, r);
                    const uint32_t byte_in_row = (uint32_t)ch << 4;
                    const uint32_t sub = byte_in_row / (uint32_t)p.span_a;
                    const uint32_t within = byte_in_row - sub * p.span_a;
                    const uint32_t row_in_tile = (uint32_t)(pw * 32 + r);
                    const uint32_t dst = a_stage + sub * (uint32_t)(TC_TILE_M * p.span_a) +
                                         swizzle_offset(row_in_tile * p.span_a + within, p.span_a);
                    const uint8_t *src = p.x + (ridx >= 0 ? (int64_t)ridx * p.xb + byte_in_row : 0);
                    cp_async_16(dst, src, ridx >= 0 ? 16u : 0u);
                }
                cp_async_mbar_arrive_noinc(&full[stage]);
                if (pw == 0 && lane == 0) {
                    const int kw = p.reverse ? p.kv - 1 - k : k;
                    mbar_arrive_expect_tx(&full[stage], (uint32_t)p.b_bytes);
                    const uint32_t b_stage = a_stage + p.a_stage_bytes;
                    for (int sb = 0; sb < p.b_subtiles; ++sb)
                        tma_load_2d(b_stage + sb * p.b_sub_bytes, &tmap_w, &full[stage],
                                    kw * p.w_inner_elems + sb * p.span_b_elems, 0);
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                k = kn; idx = idx_n;
            }
        }
    } else if (warp == 8) {
        // ================================================= MMA issuer
        int stage = 0; uint32_t phase = 0;
        int64_t local = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
            const int64_t base = tile * TC_TILE_M;
            uint32_t tm[4];
            tile_mask_or(p.mask, base, p.rows, p.words, p.kv, lane, tm);
            if (p.mask_out && lane < p.words) p.mask_out[tile * p.words + lane] = tm[lane];
            const int acc = (int)(local & 1);
            const uint32_t acc_phase = (uint32_t)((local >> 1) & 1);
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.n);
            BitIter it{{tm[0], tm[1], tm[2], tm[3]}, 0};
            bool first = true;
            for (int k = it.next(); k >= 0; k = it.next()) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                fence_proxy_async_smem();
                if (lane == 0) {
                    const uint32_t a_stage = smem_base + (uint32_t)stage * p.stage_bytes;
                    const uint32_t b_stage = a_stage + p.a_stage_bytes;
                    for (int j = 0; j < p.ksteps; ++j) {
                        const uint32_t kb = (uint32_t)j * 32u;
                        const uint32_t a_sub = kb / (uint32_t)p.span_a;
                        const uint32_t a_off = kb - a_sub * p.span_a;
                        const uint64_t a_desc = make_smem_desc(a_stage + a_sub * (uint32_t)(TC_TILE_M * p.span_a) + a_off,
                                                               16u, 8u * p.span_a, p.span_a);
                        uint64_t b_desc;
                        if (!p.b_mn_major) {
                            const uint32_t b_sub = kb / (uint32_t)p.span_b;
                            const uint32_t b_off = kb - b_sub * p.span_b;
                            b_desc = make_smem_desc(b_stage + b_sub * p.b_sub_bytes + b_off, 16u, 8u * p.span_b, p.span_b);
                        } else {
                            b_desc = make_smem_desc(b_stage + (uint32_t)j * p.umma_k_rows * p.span_b,
                                                    (uint32_t)p.b_sub_bytes, 8u * p.span_b, p.span_b);
                        }
                        umma_ss<KIND>(d_tmem, a_desc, b_desc, p.idesc, (first && j == 0) ? 0u : 1u);
                    }
                    tc_commit(&empty[stage]);        // frees the smem stage when these MMAs retire
                }
                __syncwarp();
                first = false;
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
            if (lane == 0) tc_commit(&tmem_full[acc]);   // accumulator complete
            __syncwarp();
        }
    } else {
        // ================================================= epilogue (warps 0-3 <-> TMEM lanes 32w..32w+31)
        int64_t local = 0;
        const int ybytes = dtype_bytes(p.out_dtype);
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
            const int64_t base = tile * TC_TILE_M;
            const int acc = (int)(local & 1);
            const uint32_t acc_phase = (uint32_t)((local >> 1) & 1);
            const int64_t my_row = base + warp * 32 + lane;
            int64_t dst_row = -1;
            if (my_row < p.rows) dst_row = p.argsort ? (int64_t)__ldg(p.argsort + my_row) : my_row;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * p.n);
            for (int n0 = 0; n0 < p.n; n0 += 16) {
                uint32_t v[16];
                tmem_ld_32x32b_x16(t_row + (uint32_t)n0, v);
                tc_wait_ld();
                if (dst_row >= 0) {
                    float f[16];
                    if (p.epi_mode == 0) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float val = __uint_as_float(v[j]);
                            if (p.bias) {
                                if (p.out_dtype == SPX_F32) val += ((const float *)p.bias)[n0 + j];
                                else if (p.out_dtype == SPX_F16) val += __half2float(((const __half *)p.bias)[n0 + j]);
                                else val += __bfloat162float(((const __nv_bfloat16 *)p.bias)[n0 + j]);
                            }
                            f[j] = apply_act(val, p.act, p.alpha);
                        }
                    } else {
                        // int8 inference: test/test_all_algo.py:272-287
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float val = (float)(int32_t)v[j] * p.scale[n0 + j] + (p.bias_f32 ? p.bias_f32[n0 + j] : 0.f);
                            if (p.output_add) val += (float)p.output_add[dst_row * p.n + n0 + j] * p.output_add_scale;
                            f[j] = apply_act(val, p.act, p.alpha);
                        }
                    }
                    uint8_t *dst = (uint8_t *)p.y + (dst_row * p.n + n0) * ybytes;
                    if (p.out_dtype == SPX_F32) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4)
                            *reinterpret_cast<float4 *>(dst + j * 4) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                    } else if (p.out_dtype == SPX_F16) {
                        uint32_t h[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            __half2 t = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
                            h[j] = *reinterpret_cast<uint32_t *>(&t);
                        }
                        *reinterpret_cast<uint4 *>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<uint4 *>(dst + 16) = make_uint4(h[4], h[5], h[6], h[7]);
                    } else if (p.out_dtype == SPX_BF16) {
                        uint32_t h[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                            h[j] = *reinterpret_cast<uint32_t *>(&t);
                        }
                        *reinterpret_cast<uint4 *>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<uint4 *>(dst + 16) = make_uint4(h[4], h[5], h[6], h[7]);
                    } else {   // SPX_I8: clip(rint(.)) -- numpy round-half-even
                        uint32_t q[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t w = 0;
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                float r = fminf(fmaxf(rintf(f[4 * j + b]), -128.f), 127.f);
                                w |= ((uint32_t)(uint8_t)(int8_t)(int)r) << (8 * b);
                            }
                            q[j] = w;
                        }
                        *reinterpret_cast<uint4 *>(dst) = make_uint4(q[0], q[1], q[2], q[3]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

// 2-D view of the KRSC filter: inner = kv*c_in elements, outer = c_out rows
int make_weight_tmap(CUtensorMap *tm, const void *w, int dtype, int kv, int c_in, int c_out, int span_bytes) {
    EncodeTiledFn fn = get_encode_fn();
    SPX_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available (driver too old?)");
    const int e = dtype_bytes(dtype);
    CUtensorMapDataType dt = dtype == SPX_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                           : dtype == SPX_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                           : dtype == SPX_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                              : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    cuuint64_t dims[2] = {(cuuint64_t)kv * c_in, (cuuint64_t)c_out};
    cuuint64_t strides[1] = {(cuuint64_t)kv * c_in * e};
    cuuint32_t box[2] = {(cuuint32_t)(span_bytes / e), (cuuint32_t)c_out};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = span_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : span_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                             : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = fn(tm, dt, 2, const_cast<void *>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SPX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (kv=%d C=%d K=%d span=%d)", (int)r, kv, c_in,
                c_out, span_bytes);
    return 0;
}

static bool span_ok(int bytes) { return bytes == 32 || bytes == 64 || (bytes >= 128 && bytes % 128 == 0); }

static bool tc_shape_ok(int dtype, int c_in, int c_out, int transpose_w) {
    const int e = dtype_bytes(dtype);
    if (e == 0) return false;
    if (c_in > 256 || c_out > 256) return false;
    if (c_in % 16 || c_out % 16) return false;
    if (!span_ok(c_in * e) || !span_ok(c_out * e)) return false;
    const int cy = transpose_w ? c_in : c_out;
    if (cy % 16 || cy > 256) return false;
    const int xb = (transpose_w ? c_out : c_in) * e;
    if (dtype == SPX_I8 && (c_in % 32 || c_out % 32)) return false;   // docs/INT8_GUIDE.md:10
    size_t stage = align_up((size_t)TC_TILE_M * xb, 1024) + align_up((size_t)c_in * c_out * e, 1024);
    if (TC_SMEM_BUDGET / stage < 2) return false;
    return true;
}

bool tc_gather_gemm_supported(const GatherGemmArgs &a) {
    if (a.dtype == SPX_I8) return false;
    // tf32 MN-major operands need the SWIZZLE_128B_BASE32B atom (not tiled here yet): fp32 dgrad -> SIMT
    if (a.dtype == SPX_F32 && a.transpose_w) return false;
    if (((size_t)a.kv * a.c_in * dtype_bytes(a.dtype)) % 16) return false;
    return tc_shape_ok(a.dtype, a.c_in, a.c_out, a.transpose_w);
}
bool tc_gather_gemm_int8_supported(const Int8Args &q) {
    if (((size_t)q.g.kv * q.g.c_in) % 16) return false;
    return tc_shape_ok(SPX_I8, q.g.c_in, q.g.c_out, 0);
}

static int fill_params(const GatherGemmArgs &a, TcParams &p) {
    memset(&p, 0, sizeof(p));
    const int e = dtype_bytes(a.dtype);
    const int cx = a.cx(), cy = a.cy();
    p.x = (const uint8_t *)a.x;
    p.xb = cx * e;
    p.span_a = p.xb < 128 ? p.xb : 128;
    p.ksteps = p.xb / 32;
    const int wb = a.c_in * e;                 // inner (contiguous) bytes of one weight slice row
    p.span_b = wb < 128 ? wb : 128;
    p.b_subtiles = wb / p.span_b;
    p.b_rows = a.c_out;
    p.b_sub_bytes = a.c_out * p.span_b;
    p.b_bytes = a.c_out * wb;
    p.b_mn_major = a.transpose_w;
    p.umma_k_rows = 32 / e;
    p.w_inner_elems = a.c_in;
    p.span_b_elems = p.span_b / e;
    p.n = cy;
    int c_fmt = a.dtype == SPX_I8 ? 2 : 1;
    int ab_fmt = a.dtype == SPX_F16 ? 0 : a.dtype == SPX_BF16 ? 1 : a.dtype == SPX_F32 ? 2 : 1;
    p.idesc = make_idesc(c_fmt, ab_fmt, ab_fmt, 0, a.transpose_w, TC_TILE_M, cy);
    p.a_stage_bytes = (int)align_up((size_t)TC_TILE_M * p.xb, 1024);
    p.stage_bytes = p.a_stage_bytes + (int)align_up((size_t)p.b_bytes, 1024);
    p.stages = TC_SMEM_BUDGET / p.stage_bytes;
    if (p.stages > TC_MAX_STAGES) p.stages = TC_MAX_STAGES;
    SPX_REQUIRE(p.stages >= 2, "tc_gather_gemm: tile does not fit shared memory (stage %d bytes)", p.stage_bytes);
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * cy)) cols <<= 1;
    p.tmem_cols = cols;
    p.rows = a.rows; p.pair = a.pair; p.pair_stride = a.pair_stride; p.mask = a.mask; p.argsort = a.argsort;
    p.kv = a.kv; p.words = (a.kv + 31) / 32; p.reverse = a.reverse;
    p.y = a.y; p.out_dtype = a.dtype; p.epi_mode = 0; p.bias = a.bias; p.act = a.act; p.alpha = a.alpha;
    p.mask_out = a.mask_out;
    return 0;
}

template <int KIND>
static int launch_tc(const CUtensorMap &tm, const TcParams &p, cudaStream_t stream) {
    const size_t smem = (size_t)p.stages * p.stage_bytes + 1024 /*align slack*/ + 256 /*barriers*/;
    static thread_local size_t configured[3] = {0, 0, 0};
    if (configured[KIND] < smem) {
        SPX_CHECK_CUDA(cudaFuncSetAttribute(tc_gather_gemm_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(TC_SMEM_BUDGET + 2048)));
        configured[KIND] = TC_SMEM_BUDGET + 2048;
    }
    const int64_t tiles = div_up64(p.rows, TC_TILE_M);
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    tc_gather_gemm_kernel<KIND><<<grid, TC_THREADS, smem, stream>>>(tm, p);
    SPX_CHECK_LAUNCH("tc_gather_gemm_kernel");
    return 0;
}

int tc_gather_gemm(const GatherGemmArgs &a, cudaStream_t stream) {
    TcParams p;
    if (fill_params(a, p)) return 2;
    CUtensorMap tm;
    if (make_weight_tmap(&tm, a.w, a.dtype, a.kv, a.c_in, a.c_out, p.span_b)) return 2;
    if (a.dtype == SPX_F32) return launch_tc<KIND_TF32>(tm, p, stream);
    return launch_tc<KIND_F16>(tm, p, stream);
}

int tc_gather_gemm_int8(const Int8Args &q, cudaStream_t stream) {
    TcParams p;
    if (fill_params(q.g, p)) return 2;
    p.out_dtype = q.out_dtype;
    p.epi_mode = 1;
    p.scale = q.scale; p.bias_f32 = q.bias_f32; p.output_add = q.output_add;
    p.output_add_scale = q.output_add_scale;
    CUtensorMap tm;
    if (make_weight_tmap(&tm, q.g.w, SPX_I8, q.g.kv, q.g.c_in, q.g.c_out, p.span_b)) return 2;
    return launch_tc<KIND_I8>(tm, p, stream);
}

}  // namespace spx
