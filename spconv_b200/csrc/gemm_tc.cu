// tcgen05 / TMEM / TMA masked implicit GEMM: forward and input-gradient (and int8 forward).
//
// Persistent CTAs (two per SM when the tile fits 104 KB of shared memory and 256 TMEM columns)
// work on 128-row output tiles (rows in mask_argsort order).  Tiles are handed out DYNAMICALLY:
// a scheduler warp draws tickets from an atomic counter and takes tiles from the schedule records
// behind the tile table (heaviest first = LPT list scheduling, see gemm.cuh / rulebook.cu); the
// tile's gather indices arrive as ONE bulk async copy (cp.async.bulk) of its block of the tile
// table (spx_build_tile_table), one tile ahead.  For each kernel offset whose bit is set in the
// tile's OR-mask:
//   * 8 producer warps gather the 128 input rows into a swizzled K-major shared-memory tile with
//     16-byte cp.async (zero-fill for "-1 = no neighbour"); all per-lane addressing is hoisted
//     out of the pipeline (compile-time chunk count CPR), the row indices of the next offset are
//     read before the wait for the next free stage;
//   * the TMA warp loads that offset's KRSC weight slice W[:, k, :] (a strided 2-D box);
//   * the MMA warp issues tcgen05.mma (M = 128, N = out channels, K = 32 bytes per instruction;
//     warp-uniform descriptor arithmetic, an elected lane fires) accumulating the whole offset sum
//     in TMEM;
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
#include <atomic>
#include <mutex>
#include <stdlib.h>

namespace spx {

constexpr int TC_TILE_M = 128;
constexpr int TC_PROD_WARPS = 8;         // 16 tile rows per producer warp
constexpr int TC_PROD_THREADS = TC_PROD_WARPS * 32;
constexpr int TC_MMA_WARP = 4 + TC_PROD_WARPS;
constexpr int TC_TMA_WARP = TC_MMA_WARP + 1;
constexpr int TC_SCHED_WARP = TC_TMA_WARP + 1;
constexpr int TC_THREADS = (TC_SCHED_WARP + 1) * 32;  // warps 0-3 epilogue | 4-11 gather producers | 12 MMA issuer | 13 weight TMA | 14 tile scheduler + index copies
constexpr int TC_INFO_DEPTH = 4;         // tile-info ring (scheduler -> every role)
constexpr int TC_INFO_READERS = TC_PROD_WARPS + 1 + 1 + 4;   // producer warps, MMA warp, TMA warp, epilogue warps
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_CTAS_PER_SM = 2;        // max resident CTAs per SM the kernel is compiled for (SPX_TC_CTAS picks 1 or 2)
constexpr int TC_SMEM_BUDGET_2 = 104 * 1024;   // per CTA when two CTAs share an SM
constexpr int TC_SMEM_BUDGET = 200 * 1024;     // per CTA when a tile needs the whole SM

struct TcParams {
    // gathered operand A
    const uint8_t *x;
    int xb;                 // bytes per gathered row (contraction length in bytes) = CPR * 16
    int span_a, lg_span_a;  // swizzle span of A sub-tiles: min(128, xb)
    int a_subtiles;         // xb / span_a
    int q_a;                // k-steps (32 B) per A sub-tile = span_a / 32
    // weight operand B (TMA box = [b_rows x span_b bytes] per sub-tile)
    int span_b, b_subtiles, b_sub_bytes, b_bytes, b_mn_major, q_b, b_kstep16_mn;
    int b_base32;           // 1: MN-major 32-bit B operand in the SWIZZLE_128B_BASE32B layout (tf32 input gradient)
    int w_inner_elems;      // c_in (elements between consecutive offsets along the TMA inner dim)
    int span_b_elems;       // span_b / elem bytes
    int n;                  // UMMA N (output channels of this pass)
    uint32_t idesc;
    int stages, a_stage_bytes, stage_bytes, idx_bytes, ctas_per_sm;
    uint32_t tmem_cols;
    // rows
    int64_t rows;
    const int32_t *tile_table;   // [tiles][kv+1][128]
    const int32_t *sched_rec;    // [tiles][TT_REC_INTS] schedule records, heaviest tile first (gemm.cuh)
    int *sched_state;            // [0] ticket counter, [1] finished CTAs; zero between launches
    const int32_t *argsort;      // destination rows of the epilogue (NULL = identity)
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
    long long *dbg_ts;      // optional [8 roles][2048] clock64 stamps of CTA 0 (SPX_TC_TRACE, perf triage)
    int debug;              // SPX_TC_DEBUG ablation bits (perf triage only): 1 no gather, 2 no MMA, 4 no epilogue, 8 no weight TMA
};

// iterate set bits of a <=128-bit tile mask in ascending order (register-only: no indexed array)
struct BitIter {
    uint32_t m0, m1, m2, m3;
    __device__ __forceinline__ BitIter(const uint32_t (&t)[4]) : m0(t[0]), m1(t[1]), m2(t[2]), m3(t[3]) {}
    __device__ __forceinline__ int next() {
        if (m0) { int b = __ffs(m0) - 1; m0 &= m0 - 1; return b; }
        if (m1) { int b = __ffs(m1) - 1; m1 &= m1 - 1; return 32 + b; }
        if (m2) { int b = __ffs(m2) - 1; m2 &= m2 - 1; return 64 + b; }
        if (m3) { int b = __ffs(m3) - 1; m3 &= m3 - 1; return 96 + b; }
        return -1;
    }
};

// ------------------------------------------------------------------ epilogue, one row per thread
// 16 accumulator columns -> OUT type, 16-byte vector stores.  OUT is an spx_dtype code.
template <int OUT>
__device__ __forceinline__ void store16(uint8_t *dst, const float (&f)[16]) {
    if constexpr (OUT == SPX_F32) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4 *>(dst + j * 4) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    } else if constexpr (OUT == SPX_F16) {
        uint32_t h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            __half2 t = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
            h[j] = *reinterpret_cast<uint32_t *>(&t);
        }
        *reinterpret_cast<uint4 *>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(dst + 16) = make_uint4(h[4], h[5], h[6], h[7]);
    } else if constexpr (OUT == SPX_BF16) {
        uint32_t h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
            h[j] = *reinterpret_cast<uint32_t *>(&t);
        }
        *reinterpret_cast<uint4 *>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(dst + 16) = make_uint4(h[4], h[5], h[6], h[7]);
    } else {   // int8: clip(rint(.)) -- numpy round-half-even
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

template <int OUT> struct OutElem { static constexpr int bytes = OUT == SPX_F32 ? 4 : (OUT == SPX_I8 ? 1 : 2); };

template <int OUT>
__device__ __forceinline__ float load_bias(const void *bias, int j) {
    if constexpr (OUT == SPX_F32) return ((const float *)bias)[j];
    else if constexpr (OUT == SPX_F16) return __half2float(((const __half *)bias)[j]);
    else return __bfloat162float(((const __nv_bfloat16 *)bias)[j]);
}

// drain one accumulator (this warp's 32 TMEM lanes x n columns) to global memory
template <int OUT, bool INT8_MODE>
__device__ __forceinline__ void epilogue_tile(const TcParams &p, uint32_t t_row, int64_t dst_row) {
    constexpr int EB = OutElem<OUT>::bytes;
    uint8_t *row_ptr = (uint8_t *)p.y + dst_row * (int64_t)p.n * EB;
    const bool has_bias = INT8_MODE ? (p.bias_f32 != nullptr) : (p.bias != nullptr);
    for (int n0 = 0; n0 < p.n; n0 += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_row + (uint32_t)n0, v);
        tc_wait_ld();
        if (dst_row < 0) continue;
        float f[16];
        if constexpr (!INT8_MODE) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
            if (has_bias) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] += load_bias<OUT == SPX_I8 ? SPX_F32 : OUT>(p.bias, n0 + j);
            }
        } else {
            // int8 inference: y = acc * scale[k] + bias[k] (+ add * add_scale)   test/test_all_algo.py:272-287
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = (float)(int32_t)v[j] * __ldg(p.scale + n0 + j);
            if (has_bias) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] += __ldg(p.bias_f32 + n0 + j);
            }
            if (p.output_add) {
                const int8_t *ap = p.output_add + dst_row * (int64_t)p.n + n0;
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] += (float)ap[j] * p.output_add_scale;
            }
        }
        if (p.act != SPX_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = apply_act(f[j], p.act, p.alpha);
        }
        store16<OUT>(row_ptr + n0 * EB, f);
    }
}

// per-CTA wall-clock spans (ns, %globaltimer): rows 6..7 of the trace buffer hold [cta][4] =
// {kernel entry, prologue done, role loops done, exit}
__device__ __forceinline__ long long global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TC_SPAN(i) do { if (p.dbg_ts && threadIdx.x == 0 && blockIdx.x < 1024) p.dbg_ts[6 * 2048 + blockIdx.x * 4 + (i)] = global_ns(); } while (0)
#define TC_STAMP(role, n) do { if (p.dbg_ts && blockIdx.x == 0 && lane == 0 && (n) < 2048) p.dbg_ts[(role) * 2048 + (n)] = clock64(); } while (0)

template <int KIND, int CPR>
__global__ void __launch_bounds__(TC_THREADS, TC_CTAS_PER_SM)
tc_gather_gemm_kernel(const __grid_constant__ CUtensorMap tmap_w, const TcParams p) {
    constexpr int LG_CPR = CPR == 2 ? 1 : CPR == 4 ? 2 : CPR == 8 ? 3 : CPR == 16 ? 4 : 5;
    constexpr int RPI = 32 / CPR;            // rows covered by one warp-wide cp.async instruction
    constexpr int XB = CPR * 16;             // bytes per gathered row
    constexpr int SPAN_A = XB < 128 ? XB : 128;
    constexpr int LG_SPAN_A = SPAN_A == 128 ? 7 : (SPAN_A == 64 ? 6 : 5);
    constexpr int A_SUB_BYTES = TC_TILE_M * SPAN_A;
    constexpr int ROWS_PW = TC_TILE_M / TC_PROD_WARPS;          // tile rows per producer warp
    constexpr int ITERS = ROWS_PW / RPI > 0 ? ROWS_PW / RPI : 1;  // cp.async per thread per stage
    static_assert(ROWS_PW * CPR >= 32, "a producer warp must cover at least one full cp.async instruction");

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    TC_SPAN(0);
    // dynamic smem base is only guaranteed 16-byte aligned: align manually to 1024
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
    uint8_t *smem = smem_raw + pad;
    const uint32_t smem_base = raw_addr + pad;

    const uint32_t idx_off = (uint32_t)p.stages * p.stage_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + idx_off + 2u * p.idx_bytes);
    uint64_t *full = bars;                                // [stages]  producers + TMA -> MMA
    uint64_t *empty = bars + TC_MAX_STAGES;               // [stages]  MMA -> producers
    uint64_t *tmem_full = bars + 2 * TC_MAX_STAGES;       // [2] MMA -> epilogue
    uint64_t *tmem_empty = bars + 2 * TC_MAX_STAGES + 2;  // [2] epilogue -> MMA
    uint64_t *idx_full = bars + 2 * TC_MAX_STAGES + 4;    // [2] bulk copy -> producers
    uint64_t *idx_empty = bars + 2 * TC_MAX_STAGES + 6;   // [2] producers -> bulk copy
    uint64_t *info_full = bars + 2 * TC_MAX_STAGES + 8;                    // [TC_INFO_DEPTH] scheduler -> roles
    uint64_t *info_empty = bars + 2 * TC_MAX_STAGES + 8 + TC_INFO_DEPTH;   // [TC_INFO_DEPTH] roles -> scheduler
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(bars + 2 * TC_MAX_STAGES + 8 + 2 * TC_INFO_DEPTH);
    int32_t *info = reinterpret_cast<int32_t *>(tmem_ptr_smem + 2);        // [TC_INFO_DEPTH][8]: {tile, mask[4]}

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t num_tiles = (p.rows + TC_TILE_M - 1) / TC_TILE_M;

    // Tiles are handed out dynamically: the scheduler warp draws a ticket, looks the tile and its
    // offset set up in the schedule records (heaviest first => LPT list scheduling) and publishes
    // them through the info ring; every other role consumes ring entries in order.  Entry n uses
    // ring slot n % TC_INFO_DEPTH and index buffer n % 2; tile < 0 ends the stream.
    auto read_info = [&](int n, uint32_t (&tm)[4]) -> int {
        const int e = n & (TC_INFO_DEPTH - 1);
        mbar_wait(&info_full[e], (uint32_t)((n / TC_INFO_DEPTH) & 1));
        const volatile int32_t *r = info + e * 8;
        const int tile = r[0];
        tm[0] = (uint32_t)r[1]; tm[1] = (uint32_t)r[2]; tm[2] = (uint32_t)r[3]; tm[3] = (uint32_t)r[4];
        __syncwarp();
        if (lane == 0) mbar_arrive(&info_empty[e]);
        return tile;
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], TC_PROD_THREADS + 1);    // cp.async arrivals + 1 expect_tx arrival
            mbar_init(&empty[s], 1);         // one tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 4);    // one arrival per epilogue warp
            mbar_init(&idx_full[a], 1);      // expect_tx arrival of the bulk copy
            mbar_init(&idx_empty[a], TC_PROD_WARPS);     // one release per producer warp
        }
        for (int e = 0; e < TC_INFO_DEPTH; ++e) {
            mbar_init(&info_full[e], 1);
            mbar_init(&info_empty[e], TC_INFO_READERS);  // one release per reading warp
        }
        mbar_fence_init();
        tma_prefetch_desc(&tmap_w);
    }
    if (warp == TC_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_ptr_smem)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (warp == 0) TC_STAMP(3, 0);
    TC_SPAN(1);

    if (warp >= 4 && warp < TC_MMA_WARP) {
        // ================================================= gather producers
        const int pw = warp - 4;
        // per-lane constants: chunk ch of rows r0 + itc*RPI (itc = 0..CPR-1) of this warp's 32 rows
        const int r0 = lane >> LG_CPR;
        const uint32_t byte_in_row = (uint32_t)(lane & (CPR - 1)) << 4;
        const uint8_t *x_lane = p.x + byte_in_row;
        uint32_t dst_off[ITERS];
#pragma unroll
        for (int itc = 0; itc < ITERS; ++itc) {
            const uint32_t row_in_tile = (uint32_t)(pw * ROWS_PW + r0 + itc * RPI);
            const uint32_t sub = byte_in_row >> LG_SPAN_A;
            const uint32_t within = byte_in_row & (uint32_t)(SPAN_A - 1);
            dst_off[itc] = sub * (uint32_t)A_SUB_BYTES + swizzle_offset((row_in_tile << LG_SPAN_A) + within, SPAN_A);
        }
        int stage = 0; uint32_t phase = 0;
        int nstamp = 0;
        for (int n = 0;; ++n) {
            uint32_t tm[4];
            if (read_info(n, tm) < 0) break;
            const int buf = n & 1;
            mbar_wait(&idx_full[buf], (uint32_t)((n >> 1) & 1));
            const int32_t *idx_lane = reinterpret_cast<const int32_t *>(smem + idx_off + (size_t)buf * p.idx_bytes) +
                                      pw * ROWS_PW + r0;
            BitIter it(tm);
            // the index loads of offset k+1 are issued before the wait on the next free stage
            int k = it.next();
            int32_t ridx[ITERS];
            if (k >= 0) {
#pragma unroll
                for (int itc = 0; itc < ITERS; ++itc) ridx[itc] = idx_lane[k * 128 + itc * RPI];
            }
            while (k >= 0) {
                mbar_wait(&empty[stage], phase ^ 1u);
                if (pw == 0) TC_STAMP(0, 2 * nstamp);
                const uint32_t a_stage = smem_base + (uint32_t)stage * p.stage_bytes;
                if (!(p.debug & 1)) {
#pragma unroll
                    for (int itc = 0; itc < ITERS; ++itc)
                        cp_async_16(a_stage + dst_off[itc], x_lane + (int64_t)max(ridx[itc], 0) * XB,
                                    ridx[itc] >= 0 ? 16u : 0u);
                }
                cp_async_mbar_arrive_noinc(&full[stage]);
                if (pw == 0) TC_STAMP(0, 2 * nstamp + 1);
                ++nstamp;
                k = it.next();
                if (k >= 0) {
#pragma unroll
                    for (int itc = 0; itc < ITERS; ++itc) ridx[itc] = idx_lane[k * 128 + itc * RPI];
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&idx_empty[buf]);
        }
    } else if (warp == TC_TMA_WARP) {
        // ================================================= TMA warp: one weight box per (tile, offset) stage
        // all 32 lanes walk the loops together (the CTA-wide barrier at the end must be reached
        // convergently); lane 0 issues the copies
        int stage = 0; uint32_t phase = 0;
        for (int n = 0;; ++n) {
            uint32_t tm[4];
            if (read_info(n, tm) < 0) break;
            BitIter it(tm);
            for (int k = it.next(); k >= 0; k = it.next()) {
                mbar_wait(&empty[stage], phase ^ 1u);
                if (lane == 0) {
                    const int kw = p.reverse ? p.kv - 1 - k : k;
                    if (p.debug & 8) {
                        mbar_arrive(&full[stage]);
                    } else {
                        mbar_arrive_expect_tx(&full[stage], (uint32_t)p.b_bytes);
                        const uint32_t b_stage = smem_base + (uint32_t)stage * p.stage_bytes + p.a_stage_bytes;
                        for (int sb = 0; sb < p.b_subtiles; ++sb)
                            tma_load_2d(b_stage + sb * p.b_sub_bytes, &tmap_w, &full[stage],
                                        kw * p.w_inner_elems + sb * p.span_b_elems, 0);
                    }
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == TC_SCHED_WARP) {
        // ================================================= tile scheduler + gather-index blocks
        // A free index buffer is the permission to look ONE tile ahead: the ticket for entry n is
        // drawn only when the buffer of entry n-2 has been released, so a CTA never hoards tiles
        // and the tail of the kernel is at most one (light) tile long.
        const uint32_t blk_bytes = (uint32_t)(p.kv + 1) * 512u;
        for (int n = 0;; ++n) {
            const int b = n & 1, e = n & (TC_INFO_DEPTH - 1);
            mbar_wait(&idx_empty[b], (uint32_t)(((n >> 1) & 1) ^ 1));
            mbar_wait(&info_empty[e], (uint32_t)(((n / TC_INFO_DEPTH) & 1) ^ 1));
            int tile = -1;
            if (lane == 0) {
                const int ticket = atomicAdd(p.sched_state, 1);
                int4 r0 = make_int4(-1, 1, 0, 0), r1 = make_int4(0, 0, 0, 0);
                if ((int64_t)ticket < num_tiles) {
                    const int4 *rp = reinterpret_cast<const int4 *>(p.sched_rec + (int64_t)ticket * TT_REC_INTS);
                    r0 = __ldg(rp); r1 = __ldg(rp + 1);
                }
                tile = r0.x;
                int32_t *dst = info + e * 8;
                dst[0] = r0.x; dst[1] = r0.y; dst[2] = r0.z; dst[3] = r0.w; dst[4] = r1.x;
                if (tile >= 0) {
                    mbar_arrive_expect_tx(&idx_full[b], blk_bytes);
                    bulk_copy_g2s(smem_base + idx_off + (uint32_t)b * p.idx_bytes,
                                  p.tile_table + (int64_t)tile * (p.kv + 1) * 128, blk_bytes, &idx_full[b]);
                }
                mbar_arrive(&info_full[e]);          // release: publishes the record written above
            }
            tile = __shfl_sync(0xffffffffu, tile, 0);
            if (tile < 0) break;
        }
        // the last CTA to run dry leaves the scheduler state zeroed for the next launch
        if (lane == 0) {
            __threadfence();
            const int done = atomicAdd(p.sched_state + 1, 1);
            if (done == (int)gridDim.x - 1) {
                p.sched_state[0] = 0;
                p.sched_state[1] = 0;
                __threadfence();
            }
        }
        __syncwarp();
    } else if (warp == TC_MMA_WARP) {
        // ================================================= MMA issuer
        int nstamp = 0;
        int stage = 0; uint32_t phase = 0;
        const uint64_t a_hi = smem_desc_hi(16u, 8u * SPAN_A, SPAN_A);
        // MN-major B: LBO = distance between 128-byte column atoms, SBO = one swizzle atom of K rows
        // (8 rows for the 16-byte-base swizzles, 4 rows for the 32-byte-base layout of 32-bit operands)
        const uint64_t b_hi = p.b_base32 ? smem_desc_hi_layout((uint32_t)p.b_sub_bytes, 4u * p.span_b, 1ull)
                              : p.b_mn_major ? smem_desc_hi((uint32_t)p.b_sub_bytes, 8u * p.span_b, p.span_b)
                                             : smem_desc_hi(16u, 8u * p.span_b, p.span_b);
        const uint32_t b_sub16 = (uint32_t)p.b_sub_bytes >> 4;
        for (int local = 0;; ++local) {
            uint32_t tm[4];
            if (read_info(local, tm) < 0) break;
            const int acc = (int)(local & 1);
            const uint32_t acc_phase = (uint32_t)((local >> 1) & 1);
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.n);
            BitIter it(tm);
            uint32_t accumulate = 0;
            for (int k = it.next(); k >= 0; k = it.next()) {
                mbar_wait(&full[stage], phase);
                TC_STAMP(1, 2 * nstamp);
                tc_fence_after();
                TC_STAMP(4, 2 * nstamp);     // no fence.proxy.async: cp.async data is published by the mbarrier it
                                      // arrives on (same hand-off as CUTLASS' sm100 cp.async mainloop)
                {
                    // warp-uniform issue: every lane runs the same arithmetic, an elected lane fires
                    const uint32_t a16 = (smem_base + (uint32_t)stage * p.stage_bytes) >> 4;
                    const uint32_t b16 = a16 + ((uint32_t)p.a_stage_bytes >> 4);
                    if (p.debug & 2) {
                        // ablation: no tensor work
                    } else if (!p.b_mn_major) {
                        for (int sub = 0; sub < p.a_subtiles; ++sub) {
                            const uint32_t as = a16 + (uint32_t)sub * (uint32_t)(A_SUB_BYTES >> 4);
                            const uint32_t bs = b16 + (uint32_t)sub * b_sub16;
                            for (int jr = 0; jr < p.q_a; ++jr) {
                                umma_ss_elect<KIND>(d_tmem, a_hi | (uint64_t)((as + 2u * jr) & 0x3FFFu),
                                                    b_hi | (uint64_t)((bs + 2u * jr) & 0x3FFFu), p.idesc, accumulate);
                                accumulate = 1u;
                            }
                        }
                    } else {
                        int j = 0;
                        for (int sub = 0; sub < p.a_subtiles; ++sub) {
                            const uint32_t as = a16 + (uint32_t)sub * (uint32_t)(A_SUB_BYTES >> 4);
                            for (int jr = 0; jr < p.q_a; ++jr, ++j) {
                                umma_ss_elect<KIND>(d_tmem, a_hi | (uint64_t)((as + 2u * jr) & 0x3FFFu),
                                                    b_hi | (uint64_t)((b16 + (uint32_t)j * p.b_kstep16_mn) & 0x3FFFu),
                                                    p.idesc, accumulate);
                                accumulate = 1u;
                            }
                        }
                    }
                    TC_STAMP(4, 2 * nstamp + 1);
                    tc_commit_elect(&empty[stage]);        // frees the smem stage when these MMAs retire
                    TC_STAMP(5, 2 * nstamp);
                }
                __syncwarp();
                TC_STAMP(1, 2 * nstamp + 1);
                ++nstamp;
                accumulate = 1u;
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
            tc_commit_elect(&tmem_full[acc]);   // accumulator complete
            __syncwarp();
        }
    } else {
        // ================================================= epilogue (warps 0-3 <-> TMEM lanes 32w..32w+31)
        for (int local = 0;; ++local) {
            uint32_t tm_unused[4];
            const int tile = read_info(local, tm_unused);
            if (tile < 0) break;
            const int64_t base = (int64_t)tile * TC_TILE_M;
            const int acc = (int)(local & 1);
            const uint32_t acc_phase = (uint32_t)((local >> 1) & 1);
            const int64_t my_row = base + warp * 32 + lane;
            int64_t dst_row = -1;
            if (my_row < p.rows) dst_row = p.argsort ? (int64_t)__ldg(p.argsort + my_row) : my_row;
            mbar_wait(&tmem_full[acc], acc_phase);
            if (warp == 0) TC_STAMP(2, 2 * (int)local);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * p.n);
            if (p.debug & 4) {
                // ablation: accumulator is released untouched
            } else if constexpr (KIND == KIND_F16) {
                if (p.out_dtype == SPX_F16) epilogue_tile<SPX_F16, false>(p, t_row, dst_row);
                else epilogue_tile<SPX_BF16, false>(p, t_row, dst_row);
            } else if constexpr (KIND == KIND_TF32) {
                epilogue_tile<SPX_F32, false>(p, t_row, dst_row);
            } else {
                if (p.out_dtype == SPX_I8) epilogue_tile<SPX_I8, true>(p, t_row, dst_row);
                else if (p.out_dtype == SPX_F32) epilogue_tile<SPX_F32, true>(p, t_row, dst_row);
                else epilogue_tile<SPX_F16, true>(p, t_row, dst_row);
            }
            tc_fence_before();
            __syncwarp();
            if (warp == 0) TC_STAMP(2, 2 * (int)local + 1);
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) TC_STAMP(3, 1);
    TC_SPAN(2);
    if (warp == TC_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
    TC_SPAN(3);
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
int make_weight_tmap(CUtensorMap *tm, const void *w, int dtype, int kv, int c_in, int c_out, int span_bytes,
                     bool base32 = false) {
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
    CUtensorMapSwizzle sw = base32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                          : span_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : span_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                             : CU_TENSOR_MAP_SWIZZLE_32B;
    CUresult r = fn(tm, dt, 2, const_cast<void *>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SPX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d (kv=%d C=%d K=%d span=%d)", (int)r, kv, c_in,
                c_out, span_bytes);
    return 0;
}

// row bytes the kernels tile: one swizzle span (32/64/128 B) or a power-of-two multiple of 128 B
static bool span_ok(int bytes) { return bytes == 32 || bytes == 64 || bytes == 128 || bytes == 256 || bytes == 512; }

static bool tc_shape_ok(int dtype, int kv, int c_in, int c_out, int transpose_w) {
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
    const size_t idx = align_up((size_t)(kv + 1) * 512, 1024);
    if (2 * idx + 2 * stage > (size_t)TC_SMEM_BUDGET) return false;
    return true;
}

bool tc_gather_gemm_supported(const GatherGemmArgs &a) {
    if (a.dtype == SPX_I8) return false;
    if (!a.tile_table || !a.tile_mask) return false;   // built by spx_build_tile_table
    // tf32 input gradient: the filter box is an MN-major 32-bit operand = SWIZZLE_128B_BASE32B layout (TMA
    // swizzle mode 128B_ATOM_32B).  Needs whole 128-byte filter rows, otherwise fp32 dgrad runs on the FMA
    // kernel; spx_debug_configure bit 256 switches the tensor-core route off (A/B against the FMA kernel).
    if (a.dtype == SPX_F32 && a.transpose_w && ((runtime_cfg().debug & 256) || (a.c_in * 4) % 128)) return false;
    if (((size_t)a.kv * a.c_in * dtype_bytes(a.dtype)) % 16) return false;
    return tc_shape_ok(a.dtype, a.kv, a.c_in, a.c_out, a.transpose_w);
}
bool tc_gather_gemm_int8_supported(const Int8Args &q) {
    if (!q.g.tile_table || !q.g.tile_mask) return false;
    if (((size_t)q.g.kv * q.g.c_in) % 16) return false;
    return tc_shape_ok(SPX_I8, q.g.kv, q.g.c_in, q.g.c_out, 0);
}

static int fill_params(const GatherGemmArgs &a, TcParams &p) {
    memset(&p, 0, sizeof(p));
    const int e = dtype_bytes(a.dtype);
    const int cx = a.cx(), cy = a.cy();
    p.x = (const uint8_t *)a.x;
    p.xb = cx * e;
    p.span_a = p.xb < 128 ? p.xb : 128;
    p.lg_span_a = p.span_a == 128 ? 7 : (p.span_a == 64 ? 6 : 5);
    p.a_subtiles = p.xb / p.span_a;
    p.q_a = p.span_a / 32;
    const int wb = a.c_in * e;                 // inner (contiguous) bytes of one weight slice row
    p.span_b = wb < 128 ? wb : 128;
    p.b_subtiles = wb / p.span_b;
    p.b_sub_bytes = a.c_out * p.span_b;
    p.b_bytes = a.c_out * wb;
    p.b_mn_major = a.transpose_w;
    p.q_b = p.span_b / 32;
    p.b_kstep16_mn = ((32 / e) * p.span_b) >> 4;   // one k-step = UMMA_K rows of the weight box
    p.b_base32 = (a.dtype == SPX_F32 && a.transpose_w) ? 1 : 0;
    p.w_inner_elems = a.c_in;
    p.span_b_elems = p.span_b / e;
    p.n = cy;
    int c_fmt = a.dtype == SPX_I8 ? 2 : 1;
    int ab_fmt = a.dtype == SPX_F16 ? 0 : a.dtype == SPX_BF16 ? 1 : a.dtype == SPX_F32 ? 2 : 1;
    p.idesc = make_idesc(c_fmt, ab_fmt, ab_fmt, 0, a.transpose_w, TC_TILE_M, cy);
    p.a_stage_bytes = (int)align_up((size_t)TC_TILE_M * p.xb, 1024);
    p.stage_bytes = p.a_stage_bytes + (int)align_up((size_t)p.b_bytes, 1024);
    p.idx_bytes = (int)align_up((size_t)(a.kv + 1) * 512, 1024);
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * cy)) cols <<= 1;
    p.tmem_cols = cols;
    // two CTAs per SM when >= 3 pipeline stages and both TMEM allocations fit; else one big CTA
    const RuntimeCfg &cfg = runtime_cfg();
    const int want_ctas = cfg.tc_ctas;
    p.ctas_per_sm = (want_ctas >= 2 && (TC_SMEM_BUDGET_2 - 2 * p.idx_bytes) / p.stage_bytes >= 3 && cols <= 256) ? 2 : 1;
    const int budget = p.ctas_per_sm == 2 ? TC_SMEM_BUDGET_2 : TC_SMEM_BUDGET;
    p.stages = (budget - 2 * p.idx_bytes) / p.stage_bytes;
    if (p.stages > TC_MAX_STAGES) p.stages = TC_MAX_STAGES;
    SPX_REQUIRE(p.stages >= 2, "tc_gather_gemm: tile does not fit shared memory (stage %d bytes)", p.stage_bytes);
    p.rows = a.rows; p.tile_table = a.tile_table; p.argsort = a.argsort;
    {
        const int64_t tiles = div_up64(a.rows, TC_TILE_M);
        p.sched_rec = a.tile_table + tt_blocks_elems(tiles, a.kv);
        // scheduler scratch lives in the caller's tile-table buffer (include/spconv_b200.h): TT_STATE_INTS / 2
        // {ticket, finished} pairs.  Every launch takes the next pair, so launches that overlap on the
        // same rulebook (two layers sharing an indice_key on different streams, graph branches) never
        // draw from one counter; a pair is zero when its launch ends.
        static std::atomic<unsigned> next_slot{0};
        const unsigned slot = next_slot.fetch_add(1, std::memory_order_relaxed) % (TT_STATE_INTS / 2);
        p.sched_state = const_cast<int *>(reinterpret_cast<const int *>(p.sched_rec + tiles * TT_REC_INTS)) + 2 * slot;
    }
    p.kv = a.kv; p.words = (a.kv + 31) / 32; p.reverse = a.reverse;
    p.y = a.y; p.out_dtype = a.dtype; p.epi_mode = 0; p.bias = a.bias; p.act = a.act; p.alpha = a.alpha;
    p.debug = cfg.debug;          // perf-triage hooks, set through spx_debug_configure() only
    p.dbg_ts = cfg.trace;
    return 0;
}

template <int KIND, int CPR>
static int launch_tc_cpr(const CUtensorMap &tm, const TcParams &p, cudaStream_t stream) {
    const size_t smem = (size_t)p.stages * p.stage_bytes + 2 * (size_t)p.idx_bytes + 1024 /*align slack*/ + 1024 /*barriers, tile-info ring*/;
    // the opt-in is a per-DEVICE attribute: one process may drive several GPUs
    if (!func_configured((const void *)tc_gather_gemm_kernel<KIND, CPR>, current_device())) {
        SPX_CHECK_CUDA(cudaFuncSetAttribute(tc_gather_gemm_kernel<KIND, CPR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(TC_SMEM_BUDGET + 2048)));
        SPX_CHECK_CUDA(cudaFuncSetAttribute(tc_gather_gemm_kernel<KIND, CPR>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    }
    const int64_t tiles = div_up64(p.rows, TC_TILE_M);
    const int64_t max_ctas = (int64_t)sm_count() * p.ctas_per_sm;
    const int grid = (int)(tiles < max_ctas ? tiles : max_ctas);
    tc_gather_gemm_kernel<KIND, CPR><<<grid, TC_THREADS, smem, stream>>>(tm, p);
    SPX_CHECK_LAUNCH("tc_gather_gemm_kernel");
    return 0;
}

template <int KIND>
static int launch_tc(const CUtensorMap &tm, const TcParams &p, cudaStream_t stream) {
    switch (p.xb >> 4) {
        case 2: return launch_tc_cpr<KIND, 2>(tm, p, stream);
        case 4: return launch_tc_cpr<KIND, 4>(tm, p, stream);
        case 8: return launch_tc_cpr<KIND, 8>(tm, p, stream);
        case 16: return launch_tc_cpr<KIND, 16>(tm, p, stream);
        case 32: return launch_tc_cpr<KIND, 32>(tm, p, stream);
    }
    set_error("tc_gather_gemm: unsupported row bytes %d", p.xb);
    return 2;
}

static int copy_mask_out(const GatherGemmArgs &a, cudaStream_t stream) {
    if (!a.mask_out) return 0;      // mask_output_fwd == the per-tile OR masks of the tile table
    const size_t bytes = (size_t)div_up64(a.rows, TC_TILE_M) * ((a.kv + 31) / 32) * sizeof(uint32_t);
    SPX_CHECK_CUDA(cudaMemcpyAsync(a.mask_out, a.tile_mask, bytes, cudaMemcpyDeviceToDevice, stream));
    return 0;
}

int tc_gather_gemm(const GatherGemmArgs &a, cudaStream_t stream) {
    TcParams p;
    if (fill_params(a, p)) return 2;
    if (copy_mask_out(a, stream)) return 1;
    CUtensorMap tm;
    if (make_weight_tmap(&tm, a.w, a.dtype, a.kv, a.c_in, a.c_out, p.span_b, p.b_base32 != 0)) return 2;
    if (a.dtype == SPX_F32) return launch_tc<KIND_TF32>(tm, p, stream);
    return launch_tc<KIND_F16>(tm, p, stream);
}

int tc_gather_gemm_int8(const Int8Args &q, cudaStream_t stream) {
    TcParams p;
    if (fill_params(q.g, p)) return 2;
    if (copy_mask_out(q.g, stream)) return 1;
    p.out_dtype = q.out_dtype;
    p.epi_mode = 1;
    p.scale = q.scale; p.bias_f32 = q.bias_f32; p.output_add = q.output_add;
    p.output_add_scale = q.output_add_scale;
    CUtensorMap tm;
    if (make_weight_tmap(&tm, q.g.w, SPX_I8, q.g.kv, q.g.c_in, q.g.c_out, p.span_b)) return 2;
    return launch_tc<KIND_I8>(tm, p, stream);
}

}  // namespace spx
