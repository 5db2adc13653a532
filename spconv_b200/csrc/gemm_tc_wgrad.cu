// tcgen05 weight gradient:  dW[:, k, :] = sum_o dout[o, :]^T x[pair_fwd[k][o], :].
//
// GEMM view per 128-voxel tile:  D_g[128 x c_out] += A_g[128 x 128 voxels] * B[c_out x 128 voxels]^T
//   * the contraction runs over VOXELS, so both operands are MN-major UMMA operands built from
//     the very same "rows of channels" shared-memory image the forward pass uses (a gathered
//     row of C channels is one K-slice of an MN-major tile);
//   * A_g stacks as many kernel offsets as fit M = 128 (two offsets for C_in = 64 fp16),
//     B is the dout tile, loaded once per voxel tile and shared by all offsets;
//   * every group's fp32 accumulator stays resident in TMEM for ALL tiles a CTA visits
//     (512 columns = 8 groups at c_out = 64); the kernel offsets that do not fit are handled
//     by a second "pass" CTA column (gridDim.y), so the filter gradient is never spilled;
//   * each CTA finally dumps its TMEM to an fp32 partial buffer and a small second kernel sums
//     the partials in a fixed order (deterministic; the reference's split-K does the same
//     with fp32 workspaces, spconv/csrc/sparse/convops.py:1236-1243, :2421-2436).
#include "gemm.cuh"

namespace spx {

constexpr int WG_TILE = 128;
constexpr int WG_THREADS = 288;          // warps 0-3 epilogue | 4-7 producers | 8 MMA issuer
constexpr int WG_MAX_STAGES = 6;
constexpr int WG_SMEM_BUDGET = 200 * 1024;

struct WgParams {
    const uint8_t *x; int xb, span_x, apo, apg, atom_elems;
    const uint8_t *d; int db, span_d;
    int n; uint32_t idesc; int ksteps, rows_per_kstep;
    int groups_total, groups_per_pass;
    int stages, a_stage_bytes, b_buf_bytes;
    int64_t rows;
    const int32_t *pair; int64_t pair_stride;
    const uint32_t *mask; const int32_t *argsort;
    int kv, words, c_in;
    float *partial; int64_t partial_stride;
};

struct BitIter4 { uint32_t m[4]; };

__device__ __forceinline__ bool bit_set(const uint32_t (&m)[4], int k) { return (m[k >> 5] >> (k & 31)) & 1u; }

// offsets covered by group g: atoms [g*apg, (g+1)*apg) -> offsets a / apo
__device__ __forceinline__ bool group_active(const uint32_t (&tm)[4], int g, const WgParams &p) {
    const int a0 = g * p.apg, a1 = a0 + p.apg;
    int k0 = a0 / p.apo, k1 = (a1 - 1) / p.apo;
    for (int k = k0; k <= k1 && k < p.kv; ++k)
        if (bit_set(tm, k)) return true;
    return false;
}

__device__ __forceinline__ void wg_tile_mask(const uint32_t *__restrict__ mask, int64_t base, int64_t rows, int words,
                                             int kv, int lane, uint32_t (&out)[4]) {
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
    }
}

__global__ void __launch_bounds__(WG_THREADS, 1)
tc_wgrad_kernel(const WgParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
    uint8_t *smem = smem_raw + pad;
    const uint32_t smem_base = raw_addr + pad;
    // layout: [2 x B buffer][stages x A stage][barriers]
    const uint32_t b_base = smem_base;
    const uint32_t a_base = smem_base + 2u * p.b_buf_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * (size_t)p.b_buf_bytes + (size_t)p.stages * p.a_stage_bytes);
    uint64_t *full_a = bars;                          // [stages]
    uint64_t *empty_a = bars + WG_MAX_STAGES;         // [stages]
    uint64_t *full_b = bars + 2 * WG_MAX_STAGES;      // [2]
    uint64_t *empty_b = bars + 2 * WG_MAX_STAGES + 2; // [2]
    uint64_t *acc_done = bars + 2 * WG_MAX_STAGES + 4;
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(bars + 2 * WG_MAX_STAGES + 5);
    uint32_t *used_smem = tmem_ptr_smem + 1;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t num_tiles = (p.rows + WG_TILE - 1) / WG_TILE;
    const int chunk = blockIdx.x, chunks = gridDim.x;
    const int g_begin = blockIdx.y * p.groups_per_pass;
    const int g_end = min(p.groups_total, g_begin + p.groups_per_pass);

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_a[s], 128); mbar_init(&empty_a[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&full_b[b], 128); mbar_init(&empty_b[b], 1); }
        mbar_init(acc_done, 1);
        *used_smem = 0;
        mbar_fence_init();
    }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp >= 4 && warp < 8) {
        // ================================================= producers
        const int pw = warp - 4;
        int stage = 0; uint32_t phase = 0;
        int64_t nb = 0;                                  // B buffers filled so far
        const int cpr_d = p.db >> 4;                      // 16-byte chunks per dout row
        const int cpa = p.span_x >> 4;                    // chunks per atom row
        for (int64_t tile = chunk; tile < num_tiles; tile += chunks) {
            const int64_t base = tile * WG_TILE;
            uint32_t tm[4];
            wg_tile_mask(p.mask, base, p.rows, p.words, p.kv, lane, tm);
            bool any = false;
            for (int g = g_begin; g < g_end; ++g) any = any || group_active(tm, g, p);
            if (!any) continue;
            const int64_t my_row = base + pw * 32 + lane;
            int32_t src_row = -1;
            if (my_row < p.rows) src_row = p.argsort ? __ldg(p.argsort + my_row) : (int32_t)my_row;
            // ---- dout tile (MN-major B operand)
            {
                const int bb = (int)(nb & 1);
                mbar_wait(&empty_b[bb], (uint32_t)(((nb >> 1) & 1) ^ 1));
                const uint32_t dstb = b_base + (uint32_t)bb * p.b_buf_bytes;
                for (int itc = 0; itc < cpr_d; ++itc) {
                    const int flat = itc * 32 + lane;
                    const int r = flat / cpr_d;
                    const int ch = flat - r * cpr_d;
                    const int32_t rsrc = __shfl_sync(0xffffffffu, src_row, r);
                    const uint32_t byte_in_row = (uint32_t)ch << 4;
                    const uint32_t sub = byte_in_row / (uint32_t)p.span_d;
                    const uint32_t within = byte_in_row - sub * p.span_d;
                    const uint32_t row_in_tile = (uint32_t)(pw * 32 + r);
                    const uint32_t dst = dstb + sub * (uint32_t)(WG_TILE * p.span_d) +
                                         swizzle_offset(row_in_tile * p.span_d + within, p.span_d);
                    const uint8_t *src = p.d + (rsrc >= 0 ? (int64_t)rsrc * p.db + byte_in_row : 0);
                    cp_async_16(dst, src, rsrc >= 0 ? 16u : 0u);
                }
                cp_async_mbar_arrive_noinc(&full_b[bb]);
                ++nb;
            }
            // ---- gathered x atoms, one stage per active group
            for (int g = g_begin; g < g_end; ++g) {
                if (!group_active(tm, g, p)) continue;
                mbar_wait(&empty_a[stage], phase ^ 1u);
                const uint32_t a_stage = a_base + (uint32_t)stage * p.a_stage_bytes;
                for (int s = 0; s < p.apg; ++s) {
                    const int a = g * p.apg + s;
                    const int k = a / p.apo;
                    const int cb = a - k * p.apo;
                    int32_t idx = -1;
                    if (k < p.kv && bit_set(tm, k) && src_row >= 0)
                        idx = __ldg(p.pair + (int64_t)k * p.pair_stride + src_row);
                    const uint32_t atom_base = a_stage + (uint32_t)s * (uint32_t)(WG_TILE * p.span_x);
                    for (int itc = 0; itc < cpa; ++itc) {
                        const int flat = itc * 32 + lane;
                        const int r = flat / cpa;
                        const int ch = flat - r * cpa;
                        const int32_t ridx = __shfl_sync(0xffffffffu, idx, r);
                        const uint32_t row_in_tile = (uint32_t)(pw * 32 + r);
                        const uint32_t dst = atom_base + swizzle_offset(row_in_tile * p.span_x + ((uint32_t)ch << 4), p.span_x);
                        const uint8_t *src = p.x + (ridx >= 0 ? (int64_t)ridx * p.xb + (int64_t)cb * p.span_x + (ch << 4) : 0);
                        cp_async_16(dst, src, ridx >= 0 ? 16u : 0u);
                    }
                }
                cp_async_mbar_arrive_noinc(&full_a[stage]);
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 8) {
        // ================================================= MMA issuer
        int stage = 0; uint32_t phase = 0;
        int64_t nb = 0;
        uint32_t used = 0;
        for (int64_t tile = chunk; tile < num_tiles; tile += chunks) {
            const int64_t base = tile * WG_TILE;
            uint32_t tm[4];
            wg_tile_mask(p.mask, base, p.rows, p.words, p.kv, lane, tm);
            bool any = false;
            for (int g = g_begin; g < g_end; ++g) any = any || group_active(tm, g, p);
            if (!any) continue;
            const int bb = (int)(nb & 1);
            mbar_wait(&full_b[bb], (uint32_t)((nb >> 1) & 1));
            const uint32_t b_buf = b_base + (uint32_t)bb * p.b_buf_bytes;
            for (int g = g_begin; g < g_end; ++g) {
                if (!group_active(tm, g, p)) continue;
                mbar_wait(&full_a[stage], phase);
                tc_fence_after();
                fence_proxy_async_smem();
                const int gl = g - g_begin;
                if (lane == 0) {
                    const uint32_t a_stage = a_base + (uint32_t)stage * p.a_stage_bytes;
                    const uint32_t d_tmem = tmem_base + (uint32_t)(gl * p.n);
                    for (int j = 0; j < p.ksteps; ++j) {
                        const uint64_t a_desc = make_smem_desc(a_stage + (uint32_t)j * p.rows_per_kstep * p.span_x,
                                                               (uint32_t)(WG_TILE * p.span_x), 8u * p.span_x, p.span_x);
                        const uint64_t b_desc = make_smem_desc(b_buf + (uint32_t)j * p.rows_per_kstep * p.span_d,
                                                               (uint32_t)(WG_TILE * p.span_d), 8u * p.span_d, p.span_d);
                        umma_ss<KIND_F16>(d_tmem, a_desc, b_desc, p.idesc, (((used >> gl) & 1u) || j > 0) ? 1u : 0u);
                    }
                    tc_commit(&empty_a[stage]);
                }
                __syncwarp();
                used |= 1u << gl;
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
            if (lane == 0) tc_commit(&empty_b[bb]);
            __syncwarp();
            ++nb;
        }
        if (lane == 0) {
            *used_smem = used;
            tc_commit(acc_done);
        }
        __syncwarp();
    } else {
        // ================================================= epilogue: TMEM -> fp32 partials
        mbar_wait(acc_done, 0);
        tc_fence_after();
        const uint32_t used = *reinterpret_cast<volatile uint32_t *>(used_smem);
        const int L = warp * 32 + lane;                    // TMEM lane = M index inside the group
        const int s = L / p.atom_elems;
        const int ce = L - s * p.atom_elems;
        float *part = p.partial + (int64_t)chunk * p.partial_stride;
        for (int g = g_begin; g < g_end; ++g) {
            const int gl = g - g_begin;
            const int a = g * p.apg + s;
            const int k = a / p.apo;
            const int c = (a - k * p.apo) * p.atom_elems + ce;
            const bool valid = k < p.kv;
            const bool has = (used >> gl) & 1u;
            const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(gl * p.n);
            for (int n0 = 0; n0 < p.n; n0 += 16) {
                uint32_t v[16];
                if (has) {
                    tmem_ld_32x32b_x16(t_row + (uint32_t)n0, v);
                    tc_wait_ld();
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0u;
                }
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        part[((int64_t)(n0 + j) * p.kv + k) * p.c_in + c] = __uint_as_float(v[j]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <typename T>
__global__ void wgrad_reduce_kernel(const float *__restrict__ partial, int64_t stride, int chunks, int64_t total,
                                    T *__restrict__ dw) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += partial[(int64_t)c * stride + i];
    dw[i] = from_float<T>(s);
}

// ------------------------------------------------------------------ host side
static bool wg_span_ok(int bytes) { return bytes == 32 || bytes == 64 || (bytes >= 128 && bytes % 128 == 0); }

struct WgPlan { WgParams p; int passes, chunks; size_t smem; };

static bool make_plan(const WgradArgs &a, WgPlan &pl) {
    if (a.dtype != SPX_F16 && a.dtype != SPX_BF16) return false;     // tf32 MN-major needs SW128_32B atoms
    const int e = 2;
    if (a.c_in % 16 || a.c_out % 16 || a.c_in > 256 || a.c_out > 256) return false;
    if (!wg_span_ok(a.c_in * e) || !wg_span_ok(a.c_out * e)) return false;
    WgParams &p = pl.p;
    memset(&p, 0, sizeof(p));
    p.xb = a.c_in * e; p.span_x = p.xb < 128 ? p.xb : 128;
    p.apo = p.xb / p.span_x;
    p.atom_elems = p.span_x / e;
    p.apg = 128 / p.atom_elems;
    p.db = a.c_out * e; p.span_d = p.db < 128 ? p.db : 128;
    p.n = a.c_out;
    p.rows_per_kstep = 32 / e;
    p.ksteps = WG_TILE / p.rows_per_kstep;
    const int ab_fmt = a.dtype == SPX_F16 ? 0 : 1;
    p.idesc = make_idesc(1, ab_fmt, ab_fmt, 1, 1, 128, a.c_out);
    const int atoms_total = a.kv * p.apo;
    p.groups_total = (atoms_total + p.apg - 1) / p.apg;
    p.groups_per_pass = 512 / a.c_out;
    if (p.groups_per_pass > 32) p.groups_per_pass = 32;
    pl.passes = (p.groups_total + p.groups_per_pass - 1) / p.groups_per_pass;
    p.a_stage_bytes = p.apg * WG_TILE * p.span_x;
    p.b_buf_bytes = WG_TILE * p.db;
    int avail = WG_SMEM_BUDGET - 2 * p.b_buf_bytes;
    if (avail < 2 * p.a_stage_bytes) return false;
    p.stages = avail / p.a_stage_bytes;
    if (p.stages > WG_MAX_STAGES) p.stages = WG_MAX_STAGES;
    pl.smem = 2 * (size_t)p.b_buf_bytes + (size_t)p.stages * p.a_stage_bytes + 1024 + 256;
    int64_t tiles = div_up64(a.n_out, WG_TILE);
    int chunks = sm_count() / pl.passes;
    if (chunks < 1) chunks = 1;
    if (chunks > tiles) chunks = (int)tiles;
    if (chunks < 1) chunks = 1;
    pl.chunks = chunks;
    p.rows = a.n_out; p.pair = a.pair; p.pair_stride = a.pair_stride; p.mask = a.mask; p.argsort = a.argsort;
    p.kv = a.kv; p.words = (a.kv + 31) / 32; p.c_in = a.c_in;
    p.x = (const uint8_t *)a.x; p.d = (const uint8_t *)a.dout;
    p.partial_stride = (int64_t)a.kv * a.c_in * a.c_out;
    return true;
}

bool tc_wgrad_supported(const WgradArgs &a) {
    WgPlan pl;
    return make_plan(a, pl);
}

size_t tc_wgrad_workspace_size(const WgradArgs &a) {
    WgPlan pl;
    if (!make_plan(a, pl)) return 256;
    // chunk count depends on the SM count only through an upper bound; size for the bound
    return (size_t)pl.chunks * pl.p.partial_stride * sizeof(float) + 256;
}

int tc_wgrad(const WgradArgs &a, cudaStream_t stream) {
    WgPlan pl;
    SPX_REQUIRE(make_plan(a, pl), "tc_wgrad: unsupported shape");
    pl.p.partial = (float *)a.workspace;
    SPX_REQUIRE((size_t)pl.chunks * pl.p.partial_stride * sizeof(float) <= a.workspace_bytes,
                "tc_wgrad: workspace too small");
    static thread_local bool configured = false;
    if (!configured) {
        SPX_CHECK_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            WG_SMEM_BUDGET + 2048));
        configured = true;
    }
    dim3 grid(pl.chunks, pl.passes);
    tc_wgrad_kernel<<<grid, WG_THREADS, pl.smem, stream>>>(pl.p);
    SPX_CHECK_LAUNCH("tc_wgrad_kernel");
    const int64_t total = pl.p.partial_stride;
    unsigned nblk = (unsigned)div_up64(total, 256);
    if (a.dtype == SPX_F16)
        wgrad_reduce_kernel<__half><<<nblk, 256, 0, stream>>>(pl.p.partial, total, pl.chunks, total, (__half *)a.dw);
    else
        wgrad_reduce_kernel<__nv_bfloat16><<<nblk, 256, 0, stream>>>(pl.p.partial, total, pl.chunks, total, (__nv_bfloat16 *)a.dw);
    SPX_CHECK_LAUNCH("wgrad_reduce_kernel");
    return 0;
}

}  // namespace spx
