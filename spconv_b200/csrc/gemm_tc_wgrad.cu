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
// Roles: warps 0-3 drain only, 4-11 gather producers (16-byte cp.async, 16 tile rows per warp;
// they join the drain at the end), 12 MMA issuer, 13 tile feeder (index-block ring + per-tile
// group sets).  Tiles are assigned STATICALLY so the summation order of dW -- and with it the
// result -- is reproducible bit for bit, but not round-robin: the schedule records list the tiles
// by decreasing offset count, and CTA c takes records c, 2C-1-c, 2C+c, 4C-1-c, ... (a snake over
// rows of C = chunks records).  Every CTA then holds one tile of every cost rank, which is LPT
// list scheduling without a run-time counter (round-robin over the unsorted tiles left the
// slowest CTA ~15 us behind the median on the 100 k-voxel cloud).
#include "gemm.cuh"
#include "peer.cuh"
#include <stdlib.h>

namespace spx {

constexpr int WG_TILE = 128;
constexpr int WG_PROD_WARPS = 8;        // 16 tile rows per producer warp
constexpr int WG_PROD_THREADS = WG_PROD_WARPS * 32;
constexpr int WG_MMA_WARP = 4 + WG_PROD_WARPS;
constexpr int WG_SCHED_WARP = WG_MMA_WARP + 1;
constexpr int WG_THREADS = (WG_SCHED_WARP + 1) * 32; // warps 0-3 drain | 4-11 producers | 12 MMA issuer | 13 tile feeder
constexpr int WG_MAX_STAGES = 6;
constexpr int WG_SMEM_BUDGET = 200 * 1024;    // operand stages are sized inside this ...
constexpr int WG_SMEM_MAX = 224 * 1024;       // ... what is left up to here buys deeper index prefetch
constexpr int WG_MAX_IDX = 4;                 // index-block ring depth (prefetch distance = depth - 1 tiles)

struct WgParams {
    const uint8_t *x; int xb, span_x, lg_span_x, apo, apg, atom_elems;
    const uint8_t *d; int db, span_d, lg_span_d, lg_cpr_d;
    int n; uint32_t idesc; int ksteps, rows_per_kstep;
    int groups_total, groups_per_pass;
    int stages, a_stage_bytes, b_buf_bytes, idx_bytes, idx_bufs;
    int64_t rows;
    const int32_t *tile_table;   // [tiles][kv+1][128]
    const uint32_t *tile_mask;   // [tiles][words]
    const int32_t *sched_rec;    // [tiles][TT_REC_INTS] {tile, mask[4]}: tiles by decreasing offset count (gemm.cuh)
    int kv, words, c_in;
    float *partial; int64_t partial_stride;
    long long *dbg_ts;           // optional [8][2048] clock64 stamps of CTA (0,0) (SPX_TC_TRACE)
    int debug;                   // SPX_TC_DEBUG ablation bits (perf triage only): 16 no partial stores, 32 no TMEM loads
};

__device__ __forceinline__ bool bit_set(const uint32_t (&m)[4], int k) { return (m[k >> 5] >> (k & 31)) & 1u; }

// Offset slots are filled in the order 0, kv-1, 1, kv-2, ...: a group (one M = 128 accumulator)
// then stacks an offset with its point mirror.  On surface-like clouds a voxel that has the
// neighbour +d usually has -d too, so the two halves of a group are active together and almost no
// all-zero atom is gathered (with the natural order k, k+1 about half of the gathered atoms of
// config 2 were zero fill).  Any permutation is valid; this one only changes which TMEM rows an
// offset owns.
__device__ __forceinline__ int slot_offset(int slot, int kv) {
    if (slot >= kv) return kv;                       // padding slot of the last group
    return (slot & 1) ? kv - 1 - (slot >> 1) : (slot >> 1);
}

// bit gl set iff group gl of this pass has an active offset in the tile
__device__ __forceinline__ uint32_t active_groups(const uint32_t (&tm)[4], const uint32_t *gmask, int ng, int words) {
    uint32_t act = 0;
    if (words == 1) {                                   // kv <= 32: the common 3x3x3 case
        for (int gl = 0; gl < ng; ++gl) act |= (tm[0] & gmask[gl * 4]) ? 1u << gl : 0u;
        return act;
    }
    for (int gl = 0; gl < ng; ++gl) {
        const uint32_t hit = (tm[0] & gmask[gl * 4]) | (tm[1] & gmask[gl * 4 + 1]) | (tm[2] & gmask[gl * 4 + 2]) |
                             (tm[3] & gmask[gl * 4 + 3]);
        if (hit) act |= 1u << gl;
    }
    return act;
}

__device__ __forceinline__ void wg_load_tile_mask(const uint32_t *__restrict__ tile_mask, int64_t tile, int words,
                                                  uint32_t (&out)[4]) {
#pragma unroll
    for (int w = 0; w < 4; ++w) out[w] = w < words ? __ldg(tile_mask + tile * words + w) : 0u;
}

// record visited by CTA `chunk` at step i (snake order over the cost-sorted schedule records)
__device__ __forceinline__ int64_t wg_rec_index(int64_t i, int chunk, int chunks) {
    return i * chunks + ((i & 1) ? (chunks - 1 - chunk) : chunk);
}
__device__ __forceinline__ void wg_load_rec(const int32_t *__restrict__ rec, int64_t r, int64_t &tile, uint32_t (&m)[4]) {
    const int4 a = __ldg(reinterpret_cast<const int4 *>(rec + r * TT_REC_INTS));
    const int b = __ldg(rec + r * TT_REC_INTS + 4);
    tile = a.x; m[0] = (uint32_t)a.y; m[1] = (uint32_t)a.z; m[2] = (uint32_t)a.w; m[3] = (uint32_t)b;
}

__device__ __forceinline__ long long wg_global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// per-CTA wall-clock spans (ns): row 2 of the trace buffer holds [pass * chunks + chunk][4]
#define WG_SPAN(i) do { if (p.dbg_ts && threadIdx.x == 0) { const int c_ = blockIdx.y * gridDim.x + blockIdx.x; if (c_ < 512) p.dbg_ts[2 * 2048 + c_ * 4 + (i)] = wg_global_ns(); } } while (0)
#define WG_STAMP(role, n) do { if (p.dbg_ts && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (n) < 2048) p.dbg_ts[(role) * 2048 + (n)] = clock64(); } while (0)

__device__ __forceinline__ uint32_t pick_word(const uint32_t (&m)[4], int w) {
    return w == 0 ? m[0] : (w == 1 ? m[1] : (w == 2 ? m[2] : m[3]));     // selects, no local-memory indexing
}

// CPA = 16-byte chunks per x-atom row (span_x / 16), CPD = 16-byte chunks per dout row (db / 16)
// 16-byte-chunk address swizzle of an MN-major operand row: SWIZZLE_{32,64,128}B for 16-bit operands; 32-bit
// (tf32) MN-major operands use SWIZZLE_128B_BASE32B = CuTe Swizzle<2,5,2>: 32-byte units XORed with the row
// inside its 4-row atom (the pattern TMA's 128B_ATOM_32B mode writes for the tf32 input-gradient filter box)
template <bool TF32>
__device__ __forceinline__ uint32_t wg_swizzle(uint32_t off, uint32_t span) {
    if constexpr (TF32) return off ^ (((off >> 7) & 3u) << 5);
    else return swizzle_offset(off, span);
}

template <int CPA, int CPD, bool TF32 = false>
__global__ void __launch_bounds__(WG_THREADS, 1)
tc_wgrad_kernel(const WgParams p) {
    constexpr int LG_CPA = CPA == 2 ? 1 : (CPA == 4 ? 2 : 3);
    constexpr int RPI = 32 / CPA;
    constexpr int LG_CPD = CPD == 2 ? 1 : (CPD == 4 ? 2 : (CPD == 8 ? 3 : (CPD == 16 ? 4 : 5)));
    constexpr int RPI_D = 32 / CPD;                              // dout rows covered by one warp-wide cp.async
    constexpr int DB = CPD * 16;
    constexpr int SPAN_D = DB < 128 ? DB : 128;
    constexpr int LG_SPAN_D = SPAN_D == 128 ? 7 : (SPAN_D == 64 ? 6 : 5);
    constexpr int SPAN_X = CPA * 16;
    constexpr int LG_SPAN_X = LG_CPA + 4;
    constexpr int ROWS_PW = WG_TILE / WG_PROD_WARPS;           // tile rows per producer warp
    constexpr int ITERS = ROWS_PW / RPI > 0 ? ROWS_PW / RPI : 1;   // cp.async per thread per atom
    constexpr int ITERS_D = ROWS_PW / RPI_D;                       // cp.async per thread per dout tile
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    WG_SPAN(0);
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
    uint8_t *smem = smem_raw + pad;
    const uint32_t smem_base = raw_addr + pad;
    // layout: [2 x B buffer][stages x A stage][idx_bufs x index block][barriers]
    const uint32_t b_base = smem_base;
    const uint32_t a_base = smem_base + 2u * p.b_buf_bytes;
    const uint32_t idx_off = 2u * p.b_buf_bytes + (uint32_t)p.stages * p.a_stage_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + idx_off + (uint32_t)p.idx_bufs * p.idx_bytes);
    uint64_t *full_a = bars;                          // [stages]
    uint64_t *empty_a = bars + WG_MAX_STAGES;         // [stages]
    uint64_t *full_b = bars + 2 * WG_MAX_STAGES;      // [2]
    uint64_t *empty_b = bars + 2 * WG_MAX_STAGES + 2; // [2]
    uint64_t *idx_full = bars + 2 * WG_MAX_STAGES + 4;                // [WG_MAX_IDX]
    uint64_t *idx_empty = bars + 2 * WG_MAX_STAGES + 4 + WG_MAX_IDX;  // [WG_MAX_IDX]
    uint64_t *acc_done = bars + 2 * WG_MAX_STAGES + 4 + 2 * WG_MAX_IDX;
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(acc_done + 1);
    uint32_t *used_smem = tmem_ptr_smem + 1;
    uint32_t *gmask = used_smem + 1;                  // [32][4] offsets covered by each group of this pass
    uint32_t *slot_info = gmask + 32 * 4;             // [WG_MAX_IDX][8]: {active groups, tile mask[4]} per ring slot

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t num_tiles = (p.rows + WG_TILE - 1) / WG_TILE;
    const int chunk = blockIdx.x, chunks = gridDim.x;
    // Groups are dealt to the passes round-robin (pass y owns groups y, y + passes, ...): with the
    // mirror-paired slot order the frequently active in-plane offsets sit in neighbouring groups,
    // and a contiguous split gave one pass 70 % of the stages of the 100 k-voxel cloud.
    const int g_first = blockIdx.y, g_step = gridDim.y;
    const int ng = g_first < p.groups_total ? (p.groups_total - g_first + g_step - 1) / g_step : 0;

    if (threadIdx.x < 32) {
        // group gl covers atoms [g*apg, (g+1)*apg) -> offset slots a / apo; one thread per group
        const int g = g_first + (int)threadIdx.x * g_step;
        uint32_t m[4] = {0, 0, 0, 0};
        if ((int)threadIdx.x < ng) {
            for (int a = g * p.apg; a < (g + 1) * p.apg; ++a) {
                const int k = slot_offset(a / p.apo, p.kv);
                if (k < p.kv) m[k >> 5] |= 1u << (k & 31);
            }
        }
#pragma unroll
        for (int w = 0; w < 4; ++w) gmask[threadIdx.x * 4 + w] = m[w];
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_a[s], WG_PROD_THREADS); mbar_init(&empty_a[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&full_b[b], WG_PROD_THREADS); mbar_init(&empty_b[b], 1); }
        for (int b = 0; b < p.idx_bufs; ++b) { mbar_init(&idx_full[b], 1); mbar_init(&idx_empty[b], WG_PROD_WARPS); }
        mbar_init(acc_done, 1);
        *used_smem = 0;
        mbar_fence_init();
    }
    if (warp == WG_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (warp == 0) WG_STAMP(3, 0);
    WG_SPAN(1);

    if (warp >= 4 && warp < WG_MMA_WARP) {
        // ================================================= producers
        const int pw = warp - 4;
        int stage = 0; uint32_t phase = 0;
        int64_t nb = 0;                                  // B buffers filled so far
        int nst = 0, ntile = 0;
        // per-lane constants of the atom gather: chunk chb of rows r0 + itc*RPI of this warp's 32 rows
        const int r0 = lane >> LG_CPA;
        const uint32_t chb = (uint32_t)(lane & (CPA - 1)) << 4;
        const uint8_t *x_lane = p.x + chb;
        uint32_t dst_off[ITERS];
#pragma unroll
        for (int itc = 0; itc < ITERS; ++itc)
            dst_off[itc] = wg_swizzle<TF32>(((uint32_t)(pw * ROWS_PW + r0 + itc * RPI) << LG_SPAN_X) + chb, SPAN_X);
        const int lg_apo = p.apo == 1 ? 0 : (p.apo == 2 ? 1 : 2);
        // per-lane constants of the dout gather: chunk chd of rows rd0 + itc*RPI_D of this warp's rows
        const int rd0 = lane >> LG_CPD;
        const uint32_t chd = (uint32_t)(lane & (CPD - 1)) << 4;
        const uint8_t *d_lane = p.d + chd;
        uint32_t dstd_off[ITERS_D];
#pragma unroll
        for (int itc = 0; itc < ITERS_D; ++itc) {
            const uint32_t row_in_tile = (uint32_t)(pw * ROWS_PW + rd0 + itc * RPI_D);
            dstd_off[itc] = (chd >> LG_SPAN_D) * (uint32_t)(WG_TILE * SPAN_D) +
                            wg_swizzle<TF32>((row_in_tile << LG_SPAN_D) + (chd & (uint32_t)(SPAN_D - 1)), SPAN_D);
        }
        // index-block ring (filled by the feeder warp): slot / use count advance with the tiles
        const int nring = p.idx_bufs;
        // dout tile (MN-major B operand) of the tile whose index block is idx_s; source rows = block row kv
        auto issue_b = [&](const int32_t *idx_s) {
            const int bb = (int)(nb & 1);
            mbar_wait(&empty_b[bb], (uint32_t)(((nb >> 1) & 1) ^ 1));
            if (pw == 0) WG_STAMP(6, 2 * ntile);
            const uint32_t dstb = b_base + (uint32_t)bb * p.b_buf_bytes;
            const int32_t *rows_s = idx_s + p.kv * 128 + pw * ROWS_PW + rd0;
            int32_t rsrc[ITERS_D];
#pragma unroll
            for (int itc = 0; itc < ITERS_D; ++itc) rsrc[itc] = rows_s[itc * RPI_D];
#pragma unroll
            for (int itc = 0; itc < ITERS_D; ++itc)
                cp_async_16(dstb + dstd_off[itc], d_lane + (int64_t)max(rsrc[itc], 0) * DB, rsrc[itc] >= 0 ? 16u : 0u);
            cp_async_mbar_arrive_noinc(&full_b[bb]);
            if (pw == 0) WG_STAMP(6, 2 * ntile + 1);
            ++ntile;
            ++nb;
        };
        auto idx_block = [&](int b) {
            return reinterpret_cast<const int32_t *>(smem + idx_off + (size_t)b * p.idx_bytes);
        };
        // Software pipeline over tiles: the feeder warp keeps the ring of index blocks (and each
        // tile's group set) filled nring-1 tiles ahead; right after the first x stage of tile t
        // the dout tile of t+1 is issued -- nothing but the first x stage sits on the boundary.
        auto read_slot = [&](int slot, uint32_t use, uint32_t (&m)[4]) -> uint32_t {
            mbar_wait(&idx_full[slot], use & 1u);
            const volatile uint32_t *r = slot_info + slot * 8;
            m[0] = r[1]; m[1] = r[2]; m[2] = r[3]; m[3] = r[4];
            return r[0];
        };
        int cur_slot = 0; uint32_t cur_use = 0;          // ring position of the tile being gathered
        uint32_t tm[4] = {0, 0, 0, 0}, tm1[4] = {0, 0, 0, 0};
        uint32_t act = 0;
        if (wg_rec_index(0, chunk, chunks) < num_tiles) {
            act = read_slot(0, 0u, tm);
            if (act) issue_b(idx_block(0));
        }
        for (int64_t step = 0; wg_rec_index(step, chunk, chunks) < num_tiles; ++step) {
            const bool has_next = wg_rec_index(step + 1, chunk, chunks) < num_tiles;
            int nxt_slot = cur_slot + 1; uint32_t nxt_use = cur_use;
            if (nxt_slot == nring) { nxt_slot = 0; ++nxt_use; }
            uint32_t act_next = 0;
            bool next_ready = !has_next;
            auto prepare_next = [&]() {
                act_next = read_slot(nxt_slot, nxt_use, tm1);
                if (act_next) issue_b(idx_block(nxt_slot));
                next_ready = true;
            };
            const int32_t *idx_s = idx_block(cur_slot);
            // ---- gathered x atoms, one stage per active group
            for (uint32_t rem = act; rem; rem &= rem - 1) {
                const int g = g_first + (__ffs(rem) - 1) * g_step;
                mbar_wait(&empty_a[stage], phase ^ 1u);
                if (pw == 0) WG_STAMP(0, 2 * nst);
                const uint32_t a_stage = a_base + (uint32_t)stage * p.a_stage_bytes;
                for (int s = 0; s < p.apg; ++s) {
                    const int a = g * p.apg + s;
                    const int k = slot_offset(a >> lg_apo, p.kv);
                    const int cb = a & (p.apo - 1);
                    const bool active = k < p.kv && ((pick_word(tm, k >> 5) >> (k & 31)) & 1u);
                    const int32_t *idx_k = idx_s + (active ? k : 0) * 128 + pw * ROWS_PW + r0;
                    const uint32_t atom_base = a_stage + (uint32_t)s * (uint32_t)(WG_TILE * SPAN_X);
                    const uint8_t *x_atom = x_lane + cb * SPAN_X;
                    int32_t ridx[ITERS];                      // all index loads first, then the copies
#pragma unroll
                    for (int itc = 0; itc < ITERS; ++itc) ridx[itc] = active ? idx_k[itc * RPI] : -1;
#pragma unroll
                    for (int itc = 0; itc < ITERS; ++itc)
                        cp_async_16(atom_base + dst_off[itc], x_atom + (int64_t)max(ridx[itc], 0) * p.xb,
                                    ridx[itc] >= 0 ? 16u : 0u);
                }
                cp_async_mbar_arrive_noinc(&full_a[stage]);
                if (pw == 0) WG_STAMP(0, 2 * nst + 1);
                ++nst;
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                if (!next_ready) prepare_next();
            }
            if (!next_ready) prepare_next();
            __syncwarp();
            if (lane == 0) mbar_arrive(&idx_empty[cur_slot]);
#pragma unroll
            for (int w = 0; w < 4; ++w) tm[w] = tm1[w];
            act = act_next;
            cur_slot = nxt_slot; cur_use = nxt_use;
        }
    } else if (warp == WG_SCHED_WARP) {
        // ================================================= tile feeder
        // Per tile of this CTA (static round-robin, so the summation order of dW is fixed): compute
        // the active groups of this pass from the tile mask, publish them with the mask through
        // the ring slot, and bulk-copy the tile's index block -- only when the pass has work for
        // the tile.  Tile masks are fetched 32 tiles at a time, one per lane.
        const uint32_t blk_bytes = (uint32_t)(p.kv + 1) * 512u;
        const int nring = p.idx_bufs;
        int slot = 0; uint32_t use = 0;
        uint32_t mw[4] = {0, 0, 0, 0};
        int64_t mtile = 0;
        for (int64_t i = 0; wg_rec_index(i, chunk, chunks) < num_tiles; ++i) {
            if ((i & 31) == 0) {                         // the next 32 schedule records, one per lane
                const int64_t r = wg_rec_index(i + lane, chunk, chunks);
                if (r < num_tiles) wg_load_rec(p.sched_rec, r, mtile, mw);
            }
            uint32_t tm[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) tm[w] = __shfl_sync(0xffffffffu, mw[w], (int)(i & 31));
            const int64_t tile = __shfl_sync(0xffffffffu, mtile, (int)(i & 31));
            const uint32_t act = active_groups(tm, gmask, ng, p.words);
            mbar_wait(&idx_empty[slot], (use & 1u) ^ 1u);
            if (lane == 0) {
                uint32_t *r = slot_info + slot * 8;
                r[0] = act; r[1] = tm[0]; r[2] = tm[1]; r[3] = tm[2]; r[4] = tm[3];
                if (act) {
                    mbar_arrive_expect_tx(&idx_full[slot], blk_bytes);
                    bulk_copy_g2s(smem_base + idx_off + (uint32_t)slot * p.idx_bytes,
                                  p.tile_table + tile * (int64_t)(p.kv + 1) * 128, blk_bytes, &idx_full[slot]);
                } else {
                    mbar_arrive(&idx_full[slot]);
                }
            }
            __syncwarp();
            if (++slot == nring) { slot = 0; ++use; }
        }
    } else if (warp == WG_MMA_WARP) {
        // ================================================= MMA issuer
        int stage = 0; uint32_t phase = 0;
        int64_t nb = 0;
        uint32_t used = 0;
        int nst = 0, ntile = 0;
        uint32_t tm[4] = {0, 0, 0, 0};
        int64_t tile_unused = 0;
        if (wg_rec_index(0, chunk, chunks) < num_tiles)
            wg_load_rec(p.sched_rec, wg_rec_index(0, chunk, chunks), tile_unused, tm);
        // both operands are MN-major: LBO = distance between 128-row atoms, SBO = 8 rows
        // (tf32: 4-row atoms of the SWIZZLE_128B_BASE32B layout, descriptor layout code 1 -- as in gemm_tc.cu)
        const uint64_t a_hi = TF32 ? smem_desc_hi_layout((uint32_t)(WG_TILE * p.span_x), 4u * p.span_x, 1ull)
                                   : smem_desc_hi((uint32_t)(WG_TILE * p.span_x), 8u * p.span_x, p.span_x);
        const uint64_t b_hi = TF32 ? smem_desc_hi_layout((uint32_t)(WG_TILE * SPAN_D), 4u * SPAN_D, 1ull)
                                   : smem_desc_hi((uint32_t)(WG_TILE * SPAN_D), 8u * SPAN_D, SPAN_D);
        const uint32_t a_step16 = (uint32_t)(p.rows_per_kstep * p.span_x) >> 4;
        const uint32_t b_step16 = (uint32_t)(p.rows_per_kstep * SPAN_D) >> 4;
        for (int64_t step = 0; wg_rec_index(step, chunk, chunks) < num_tiles; ++step) {
            const int64_t next = wg_rec_index(step + 1, chunk, chunks);
            uint32_t tm_next[4] = {0, 0, 0, 0};
            if (next < num_tiles) wg_load_rec(p.sched_rec, next, tile_unused, tm_next);
            const uint32_t act = active_groups(tm, gmask, ng, p.words);
            if (act) {
                const int bb = (int)(nb & 1);
                mbar_wait(&full_b[bb], (uint32_t)((nb >> 1) & 1));
                WG_STAMP(7, ntile); ++ntile;
                const uint32_t b16 = (b_base + (uint32_t)bb * p.b_buf_bytes) >> 4;
                for (uint32_t rem = act; rem; rem &= rem - 1) {
                    const int gl = __ffs(rem) - 1;
                    mbar_wait(&full_a[stage], phase);
                    WG_STAMP(1, 2 * nst);
                    tc_fence_after();
                    {
                        // warp-uniform issue (see gemm_tc.cu): all lanes compute, an elected lane fires
                        const uint32_t a16 = (a_base + (uint32_t)stage * p.a_stage_bytes) >> 4;
                        const uint32_t d_tmem = tmem_base + (uint32_t)(gl * p.n);
                        uint32_t acc_flag = (used >> gl) & 1u;
                        for (int j = 0; j < p.ksteps; ++j) {
                            const uint64_t a_desc = a_hi | (uint64_t)((a16 + (uint32_t)j * a_step16) & 0x3FFFu);
                            const uint64_t b_desc = b_hi | (uint64_t)((b16 + (uint32_t)j * b_step16) & 0x3FFFu);
                            umma_ss_elect<TF32 ? KIND_TF32 : KIND_F16>(d_tmem, a_desc, b_desc, p.idesc, acc_flag);
                            acc_flag = 1u;
                        }
                        tc_commit_elect(&empty_a[stage]);
                    }
                    WG_STAMP(1, 2 * nst + 1); ++nst;
                    __syncwarp();
                    used |= 1u << gl;
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
                tc_commit_elect(&empty_b[bb]);
                __syncwarp();
                ++nb;
            }
#pragma unroll
            for (int w = 0; w < 4; ++w) tm[w] = tm_next[w];
        }
        if (lane == 0) *used_smem = used;
        __syncwarp();
        tc_commit_elect(acc_done);
        __syncwarp();
    }

    // ================================================= drain: TMEM -> fp32 partials
    // Twelve warps share it (a warp may touch the TMEM lane quarter warp % 4, so the eight
    // producer warps, idle by now, drain next to the four epilogue warps); work items are
    // (group, 32-column block) pairs with both tcgen05.ld in flight before the single wait.
    if (warp < WG_MMA_WARP) {
        mbar_wait(acc_done, 0);
        if (warp == 0) WG_SPAN(2);
        tc_fence_after();
        const uint32_t used = *reinterpret_cast<volatile uint32_t *>(used_smem);
        const int q = warp & 3, helper = warp >> 2;
        const int L = q * 32 + lane;                       // TMEM lane = M index inside the group
        const int s = L / p.atom_elems;
        const int ce = L - s * p.atom_elems;
        float *part = p.partial + (int64_t)chunk * p.partial_stride;
        const int nblocks = p.n >> 5;                      // 32-column blocks per group (n % 32 == 0 here)
        const int tail16 = p.n & 16;                       // one 16-column remainder when n % 32 == 16
        const int per_group = nblocks + (tail16 ? 1 : 0);
        int item = 0;
        for (int gl = 0; gl < ng; ++gl) {
            const int g = g_first + gl * g_step;
            const int a = g * p.apg + s;
            const int ks = a / p.apo;
            const int k = slot_offset(ks, p.kv);
            const int c = (a - ks * p.apo) * p.atom_elems + ce;
            const bool valid = k < p.kv;
            const bool has = (used >> gl) & 1u;
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(gl * p.n);
            for (int blk = 0; blk < per_group; ++blk, ++item) {
                if (item % 3 != helper) continue;
                const int n0 = blk * 32;
                const int cols = blk < nblocks ? 32 : 16;
                uint32_t v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
                if (has && !(p.debug & 32)) {
                    uint32_t lo[16], hi[16];
                    tmem_ld_32x32b_x16(t_row + (uint32_t)n0, lo);
                    if (cols == 32) tmem_ld_32x32b_x16(t_row + (uint32_t)n0 + 16u, hi);
                    tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) { v[j] = lo[j]; v[16 + j] = cols == 32 ? hi[j] : 0u; }
                }
                if (valid && !(p.debug & 16)) {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < cols) part[((int64_t)(n0 + j) * p.kv + k) * p.c_in + c] = __uint_as_float(v[j]);
                }
            }
        }
    }

    if (warp == 0) WG_STAMP(3, 1);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) WG_STAMP(3, 2);
    WG_SPAN(3);
    if (warp == WG_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// 4 outputs per thread (float4 partial reads), the chunk range split over the 8 warps of a block,
// partial sums combined through shared memory in warp order (fixed order => deterministic)
constexpr int RED_WARPS = 8;
template <typename T>
__global__ void __launch_bounds__(RED_WARPS * 32)
wgrad_reduce_kernel(const float *__restrict__ partial, int64_t stride, int chunks, int64_t total,
                    T *__restrict__ dw) {
    __shared__ float4 acc_s[RED_WARPS][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * 32 + lane) * 4;      // total % 4 == 0 (channels % 16 == 0)
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < total) {
        const int per = (chunks + RED_WARPS - 1) / RED_WARPS;
        const int c0 = warp * per, c1 = min(chunks, c0 + per);
#pragma unroll 4
        for (int c = c0; c < c1; ++c) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(partial + (int64_t)c * stride + i));
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    }
    acc_s[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && i < total) {
        float4 t = acc_s[0][lane];
#pragma unroll
        for (int w = 1; w < RED_WARPS; ++w) {
            const float4 v = acc_s[w][lane];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        dw[i] = from_float<T>(t.x); dw[i + 1] = from_float<T>(t.y);
        dw[i + 2] = from_float<T>(t.z); dw[i + 3] = from_float<T>(t.w);
    }
}

// ------------------------------------------------------------------ host side
static bool wg_span_ok(int bytes) { return bytes == 32 || bytes == 64 || (bytes >= 128 && bytes % 128 == 0); }

struct WgPlan { WgParams p; int passes, chunks; size_t smem; };

static bool make_plan(const WgradArgs &a, WgPlan &pl) {
    if (a.dtype != SPX_F16 && a.dtype != SPX_BF16 && a.dtype != SPX_F32) return false;
    if (!a.tile_table || !a.tile_mask) return false;                 // built by spx_build_tile_table
    const bool tf32 = a.dtype == SPX_F32;       // fp32 reaches here only in TF32 mode (api_gemm.cu)
    const int e = tf32 ? 4 : 2;
    if (a.c_in % 16 || a.c_out % 16 || a.c_in > 256 || a.c_out > 256) return false;
    // tf32: whole 128-byte atoms of 32 channels on both operands, dout rows of at most 512 bytes;
    // spx_debug_configure bit 4096 sends it back to the FMA kernel (A/B)
    if (tf32 && (a.c_in % 32 || a.c_out % 32 || a.c_out > 128 || (runtime_cfg().debug & 4096))) return false;
    if (!wg_span_ok(a.c_in * e) || !wg_span_ok(a.c_out * e)) return false;
    WgParams &p = pl.p;
    memset(&p, 0, sizeof(p));
    p.xb = a.c_in * e; p.span_x = p.xb < 128 ? p.xb : 128;
    p.lg_span_x = p.span_x == 128 ? 7 : (p.span_x == 64 ? 6 : 5);
    p.apo = p.xb / p.span_x;
    p.atom_elems = p.span_x / e;
    p.apg = 128 / p.atom_elems;
    p.db = a.c_out * e; p.span_d = p.db < 128 ? p.db : 128;
    p.lg_span_d = p.span_d == 128 ? 7 : (p.span_d == 64 ? 6 : 5);
    if (p.db & (p.db - 1)) return false;
    p.lg_cpr_d = 0;
    while ((1 << p.lg_cpr_d) < (p.db >> 4)) ++p.lg_cpr_d;
    p.n = a.c_out;
    p.rows_per_kstep = 32 / e;
    p.ksteps = WG_TILE / p.rows_per_kstep;
    const int ab_fmt = tf32 ? 2 : (a.dtype == SPX_F16 ? 0 : 1);
    p.idesc = make_idesc(1, ab_fmt, ab_fmt, 1, 1, 128, a.c_out);
    const int atoms_total = a.kv * p.apo;
    p.groups_total = (atoms_total + p.apg - 1) / p.apg;
    p.groups_per_pass = 512 / a.c_out;
    if (p.groups_per_pass > 32) p.groups_per_pass = 32;
    // A/B knob (spx_debug_configure bit 1024): half the accumulators per CTA => twice the passes, half the
    // chunks.  The fp32 partial volume (chunks x all groups) halves, the dout tile is gathered by twice as
    // many passes -- see profiles/README.md for the measured trade.
    if ((runtime_cfg().debug & 1024) && p.groups_per_pass >= 2) p.groups_per_pass /= 2;
    pl.passes = (p.groups_total + p.groups_per_pass - 1) / p.groups_per_pass;
    p.a_stage_bytes = p.apg * WG_TILE * p.span_x;
    p.b_buf_bytes = WG_TILE * p.db;
    p.idx_bytes = (int)align_up((size_t)(a.kv + 1) * 512, 1024);
    // tf32 stages are twice as large (64 KB per group of four 32-channel atoms): two of them only fit without the
    // head-room the 16-bit plans keep for a deeper index ring
    int avail = (tf32 ? WG_SMEM_MAX - 2048 : WG_SMEM_BUDGET) - 2 * p.b_buf_bytes - 2 * p.idx_bytes;
    if (avail < 2 * p.a_stage_bytes) return false;
    p.stages = avail / p.a_stage_bytes;
    if (p.stages > WG_MAX_STAGES) p.stages = WG_MAX_STAGES;
    size_t fixed = 2 * (size_t)p.b_buf_bytes + (size_t)p.stages * p.a_stage_bytes + 1024 + 1024;
    p.idx_bufs = 2;
    while (p.idx_bufs < WG_MAX_IDX && fixed + (size_t)(p.idx_bufs + 1) * p.idx_bytes <= (size_t)WG_SMEM_MAX) ++p.idx_bufs;
    pl.smem = fixed + (size_t)p.idx_bufs * p.idx_bytes;
    int64_t tiles = div_up64(a.n_out, WG_TILE);
    int chunks = sm_count() / pl.passes;
    if (chunks < 1) chunks = 1;
    if (chunks > tiles) chunks = (int)tiles;
    if (chunks < 1) chunks = 1;
    pl.chunks = chunks;
    p.rows = a.n_out; p.tile_table = a.tile_table; p.tile_mask = a.tile_mask;
    p.sched_rec = a.tile_table + tt_blocks_elems(tiles, a.kv);
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
    pl.p.dbg_ts = runtime_cfg().trace;      // perf-triage hooks (spx_debug_configure)
    pl.p.debug = runtime_cfg().debug;
    SPX_REQUIRE((size_t)pl.chunks * pl.p.partial_stride * sizeof(float) <= a.workspace_bytes,
                "tc_wgrad: workspace too small");
    SPX_REQUIRE(((uintptr_t)a.workspace & 15u) == 0, "tc_wgrad: workspace must be 16-byte aligned");
    dim3 grid(pl.chunks, pl.passes);
    const int cpa = pl.p.span_x >> 4, cpd = pl.p.db >> 4;
    using KernelFn = void (*)(const WgParams);
    KernelFn fn = nullptr;
    if (a.dtype == SPX_F32) {
        if (cpa == 8 && cpd == 8) fn = tc_wgrad_kernel<8, 8, true>;
        if (cpa == 8 && cpd == 16) fn = tc_wgrad_kernel<8, 16, true>;
        if (cpa == 8 && cpd == 32) fn = tc_wgrad_kernel<8, 32, true>;
    } else {
#define WG_PICK(A, D) if (cpa == A && cpd == D) fn = tc_wgrad_kernel<A, D>;
#define WG_PICK_ROW(A) WG_PICK(A, 2) WG_PICK(A, 4) WG_PICK(A, 8) WG_PICK(A, 16) WG_PICK(A, 32)
        WG_PICK_ROW(2) WG_PICK_ROW(4) WG_PICK_ROW(8)
#undef WG_PICK_ROW
#undef WG_PICK
    }
    SPX_REQUIRE(fn != nullptr, "tc_wgrad: no kernel instance for this channel layout");
    if (!func_configured((const void *)fn, current_device()))      // per-device attribute
        SPX_CHECK_CUDA(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            WG_SMEM_MAX));
    fn<<<grid, WG_THREADS, pl.smem, stream>>>(pl.p);
    SPX_CHECK_LAUNCH("tc_wgrad_kernel");
    const int64_t total = pl.p.partial_stride;
    if (a.peers)        // data-parallel: the reduction of the partials IS the send side of the exchange (peer.cu); the
                        // caller finishes it (peer_finish writes dW) after the work it wants to overlap
        return peer_push(pl.p.partial, total, pl.chunks, nullptr, total, a.dtype, a.peers, stream);
    unsigned nblk = (unsigned)div_up64(total, 128);
    if (a.dtype == SPX_F32)
        wgrad_reduce_kernel<float><<<nblk, RED_WARPS * 32, 0, stream>>>(pl.p.partial, total, pl.chunks, total, (float *)a.dw);
    else if (a.dtype == SPX_F16)
        wgrad_reduce_kernel<__half><<<nblk, RED_WARPS * 32, 0, stream>>>(pl.p.partial, total, pl.chunks, total, (__half *)a.dw);
    else
        wgrad_reduce_kernel<__nv_bfloat16><<<nblk, RED_WARPS * 32, 0, stream>>>(pl.p.partial, total, pl.chunks, total, (__nv_bfloat16 *)a.dw);
    SPX_CHECK_LAUNCH("wgrad_reduce_kernel");
    return 0;
}

}  // namespace spx
