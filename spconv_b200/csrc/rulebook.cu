// Rulebook (indice-pair) generation for SubM and regular/transposed sparse convolution.
//
// Replaces the reference's hash + atomic-append kernels (spconv/csrc/sparse/indices.py:292-939)
// and their host drivers (spconv/csrc/sparse/all.py:1660-2218).  Design differences, on purpose:
//   * one packed 64-bit slot {key:32 | value:32} per hash entry so a probe is ONE 8-byte load
//     (the reference probes split key/value arrays: "performance bound", indices.py:791);
//   * the dense tables pair_fwd / pair_bwd are written exactly once, coalesced, -1 included
//     (no torch.full(-1) pre-pass + scattered writes);
//   * every ordering decision is deterministic and equal to the reference *CPU* rulebook
//     (indices.py:1640-1778): outputs of a regular conv are ranked by first touch in
//     offset-major order (atomicMin of k*N+i, then a radix sort of the minima), compact
//     "Native" pairs come from a stable per-offset scan instead of atomicAggInc.
#include "common.cuh"
#include "gemm.cuh"
#include "hash.cuh"
#include <cub/cub.cuh>

namespace spx {
size_t radix_argsort_workspace_bytes(int64_t n);
int radix_argsort_pair(uint32_t *mask0, int32_t *argsort0, int64_t n0, uint32_t *mask1, int32_t *argsort1, int64_t n1,
                       int key_bits, void *ws0, size_t ws0_bytes, void *ws1, size_t ws1_bytes, cudaStream_t stream);
int radix_argsort(uint32_t *mask, int32_t *argsort, int64_t n, int key_bits, void *workspace, size_t workspace_bytes,
                  cudaStream_t stream);
}

namespace spx {

// ------------------------------------------------------------------ geometry
struct Geom {
    int ndim, batch, kv;
    int in_dims[SPX_MAX_NDIM], out_dims[SPX_MAX_NDIM], ksize[SPX_MAX_NDIM];
    int stride[SPX_MAX_NDIM], padding[SPX_MAX_NDIM], dilation[SPX_MAX_NDIM];
    int transposed;
};

static Geom make_geom(const spx_conv_geometry *g, bool subm) {
    Geom r;
    memset(&r, 0, sizeof(r));
    r.ndim = g->ndim;
    r.batch = g->batch_size;
    r.kv = 1;
    r.transposed = g->transposed;
    for (int a = 0; a < g->ndim; ++a) {
        r.in_dims[a] = g->in_dims[a];
        r.ksize[a] = g->ksize[a];
        r.dilation[a] = g->dilation[a];
        r.kv *= g->ksize[a];
        if (subm) {  // indices.py:1648-1657: stride 1, pad = (k/2)*dil, out dims = in dims
            r.out_dims[a] = g->in_dims[a];
            r.stride[a] = 1;
            r.padding[a] = (g->ksize[a] / 2) * g->dilation[a];
        } else {
            r.out_dims[a] = g->out_dims[a];
            r.stride[a] = g->stride[a];
            r.padding[a] = g->padding[a];
        }
    }
    return r;
}

static bool needs_i64(const Geom &g, const int *dims) {
    // same rule as the reference: int64 keys once batch * prod(dims) reaches 2^31
    // (spconv/pytorch/ops.py:188-190, ConvProblem::check_npq_not_overflow)
    double v = (double)g.batch;
    for (int a = 0; a < g.ndim; ++a) v *= (double)dims[a];
    return v >= 2147483647.0;
}

// ------------------------------------------------------------------ coordinate helpers
__device__ __forceinline__ void load_coord(const int32_t *indices, int64_t i, int ndim, int (&c)[SPX_MAX_NDIM + 1]) {
    if (ndim == 3) {
        int4 v = __ldg(reinterpret_cast<const int4 *>(indices) + i);
        c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
    } else {
        const int32_t *p = indices + i * (ndim + 1);
#pragma unroll
        for (int a = 0; a <= SPX_MAX_NDIM; ++a) if (a <= ndim) c[a] = __ldg(p + a);
    }
}

__device__ __forceinline__ int64_t linear_key(const int (&c)[SPX_MAX_NDIM + 1], const int *dims, int ndim) {
    int64_t k = c[0];
#pragma unroll
    for (int a = 0; a < SPX_MAX_NDIM; ++a) if (a < ndim) k = k * dims[a] + c[a + 1];
    return k;
}

__device__ __forceinline__ void offset_taps(int k, const int *ksize, int ndim, int (&r)[SPX_MAX_NDIM]) {
#pragma unroll
    for (int a = SPX_MAX_NDIM - 1; a >= 0; --a) if (a < ndim) { r[a] = k % ksize[a]; k /= ksize[a]; }
}

// ------------------------------------------------------------------ SubM
template <typename Table>
__global__ void subm_insert_kernel(Table table, Geom g, const int32_t *__restrict__ indices, int64_t N) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    int c[SPX_MAX_NDIM + 1];
    load_coord(indices, i, g.ndim, c);
    table.insert_min(linear_key(c, g.in_dims, g.ndim), (int32_t)i);
}

// one thread per voxel, all kv offsets: pair_fwd[k][o] = index of the voxel at
// coord(o) - pad + r_k * dil (query_nhw, indices.py:222-236), coalesced writes, no atomics.
template <typename Table>
__global__ void subm_probe_kernel(Table table, Geom g, const int32_t *__restrict__ indices, int64_t N,
                                  int32_t *__restrict__ pair_fwd, int32_t *__restrict__ pair_bwd,
                                  uint32_t *__restrict__ mask, int words) {
    int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (o >= N) return;
    int c[SPX_MAX_NDIM + 1];
    load_coord(indices, o, g.ndim, c);
    const int kv = g.kv;
    uint32_t mword = 0;
    int r[SPX_MAX_NDIM] = {0, 0, 0, 0};
    for (int k = 0; k < kv; ++k) {
        int32_t found = -1;
        if (k == kv / 2) {
            found = (int32_t)o;   // centre: identity (indices.py:1671-1676)
        } else {
            int q[SPX_MAX_NDIM + 1];
            q[0] = c[0];
            bool valid = c[0] >= 0 && c[0] < g.batch;
#pragma unroll
            for (int a = 0; a < SPX_MAX_NDIM; ++a) {
                if (a < g.ndim) {
                    q[a + 1] = c[a + 1] - g.padding[a] + r[a] * g.dilation[a];
                    valid = valid && q[a + 1] >= 0 && q[a + 1] < g.in_dims[a];
                }
            }
            if (valid) {
                int32_t v;
                if (table.find_slot(linear_key(q, g.in_dims, g.ndim), v) >= 0) found = v;
            }
        }
        pair_fwd[(int64_t)k * N + o] = found;
        if (pair_bwd) pair_bwd[(int64_t)(kv - 1 - k) * N + o] = found;
        if (found >= 0) mword |= 1u << (k & 31);
        if (mask && ((k & 31) == 31 || k == kv - 1)) {
            mask[o * words + (k >> 5)] = mword;
            mword = 0;
        }
        // advance taps row-major, last axis fastest (ConvOutLocIter::operator++, indices.py:117-127)
#pragma unroll
        for (int a = SPX_MAX_NDIM - 1; a >= 0; --a) {
            if (a < g.ndim) {
                if (++r[a] < g.ksize[a]) break;
                r[a] = 0;
            }
        }
    }
}

// 3-D, 3x3x3 (any dilation), 32-bit keys -- the shape of every SubMConv3d in SECOND-style nets.
// (The generic kernel above exposes one L2 round trip per offset and spends ~150 instructions per
// probe on generic n-d index arithmetic.)
// A warp probes ONE kernel offset for K3_VOX = 128 consecutive voxels (four per lane), a block
// (27 warps) covers all offsets of those voxels.  The kernel is a chain of dependent L2 round
// trips (coordinates -> slot -> collision chain -> stores); per-offset state is one key / slot per
// voxel, so every thread keeps four independent probes in flight at full occupancy, and chains
// advance in rounds (one extra round trip per round for the whole warp, not per probe).  The
// warp's hit ballots are the columns of mask bits of its offset; thread v assembles voxel v's mask
// word from the 27 ballots.  Rows of the table are staged in shared memory for the row-major copy.
constexpr int K3_VOX = 128;                    // voxels per block; block = 27 warps
constexpr int K3_PER_LANE = K3_VOX / 32;
__global__ void __launch_bounds__(27 * 32, 2)
subm_probe_k3_kernel(Table32 table, Geom g, const int32_t *__restrict__ indices, int64_t N,
                     int32_t *__restrict__ pair_fwd, int32_t *__restrict__ pair_bwd,
                     uint32_t *__restrict__ mask, int32_t *__restrict__ row_table) {
    __shared__ uint32_t hit_col[27][K3_PER_LANE];
    __shared__ int32_t row_stage[K3_VOX][27];     // odd row stride: conflict-free both ways
    const int lane = threadIdx.x & 31;
    const int k = threadIdx.x >> 5;               // warp-uniform kernel offset, k = (rz*3 + ry)*3 + rx
    const int rz = k / 9, ry = (k / 3) % 3, rx = k % 3;
    const int64_t vbase = blockIdx.x * (int64_t)K3_VOX;
    const int D0 = g.in_dims[0], D1 = g.in_dims[1], D2 = g.in_dims[2];
    // q_a = c_a + (r_a - 1) * dil_a   (pad = dil for ksize 3), r_a in {0,1,2}
    const int oz = (rz - 1) * g.dilation[0], oy = (ry - 1) * g.dilation[1], ox = (rx - 1) * g.dilation[2];
    int4 c[K3_PER_LANE];
#pragma unroll
    for (int j = 0; j < K3_PER_LANE; ++j) {
        const int64_t o = vbase + j * 32 + lane;
        c[j] = o < N ? __ldg(reinterpret_cast<const int4 *>(indices) + o) : make_int4(-1, 0, 0, 0);
    }
    uint32_t key[K3_PER_LANE], h[K3_PER_LANE];
    unsigned long long cur[K3_PER_LANE];
    uint32_t pending = 0;
#pragma unroll
    for (int j = 0; j < K3_PER_LANE; ++j) {
        const int qz = c[j].y + oz, qy = c[j].z + oy, qx = c[j].w + ox;
        const bool valid = k != 13 && c[j].x >= 0 && c[j].x < g.batch && qz >= 0 && qz < D0 && qy >= 0 && qy < D1 &&
                           qx >= 0 && qx < D2;
        key[j] = (uint32_t)(((c[j].x * D0 + qz) * D1 + qy) * D2 + qx);
        h[j] = mix32(key[j]) & table.cap_mask;
        cur[j] = valid ? __ldg(&table.slots[h[j]]) : Table32::EMPTY;
    }
#pragma unroll
    for (int j = 0; j < K3_PER_LANE; ++j)
        if (cur[j] != Table32::EMPTY && (uint32_t)(cur[j] >> 32) != key[j]) pending |= 1u << j;
    while (pending) {                             // collision chains (short at load factor <= 0.25)
#pragma unroll
        for (int j = 0; j < K3_PER_LANE; ++j)
            if (pending & (1u << j)) { h[j] = (h[j] + 1) & table.cap_mask; cur[j] = __ldg(&table.slots[h[j]]); }
#pragma unroll
        for (int j = 0; j < K3_PER_LANE; ++j)
            if ((pending & (1u << j)) && (cur[j] == Table32::EMPTY || (uint32_t)(cur[j] >> 32) == key[j]))
                pending &= ~(1u << j);
    }
#pragma unroll
    for (int j = 0; j < K3_PER_LANE; ++j) {
        const int64_t o = vbase + j * 32 + lane;
        int32_t found = cur[j] != Table32::EMPTY ? (int32_t)(uint32_t)cur[j] : -1;
        if (k == 13 && o < N) found = (int32_t)o;     // centre: identity
        if (o < N) {
            pair_fwd[(int64_t)k * N + o] = found;
            if (pair_bwd) pair_bwd[(int64_t)(26 - k) * N + o] = found;
        }
        const uint32_t hits = __ballot_sync(0xffffffffu, found >= 0);
        if (lane == 0) hit_col[k][j] = hits;
        row_stage[j * 32 + lane][k] = found;
    }
    __syncthreads();
    if (mask && threadIdx.x < K3_VOX && vbase + threadIdx.x < N) {
        const int v = threadIdx.x;
        uint32_t m = 0;
#pragma unroll
        for (int kk = 0; kk < 27; ++kk) m |= ((hit_col[kk][v >> 5] >> (v & 31)) & 1u) << kk;
        mask[vbase + v] = m;
    }
    if (row_table) {
        // row-major copy [N][32] (one 128-byte line per voxel, -1 padded) for spx_build_tile_table:
        // the permuted re-read then costs 4 sectors per row instead of one per (row, offset)
        for (int e = threadIdx.x; e < K3_VOX * 32; e += 27 * 32) {
            const int v = e >> 5, kk = e & 31;
            if (vbase + v < N) row_table[(vbase + v) * 32 + kk] = kk < 27 ? row_stage[v][kk] : -1;
        }
    }
}

// ------------------------------------------------------------------ regular / transposed conv
__device__ __forceinline__ bool conv_out_coord(const Geom &g, const int (&c)[SPX_MAX_NDIM + 1],
                                               const int (&r)[SPX_MAX_NDIM], int (&o)[SPX_MAX_NDIM + 1]) {
    bool valid = c[0] >= 0 && c[0] < g.batch;
    o[0] = c[0];
#pragma unroll
    for (int a = 0; a < SPX_MAX_NDIM; ++a) {
        if (a < g.ndim) {
            if (g.transposed) {   // query_nhw_out, indices.py:253-269
                o[a + 1] = c[a + 1] * g.stride[a] - g.padding[a] + r[a] * g.dilation[a];
                valid = valid && o[a + 1] >= 0 && o[a + 1] < g.out_dims[a];
            } else {              // query_npq, indices.py:141-203
                int h = c[a + 1] + g.padding[a] - r[a] * g.dilation[a];
                o[a + 1] = h / g.stride[a];
                valid = valid && o[a + 1] >= 0 && o[a + 1] < g.out_dims[a] && (h % g.stride[a] == 0);
            }
        }
    }
    return valid;
}

// 3-D, non-transposed fast path of query_npq (indices.py:141-203): the taps of the block's offset
// are decoded once per block, the stride division is a shift for strides 1 and 2, everything else
// is three fused range tests.  (The generic helpers spend ~150 instructions per (input, offset)
// on run-time-ndim loops and two integer divisions per axis.)
struct Taps3 { int r0, r1, r2; };
__device__ __forceinline__ Taps3 block_taps3(const Geom &g, int k) {
    Taps3 t;
    t.r2 = k % g.ksize[2]; k /= g.ksize[2];
    t.r1 = k % g.ksize[1]; k /= g.ksize[1];
    t.r0 = k;
    return t;
}
__device__ __forceinline__ bool axis_out(int c, int pad, int r, int dil, int stride, int odim, int &o) {
    const int h = c + pad - r * dil;
    if (h < 0) return false;
    if (stride == 1) o = h;
    else if (stride == 2) { if (h & 1) return false; o = h >> 1; }
    else { o = h / stride; if (o * stride != h) return false; }
    return o < odim;
}
__device__ __forceinline__ bool conv3_out_key(const Geom &g, const int4 c, const Taps3 &t, int64_t &key) {
    int o0, o1, o2;
    if (c.x < 0 || c.x >= g.batch) return false;
    if (!axis_out(c.y, g.padding[0], t.r0, g.dilation[0], g.stride[0], g.out_dims[0], o0)) return false;
    if (!axis_out(c.z, g.padding[1], t.r1, g.dilation[1], g.stride[1], g.out_dims[1], o1)) return false;
    if (!axis_out(c.w, g.padding[2], t.r2, g.dilation[2], g.stride[2], g.out_dims[2], o2)) return false;
    key = (((int64_t)c.x * g.out_dims[0] + o0) * g.out_dims[1] + o1) * g.out_dims[2] + o2;
    return true;
}

// grid (ceil(N/T), kv): hash every hit, payload = k*N + i (first touch in offset-major order)
template <typename Table, bool FAST3>
__global__ void conv_insert_kernel(Table table, Geom g, const int32_t *__restrict__ indices, int64_t N) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int k = blockIdx.y;
    if constexpr (FAST3) {
        __shared__ Taps3 taps;
        if (threadIdx.x == 0) taps = block_taps3(g, k);
        __syncthreads();
        if (i >= N) return;
        const int4 c = __ldg(reinterpret_cast<const int4 *>(indices) + i);
        int64_t key;
        if (conv3_out_key(g, c, taps, key)) table.insert_min(key, (int32_t)((int64_t)k * N + i));
    } else {
        if (i >= N) return;
        int c[SPX_MAX_NDIM + 1], o[SPX_MAX_NDIM + 1], r[SPX_MAX_NDIM];
        load_coord(indices, i, g.ndim, c);
        offset_taps(k, g.ksize, g.ndim, r);
        if (conv_out_coord(g, c, r, o))
            table.insert_min(linear_key(o, g.out_dims, g.ndim), (int32_t)((int64_t)k * N + i));
    }
}

// 3-D, 3x3x3, non-transposed: one thread per INPUT voxel walks the 27 offsets.  The per-axis output
// coordinates of the three taps are computed once (9 divisions-by-stride instead of 81), and with a
// stride > 1 most (axis, tap) pairs fail the divisibility test, so only the surviving combinations
// (3.4 of 27 on average at stride 2) reach the hash table.  The grid-(N, kv) kernels above re-read
// the coordinates 27 times and spend a thread per rejected combination.
struct Axis3 { int o[3]; };
__device__ __forceinline__ Axis3 axis_taps3(int c, int pad, int dil, int stride, int odim) {
    Axis3 a;
#pragma unroll
    for (int r = 0; r < 3; ++r) { int o; a.o[r] = axis_out(c, pad, r, dil, stride, odim, o) ? o : -1; }
    return a;
}

template <typename Table>
__global__ void conv_insert_k3_kernel(Table table, Geom g, const int32_t *__restrict__ indices, int64_t N) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int4 c = __ldg(reinterpret_cast<const int4 *>(indices) + i);
    if (c.x < 0 || c.x >= g.batch) return;
    const Axis3 az = axis_taps3(c.y, g.padding[0], g.dilation[0], g.stride[0], g.out_dims[0]);
    const Axis3 ay = axis_taps3(c.z, g.padding[1], g.dilation[1], g.stride[1], g.out_dims[1]);
    const Axis3 ax = axis_taps3(c.w, g.padding[2], g.dilation[2], g.stride[2], g.out_dims[2]);
#pragma unroll
    for (int r0 = 0; r0 < 3; ++r0) {
        if (az.o[r0] < 0) continue;
        const int64_t kz = (int64_t)c.x * g.out_dims[0] + az.o[r0];
#pragma unroll
        for (int r1 = 0; r1 < 3; ++r1) {
            if (ay.o[r1] < 0) continue;
            const int64_t kzy = kz * g.out_dims[1] + ay.o[r1];
#pragma unroll
            for (int r2 = 0; r2 < 3; ++r2) {
                if (ax.o[r2] < 0) continue;
                const int k = (r0 * 3 + r1) * 3 + r2;
                table.insert_min(kzy * g.out_dims[2] + ax.o[r2], (int32_t)((int64_t)k * N + i));
            }
        }
    }
}

// pair_bwd[k][i] = output of (input i, offset k) or -1 (every element written, coalesced over i);
// pair_fwd[k][o] = i scattered; mask_bwd[i] (optional) = bits of the offsets that hit an output
template <typename Table>
__global__ void conv_pairs_k3_kernel(Table table, Geom g, const int32_t *__restrict__ indices, int64_t N, int64_t M,
                                     int32_t *__restrict__ pair_fwd, int32_t *__restrict__ pair_bwd,
                                     uint32_t *__restrict__ mask_bwd, uint32_t *__restrict__ mask_fwd_or) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int4 c = __ldg(reinterpret_cast<const int4 *>(indices) + i);
    const bool bok = c.x >= 0 && c.x < g.batch;
    const Axis3 az = axis_taps3(c.y, g.padding[0], g.dilation[0], g.stride[0], g.out_dims[0]);
    const Axis3 ay = axis_taps3(c.z, g.padding[1], g.dilation[1], g.stride[1], g.out_dims[1]);
    const Axis3 ax = axis_taps3(c.w, g.padding[2], g.dilation[2], g.stride[2], g.out_dims[2]);
    uint32_t mword = 0;
#pragma unroll
    for (int r0 = 0; r0 < 3; ++r0) {
#pragma unroll
        for (int r1 = 0; r1 < 3; ++r1) {
#pragma unroll
            for (int r2 = 0; r2 < 3; ++r2) {
                const int k = (r0 * 3 + r1) * 3 + r2;
                int32_t out = -1;
                if (bok && az.o[r0] >= 0 && ay.o[r1] >= 0 && ax.o[r2] >= 0) {
                    const int64_t key = (((int64_t)c.x * g.out_dims[0] + az.o[r0]) * g.out_dims[1] + ay.o[r1]) *
                                            g.out_dims[2] + ax.o[r2];
                    int32_t v;
                    if (table.find_slot(key, v) >= 0) out = v;
                }
                pair_bwd[(int64_t)k * N + i] = out;
                if (out >= 0) {
                    pair_fwd[(int64_t)k * M + out] = (int32_t)i;
                    mword |= 1u << k;
                    if (mask_fwd_or) atomicOr(&mask_fwd_or[out], 1u << k);   // zeroed by conv_assign_rank_kernel
                }
            }
        }
    }
    if (mask_bwd) mask_bwd[i] = mword;
}


// ---- append-on-create variants (default path): the thread whose CAS creates a table entry records
// the slot, so the distinct outputs are known without scanning the (mostly empty) table afterwards.
// Created slots are staged in shared memory and a block reserves its range of the global list with
// ONE atomic.  `state`: [0] number of created entries, [1] overflow flag (a probe chain exceeded
// CONV_MAX_PROBES: the optimistically sized table was too small, the host re-runs with the full size).
constexpr int CONV_MAX_PROBES = 96;
constexpr int APPEND_THREADS = 128;

template <typename Table>
__global__ void __launch_bounds__(APPEND_THREADS)
conv_insert_k3_append_kernel(Table table, Geom g, const int32_t *__restrict__ indices, int64_t N,
                             uint32_t *__restrict__ slot_list, int *__restrict__ state) {
    __shared__ uint32_t stage[APPEND_THREADS * 27];
    __shared__ int cnt_s, base_s;
    if (threadIdx.x == 0) cnt_s = 0;
    __syncthreads();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < N) {
        const int4 c = __ldg(reinterpret_cast<const int4 *>(indices) + i);
        if (c.x >= 0 && c.x < g.batch) {
            const Axis3 az = axis_taps3(c.y, g.padding[0], g.dilation[0], g.stride[0], g.out_dims[0]);
            const Axis3 ay = axis_taps3(c.z, g.padding[1], g.dilation[1], g.stride[1], g.out_dims[1]);
            const Axis3 ax = axis_taps3(c.w, g.padding[2], g.dilation[2], g.stride[2], g.out_dims[2]);
#pragma unroll
            for (int r0 = 0; r0 < 3; ++r0) {
                if (az.o[r0] < 0) continue;
                const int64_t kz = (int64_t)c.x * g.out_dims[0] + az.o[r0];
#pragma unroll
                for (int r1 = 0; r1 < 3; ++r1) {
                    if (ay.o[r1] < 0) continue;
                    const int64_t kzy = kz * g.out_dims[1] + ay.o[r1];
#pragma unroll
                    for (int r2 = 0; r2 < 3; ++r2) {
                        if (ax.o[r2] < 0) continue;
                        const int k = (r0 * 3 + r1) * 3 + r2;
                        bool created;
                        const int64_t slot = table.insert_min_slot(kzy * g.out_dims[2] + ax.o[r2],
                                                                   (int32_t)((int64_t)k * N + i), created, CONV_MAX_PROBES);
                        if (slot < 0) state[1] = 1;
                        else if (created) stage[atomicAdd(&cnt_s, 1)] = (uint32_t)slot;
                    }
                }
            }
        }
    }
    __syncthreads();
    const int cnt = cnt_s;
    if (threadIdx.x == 0) base_s = cnt ? atomicAdd(state, cnt) : 0;
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += APPEND_THREADS) slot_list[base_s + j] = stage[j];
}

// grid (ceil(N/T), kv): one (input, offset) per thread -> at most one creation per thread
template <typename Table, bool FAST3>
__global__ void __launch_bounds__(APPEND_THREADS)
conv_insert_append_kernel(Table table, Geom g, const int32_t *__restrict__ indices, int64_t N,
                          uint32_t *__restrict__ slot_list, int *__restrict__ state) {
    __shared__ int warp_cnt[APPEND_THREADS / 32];
    __shared__ int base_s;
    __shared__ Taps3 taps;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (FAST3 && threadIdx.x == 0) taps = block_taps3(g, k);
    __syncthreads();
    bool created = false;
    int64_t slot = 0;
    if (i < N) {
        int64_t key = 0;
        bool valid;
        if constexpr (FAST3) {
            const int4 c = __ldg(reinterpret_cast<const int4 *>(indices) + i);
            valid = conv3_out_key(g, c, taps, key);
        } else {
            int c[SPX_MAX_NDIM + 1], o[SPX_MAX_NDIM + 1], r[SPX_MAX_NDIM];
            load_coord(indices, i, g.ndim, c);
            offset_taps(k, g.ksize, g.ndim, r);
            valid = conv_out_coord(g, c, r, o);
            if (valid) key = linear_key(o, g.out_dims, g.ndim);
        }
        if (valid) {
            slot = table.insert_min_slot(key, (int32_t)((int64_t)k * N + i), created, CONV_MAX_PROBES);
            if (slot < 0) { state[1] = 1; created = false; }
        }
    }
    const unsigned ball = __ballot_sync(0xffffffffu, created);
    if (lane == 0) warp_cnt[warp] = __popc(ball);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < APPEND_THREADS / 32; ++w) { const int c = warp_cnt[w]; warp_cnt[w] = tot; tot += c; }
        base_s = tot ? atomicAdd(state, tot) : 0;
    }
    __syncthreads();
    if (created) slot_list[base_s + warp_cnt[warp] + __popc(ball & ((1u << lane) - 1u))] = (uint32_t)slot;
}

// ---- ranking the outputs by first touch WITHOUT a sort.  Every output's final payload p = k*N + i
// (the smallest (offset, input) pair that produces it) is distinct, so the rank of an output is the
// number of outputs with a smaller payload = the number of set bits below p in a bitmap over
// [0, kv*N).  mark: one atomicOr per output (+ a count per 1024-bit tile); scan: exclusive prefix over
// the tile counts (one block); assign: tile prefix + popcount of at most 32 words.  Replaces the
// radix sort of the payloads (1 + 2 x 3 launches at 22 key bits) by one single-block kernel.
constexpr int RANK_TILE_WORDS = 32;              // bitmap words per counted tile (1024 payloads)
static size_t rank_scratch_bytes(int64_t kvn, int64_t *ntiles = nullptr) {
    const int64_t words = div_up64(kvn > 0 ? kvn : 1, 32), tiles = div_up64(words, RANK_TILE_WORDS);
    if (ntiles) *ntiles = tiles;
    return align_up((size_t)(tiles * RANK_TILE_WORDS + tiles + 1) * 4, 256);
}

// one launch clears everything stage 1 needs: hash table (0xFF), 64-bit-key value array (0x7F), ranking scratch and counters (0)
__global__ void conv_clear_kernel(uint4 *__restrict__ table, int64_t table_vec, uint4 *__restrict__ tvals, int64_t tvals_vec,
                                  uint4 *__restrict__ zero, int64_t zero_vec) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint4 ff = make_uint4(~0u, ~0u, ~0u, ~0u), sf = make_uint4(0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu),
                zz = make_uint4(0u, 0u, 0u, 0u);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < table_vec; i += stride) table[i] = ff;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < tvals_vec; i += stride) tvals[i] = sf;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < zero_vec; i += stride) zero[i] = zz;
}

constexpr int MARK_THREADS = 256;
// mark every output's final payload; the LAST block to finish turns the tile counts into an exclusive prefix
template <typename Table>
__global__ void __launch_bounds__(MARK_THREADS)
conv_mark_kernel(Table table, const uint32_t *__restrict__ slot_list, int64_t M, uint32_t *__restrict__ bitmap,
                 int *__restrict__ tile_cnt, int64_t tiles, int *__restrict__ done) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j < M) {
        const uint32_t p = (uint32_t)table.value_at(slot_list[j]);   // final: all inserts finished in an earlier kernel
        atomicOr(&bitmap[p >> 5], 1u << (p & 31));
        atomicAdd(&tile_cnt[p >> 10], 1);
    }
    __shared__ int warp_sums[MARK_THREADS / 32];
    __shared__ int carry_s, last_s;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) { last_s = atomicAdd(done, 1) == (int)gridDim.x - 1; carry_s = 0; }
    __syncthreads();
    if (!last_s) return;
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t t0 = 0; t0 < tiles; t0 += MARK_THREADS) {
        const int64_t t = t0 + threadIdx.x;
        const int v = t < tiles ? __ldcg(tile_cnt + t) : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += warp_sums[w];
        const int carry = carry_s;
        if (t < tiles) tile_cnt[t] = carry + wbase + incl - v;
        __syncthreads();
        if (threadIdx.x == MARK_THREADS - 1) carry_s = carry + wbase + incl;
        __syncthreads();
    }
}

// created slot j -> rank r of its payload: write r into the slot, decode the key into out_inds[r]
template <typename Table>
__global__ void conv_assign_rank_kernel(Table table, Geom g, const uint32_t *__restrict__ slot_list, int64_t M,
                                        const uint32_t *__restrict__ bitmap, const int *__restrict__ tile_prefix,
                                        int32_t *__restrict__ out_inds, uint32_t *__restrict__ mask_zero,
                                        int32_t *__restrict__ pair_fill, int kv) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= M) return;
    if (pair_fill)                                  // column j of the forward table: -1 until the pairs kernel fills it
        for (int k = 0; k < kv; ++k) pair_fill[(int64_t)k * M + j] = -1;
    const uint32_t s = slot_list[j];
    int64_t key; int32_t val;
    table.occupied(s, key, val);
    const uint32_t p = (uint32_t)val, word = p >> 5, first = word & ~(uint32_t)(RANK_TILE_WORDS - 1);
    int r = __ldg(tile_prefix + (p >> 10)) + __popc(__ldg(bitmap + word) & ((1u << (p & 31)) - 1u));
    for (uint32_t wd = first; wd < word; ++wd) r += __popc(__ldg(bitmap + wd));
    table.set_value(s, r);
    if (mask_zero) mask_zero[r] = 0u;              // the pairs kernel ORs the forward masks into it
    int32_t *dst = out_inds + (int64_t)r * (g.ndim + 1);
    for (int a = g.ndim - 1; a >= 0; --a) {
        dst[a + 1] = (int32_t)(key % g.out_dims[a]);
        key /= g.out_dims[a];
    }
    dst[0] = (int32_t)key;
}

// compact occupied slots -> (first-touch payload, slot); order irrelevant (sorted next).
// The table is sized for the worst-case output count, so most of it is empty: every thread scans
// COLLECT_ITEMS slots and a block reserves its output range with ONE atomic (a single global
// counter takes ~2.7 G same-address atomics/s on B200; one per warp made this kernel the slowest
// of the conv rulebook).
constexpr int COLLECT_THREADS = 256;
constexpr int COLLECT_ITEMS = 4;
template <typename Table>
__global__ void __launch_bounds__(COLLECT_THREADS)
conv_collect_kernel(Table table, uint32_t capacity, uint32_t *__restrict__ payload,
                    uint32_t *__restrict__ slot_of, int *__restrict__ counter) {
    __shared__ int warp_cnt[COLLECT_THREADS / 32];
    __shared__ int block_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t first = blockIdx.x * (uint32_t)(COLLECT_THREADS * COLLECT_ITEMS);
    bool occ[COLLECT_ITEMS];
    int32_t val[COLLECT_ITEMS];
    unsigned ballot[COLLECT_ITEMS];
    int mine = 0;                                   // occupied slots seen by this warp
#pragma unroll
    for (int j = 0; j < COLLECT_ITEMS; ++j) {
        const uint32_t s = first + j * COLLECT_THREADS + threadIdx.x;
        int64_t key;
        val[j] = 0;
        occ[j] = s < capacity && table.occupied(s, key, val[j]);
        ballot[j] = __ballot_sync(0xffffffffu, occ[j]);
        mine += __popc(ballot[j]);
    }
    if (lane == 0) warp_cnt[warp] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < COLLECT_THREADS / 32; ++w) { const int c = warp_cnt[w]; warp_cnt[w] = tot; tot += c; }
        block_base = tot ? atomicAdd(counter, tot) : 0;
    }
    __syncthreads();
    int base = block_base + warp_cnt[warp];
#pragma unroll
    for (int j = 0; j < COLLECT_ITEMS; ++j) {
        if (occ[j]) {
            const int pos = base + __popc(ballot[j] & ((1u << lane) - 1));
            payload[pos] = (uint32_t)val[j];
            slot_of[pos] = first + j * COLLECT_THREADS + threadIdx.x;
        }
        base += __popc(ballot[j]);
    }
}

// rank r (first-touch order) -> write r into its slot, decode the key into out_inds[r]
template <typename Table>
__global__ void conv_assign_kernel(Table table, Geom g, const uint32_t *__restrict__ sorted_slot,
                                   const int32_t *__restrict__ order, int64_t M, int32_t *__restrict__ out_inds) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= M) return;
    // order != NULL: sorted_slot is the UNSORTED slot list and order the argsort of its payloads
    uint32_t s = order ? sorted_slot[order[r]] : sorted_slot[r];
    int64_t key; int32_t val;
    table.occupied(s, key, val);
    table.set_value(s, (int32_t)r);
    int32_t *dst = out_inds + r * (g.ndim + 1);
    for (int a = g.ndim - 1; a >= 0; --a) {
        dst[a + 1] = (int32_t)(key % g.out_dims[a]);
        key /= g.out_dims[a];
    }
    dst[0] = (int32_t)key;
}

// grid (ceil(N/T), kv): pair_bwd[k][i] = o (every element written), pair_fwd[k][o] = i
template <typename Table, bool FAST3>
__global__ void conv_pairs_kernel(Table table, Geom g, const int32_t *__restrict__ indices, int64_t N, int64_t M,
                                  int32_t *__restrict__ pair_fwd, int32_t *__restrict__ pair_bwd) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int k = blockIdx.y;
    int32_t out = -1;
    if constexpr (FAST3) {
        __shared__ Taps3 taps;
        if (threadIdx.x == 0) taps = block_taps3(g, k);
        __syncthreads();
        if (i >= N) return;
        const int4 c = __ldg(reinterpret_cast<const int4 *>(indices) + i);
        int64_t key;
        int32_t v;
        if (conv3_out_key(g, c, taps, key) && table.find_slot(key, v) >= 0) out = v;
    } else {
        if (i >= N) return;
        int c[SPX_MAX_NDIM + 1], o[SPX_MAX_NDIM + 1], r[SPX_MAX_NDIM];
        load_coord(indices, i, g.ndim, c);
        offset_taps(k, g.ksize, g.ndim, r);
        if (conv_out_coord(g, c, r, o)) {
            int32_t v;
            if (table.find_slot(linear_key(o, g.out_dims, g.ndim), v) >= 0) out = v;
        }
    }
    pair_bwd[(int64_t)k * N + i] = out;
    if (out >= 0) pair_fwd[(int64_t)k * M + out] = (int32_t)i;
}

// mask[row] = bits of the non-negative entries of table[:, row]
__global__ void table_mask_kernel(const int32_t *__restrict__ table, int64_t rows, int kv, int words,
                                  uint32_t *__restrict__ mask) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= rows) return;
    uint32_t m = 0;
    for (int k = 0; k < kv; ++k) {
        if (table[(int64_t)k * rows + r] >= 0) m |= 1u << (k & 31);
        if ((k & 31) == 31 || k == kv - 1) { mask[r * words + (k >> 5)] = m; m = 0; }
    }
}

// ------------------------------------------------------------------ Native compact pairs (stable scan)
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// rows: regular conv -> kv rows (row k of pair_bwd); SubM -> kv/2 rows
__global__ void native_count_kernel(const int32_t *__restrict__ pair_bwd, int64_t N, int *__restrict__ block_counts,
                                    int nblk) {
    int row = blockIdx.y;
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        int64_t i = base + j * SCAN_THREADS + threadIdx.x;
        if (i < N && pair_bwd[(int64_t)row * N + i] >= 0) ++cnt;
    }
    typedef cub::BlockReduce<int, SCAN_THREADS> BR;
    __shared__ typename BR::TempStorage tmp;
    int total = BR(tmp).Sum(cnt);
    if (threadIdx.x == 0) block_counts[row * nblk + blockIdx.x] = total;
}

// one block per row: exclusive scan of the block counts, total -> num[row]
__global__ void native_scan_kernel(int *__restrict__ block_counts, int nblk, int32_t *__restrict__ num) {
    int row = blockIdx.x;
    typedef cub::BlockScan<int, SCAN_THREADS> BS;
    __shared__ typename BS::TempStorage tmp;
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nblk; b0 += SCAN_THREADS) {
        int b = b0 + threadIdx.x;
        int v = b < nblk ? block_counts[row * nblk + b] : 0;
        int excl, agg;
        BS(tmp).ExclusiveSum(v, excl, agg);
        int carry = carry_s;
        if (b < nblk) block_counts[row * nblk + b] = carry + excl;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + agg;
        __syncthreads();
    }
    if (threadIdx.x == 0) num[row] = carry_s;
}

__global__ void native_write_kernel(const int32_t *__restrict__ pair_bwd, int64_t N, int kv, int is_subm,
                                    const int *__restrict__ block_offsets, int nblk, int32_t *__restrict__ pairs) {
    int row = blockIdx.y;
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    // blocked arrangement keeps ascending-i order inside the tile
    int32_t vals[SCAN_ITEMS];
    int flags[SCAN_ITEMS];
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        int64_t i = base + (int64_t)threadIdx.x * SCAN_ITEMS + j;
        vals[j] = i < N ? pair_bwd[(int64_t)row * N + i] : -1;
        flags[j] = vals[j] >= 0;
        cnt += flags[j];
    }
    typedef cub::BlockScan<int, SCAN_THREADS> BS;
    __shared__ typename BS::TempStorage tmp;
    int excl;
    BS(tmp).ExclusiveSum(cnt, excl);
    int pos = block_offsets[row * nblk + blockIdx.x] + excl;
    int32_t *pin = pairs, *pout = pairs + (int64_t)kv * N;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        if (flags[j]) {
            int32_t i = (int32_t)(base + (int64_t)threadIdx.x * SCAN_ITEMS + j);
            int32_t o = vals[j];
            pin[(int64_t)row * N + pos] = i;
            pout[(int64_t)row * N + pos] = o;
            if (is_subm) {   // mirrored entry, indices.py:1696-1699
                pin[(int64_t)(kv - 1 - row) * N + pos] = o;
                pout[(int64_t)(kv - 1 - row) * N + pos] = i;
            }
            ++pos;
        }
    }
}

__global__ void subm_centre_kernel(int32_t *__restrict__ pairs, int64_t N, int kv) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    pairs[(int64_t)(kv / 2) * N + i] = (int32_t)i;
    pairs[(int64_t)kv * N + (int64_t)(kv / 2) * N + i] = (int32_t)i;
}

// compact pairs -> dense tables (ConvAlgo.Native operator path)
__global__ void pairs_to_table_kernel(const int32_t *__restrict__ pairs, const int32_t *__restrict__ num, int kv,
                                      int64_t pair_stride, int64_t n_in, int64_t n_out, int is_subm, int inverse,
                                      int32_t *__restrict__ table_fwd, int32_t *__restrict__ table_bwd,
                                      uint32_t *__restrict__ mask_fwd, uint32_t *__restrict__ mask_bwd, int words) {
    int k = blockIdx.y;
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t cnt;
    if (is_subm) {   // mirror rule, spconv/pytorch/ops.py:962-968
        if (k == kv / 2) cnt = n_in;
        else cnt = k < kv / 2 ? num[k] : num[kv - 1 - k];
    } else {
        cnt = num[k];
    }
    if (j >= cnt) return;
    int32_t a = pairs[(int64_t)k * pair_stride + j];
    int32_t b = pairs[(int64_t)kv * pair_stride + (int64_t)k * pair_stride + j];
    int32_t i = inverse ? b : a, o = inverse ? a : b;
    if (i < 0 || o < 0 || i >= n_in || o >= n_out) return;
    if (table_fwd) table_fwd[(int64_t)k * n_out + o] = i;
    if (table_bwd) table_bwd[(int64_t)k * n_in + i] = o;
    if (mask_fwd) atomicOr(&mask_fwd[o * words + (k >> 5)], 1u << (k & 31));
    if (mask_bwd) atomicOr(&mask_bwd[i * words + (k >> 5)], 1u << (k & 31));
}

// ------------------------------------------------------------------ argsort helpers
__global__ void iota_kernel(int32_t *__restrict__ p, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int32_t)i;
}
__global__ void gather_word_kernel(const uint32_t *__restrict__ mask, const int32_t *__restrict__ perm, int64_t n,
                                   int words, int w, uint32_t *__restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = mask[(int64_t)perm[i] * words + w];
}
__global__ void gather_rows_kernel(const uint32_t *__restrict__ src, const int32_t *__restrict__ perm, int64_t n,
                                   int words, uint32_t *__restrict__ dst) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int w = 0; w < words; ++w) dst[i * words + w] = src[(int64_t)perm[i] * words + w];
}

// ------------------------------------------------------------------ tile-blocked gather table
// one block per 128-row tile; see include/spconv_b200.h (spx_build_tile_table)
constexpr int TT_SPLIT = 4;                     // threads per tile row: offsets k = q, q + 4, ...
// (blockIdx.y picks one of up to two jobs: a regular conv builds its forward and backward tables in one launch)
struct TtJob {
    const int32_t *pair; int64_t pair_stride;
    const int32_t *argsort; const uint32_t *mask;
    int64_t rows;
    int32_t *table; uint32_t *tile_mask;
};
struct TtJobs { TtJob j[2]; };

__global__ void __launch_bounds__(128 * TT_SPLIT)
build_tile_table_kernel(const TtJobs jobs, int kv, int words) {
    const TtJob &J = jobs.j[blockIdx.y];
    const int64_t t = blockIdx.x;
    if (t * 128 >= J.rows) return;
    const int32_t *__restrict__ pair = J.pair;
    const int64_t pair_stride = J.pair_stride;
    const int32_t *__restrict__ argsort = J.argsort;
    const uint32_t *__restrict__ mask = J.mask;
    const int64_t rows = J.rows;
    int32_t *__restrict__ table = J.table;
    uint32_t *__restrict__ tile_mask = J.tile_mask;
    const int r = threadIdx.x & 127;
    const int q = threadIdx.x >> 7;             // warp-uniform
    const int64_t j = t * 128 + r;
    int32_t src = -1;
    if (j < rows) src = argsort ? __ldg(argsort + j) : (int32_t)j;
    int32_t *blk = table + t * (int64_t)(kv + 1) * 128;
    // 3N x 4 threads with <= 8 independent scattered 4-byte loads each: the kernel is a pure
    // L2-latency problem, so it is sized for memory-level parallelism, not for work per thread
    const int32_t *col = pair + (src >= 0 ? src : 0);
#pragma unroll 8
    for (int k = q; k < kv; k += TT_SPLIT)
        blk[k * 128 + r] = src >= 0 ? __ldg(col + (int64_t)k * pair_stride) : -1;
    if (q == 0) blk[kv * 128 + r] = src;
    __shared__ uint32_t red[4][4];
    if (q == 0) {
        for (int w = 0; w < words; ++w) {
            uint32_t m = 0;
            if (j < rows) {
                if (mask) m = __ldg(mask + j * words + w);
                else { int hi = kv - 32 * w; m = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u); }
            }
            m = __reduce_or_sync(0xffffffffu, m);
            if ((r & 31) == 0) red[w][r >> 5] = m;
        }
    }
    __syncthreads();
    if (threadIdx.x < words) tile_mask[t * words + threadIdx.x] = red[threadIdx.x][0] | red[threadIdx.x][1] |
                                                                  red[threadIdx.x][2] | red[threadIdx.x][3];
}

// same table from the row-major copy [rows][32] written by subm_probe_k3_kernel (kv <= 32):
// 4 threads per row, two 16-byte loads each, stores coalesced over the rows of the tile
__global__ void __launch_bounds__(128 * TT_SPLIT)
build_tile_table_rows_kernel(const int32_t *__restrict__ row_table, int kv, const int32_t *__restrict__ argsort,
                             const uint32_t *__restrict__ mask, int64_t rows, int32_t *__restrict__ table,
                             uint32_t *__restrict__ tile_mask) {
    const int64_t t = blockIdx.x;
    const int r = threadIdx.x & 127;
    const int q = threadIdx.x >> 7;
    const int64_t j = t * 128 + r;
    int32_t src = -1;
    if (j < rows) src = argsort ? __ldg(argsort + j) : (int32_t)j;
    int32_t *blk = table + t * (int64_t)(kv + 1) * 128;
    int4 v0 = make_int4(-1, -1, -1, -1), v1 = v0;
    if (src >= 0) {
        const int4 *rp = reinterpret_cast<const int4 *>(row_table + (int64_t)src * 32 + q * 8);
        v0 = __ldg(rp); v1 = __ldg(rp + 1);
    }
    const int32_t vals[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = q * 8 + i;
        if (k < kv) blk[k * 128 + r] = vals[i];
    }
    if (q == 0) blk[kv * 128 + r] = src;
    __shared__ uint32_t red[4];
    if (q == 0) {
        uint32_t m = 0;
        if (j < rows) m = mask ? __ldg(mask + j) : (kv >= 32 ? 0xffffffffu : ((1u << kv) - 1u));
        m = __reduce_or_sync(0xffffffffu, m);
        if ((r & 31) == 0) red[r >> 5] = m;
    }
    __syncthreads();
    if (threadIdx.x == 0) tile_mask[t] = red[0] | red[1] | red[2] | red[3];
}

// Schedule records for the dynamically scheduled kernels: tiles in order of decreasing stage count
// (= popcount of the tile mask), ties in ascending tile order (stable => deterministic).  Handing
// tiles out heaviest-first through an atomic ticket is LPT list scheduling: on the 100 k-voxel
// cloud the static round-robin assignment leaves the slowest CTA 1.6x over the mean, LPT 1.07x.
// One block; chunks of 1024 tiles are ranked with warp match_any + per-warp bucket counts.
constexpr int TO_THREADS = 1024;
constexpr int TO_BUCKETS = 130;                 // stage counts 0..128 (+1 spare)
struct ToJob { const uint32_t *tile_mask; int tiles; int32_t *rec; int32_t *state; };
struct ToJobs { ToJob j[2]; };
__global__ void __launch_bounds__(TO_THREADS)
tile_order_kernel(const ToJobs jobs, int words) {
    const ToJob &J = jobs.j[blockIdx.x];
    const uint32_t *__restrict__ tile_mask = J.tile_mask;
    const int tiles = J.tiles;
    int32_t *__restrict__ rec = J.rec;
    int32_t *__restrict__ state = J.state;
    __shared__ int bucket_base[TO_BUCKETS];
    __shared__ int wcnt[TO_THREADS / 32][TO_BUCKETS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < TT_STATE_INTS) state[tid] = 0;
    for (int i = tid; i < TO_BUCKETS; i += TO_THREADS) bucket_base[i] = 0;
    __syncthreads();
    auto load_mask = [&](int t, uint32_t (&m)[4]) {
        uint32_t any = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) { m[w] = w < words ? __ldg(tile_mask + (int64_t)t * words + w) : 0u; any |= m[w]; }
        if (!any) m[0] = 1u;                     // an empty tile still runs one (all-zero) stage
        return __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]);
    };
    // bucket sizes
    for (int t = tid; t < tiles; t += TO_THREADS) {
        uint32_t m[4];
        atomicAdd(&bucket_base[load_mask(t, m)], 1);
    }
    __syncthreads();
    if (warp == 0) {                              // exclusive scan, heaviest bucket first (one warp, shuffles)
        int carry = 0;
        for (int base = TO_BUCKETS - 1; base >= 0; base -= 32) {
            const int c = base - lane;
            const int n = c >= 0 ? bucket_base[c] : 0;
            int incl = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            if (c >= 0) bucket_base[c] = carry + incl - n;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
    for (int t0 = 0; t0 < tiles; t0 += TO_THREADS) {
        for (int i = tid; i < (TO_THREADS / 32) * TO_BUCKETS; i += TO_THREADS) (&wcnt[0][0])[i] = 0;
        __syncthreads();
        const int t = t0 + tid;
        const bool ok = t < tiles;
        uint32_t m[4] = {0, 0, 0, 0};
        const int c = ok ? load_mask(t, m) : TO_BUCKETS - 1;
        const unsigned peers = __match_any_sync(0xffffffffu, ok ? c : -1);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (ok && rank == 0) wcnt[warp][c] = __popc(peers);
        __syncthreads();
        const int live_warps = min(TO_THREADS / 32, (tiles - t0 + 31) / 32);     // warps that hold tiles
        for (int b = tid; b < TO_BUCKETS; b += TO_THREADS) {
            int run = bucket_base[b];
            for (int w = 0; w < live_warps; ++w) { const int n = wcnt[w][b]; wcnt[w][b] = run; run += n; }
            bucket_base[b] = run;
        }
        __syncthreads();
        if (ok) {
            int32_t *r = rec + (int64_t)(wcnt[warp][c] + rank) * TT_REC_INTS;
            *reinterpret_cast<int4 *>(r) = make_int4(t, (int)m[0], (int)m[1], (int)m[2]);
            *reinterpret_cast<int4 *>(r + 4) = make_int4((int)m[3], 0, 0, 0);
        }
        __syncthreads();
    }
}

}  // namespace spx

using namespace spx;

// ====================================================================== C ABI
static int validate_geom(const spx_conv_geometry *g) {
    SPX_REQUIRE(g != nullptr, "geometry is NULL");
    SPX_REQUIRE(g->ndim >= 1 && g->ndim <= SPX_MAX_NDIM, "ndim must be in [1, %d], got %d", SPX_MAX_NDIM, g->ndim);
    SPX_REQUIRE(g->batch_size > 0, "batch_size must be positive");
    for (int a = 0; a < g->ndim; ++a) {
        SPX_REQUIRE(g->ksize[a] > 0 && g->dilation[a] > 0 && g->in_dims[a] > 0, "bad ksize/dilation/dims on axis %d", a);
    }
    return 0;
}

struct RbLayout {   // workspace layout shared by the rulebook entry points
    size_t table_bytes, table_vals_bytes;
    uint32_t capacity;
    bool i64;
};

// `items` is exact for SubM (factor 4); for a regular conv it is the UPPER BOUND on outputs
// (typically ~7x the real count), so factor 2 already means a load factor well below 0.1
static RbLayout rb_layout(const Geom &g, int64_t items, const int *dims, int factor = 4) {
    RbLayout L;
    L.i64 = needs_i64(g, dims);
    L.capacity = table_capacity(items, factor);
    L.table_bytes = (size_t)L.capacity * 8;
    L.table_vals_bytes = L.i64 ? (size_t)L.capacity * 4 : 0;
    return L;
}

extern "C" int64_t spx_conv_max_out(const spx_conv_geometry *g, int64_t num_in) {
    // all.py:1559-1580, with the transposed case bounded by kv*N (ops.py:569-570)
    int64_t res = num_in, kv = 1;
    for (int i = 0; i < g->ndim; ++i) {
        kv *= g->ksize[i];
        if (g->ksize[i] > g->stride[i]) res *= (g->ksize[i] + g->stride[i] - 1) / g->stride[i];
    }
    if (g->transposed) res = kv * num_in;
    if (res > kv * num_in) res = kv * num_in;
    return res;
}

static size_t sort_pairs_temp_bytes(int64_t n) {
    // The size query goes through the CUDA runtime: a stale error left by an earlier failed call (e.g. a
    // refused stream capture) would make it return early with bytes = 0, and the workspace computed here
    // would then be smaller than what the same query yields a moment later.  Clear the state first and
    // never return less than a bound that covers CUB's double buffers + histograms.
    cudaGetLastError();
    size_t bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                                    (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
    const size_t floor_bytes = (size_t)(n > 0 ? n : 1) * 16 + (1u << 20);
    if (e != cudaSuccess) { cudaGetLastError(); return floor_bytes; }
    return bytes > floor_bytes ? bytes : floor_bytes;
}

extern "C" size_t spx_rulebook_workspace_size(const spx_conv_geometry *g, int64_t num_in, int64_t max_out, int is_subm) {
    if (!g || g->ndim < 1 || g->ndim > SPX_MAX_NDIM) return 0;
    Geom gg = make_geom(g, is_subm != 0);
    size_t total = 256;
    if (is_subm) {
        RbLayout L = rb_layout(gg, num_in, gg.in_dims);
        total += align_up(L.table_bytes, 256) + align_up(L.table_vals_bytes, 256);
    } else {
        max_out = spx_conv_max_out(g, num_in);   // the bound is recomputed by both stages
        RbLayout L = rb_layout(gg, max_out, gg.out_dims, 2);
        total += align_up(L.table_bytes, 256) + align_up(L.table_vals_bytes, 256);
        total += 5 * align_up((size_t)max_out * 4, 256);          // payload, slot (in + out), order
        total += align_up(sort_pairs_temp_bytes(max_out), 256);
        total += align_up(radix_argsort_workspace_bytes(max_out), 256);
        total += align_up(rank_scratch_bytes((int64_t)gg.kv * num_in), 256);
        total += 256;                                             // counter
    }
    return total + 1024;
}

static bool subm_k3_path(const Geom &gg) {
    return !needs_i64(gg, gg.in_dims) && gg.ndim == 3 && gg.ksize[0] == 3 && gg.ksize[1] == 3 && gg.ksize[2] == 3;
}

extern "C" int spx_subm_row_table_supported(const spx_conv_geometry *g) {
    if (!g || g->ndim < 1 || g->ndim > SPX_MAX_NDIM) return 0;
    return subm_k3_path(make_geom(g, true)) ? 1 : 0;
}

extern "C" int spx_subm_rulebook(const spx_conv_geometry *g, const int32_t *indices, int64_t N, int32_t *pair_fwd,
                                 int32_t *pair_bwd, uint32_t *mask, int32_t *row_table, void *workspace,
                                 size_t workspace_bytes, spx_stream_t stream_) {
    if (validate_geom(g)) return 2;
    for (int a = 0; a < g->ndim; ++a)
        SPX_REQUIRE(g->ksize[a] % 2 == 1, "subm only support odd ksize");
    SPX_REQUIRE(N >= 0 && N < 2147483647ll, "bad N");
    if (N == 0) return 0;
    SPX_REQUIRE(indices && pair_fwd && workspace, "NULL pointer argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    Geom gg = make_geom(g, true);
    SPX_REQUIRE((int64_t)gg.kv * N < 2147483647ll * 4, "kv*N too large");
    RbLayout L = rb_layout(gg, N, gg.in_dims);
    WorkspaceCarver ws(workspace, workspace_bytes);
    void *tbl = ws.take<char>(L.table_bytes);
    int32_t *tvals = L.i64 ? ws.take<int32_t>(L.capacity) : nullptr;
    SPX_REQUIRE(ws.ok(), "rulebook workspace too small: need %zu, have %zu", ws.off, workspace_bytes);
    int words = (gg.kv + 31) / 32;
    const int T = 128;
    unsigned nblk = (unsigned)div_up64(N, T);
    SPX_CHECK_CUDA(cudaMemsetAsync(tbl, 0xFF, L.table_bytes, stream));
    if (!L.i64) {
        Table32 t{(unsigned long long *)tbl, L.capacity - 1};
        subm_insert_kernel<<<nblk, T, 0, stream>>>(t, gg, indices, N);
        SPX_CHECK_LAUNCH("subm_insert_kernel");
        if (subm_k3_path(gg)) {
            subm_probe_k3_kernel<<<(unsigned)div_up64(N, K3_VOX), 27 * 32, 0, stream>>>(t, gg, indices, N, pair_fwd,
                                                                                        pair_bwd, mask, row_table);
        } else {
            SPX_REQUIRE(row_table == nullptr, "row_table is only produced when spx_subm_row_table_supported()");
            subm_probe_kernel<<<nblk, T, 0, stream>>>(t, gg, indices, N, pair_fwd, pair_bwd, mask, words);
        }
        SPX_CHECK_LAUNCH("subm_probe_kernel");
    } else {
        SPX_CHECK_CUDA(cudaMemsetAsync(tvals, 0x7F, (size_t)L.capacity * 4, stream));
        Table64 t{(long long *)tbl, tvals, L.capacity - 1};
        subm_insert_kernel<<<nblk, T, 0, stream>>>(t, gg, indices, N);
        SPX_CHECK_LAUNCH("subm_insert_kernel");
        SPX_REQUIRE(row_table == nullptr, "row_table is only produced when spx_subm_row_table_supported()");
        subm_probe_kernel<<<nblk, T, 0, stream>>>(t, gg, indices, N, pair_fwd, pair_bwd, mask, words);
        SPX_CHECK_LAUNCH("subm_probe_kernel");
    }
    return 0;
}


namespace {
struct ConvWs {
    void *tbl; int32_t *tvals;
    uint32_t *payload, *slot, *payload_sorted, *slot_sorted;
    void *sort_tmp; size_t sort_tmp_bytes;
    int32_t *order;                    // argsort of the payloads (default path)
    void *radix_ws; size_t radix_ws_bytes;
    uint32_t *rank_bitmap; int *rank_tiles;     // first-touch ranking (default path): bitmap over kv*N + tile counts
    size_t rank_bytes; int64_t rank_ntiles;
    int *counter;
    RbLayout L;
};

// which table capacity stage 1 ended up with (the optimistic size, or the full one after an overflow);
// stage 2 is called right after stage 1 on the same thread with the same workspace
struct ConvStage1Record { const void *ws; uint32_t capacity; bool legacy; };
thread_local ConvStage1Record g_stage1 = {nullptr, 0, false};

// Optimistic table size: the bound on the outputs (spx_conv_max_out, e.g. 8 N for 3^3 stride 2) is what
// isolated points would produce; LiDAR-like clouds yield ~0.5 N.  A table for a quarter of the bound
// (load <= 0.5 if M <= bound / 4) is cleared and probed 4x cheaper; a probe chain longer than
// CONV_MAX_PROBES flags an overflow and stage 1 re-runs with the full size.
uint32_t optimistic_capacity(int64_t max_out, uint32_t full_capacity) {
    uint32_t cap = table_capacity((max_out + 3) / 4, 2);
    return cap < full_capacity ? cap : full_capacity;
}
int carve_conv_ws(const spx_conv_geometry *g, const Geom &gg, int64_t N, void *workspace, size_t bytes, ConvWs &w) {
    int64_t max_out = spx_conv_max_out(g, N);
    w.L = rb_layout(gg, max_out, gg.out_dims, 2);
    WorkspaceCarver ws(workspace, bytes);
    w.tbl = ws.take<char>(w.L.table_bytes);
    w.tvals = w.L.i64 ? ws.take<int32_t>(w.L.capacity) : nullptr;
    w.payload = ws.take<uint32_t>(max_out);
    w.slot = ws.take<uint32_t>(max_out);
    w.payload_sorted = ws.take<uint32_t>(max_out);
    w.slot_sorted = ws.take<uint32_t>(max_out);
    w.sort_tmp_bytes = sort_pairs_temp_bytes(max_out);
    w.sort_tmp = ws.take<char>(w.sort_tmp_bytes);
    w.order = ws.take<int32_t>(max_out);
    w.radix_ws_bytes = radix_argsort_workspace_bytes(max_out);
    w.radix_ws = ws.take<char>(w.radix_ws_bytes);
    w.rank_bytes = rank_scratch_bytes((int64_t)gg.kv * N, &w.rank_ntiles);
    w.rank_bitmap = (uint32_t *)ws.take<char>(w.rank_bytes);
    w.rank_tiles = (int *)(w.rank_bitmap + w.rank_ntiles * RANK_TILE_WORDS);
    w.counter = ws.take<int>(64);
    SPX_REQUIRE(ws.ok(), "rulebook workspace too small: need %zu, have %zu", ws.off, bytes);
    return 0;
}
}  // namespace

extern "C" int spx_conv_rulebook_stage1(const spx_conv_geometry *g, const int32_t *indices, int64_t N,
                                        int64_t *num_out_host, void *workspace, size_t workspace_bytes,
                                        spx_stream_t stream_) {
    if (validate_geom(g)) return 2;
    SPX_REQUIRE(num_out_host != nullptr, "num_out_host is NULL");
    *num_out_host = 0;
    if (N == 0) return 0;
    SPX_REQUIRE(indices && workspace, "NULL pointer argument");
    for (int a = 0; a < g->ndim; ++a)
        SPX_REQUIRE(g->stride[a] > 0 && g->out_dims[a] > 0, "bad stride/out_dims on axis %d", a);
    cudaStream_t stream = (cudaStream_t)stream_;
    Geom gg = make_geom(g, false);
    SPX_REQUIRE((int64_t)gg.kv * N < 2000000000ll, "kv*N must stay below 2e9 (kv=%d, N=%lld)", gg.kv, (long long)N);
    ConvWs w;
    if (carve_conv_ws(g, gg, N, workspace, workspace_bytes, w)) return 2;
    const int T = 128;
    dim3 grid((unsigned)div_up64(N, T), gg.kv);
    const bool fast3 = gg.ndim == 3 && !gg.transposed;
    const bool k3 = fast3 && gg.ksize[0] == 3 && gg.ksize[1] == 3 && gg.ksize[2] == 3;
    const bool legacy = (runtime_cfg().debug & 128) != 0;
    g_stage1 = {workspace, w.L.capacity, legacy};
    if (!legacy) {
        // ---- default path: optimistic table, append-on-create, own radix sort of the first-touch payloads
        const int64_t max_out = spx_conv_max_out(g, N);
        int host_state[2] = {0, 0};
        uint32_t capacity = optimistic_capacity(max_out, w.L.capacity);
        for (int attempt = 0; attempt < 2; ++attempt) {
            // ranking scratch and counters are contiguous (carve_conv_ws): one region of zeros
            const int64_t zero_vec = (int64_t)(((char *)(w.counter + 64) - (char *)w.rank_bitmap) / 16);
            conv_clear_kernel<<<sm_count() * 4, 256, 0, stream>>>((uint4 *)w.tbl, (int64_t)capacity / 2,
                                                                  w.L.i64 ? (uint4 *)w.tvals : nullptr,
                                                                  w.L.i64 ? (int64_t)capacity / 4 : 0,
                                                                  (uint4 *)w.rank_bitmap, zero_vec);
            SPX_CHECK_LAUNCH("conv_clear_kernel");
            if (!w.L.i64) {
                Table32 t{(unsigned long long *)w.tbl, capacity - 1};
                if (k3) conv_insert_k3_append_kernel<<<(unsigned)div_up64(N, APPEND_THREADS), APPEND_THREADS, 0, stream>>>(t, gg, indices, N, w.slot, w.counter);
                else if (fast3) conv_insert_append_kernel<Table32, true><<<grid, APPEND_THREADS, 0, stream>>>(t, gg, indices, N, w.slot, w.counter);
                else conv_insert_append_kernel<Table32, false><<<grid, APPEND_THREADS, 0, stream>>>(t, gg, indices, N, w.slot, w.counter);
            } else {
                Table64 t{(long long *)w.tbl, w.tvals, capacity - 1};
                if (k3) conv_insert_k3_append_kernel<<<(unsigned)div_up64(N, APPEND_THREADS), APPEND_THREADS, 0, stream>>>(t, gg, indices, N, w.slot, w.counter);
                else if (fast3) conv_insert_append_kernel<Table64, true><<<grid, APPEND_THREADS, 0, stream>>>(t, gg, indices, N, w.slot, w.counter);
                else conv_insert_append_kernel<Table64, false><<<grid, APPEND_THREADS, 0, stream>>>(t, gg, indices, N, w.slot, w.counter);
            }
            SPX_CHECK_LAUNCH("conv_insert_append_kernel");
            SPX_CHECK_CUDA(cudaMemcpyAsync(host_state, w.counter, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
            SPX_CHECK_CUDA(cudaStreamSynchronize(stream));
            if (!host_state[1]) break;
            SPX_REQUIRE(capacity < w.L.capacity, "conv rulebook: hash table overflow at full capacity (%u slots)", capacity);
            capacity = w.L.capacity;                     // rare: far more outputs than the optimistic guess
        }
        g_stage1.capacity = capacity;
        const int m_host = host_state[0];
        *num_out_host = m_host;
        if (m_host == 0) return 0;
        const unsigned mblk = (unsigned)div_up64(m_host, MARK_THREADS);
        if (!w.L.i64) {
            Table32 t{(unsigned long long *)w.tbl, capacity - 1};
            conv_mark_kernel<<<mblk, MARK_THREADS, 0, stream>>>(t, w.slot, m_host, w.rank_bitmap, w.rank_tiles, w.rank_ntiles, w.counter + 2);
        } else {
            Table64 t{(long long *)w.tbl, w.tvals, capacity - 1};
            conv_mark_kernel<<<mblk, MARK_THREADS, 0, stream>>>(t, w.slot, m_host, w.rank_bitmap, w.rank_tiles, w.rank_ntiles, w.counter + 2);
        }
        SPX_CHECK_LAUNCH("conv_mark_kernel");
        return 0;
    }
    SPX_CHECK_CUDA(cudaMemsetAsync(w.tbl, 0xFF, w.L.table_bytes, stream));
    SPX_CHECK_CUDA(cudaMemsetAsync(w.counter, 0, sizeof(int), stream));
    unsigned cblk = (unsigned)div_up64(w.L.capacity, COLLECT_THREADS * COLLECT_ITEMS);
    if (!w.L.i64) {
        Table32 t{(unsigned long long *)w.tbl, w.L.capacity - 1};
        if (k3) conv_insert_k3_kernel<<<(unsigned)div_up64(N, T), T, 0, stream>>>(t, gg, indices, N);
        else if (fast3) conv_insert_kernel<Table32, true><<<grid, T, 0, stream>>>(t, gg, indices, N);
        else conv_insert_kernel<Table32, false><<<grid, T, 0, stream>>>(t, gg, indices, N);
        SPX_CHECK_LAUNCH("conv_insert_kernel");
        conv_collect_kernel<<<cblk, COLLECT_THREADS, 0, stream>>>(t, w.L.capacity, w.payload, w.slot, w.counter);
        SPX_CHECK_LAUNCH("conv_collect_kernel");
    } else {
        SPX_CHECK_CUDA(cudaMemsetAsync(w.tvals, 0x7F, (size_t)w.L.capacity * 4, stream));
        Table64 t{(long long *)w.tbl, w.tvals, w.L.capacity - 1};
        if (k3) conv_insert_k3_kernel<<<(unsigned)div_up64(N, T), T, 0, stream>>>(t, gg, indices, N);
        else if (fast3) conv_insert_kernel<Table64, true><<<grid, T, 0, stream>>>(t, gg, indices, N);
        else conv_insert_kernel<Table64, false><<<grid, T, 0, stream>>>(t, gg, indices, N);
        SPX_CHECK_LAUNCH("conv_insert_kernel");
        conv_collect_kernel<<<cblk, COLLECT_THREADS, 0, stream>>>(t, w.L.capacity, w.payload, w.slot, w.counter);
        SPX_CHECK_LAUNCH("conv_collect_kernel");
    }
    int m_host = 0;
    SPX_CHECK_CUDA(cudaMemcpyAsync(&m_host, w.counter, sizeof(int), cudaMemcpyDeviceToHost, stream));
    SPX_CHECK_CUDA(cudaStreamSynchronize(stream));
    *num_out_host = m_host;
    if (m_host == 0) return 0;
    // rank outputs by first touch: sort (payload, slot) by payload; payload < kv*N
    int end_bit = 1;
    while (end_bit < 32 && ((int64_t)1 << end_bit) < (int64_t)gg.kv * N) ++end_bit;
    size_t tmp_bytes = w.sort_tmp_bytes;
    SPX_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(w.sort_tmp, tmp_bytes, w.payload, w.payload_sorted, w.slot,
                                                   w.slot_sorted, m_host, 0, end_bit, stream));
    count_launch(3);
    return 0;
}

extern "C" int spx_conv_rulebook_stage2(const spx_conv_geometry *g, const int32_t *indices, int64_t N, int64_t M,
                                        int32_t *out_inds, int32_t *pair_fwd, int32_t *pair_bwd, uint32_t *mask_fwd,
                                        uint32_t *mask_bwd, void *workspace, size_t workspace_bytes,
                                        spx_stream_t stream_) {
    if (validate_geom(g)) return 2;
    if (N == 0 || M == 0) return 0;
    SPX_REQUIRE(indices && out_inds && pair_fwd && pair_bwd && workspace, "NULL pointer argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    Geom gg = make_geom(g, false);
    ConvWs w;
    if (carve_conv_ws(g, gg, N, workspace, workspace_bytes, w)) return 2;
    int words = (gg.kv + 31) / 32;
    const int T = 128;
    dim3 grid((unsigned)div_up64(N, T), gg.kv);
    const bool fast3 = gg.ndim == 3 && !gg.transposed;
    const bool k3 = fast3 && gg.ksize[0] == 3 && gg.ksize[1] == 3 && gg.ksize[2] == 3;
    SPX_REQUIRE(g_stage1.ws == workspace && g_stage1.capacity != 0,
                "conv_rulebook_stage2 must follow conv_rulebook_stage1 on the same thread with the same workspace");
    const bool legacy = g_stage1.legacy;
    if (legacy) SPX_CHECK_CUDA(cudaMemsetAsync(pair_fwd, 0xFF, (size_t)gg.kv * M * 4, stream));   // default: the assign kernel fills it
    const uint32_t capacity = g_stage1.capacity;
    // 3x3x3, one mask word: the pairs kernel ORs the forward masks too (zeroed by the assign kernel)
    uint32_t *mask_fwd_or = (k3 && !legacy && mask_fwd && words == 1) ? mask_fwd : nullptr;
    if (!w.L.i64) {
        Table32 t{(unsigned long long *)w.tbl, capacity - 1};
        if (legacy) conv_assign_kernel<<<(unsigned)div_up64(M, 256), 256, 0, stream>>>(t, gg, w.slot_sorted, nullptr, M, out_inds);
        else conv_assign_rank_kernel<<<(unsigned)div_up64(M, 256), 256, 0, stream>>>(t, gg, w.slot, M, w.rank_bitmap, w.rank_tiles, out_inds, mask_fwd_or, pair_fwd, gg.kv);
        SPX_CHECK_LAUNCH("conv_assign_kernel");
        if (k3) conv_pairs_k3_kernel<<<(unsigned)div_up64(N, T), T, 0, stream>>>(t, gg, indices, N, M, pair_fwd, pair_bwd, mask_bwd, mask_fwd_or);
        else if (fast3) conv_pairs_kernel<Table32, true><<<grid, T, 0, stream>>>(t, gg, indices, N, M, pair_fwd, pair_bwd);
        else conv_pairs_kernel<Table32, false><<<grid, T, 0, stream>>>(t, gg, indices, N, M, pair_fwd, pair_bwd);
        SPX_CHECK_LAUNCH("conv_pairs_kernel");
    } else {
        Table64 t{(long long *)w.tbl, w.tvals, capacity - 1};
        if (legacy) conv_assign_kernel<<<(unsigned)div_up64(M, 256), 256, 0, stream>>>(t, gg, w.slot_sorted, nullptr, M, out_inds);
        else conv_assign_rank_kernel<<<(unsigned)div_up64(M, 256), 256, 0, stream>>>(t, gg, w.slot, M, w.rank_bitmap, w.rank_tiles, out_inds, mask_fwd_or, pair_fwd, gg.kv);
        SPX_CHECK_LAUNCH("conv_assign_kernel");
        if (k3) conv_pairs_k3_kernel<<<(unsigned)div_up64(N, T), T, 0, stream>>>(t, gg, indices, N, M, pair_fwd, pair_bwd, mask_bwd, mask_fwd_or);
        else if (fast3) conv_pairs_kernel<Table64, true><<<grid, T, 0, stream>>>(t, gg, indices, N, M, pair_fwd, pair_bwd);
        else conv_pairs_kernel<Table64, false><<<grid, T, 0, stream>>>(t, gg, indices, N, M, pair_fwd, pair_bwd);
        SPX_CHECK_LAUNCH("conv_pairs_kernel");
    }
    if (mask_fwd && !mask_fwd_or) {
        table_mask_kernel<<<(unsigned)div_up64(M, 256), 256, 0, stream>>>(pair_fwd, M, gg.kv, words, mask_fwd);
        SPX_CHECK_LAUNCH("table_mask_kernel");
    }
    if (mask_bwd && !k3) {                      // the 3x3x3 pairs kernel has already written it
        table_mask_kernel<<<(unsigned)div_up64(N, 256), 256, 0, stream>>>(pair_bwd, N, gg.kv, words, mask_bwd);
        SPX_CHECK_LAUNCH("table_mask_kernel");
    }
    return 0;
}

extern "C" size_t spx_native_pairs_workspace_size(int64_t N, int kv) {
    int64_t nblk = div_up64(N > 0 ? N : 1, SCAN_TILE);
    return (size_t)kv * nblk * sizeof(int) + 1024;
}

extern "C" int spx_native_pairs(const int32_t *pair_bwd, int64_t N, int kv, int is_subm, int32_t *pairs,
                                int32_t *indice_pair_num, void *workspace, size_t workspace_bytes,
                                spx_stream_t stream_) {
    SPX_REQUIRE(kv > 0 && N >= 0, "bad kv / N");
    SPX_REQUIRE(pairs && indice_pair_num, "NULL pointer argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    SPX_CHECK_CUDA(cudaMemsetAsync(indice_pair_num, 0, sizeof(int32_t) * kv, stream));
    if (N == 0) return 0;
    SPX_REQUIRE(pair_bwd && workspace, "NULL pointer argument");
    SPX_CHECK_CUDA(cudaMemsetAsync(pairs, 0xFF, (size_t)2 * kv * N * 4, stream));
    int rows = is_subm ? kv / 2 : kv;
    int nblk = (int)div_up64(N, SCAN_TILE);
    WorkspaceCarver ws(workspace, workspace_bytes);
    int *counts = ws.take<int>((size_t)kv * nblk);
    SPX_REQUIRE(ws.ok(), "native-pairs workspace too small: need %zu, have %zu", ws.off, workspace_bytes);
    if (rows > 0) {
        dim3 grid(nblk, rows);
        native_count_kernel<<<grid, SCAN_THREADS, 0, stream>>>(pair_bwd, N, counts, nblk);
        SPX_CHECK_LAUNCH("native_count_kernel");
        native_scan_kernel<<<rows, SCAN_THREADS, 0, stream>>>(counts, nblk, indice_pair_num);
        SPX_CHECK_LAUNCH("native_scan_kernel");
        native_write_kernel<<<grid, SCAN_THREADS, 0, stream>>>(pair_bwd, N, kv, is_subm, counts, nblk, pairs);
        SPX_CHECK_LAUNCH("native_write_kernel");
    }
    if (is_subm) {
        subm_centre_kernel<<<(unsigned)div_up64(N, 256), 256, 0, stream>>>(pairs, N, kv);
        SPX_CHECK_LAUNCH("subm_centre_kernel");
    }
    return 0;
}

extern "C" int spx_pairs_to_table(const int32_t *pairs, const int32_t *indice_pair_num, int kv, int64_t pair_stride,
                                  int64_t n_in, int64_t n_out, int is_subm, int inverse, int32_t *table_fwd,
                                  int32_t *table_bwd, uint32_t *mask_fwd, uint32_t *mask_bwd, spx_stream_t stream_) {
    SPX_REQUIRE(pairs && indice_pair_num && kv > 0, "NULL pointer argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    int words = (kv + 31) / 32;
    if (table_fwd && n_out > 0) SPX_CHECK_CUDA(cudaMemsetAsync(table_fwd, 0xFF, (size_t)kv * n_out * 4, stream));
    if (table_bwd && n_in > 0) SPX_CHECK_CUDA(cudaMemsetAsync(table_bwd, 0xFF, (size_t)kv * n_in * 4, stream));
    if (mask_fwd && n_out > 0) SPX_CHECK_CUDA(cudaMemsetAsync(mask_fwd, 0, (size_t)n_out * words * 4, stream));
    if (mask_bwd && n_in > 0) SPX_CHECK_CUDA(cudaMemsetAsync(mask_bwd, 0, (size_t)n_in * words * 4, stream));
    int64_t span = pair_stride;
    if (span <= 0) return 0;
    dim3 grid((unsigned)div_up64(span, 256), kv);
    pairs_to_table_kernel<<<grid, 256, 0, stream>>>(pairs, indice_pair_num, kv, pair_stride, n_in, n_out, is_subm,
                                                    inverse, table_fwd, table_bwd, mask_fwd, mask_bwd, words);
    SPX_CHECK_LAUNCH("pairs_to_table_kernel");
    return 0;
}


extern "C" size_t spx_mask_argsort_workspace_size(int64_t N, int words) {
    if (N <= 0) return 256;
    if (words == 1) return radix_argsort_workspace_bytes(N);
    size_t n = (size_t)N;
    return 4 * align_up(n * 4, 256) + align_up(n * 4 * (size_t)words, 256) + align_up(sort_pairs_temp_bytes(N), 256) + 1024;
}

extern "C" int spx_mask_argsort(uint32_t *mask, int32_t *argsort, int64_t N, int words, int kv, int do_sort,
                                void *workspace, size_t workspace_bytes, spx_stream_t stream_) {
    SPX_REQUIRE(words >= 1 && words <= 4, "mask words must be in [1,4], got %d", words);
    if (N == 0) return 0;
    SPX_REQUIRE(mask && argsort, "NULL pointer argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    unsigned nblk = (unsigned)div_up64(N, 256);
    if (!do_sort) {
        iota_kernel<<<nblk, 256, 0, stream>>>(argsort, N);
        SPX_CHECK_LAUNCH("iota_kernel");
        return 0;
    }
    SPX_REQUIRE(workspace != nullptr, "workspace is NULL");
    if (words == 1)   // hand-written 9-bit-digit stable radix argsort (sort.cu)
        return radix_argsort(mask, argsort, N, kv > 0 && kv < 32 ? kv : 32, workspace, workspace_bytes, stream);
    WorkspaceCarver ws(workspace, workspace_bytes);
    uint32_t *keys_in = ws.take<uint32_t>(N);
    uint32_t *keys_out = ws.take<uint32_t>(N);
    int32_t *perm_a = ws.take<int32_t>(N);
    int32_t *perm_b = ws.take<int32_t>(N);
    uint32_t *rows_tmp = ws.take<uint32_t>((size_t)N * words);
    size_t tmp_bytes = sort_pairs_temp_bytes(N);
    void *tmp = ws.take<char>(tmp_bytes);
    SPX_REQUIRE(ws.ok(), "argsort workspace too small: need %zu, have %zu", ws.off, workspace_bytes);
    iota_kernel<<<nblk, 256, 0, stream>>>(perm_a, N);
    SPX_CHECK_LAUNCH("iota_kernel");
    if (words == 1) {
        int end_bit = kv > 0 && kv < 32 ? kv : 32;
        SPX_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const uint32_t *)mask, keys_out,
                                                       (const int32_t *)perm_a, argsort, (int)N, 0, end_bit, stream));
        count_launch(3);
        SPX_CHECK_CUDA(cudaMemcpyAsync(mask, keys_out, (size_t)N * 4, cudaMemcpyDeviceToDevice, stream));
        return 0;
    }
    // LSD over words: least significant word (last) first; stable sorts compose
    int32_t *cur = perm_a, *nxt = perm_b;
    for (int w = words - 1; w >= 0; --w) {
        gather_word_kernel<<<nblk, 256, 0, stream>>>(mask, cur, N, words, w, keys_in);
        SPX_CHECK_LAUNCH("gather_word_kernel");
        SPX_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const uint32_t *)keys_in, keys_out,
                                                       (const int32_t *)cur, nxt, (int)N, 0, 32, stream));
        count_launch(3);
        int32_t *t = cur; cur = nxt; nxt = t;
    }
    SPX_CHECK_CUDA(cudaMemcpyAsync(argsort, cur, (size_t)N * 4, cudaMemcpyDeviceToDevice, stream));
    gather_rows_kernel<<<nblk, 256, 0, stream>>>(mask, argsort, N, words, rows_tmp);
    SPX_CHECK_LAUNCH("gather_rows_kernel");
    SPX_CHECK_CUDA(cudaMemcpyAsync(mask, rows_tmp, (size_t)N * words * 4, cudaMemcpyDeviceToDevice, stream));
    return 0;
}

extern "C" size_t spx_tile_table_elems(int64_t rows, int kv) {
    return (size_t)tt_total_elems(div_up64(rows > 0 ? rows : 1, 128), kv);
}

extern "C" int spx_build_tile_table(const int32_t *pair, int64_t pair_stride, int kv, const int32_t *argsort,
                                    const uint32_t *mask, int64_t rows, const int32_t *row_table, int32_t *table,
                                    uint32_t *tile_mask, spx_stream_t stream_) {
    SPX_REQUIRE(kv >= 1 && kv <= 128, "build_tile_table: kernel volume %d not in [1,128]", kv);
    if (rows == 0) return 0;
    SPX_REQUIRE((pair || row_table) && table && tile_mask, "build_tile_table: NULL pointer argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    int words = (kv + 31) / 32;
    if (row_table) {
        SPX_REQUIRE(kv <= 32, "build_tile_table: row_table holds at most 32 offsets per row, kv = %d", kv);
        SPX_REQUIRE(((uintptr_t)row_table & 15u) == 0, "build_tile_table: row_table must be 16-byte aligned");
        build_tile_table_rows_kernel<<<(unsigned)div_up64(rows, 128), 128 * TT_SPLIT, 0, stream>>>(
            row_table, kv, argsort, mask, rows, table, tile_mask);
        SPX_CHECK_LAUNCH("build_tile_table_rows_kernel");
    } else {
        TtJobs jobs;
        memset(&jobs, 0, sizeof(jobs));
        jobs.j[0] = TtJob{pair, pair_stride, argsort, mask, rows, table, tile_mask};
        build_tile_table_kernel<<<dim3((unsigned)div_up64(rows, 128), 1), 128 * TT_SPLIT, 0, stream>>>(jobs, kv, words);
        SPX_CHECK_LAUNCH("build_tile_table_kernel");
    }
    const int64_t tiles = div_up64(rows, 128);
    SPX_REQUIRE(tiles < 2147483647ll, "build_tile_table: too many tiles");
    SPX_REQUIRE(((uintptr_t)table & 15u) == 0, "build_tile_table: table must be 16-byte aligned");
    int32_t *rec = table + tt_blocks_elems(tiles, kv);
    ToJobs oj;
    memset(&oj, 0, sizeof(oj));
    oj.j[0] = ToJob{tile_mask, (int)tiles, rec, rec + tiles * TT_REC_INTS};
    tile_order_kernel<<<1, TO_THREADS, 0, stream>>>(oj, words);
    SPX_CHECK_LAUNCH("tile_order_kernel");
    return 0;
}

// forward + backward tile tables of a regular conv in one launch each (gather kernel, schedule records)
static int build_tile_tables_pair(int kv, const int32_t *pair0, const int32_t *argsort0, const uint32_t *mask0, int64_t rows0,
                                  int32_t *table0, uint32_t *tmask0, const int32_t *pair1, const int32_t *argsort1,
                                  const uint32_t *mask1, int64_t rows1, int32_t *table1, uint32_t *tmask1,
                                  cudaStream_t stream) {
    SPX_REQUIRE(kv >= 1 && kv <= 128, "build_tile_table: kernel volume %d not in [1,128]", kv);
    const int words = (kv + 31) / 32;
    const int64_t tiles0 = div_up64(rows0, 128), tiles1 = div_up64(rows1, 128);
    SPX_REQUIRE(tiles0 < 2147483647ll && tiles1 < 2147483647ll, "build_tile_table: too many tiles");
    SPX_REQUIRE((((uintptr_t)table0 | (uintptr_t)table1) & 15u) == 0, "build_tile_table: table must be 16-byte aligned");
    TtJobs jobs;
    memset(&jobs, 0, sizeof(jobs));
    jobs.j[0] = TtJob{pair0, rows0, argsort0, mask0, rows0, table0, tmask0};
    jobs.j[1] = TtJob{pair1, rows1, argsort1, mask1, rows1, table1, tmask1};
    const int64_t tmax = tiles0 > tiles1 ? tiles0 : tiles1;
    build_tile_table_kernel<<<dim3((unsigned)tmax, 2), 128 * TT_SPLIT, 0, stream>>>(jobs, kv, words);
    SPX_CHECK_LAUNCH("build_tile_table_kernel");
    int32_t *rec0 = table0 + tt_blocks_elems(tiles0, kv), *rec1 = table1 + tt_blocks_elems(tiles1, kv);
    ToJobs oj;
    memset(&oj, 0, sizeof(oj));
    oj.j[0] = ToJob{tmask0, (int)tiles0, rec0, rec0 + tiles0 * TT_REC_INTS};
    oj.j[1] = ToJob{tmask1, (int)tiles1, rec1, rec1 + tiles1 * TT_REC_INTS};
    tile_order_kernel<<<2, TO_THREADS, 0, stream>>>(oj, words);
    SPX_CHECK_LAUNCH("tile_order_kernel");
    return 0;
}

// ====================================================================== fused host entry points
// One C-ABI call per rulebook (the eager Python path was paying ~10 us of interpreter + ctypes +
// allocator time for every separate call, workspace query and scratch tensor).  Pure orchestration:
// the launches are exactly those of the separate entry points, scratch regions are carved from ONE
// caller-provided workspace.

static size_t subm_row_table_bytes(const spx_conv_geometry *g, int64_t N) {
    return spx_subm_row_table_supported(g) ? align_up((size_t)N * 32 * sizeof(int32_t), 256) : 0;
}

extern "C" size_t spx_subm_rulebook_all_workspace_size(const spx_conv_geometry *g, int64_t N) {
    if (!g || g->ndim < 1 || g->ndim > SPX_MAX_NDIM) return 0;
    int kv = 1;
    for (int a = 0; a < g->ndim; ++a) kv *= g->ksize[a];
    const size_t a = spx_rulebook_workspace_size(g, N, 0, 1);
    const size_t b = spx_mask_argsort_workspace_size(N, (kv + 31) / 32);
    return align_up(a > b ? a : b, 256) + subm_row_table_bytes(g, N) + 256;
}

extern "C" int spx_subm_rulebook_all(const spx_conv_geometry *g, const int32_t *indices, int64_t N, int32_t *pair_fwd,
                                     int32_t *pair_bwd, uint32_t *mask, int32_t *argsort, int do_sort,
                                     int32_t *tile_table, uint32_t *tile_mask, void *workspace, size_t workspace_bytes,
                                     spx_stream_t stream) {
    if (validate_geom(g)) return 2;
    if (N == 0) return 0;
    SPX_REQUIRE(mask && argsort && workspace, "subm_rulebook_all: NULL pointer argument");
    SPX_REQUIRE(workspace_bytes >= spx_subm_rulebook_all_workspace_size(g, N), "subm_rulebook_all: workspace too small");
    int kv = 1;
    for (int a = 0; a < g->ndim; ++a) kv *= g->ksize[a];
    const int words = (kv + 31) / 32;
    const size_t rb = spx_rulebook_workspace_size(g, N, 0, 1), as = spx_mask_argsort_workspace_size(N, words);
    const size_t shared = align_up(rb > as ? rb : as, 256);          // rulebook and sort scratch are used one after the other
    int32_t *row_table = subm_row_table_bytes(g, N) ? (int32_t *)((char *)workspace + shared) : nullptr;
    if (int rc = spx_subm_rulebook(g, indices, N, pair_fwd, pair_bwd, mask, row_table, workspace, rb, stream)) return rc;
    if (int rc = spx_mask_argsort(mask, argsort, N, words, kv, do_sort, workspace, as, stream)) return rc;
    if (tile_table)
        return spx_build_tile_table(pair_fwd, N, kv, argsort, mask, N, row_table, tile_table, tile_mask, stream);
    return 0;
}

extern "C" size_t spx_conv_rulebook_all_workspace_size(const spx_conv_geometry *g, int64_t N) {
    if (!g || g->ndim < 1 || g->ndim > SPX_MAX_NDIM) return 0;
    int kv = 1;
    for (int a = 0; a < g->ndim; ++a) kv *= g->ksize[a];
    const int64_t max_rows = spx_conv_max_out(g, N) > N ? spx_conv_max_out(g, N) : N;
    return align_up(spx_rulebook_workspace_size(g, N, 0, 0), 256) +
           2 * align_up(spx_mask_argsort_workspace_size(max_rows, (kv + 31) / 32), 256) + 256;   // two sorts side by side
}

// stage 2 + both mask argsorts + both tile tables (argsort_bwd / table_bwd may be NULL: inference)
extern "C" int spx_conv_rulebook_stage2_all(const spx_conv_geometry *g, const int32_t *indices, int64_t N, int64_t M,
                                            int32_t *out_inds, int32_t *pair_fwd, int32_t *pair_bwd, uint32_t *mask_fwd,
                                            uint32_t *mask_bwd, int32_t *argsort_fwd, int32_t *argsort_bwd, int do_sort,
                                            int32_t *table_fwd, uint32_t *tmask_fwd, int32_t *table_bwd,
                                            uint32_t *tmask_bwd, void *workspace, size_t workspace_bytes,
                                            spx_stream_t stream) {
    if (validate_geom(g)) return 2;
    if (N == 0 || M == 0) return 0;
    SPX_REQUIRE(mask_fwd && mask_bwd && argsort_fwd && workspace, "conv_rulebook_stage2_all: NULL pointer argument");
    SPX_REQUIRE(workspace_bytes >= spx_conv_rulebook_all_workspace_size(g, N), "conv_rulebook_stage2_all: workspace too small");
    int kv = 1;
    for (int a = 0; a < g->ndim; ++a) kv *= g->ksize[a];
    const int words = (kv + 31) / 32;
    const size_t rb = spx_rulebook_workspace_size(g, N, 0, 0);
    void *sort_ws = (char *)workspace + align_up(rb, 256);
    const size_t sort_bytes = workspace_bytes - align_up(rb, 256);
    if (int rc = spx_conv_rulebook_stage2(g, indices, N, M, out_inds, pair_fwd, pair_bwd, mask_fwd, mask_bwd, workspace, rb, stream)) return rc;
    if (words == 1 && do_sort && argsort_bwd && !(runtime_cfg().debug & 2048)) {
        // both mask sorts, then both tile tables, two jobs per launch (debug bit 2048: one after the other)
        const size_t half = (sort_bytes / 2) & ~(size_t)255;
        const int key_bits = kv < 32 ? kv : 32;
        if (int rc = radix_argsort_pair(mask_fwd, argsort_fwd, M, mask_bwd, argsort_bwd, N, key_bits, sort_ws, half,
                                        (char *)sort_ws + half, half, (cudaStream_t)stream)) return rc;
        if (table_fwd && table_bwd)
            return build_tile_tables_pair(kv, pair_fwd, argsort_fwd, mask_fwd, M, table_fwd, tmask_fwd, pair_bwd, argsort_bwd,
                                          mask_bwd, N, table_bwd, tmask_bwd, (cudaStream_t)stream);
        if (table_fwd)
            if (int rc = spx_build_tile_table(pair_fwd, M, kv, argsort_fwd, mask_fwd, M, nullptr, table_fwd, tmask_fwd, stream)) return rc;
        if (table_bwd)
            if (int rc = spx_build_tile_table(pair_bwd, N, kv, argsort_bwd, mask_bwd, N, nullptr, table_bwd, tmask_bwd, stream)) return rc;
        return 0;
    }
    if (int rc = spx_mask_argsort(mask_fwd, argsort_fwd, M, words, kv, do_sort, sort_ws, sort_bytes, stream)) return rc;
    if (table_fwd)
        if (int rc = spx_build_tile_table(pair_fwd, M, kv, argsort_fwd, mask_fwd, M, nullptr, table_fwd, tmask_fwd, stream)) return rc;
    if (argsort_bwd) {
        if (int rc = spx_mask_argsort(mask_bwd, argsort_bwd, N, words, kv, do_sort, sort_ws, sort_bytes, stream)) return rc;
        if (table_bwd)
            if (int rc = spx_build_tile_table(pair_bwd, N, kv, argsort_bwd, mask_bwd, N, nullptr, table_bwd, tmask_bwd, stream)) return rc;
    }
    return 0;
}
