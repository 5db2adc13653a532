// Weight-gradient exchange between data-parallel ranks over NVLink peer memory.
//
// The only tensor that crosses GPUs on this path is dW (the reference itself has no
// distributed code).  dW is small (27*C*K values: 0.4 MB fp32 at C = K = 64), so the exchange is
// latency-bound and a library all-reduce costs more in launches and protocol than in bytes.  Here it
// is the tail of the weight-gradient reduction itself:
//
//   phase 1  every CTA sums its slice of the split-K partials (or reads its slice of an existing
//            gradient) in fp32 and PUSHES the slice into every rank's exchange buffer
//            [slot][source rank][element] with plain vector stores through the peer mapping;
//            then one release-add per CTA on every rank's arrival counter for this source.
//   phase 2  every CTA waits until all sources have fully arrived (acquire loads on its OWN
//            counters), then sums the world's slices from LOCAL memory in rank order -- the same order
//            on every rank, so all replicas get bit-identical gradients -- scales, rounds once and
//            writes dW.
//
// One kernel, no grid-wide barrier, no host involvement, CUDA-graph capturable (the epoch lives in
// device memory).  Phase 1 never waits, and the grid never exceeds what is co-resident on an empty
// GPU, so ranks cannot deadlock each other; a peer that never shows up trips the group's timeout (error
// word + NaN result) instead of hanging the GPU.  Two slots alternate by epoch: a rank can only reach
// epoch e+2 after every peer has started epoch e+1, i.e. finished reading epoch e (exchange calls of
// one peer group must be stream-ordered on each rank, and issued in the same order on all ranks).
#include "common.cuh"
#include "gemm.cuh"
#include "peer.cuh"

namespace spx {

// ---- exchange buffer layout (bytes from the base of each rank's buffer)
//   [0, 64)      local state: epoch, finished CTAs of the running call, arrivals expected so far, error
//   [256, 320)   arrival counters, one per source rank (written by the peers)
//   [4096, ...)  data [2 slots][world][capacity] fp32
constexpr size_t PEER_COUNTERS = 256, PEER_DATA = 4096;

struct PeerState { unsigned epoch, finished, expected, error; };

__device__ __forceinline__ void red_release_sys_add(unsigned *addr, unsigned v) {
    asm volatile("red.release.sys.global.add.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *addr) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct PeerPtrs { char *buf[SPX_MAX_PEERS]; };

constexpr int PX_THREADS = 256, PX_WARPS = PX_THREADS / 32;

// partial != nullptr: slice sums of [chunks] split-K partials (the fused weight-gradient tail);
// otherwise the contribution is read from `src` (T).
template <typename T>
__global__ void __launch_bounds__(PX_THREADS)
peer_reduce_exchange_kernel(const float *__restrict__ partial, int64_t stride, int chunks, const T *src, int64_t total,
                            T *dst, PeerPtrs peers, int world, int rank, int64_t capacity, float scale,
                            unsigned long long timeout_ns) {
    __shared__ float4 acc_s[PX_WARPS][32];
    __shared__ unsigned s_epoch, s_expected;
    char *mine = peers.buf[rank];
    PeerState *st = reinterpret_cast<PeerState *>(mine);
    if (threadIdx.x == 0) { s_epoch = st->epoch; s_expected = st->expected; }
    __syncthreads();
    const unsigned epoch = s_epoch;
    const int slot = epoch & 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t slot_off = (int64_t)slot * world * capacity;

    // ---- phase 1: my slices -> every rank's buffer
    if (partial) {
        const int64_t items = (total + 127) / 128;                  // one item = 32 lanes x float4
        const int per = (chunks + PX_WARPS - 1) / PX_WARPS;
        const int c0 = warp * per, c1 = min(chunks, c0 + per);
        for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
            const int64_t i = (it * 32 + lane) * 4;
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < total) {
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
                for (int w = 1; w < PX_WARPS; ++w) {
                    const float4 v = acc_s[w][lane];
                    t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
                }
                for (int p = 0; p < world; ++p) {
                    const int q = (rank + p) % world;               // start at home, spread the link load
                    float *data = reinterpret_cast<float *>(peers.buf[q] + PEER_DATA);
                    *reinterpret_cast<float4 *>(data + slot_off + (int64_t)rank * capacity + i) = t;
                }
            }
            __syncthreads();
        }
    } else {
        const int64_t n4 = total / 4;
        for (int64_t g = (int64_t)blockIdx.x * PX_THREADS + threadIdx.x; g < n4; g += (int64_t)gridDim.x * PX_THREADS) {
            const int64_t i = g * 4;
            const float4 t = make_float4(to_float<T>(src[i]), to_float<T>(src[i + 1]), to_float<T>(src[i + 2]),
                                         to_float<T>(src[i + 3]));
            for (int p = 0; p < world; ++p) {
                const int q = (rank + p) % world;
                float *data = reinterpret_cast<float *>(peers.buf[q] + PEER_DATA);
                *reinterpret_cast<float4 *>(data + slot_off + (int64_t)rank * capacity + i) = t;
            }
        }
        if (blockIdx.x == 0 && threadIdx.x < (int)(total - n4 * 4)) {       // tail of an odd-sized tensor
            const int64_t i = n4 * 4 + threadIdx.x;
            const float t = to_float<T>(src[i]);
            for (int q = 0; q < world; ++q)
                reinterpret_cast<float *>(peers.buf[q] + PEER_DATA)[slot_off + (int64_t)rank * capacity + i] = t;
        }
    }
    __threadfence_system();                                          // my stores before my CTA's arrival
    __syncthreads();
    if (threadIdx.x < world) {
        unsigned *ctr = reinterpret_cast<unsigned *>(peers.buf[threadIdx.x] + PEER_COUNTERS) + rank;
        red_release_sys_add(ctr, 1u);
    }

    // ---- phase 2: wait for every source, then reduce locally in rank order
    const unsigned target = s_expected + gridDim.x;
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    if (threadIdx.x < world) {
        const unsigned *ctr = reinterpret_cast<const unsigned *>(mine + PEER_COUNTERS) + threadIdx.x;
        const unsigned long long t0 = globaltimer_ns();
        unsigned spins = 0;
        while ((int)(ld_acquire_sys(ctr) - target) < 0) {
            if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > timeout_ns) { s_bad = 1; break; }
        }
    }
    __syncthreads();
    const bool bad = s_bad != 0;
    if (bad && threadIdx.x == 0) st->error = 1u;
    {
        const float *data = reinterpret_cast<const float *>(mine + PEER_DATA) + slot_off;
        const int64_t n4 = total / 4;
        const float nan = __int_as_float(0x7fc00000);
        for (int64_t g = (int64_t)blockIdx.x * PX_THREADS + threadIdx.x; g < n4; g += (int64_t)gridDim.x * PX_THREADS) {
            const int64_t i = g * 4;
            float4 t = __ldcg(reinterpret_cast<const float4 *>(data + i));
            for (int r = 1; r < world; ++r) {
                const float4 v = __ldcg(reinterpret_cast<const float4 *>(data + (int64_t)r * capacity + i));
                t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
            }
            if (bad) t = make_float4(nan, nan, nan, nan);
            dst[i] = from_float<T>(t.x * scale); dst[i + 1] = from_float<T>(t.y * scale);
            dst[i + 2] = from_float<T>(t.z * scale); dst[i + 3] = from_float<T>(t.w * scale);
        }
        if (blockIdx.x == 0 && threadIdx.x < (int)(total - n4 * 4)) {
            const int64_t i = n4 * 4 + threadIdx.x;
            float t = __ldcg(data + i);
            for (int r = 1; r < world; ++r) t += __ldcg(data + (int64_t)r * capacity + i);
            dst[i] = from_float<T>(bad ? nan : t * scale);
        }
    }

    // ---- the last CTA out advances the epoch (kernels of one group are stream-ordered)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&st->finished, 1u) == gridDim.x - 1) {
            st->finished = 0;
            st->expected = target;
            st->epoch = epoch + 1;
            __threadfence();
        }
    }
}

static int px_grid(int64_t total, int colocated) {
    // co-resident by construction: 2 CTAs of 256 threads per SM at most (an SM holds 8).  When several ranks
    // share this device (single-GPU protocol tests only) a waiting exchange CTA could keep another rank's
    // weight-gradient CTA (57 k registers) off its SM, so all of them together stay on a quarter of the SMs.
    int64_t items = div_up64(total, 128);
    int64_t cap = colocated > 1 ? sm_count() / (4 * colocated) : (int64_t)sm_count() * 2;
    if (cap < 1) cap = 1;
    return (int)(items < cap ? items : cap);
}

static int check_group(const spx_peer_group *pg, int64_t total, const char *who) {
    SPX_REQUIRE(pg != nullptr, "%s: peer group is NULL", who);
    SPX_REQUIRE(pg->world >= 1 && pg->world <= SPX_MAX_PEERS && pg->rank >= 0 && pg->rank < pg->world,
                "%s: bad peer group (world %d, rank %d)", who, pg->world, pg->rank);
    SPX_REQUIRE((uint64_t)total * 4 <= pg->capacity_bytes, "%s: %lld fp32 values exceed the exchange capacity of %llu bytes",
                who, (long long)total, (unsigned long long)pg->capacity_bytes);
    for (int r = 0; r < pg->world; ++r) SPX_REQUIRE(pg->buffers[r] != nullptr, "%s: buffer of rank %d is NULL", who, r);
    return 0;
}

int peer_reduce_exchange(const float *partial, int64_t stride, int chunks, const void *src, int64_t total, void *dst,
                         int dtype, const spx_peer_group *pg, float scale, cudaStream_t stream) {
    if (int rc = check_group(pg, total, "peer exchange")) return rc;
    PeerPtrs pp;
    memset(&pp, 0, sizeof(pp));
    for (int r = 0; r < pg->world; ++r) pp.buf[r] = (char *)pg->buffers[r];
    const int64_t cap = (int64_t)(pg->capacity_bytes / 4);
    const int grid = px_grid(total, pg->colocated);
    const unsigned long long timeout_ns = (unsigned long long)(pg->timeout_ms > 0 ? pg->timeout_ms : 20000) * 1000000ull;
    if (total == 0) return 0;
#define PX_LAUNCH(T)                                                                                                   \
    peer_reduce_exchange_kernel<T><<<grid, PX_THREADS, 0, stream>>>(partial, stride, chunks, (const T *)src, total,    \
                                                                    (T *)dst, pp, pg->world, pg->rank, cap, scale, timeout_ns)
    if (dtype == SPX_F16) PX_LAUNCH(__half);
    else if (dtype == SPX_BF16) PX_LAUNCH(__nv_bfloat16);
    else if (dtype == SPX_F32) PX_LAUNCH(float);
    else { set_error("peer exchange: dtype %d not supported", dtype); return 2; }
#undef PX_LAUNCH
    SPX_CHECK_LAUNCH("peer_reduce_exchange_kernel");
    return 0;
}

}  // namespace spx

using namespace spx;

extern "C" size_t spx_peer_buffer_bytes(size_t capacity_bytes, int world) {
    if (world < 1 || world > SPX_MAX_PEERS) return 0;
    return PEER_DATA + 2 * (size_t)world * align_up(capacity_bytes, 16);
}

extern "C" int spx_peer_buffer_create(size_t capacity_bytes, int world, void **buffer, unsigned char handle[64]) {
    SPX_REQUIRE(buffer && handle, "peer_buffer_create: NULL output");
    SPX_REQUIRE(world >= 1 && world <= SPX_MAX_PEERS, "peer_buffer_create: world %d out of range (1..%d)", world, SPX_MAX_PEERS);
    SPX_REQUIRE(capacity_bytes % 16 == 0 && capacity_bytes > 0, "peer_buffer_create: capacity must be a positive multiple of 16");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    void *p = nullptr;
    const size_t bytes = spx_peer_buffer_bytes(capacity_bytes, world);
    SPX_CHECK_CUDA(cudaMalloc(&p, bytes));
    SPX_CHECK_CUDA(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(p);
        set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return 1;
    }
    memcpy(handle, &h, 64);
    *buffer = p;
    return 0;
}

extern "C" int spx_peer_buffer_open(const unsigned char handle[64], void **mapped) {
    SPX_REQUIRE(handle && mapped, "peer_buffer_open: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    SPX_CHECK_CUDA(cudaIpcOpenMemHandle(mapped, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int spx_peer_buffer_close(void *mapped) {
    if (mapped) SPX_CHECK_CUDA(cudaIpcCloseMemHandle(mapped));
    return 0;
}

extern "C" int spx_peer_buffer_destroy(void *buffer) {
    if (buffer) SPX_CHECK_CUDA(cudaFree(buffer));
    return 0;
}

extern "C" int spx_peer_error(const spx_peer_group *pg, int *error) {
    SPX_REQUIRE(pg && error, "peer_error: NULL argument");
    PeerState st;
    SPX_CHECK_CUDA(cudaMemcpy(&st, pg->buffers[pg->rank], sizeof(st), cudaMemcpyDeviceToHost));
    *error = (int)st.error;
    return 0;
}

extern "C" int spx_peer_allreduce(const spx_peer_group *pg, void *data, int64_t count, int dtype, float scale,
                                  spx_stream_t stream) {
    SPX_REQUIRE(data != nullptr || count == 0, "peer_allreduce: data is NULL");
    return peer_reduce_exchange(nullptr, 0, 0, data, count, data, dtype, pg, scale, (cudaStream_t)stream);
}
