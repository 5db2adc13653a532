// Shared host/device helpers for the sm_100a kernels behind include/spconv_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>
#include <type_traits>

#include "../../include/spconv_b200.h"

namespace spx {

// ---------------------------------------------------------------- error plumbing
void set_error(const char *fmt, ...);
void count_launch(int n = 1);
void set_family(int f);
int sm_count();          // SMs of the current device (cached)
int current_device();    // cudaGetDevice (0 on failure)

// Process-wide switches.  SPX_FORCE_SIMT / SPX_FORCE_TC / SPX_TC_CTAS are read from the
// environment ONCE, when the library is loaded; the perf-triage hooks (ablation bits, timeline
// buffer) are only reachable through spx_debug_configure(), which validates the buffer.
struct RuntimeCfg {
    int force_simt = 0, force_tc = 0, tc_ctas = 2;
    int debug = 0;                 // ablation bits, see gemm_tc.cu / gemm_tc_wgrad.cu
    long long *trace = nullptr;    // [8][2048] int64 device buffer or NULL
};
RuntimeCfg &runtime_cfg();

// cudaFuncSetAttribute is per DEVICE: remember which (function, device) pairs were configured
bool func_configured(const void *fn, int dev);   // returns the previous state and marks it

#define SPX_CHECK_CUDA(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            spx::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                           __FILE__, __LINE__);                                           \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

#define SPX_CHECK_LAUNCH(name)                                                            \
    do {                                                                                  \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            spx::set_error("launch of %s failed: %s (%s:%d)", name,                       \
                           cudaGetErrorString(_e), __FILE__, __LINE__);                   \
            return 1;                                                                     \
        }                                                                                 \
        spx::count_launch();                                                              \
    } while (0)

#define SPX_REQUIRE(cond, ...)                                                            \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            spx::set_error(__VA_ARGS__);                                                  \
            return 2;                                                                     \
        }                                                                                 \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t div_up64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// carve a caller-provided workspace
struct WorkspaceCarver {
    char *base;
    size_t off = 0, cap;
    WorkspaceCarver(void *p, size_t bytes) : base((char *)p), cap(bytes) {}
    template <typename T> T *take(size_t n) {
        off = align_up(off, 256);
        T *r = (T *)(base + off);
        off += n * sizeof(T);
        return r;
    }
    bool ok() const { return off <= cap; }
};

__host__ __device__ static inline int dtype_bytes(int dt) {
    switch (dt) {
        case SPX_F32: return 4;
        case SPX_F16: return 2;
        case SPX_BF16: return 2;
        case SPX_I8: return 1;
    }
    return 0;
}

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

template <typename T> __device__ __forceinline__ float to_float(T v);
template <> __device__ __forceinline__ float to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_float<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_float<int8_t>(int8_t v) { return (float)v; }

template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
    // InferenceOps activations: spconv/csrc/sparse/inference.py:26-146
    switch (act) {
        case SPX_ACT_RELU: return v > 0.f ? v : 0.f;
        case SPX_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
        case SPX_ACT_LEAKY_RELU: return v >= 0.f ? v : v * alpha;
        default: return v;
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Spin on try_wait (which itself suspends in hardware for a while).  A protocol bug must become
// an error, not a hung GPU: after ~2 s of waiting the kernel traps (cudaErrorLaunchFailure).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3FFFu) == 0u && clock64() - t0 > 4000000000ll) {
            printf("spconv_b200: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n",
                   (int)blockIdx.x, (int)threadIdx.x, smem_u32(bar), parity);
            __trap();
        }
    }
}
// arrive (no pending-count increment) once all prior cp.async of this thread have landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- cp.async 16 B with zero fill (src_bytes = 0 -> writes 16 zero bytes, reads nothing)
__device__ __forceinline__ void cp_async_16(uint32_t dst_smem, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// generic-proxy writes -> visible to the async proxy (UMMA / TMA reads of smem)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- TMA (tiled 2D load, mbarrier completion)
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const void *tmap, uint64_t *bar,
                                            int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst_smem), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// 1-D bulk async copy global -> shared (UBLKCP), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void *tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// ---- tcgen05
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS> __device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// signal an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive 32-bit columns: thread t gets lane (base_lane + t), columns c..c+31
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

// UMMA kinds
enum MmaKind { KIND_F16 = 0, KIND_TF32 = 1, KIND_I8 = 2 };

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread
template <int KIND>
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
    if constexpr (KIND == KIND_F16) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
    } else if constexpr (KIND == KIND_TF32) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
    }
}

// Warp-uniform variants: executed by ALL 32 lanes with identical operands, one elected lane
// issues.  Keeping the surrounding control flow uniform lets the compiler hold descriptors in
// uniform registers (no per-MMA R2UR / ELECT retry loop on the single issuing thread).
template <int KIND>
__device__ __forceinline__ void umma_ss_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
    if constexpr (KIND == KIND_F16) {
        asm volatile(
            "{\n\t.reg .pred pe, pa;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %4, 0;\n\t"
            "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
    } else if constexpr (KIND == KIND_TF32) {
        asm volatile(
            "{\n\t.reg .pred pe, pa;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %4, 0;\n\t"
            "@pe tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, pa;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred pe, pa;\n\telect.sync _|pe, 0xffffffff;\n\tsetp.ne.b32 pa, %4, 0;\n\t"
            "@pe tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, pa;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
    }
}
__device__ __forceinline__ void tc_commit_elect(uint64_t *bar) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar)) : "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1"); byte quantities, 16-byte granular.
//   swizzle_bytes in {32, 64, 128}; layout_type: 128B=2, 64B=4, 32B=6
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t swizzle_bytes) {
    uint64_t layout = swizzle_bytes == 128 ? 2ull : (swizzle_bytes == 64 ? 4ull : (swizzle_bytes == 32 ? 6ull : 0ull));
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;                 // descriptor version (Blackwell)
    d |= layout << 61;
    return d;
}

// Constant (address-independent) part of a descriptor; OR it with ((smem_addr >> 4) & 0x3FFF).
__host__ __device__ __forceinline__ uint64_t smem_desc_hi(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t swizzle_bytes) {
    uint64_t layout = swizzle_bytes == 128 ? 2ull : (swizzle_bytes == 64 ? 4ull : (swizzle_bytes == 32 ? 6ull : 0ull));
    return ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           (1ull << 46) | (layout << 61);
}

// same with an explicit layout code (1 = SWIZZLE_128B_BASE32B: 128-byte rows, 32-byte swizzle atoms -- the
// MN-major layout of 32-bit operands)
__host__ __device__ __forceinline__ uint64_t smem_desc_hi_layout(uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout) {
    return ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           (1ull << 46) | (layout << 61);
}

// Instruction descriptor, dense, no negate, no saturate.
//   c_format: 0 f16, 1 f32, 2 s32;  ab_format: kind::f16 -> 0 f16 / 1 bf16; tf32 -> 2; i8 -> 1 (signed)
__host__ __device__ __forceinline__ uint32_t make_idesc(int c_format, int a_format, int b_format, int a_mn_major,
                                                        int b_mn_major, int M, int N) {
    uint32_t d = 0;
    d |= (uint32_t)(c_format & 3) << 4;
    d |= (uint32_t)(a_format & 7) << 7;
    d |= (uint32_t)(b_format & 7) << 10;
    d |= (uint32_t)(a_mn_major & 1) << 15;
    d |= (uint32_t)(b_mn_major & 1) << 16;
    d |= (uint32_t)((N >> 3) & 0x3F) << 17;
    d |= (uint32_t)((M >> 4) & 0x1F) << 24;
    return d;
}

// byte offset inside a swizzled tile whose rows are `swizzle_bytes` long and whose base is
// 1024-byte aligned: XOR the 16-byte chunk index with the row index (Swizzle<B,4,3>).
__device__ __forceinline__ uint32_t swizzle_offset(uint32_t off, uint32_t swizzle_bytes) {
    uint32_t bits = swizzle_bytes == 128 ? 7u : (swizzle_bytes == 64 ? 3u : 1u);
    return off ^ (((off >> 7) & bits) << 4);
}

#endif  // __CUDACC__

}  // namespace spx
