// Probe: TMA tile::gather4 (sm_100) as the row-gather engine of the implicit GEMM A operand.
//   (1) layout check: 128 rows x 128 B gathered 4 rows per instruction into a SWIZZLE_128B atom,
//       invalid rows (index -1 / >= N) must come back as zeros;
//   (2) throughput: stages of 16 KB per CTA, gather4 vs 16-byte cp.async, all SMs busy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gather4_probe gather4_probe.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 2000000000ll) return false;
    }
    return true;
}
__device__ __forceinline__ void gather4(uint32_t dst, const void *tmap, uint64_t *bar, int col, int r0, int r1, int r2, int r3) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void *src, uint32_t bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- (1) layout check
__global__ void layout_kernel(const __grid_constant__ CUtensorMap tm, const int32_t *idx, uint8_t *out, int *status) {
    extern __shared__ __align__(1024) uint8_t raw[];
    const uint32_t a = smem_u32(raw);
    const uint32_t pad = (1024u - (a & 1023u)) & 1023u;
    uint8_t *buf = raw + pad;
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) mbar_expect_tx(&bar, 128 * 128);
    __syncwarp();
    const int l = threadIdx.x;
    gather4(a + pad + l * 512, &tm, &bar, 0, idx[4 * l], idx[4 * l + 1], idx[4 * l + 2], idx[4 * l + 3]);
    const bool ok = mbar_wait(&bar, 0);
    if (threadIdx.x == 0) *status = ok ? 1 : -1;
    __syncwarp();
    for (int i = threadIdx.x; i < 128 * 128 / 16; i += 32)
        reinterpret_cast<uint4 *>(out)[i] = reinterpret_cast<const uint4 *>(buf)[i];
}

// ---------------------------------------------------------------- (2) throughput
constexpr int STAGE_BYTES = 128 * 128;
template <int MODE>   // 0 = gather4 by ISSUERS warps, 1 = cp.async by 8 warps
__global__ void __launch_bounds__(288) thr_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t *x, const int32_t *idx,
                                                  int64_t tiles, int stages, int issuers, long long *cycles, int *fail) {
    extern __shared__ __align__(1024) uint8_t raw[];
    const uint32_t a = smem_u32(raw);
    const uint32_t pad = (1024u - (a & 1023u)) & 1023u;
    const uint32_t base = a + pad;
    __shared__ uint64_t full[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&full[s], MODE == 0 ? 1 : 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    if (MODE == 0) {
        if (warp < issuers) {
            const int per = 32 / issuers;            // gather4 instructions per issuing warp per stage
            int it = 0;
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
                const int s = it % stages;
                if (it >= stages) { if (!mbar_wait(&full[s], (uint32_t)((it / stages - 1) & 1))) { *fail = 1; return; } }
                if (warp == 0 && lane == 0) mbar_expect_tx(&full[s], STAGE_BYTES);
                __syncwarp();
                if (lane < per) {
                    const int g = warp * per + lane;
                    const int4 r = *reinterpret_cast<const int4 *>(idx + t * 128 + g * 4);
                    gather4(base + s * STAGE_BYTES + g * 512, &tm, &full[s], 0, r.x, r.y, r.z, r.w);
                }
            }
            // drain
            for (int k = 0; k < stages && k < it; ++k) {
                const int j = it - 1 - k; const int s = j % stages;
                if (!mbar_wait(&full[s], (uint32_t)((j / stages) & 1))) { *fail = 1; return; }
            }
        }
    } else {
        if (warp < 8) {
            const int r0 = lane >> 3; const uint32_t chb = (lane & 7) << 4;
            int it = 0;
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
                const int s = it % stages;
                if (it >= stages) { if (!mbar_wait(&full[s], (uint32_t)((it / stages - 1) & 1))) { *fail = 1; return; } }
                int32_t ri[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) ri[q] = idx[t * 128 + warp * 16 + r0 + q * 4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t row = warp * 16 + r0 + q * 4;
                    const uint32_t off = row * 128 + (chb ^ ((row & 7) << 4));
                    cp_async_16(base + s * STAGE_BYTES + off, x + (int64_t)max(ri[q], 0) * 128 + chb, ri[q] >= 0 ? 16u : 0u);
                }
                cp_async_arrive(&full[s]);
            }
            for (int k = 0; k < stages && k < it; ++k) {
                const int j = it - 1 - k; const int s = j % stages;
                if (!mbar_wait(&full[s], (uint32_t)((j / stages) & 1))) { *fail = 1; return; }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int64_t N = 100000;
    const int C = 64;
    void *fnp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    std::vector<__half> hx(N * C);
    for (int64_t i = 0; i < N * C; ++i) hx[i] = __float2half((float)((i * 7 + i / C) % 2039) - 1000.f);
    __half *dx; CK(cudaMalloc(&dx, N * C * 2)); CK(cudaMemcpy(dx, hx.data(), N * C * 2, cudaMemcpyHostToDevice));

    for (int boxrows = 1; boxrows <= 4; boxrows += 3) {
        CUtensorMap tm;
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)N};
        cuuint64_t strides[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {(cuuint32_t)C, (cuuint32_t)boxrows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dx, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode box rows %d -> %d\n", boxrows, (int)r);
        if (r != CUDA_SUCCESS) continue;
        std::vector<int32_t> hidx(128);
        srand(7);
        for (int i = 0; i < 128; ++i) hidx[i] = (i % 5 == 3) ? -1 : (i % 17 == 11 ? (int)N + 5 : rand() % (int)N);
        int32_t *didx; uint8_t *dout; int *dstat;
        CK(cudaMalloc(&didx, 512)); CK(cudaMalloc(&dout, 16384)); CK(cudaMalloc(&dstat, 4));
        CK(cudaMemcpy(didx, hidx.data(), 512, cudaMemcpyHostToDevice));
        CK(cudaMemset(dout, 0xAB, 16384)); CK(cudaMemset(dstat, 0, 4));
        CK(cudaFuncSetAttribute(layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
        layout_kernel<<<1, 32, 16384 + 1024>>>(tm, didx, dout, dstat);
        cudaError_t e = cudaDeviceSynchronize();
        int st = 0; cudaMemcpy(&st, dstat, 4, cudaMemcpyDeviceToHost);
        printf("  layout kernel: %s, barrier %s\n", cudaGetErrorString(e), st == 1 ? "completed" : "TIMED OUT");
        if (e != cudaSuccess) { printf("  (sticky error, stopping)\n"); return 1; }
        std::vector<uint8_t> img(16384); CK(cudaMemcpy(img.data(), dout, 16384, cudaMemcpyDeviceToHost));
        int64_t bad = 0, firstbad = -1;
        for (int row = 0; row < 128; ++row)
            for (int c = 0; c < 8; ++c) {
                const uint8_t *got = img.data() + row * 128 + ((c ^ (row & 7)) * 16);
                uint8_t want[16]; memset(want, 0, 16);
                const int ri = hidx[row];
                if (ri >= 0 && ri < N) memcpy(want, (const uint8_t *)(hx.data() + (int64_t)ri * C) + c * 16, 16);
                if (memcmp(got, want, 16)) { if (firstbad < 0) firstbad = row * 8 + c; ++bad; }
            }
        printf("  swizzled-image mismatches (16-byte chunks): %lld of 1024 (first %lld)\n", (long long)bad, (long long)firstbad);
        if (boxrows == 1 && bad == 0) {
            // ---- throughput
            const int64_t tiles = 148 * 2 * 64;
            for (int pass = 0; pass < 2; ++pass) {
                const double keep = pass == 0 ? 1.0 : 0.6;
                std::vector<int32_t> tidx(tiles * 128);
                for (size_t i = 0; i < tidx.size(); ++i) tidx[i] = ((double)rand() / RAND_MAX) < keep ? rand() % (int)N : -1;
                int32_t *dt; CK(cudaMalloc(&dt, tidx.size() * 4)); CK(cudaMemcpy(dt, tidx.data(), tidx.size() * 4, cudaMemcpyHostToDevice));
                long long *dc; CK(cudaMalloc(&dc, 8 * 1024)); int *df; CK(cudaMalloc(&df, 4)); CK(cudaMemset(df, 0, 4));
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                for (int mode = 0; mode < 2; ++mode)
                    for (int cps = 1; cps <= 2; ++cps)
                        for (int issuers = 1; issuers <= (mode == 0 ? 4 : 1); issuers *= 2) {
                            const int stages = 4;
                            const size_t smem = stages * STAGE_BYTES + 1024;
                            auto kern = mode == 0 ? thr_kernel<0> : thr_kernel<1>;
                            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                            const int grid = 148 * cps;
                            float best = 1e9f;
                            for (int rep = 0; rep < 5; ++rep) {
                                cudaEventRecord(e0);
                                kern<<<grid, 288, smem>>>(tm, (const uint8_t *)dx, dt, tiles, stages, issuers, dc, df);
                                cudaEventRecord(e1);
                                cudaError_t er = cudaDeviceSynchronize();
                                if (er != cudaSuccess) { printf("thr kernel error %s\n", cudaGetErrorString(er)); return 1; }
                                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
                            }
                            int f = 0; cudaMemcpy(&f, df, 4, cudaMemcpyDeviceToHost);
                            std::vector<long long> cyc(grid); cudaMemcpy(cyc.data(), dc, grid * 8, cudaMemcpyDeviceToHost);
                            long long mx = 0; for (auto v : cyc) if (v > mx) mx = v;
                            const double per_cta_stages = (double)tiles / grid;
                            printf("  keep %.1f %s ctas/SM %d issuers %d: %.1f us, %.0f cyc/stage/CTA, %.1f B/cyc/SM (slot bytes), %.2f TB/s valid%s\n",
                                   keep, mode == 0 ? "gather4 " : "cp.async", cps, issuers, best * 1e3, mx / per_cta_stages,
                                   (double)tiles * STAGE_BYTES / 148 / mx, tiles * 128.0 * 128 * keep / (best * 1e-3) / 1e12,
                                   f ? "  [BARRIER TIMEOUT]" : "");
                        }
                cudaFree(dt); cudaFree(dc); cudaFree(df);
            }
        }
        cudaFree(didx); cudaFree(dout); cudaFree(dstat);
    }
    return 0;
}
