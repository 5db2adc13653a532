// Library core: thread-local error string, launch counter, device check, elementwise epilogue.
#include "common.cuh"

namespace spx {

static thread_local char g_err[1024] = "";
static thread_local int64_t g_launches = 0;
static thread_local int g_family = 0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }
void set_family(int f) { g_family = f; }

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
        cached = p.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}

// x[r, j] = act(x[r, j] + bias[j])  -- InferenceOps.bias_add_act_inplace, inference.py:166-252
template <typename T>
__global__ void bias_act_kernel(T *__restrict__ x, const T *__restrict__ bias, int64_t total, int cols, int act,
                                float alpha) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        float v = to_float(x[i]);
        if (bias) v += to_float(bias[i % cols]);
        x[i] = from_float<T>(apply_act(v, act, alpha));
    }
}

}  // namespace spx

using namespace spx;

extern "C" const char *spx_last_error(void) { return g_err; }
extern "C" int spx_version(void) { return 100; }
extern "C" int spx_last_kernel_family(void) { return g_family; }
extern "C" int64_t spx_launch_count(int reset) {
    int64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

extern "C" int spx_device_check(int dev, int *sm_count_out, int *cc_major, int *cc_minor) {
    cudaDeviceProp p;
    SPX_CHECK_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count_out) *sm_count_out = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    SPX_REQUIRE(p.major == 10, "spconv_b200 is built for sm_100a only; device %d is sm_%d%d", dev, p.major, p.minor);
    return 0;
}

extern "C" int spx_bias_act_inplace(void *x, const void *bias, int64_t rows, int cols, int dtype, int act,
                                    float act_alpha, spx_stream_t stream_) {
    if (rows == 0 || cols == 0) return 0;
    SPX_REQUIRE(x != nullptr, "x is NULL");
    cudaStream_t stream = (cudaStream_t)stream_;
    int64_t total = rows * cols;
    unsigned nblk = (unsigned)(div_up64(total, 256) < 148 * 16 ? div_up64(total, 256) : 148 * 16);
    switch (dtype) {
        case SPX_F32: bias_act_kernel<float><<<nblk, 256, 0, stream>>>((float *)x, (const float *)bias, total, cols, act, act_alpha); break;
        case SPX_F16: bias_act_kernel<__half><<<nblk, 256, 0, stream>>>((__half *)x, (const __half *)bias, total, cols, act, act_alpha); break;
        case SPX_BF16: bias_act_kernel<__nv_bfloat16><<<nblk, 256, 0, stream>>>((__nv_bfloat16 *)x, (const __nv_bfloat16 *)bias, total, cols, act, act_alpha); break;
        default: SPX_REQUIRE(false, "bias_act: unsupported dtype %d", dtype);
    }
    SPX_CHECK_LAUNCH("bias_act_kernel");
    return 0;
}
