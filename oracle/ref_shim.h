/*
 * ref_shim.h -- the few tv:: / cumm types the reference's CPU rulebook and gather/scatter text
 * uses, so that text (extracted by oracle/make_ref.py from /root/reference) compiles on its own.
 * TEST INFRASTRUCTURE ONLY; never linked into the product library.
 *
 * cumm (the reference's tensorview / layout dependency, cumm>=0.7.11,<0.8.0) is not vendored in
 * /root/reference, so these restate its published semantics:
 *   tv::array<T, N>            fixed array with brace initialisation and op<arrayops::prod>()
 *   tv::Tensor                 non-owning (pointer, shape, dtype) view; dim(), data_ptr<T>()
 *   TensorGeneric<N, Index>    row-major layout: from_shape(), operator()(idx) -> linear index,
 *                              inverse(index, out) -> coordinates        (cumm/gemm/layout.py)
 *   ConvProblem<NDim>          (N, C, K, input_dims, output_dims, ksize, padding, stride,
 *                              dilation) + check_npq_not_overflow()      (cumm/conv/params.py;
 *                              same rule as spconv/pytorch/ops.py:188-190: int64 keys once
 *                              N * prod(output_dims) reaches INT32_MAX)
 *   tv::dispatch / kernel_1d   dtype dispatch (fp32 / fp64 only here) and the 1-D CPU loop helper
 *                              (one serial range: the reference's default build has no OpenMP in
 *                              GatherCPU, gather.py:25-26; -fopenmp would split it into ranges)
 */
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <tuple>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <type_traits>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TV_HOST_DEVICE_INLINE inline
#define TV_DECLTYPE(x) std::decay_t<decltype(x)>
#define TV_IF_CONSTEXPR constexpr
#define TV_ASSERT_RT_ERR(cond, ...)                                                       \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            std::stringstream ss_;                                                        \
            ss_ << #cond << " assert failed. ";                                           \
            tv::detail::sstream_print(ss_, __VA_ARGS__);                                  \
            throw std::runtime_error(ss_.str());                                          \
        }                                                                                 \
    } while (0)

namespace tv {
namespace detail {
inline void sstream_print(std::stringstream &) {}
template <class T, class... Ts> void sstream_print(std::stringstream &ss, T &&v, Ts &&...rest) {
    ss << v << " ";
    sstream_print(ss, std::forward<Ts>(rest)...);
}
}  // namespace detail

namespace arrayops { struct prod {}; }

template <typename T, size_t N> struct array {
    T d_[N];
    T &operator[](size_t i) { return d_[i]; }
    const T &operator[](size_t i) const { return d_[i]; }
    template <class Op> T op() const {
        static_assert(std::is_same<Op, arrayops::prod>::value, "only prod is used by the reference text");
        T r = T(1);
        for (size_t i = 0; i < N; ++i) r *= d_[i];
        return r;
    }
};

enum DType { float32 = 0, float64 = 1, int32 = 2, int64 = 3 };
struct half_t {};
struct bfloat16_t {};

// row-major N-d accessor: view(i, j, k)
template <typename T, int N> struct TensorView {
    T *p;
    int64_t shape[N];
    template <class... I> T &operator()(I... idx) const {
        const int64_t ii[sizeof...(I)] = {(int64_t)idx...};
        int64_t off = 0;
        for (int a = 0; a < (int)sizeof...(I); ++a) off = off * shape[a] + ii[a];
        return p[off];
    }
};

struct Tensor {
    void *ptr = nullptr;
    std::vector<int64_t> shape;
    DType dt = int32;
    Tensor() = default;
    Tensor(void *p, std::vector<int64_t> s, DType d) : ptr(p), shape(std::move(s)), dt(d) {}
    int64_t dim(int i) const { return shape[i]; }
    int64_t stride(int i) const {                 // contiguous row-major views only
        int64_t s = 1;
        for (size_t j = i + 1; j < shape.size(); ++j) s *= shape[j];
        return s;
    }
    template <typename T> T *data_ptr() const { return reinterpret_cast<T *>(const_cast<void *>(ptr)); }
    template <typename T, int N> TensorView<T, N> tview() const {
        TensorView<T, N> v;
        v.p = data_ptr<T>();
        for (int a = 0; a < N; ++a) v.shape[a] = shape[a];
        return v;
    }
    size_t itemsize() const { return dt == float64 || dt == int64 ? 8 : 4; }
    int64_t numel() const { int64_t n = 1; for (auto s : shape) n *= s; return n; }
    void zero_() { std::memset(ptr, 0, (size_t)numel() * itemsize()); }
    Tensor slice_first_axis(int64_t b, int64_t e) const {
        Tensor r = *this;
        int64_t inner = shape.empty() ? 1 : numel() / (shape[0] ? shape[0] : 1);
        r.ptr = (char *)ptr + (size_t)b * inner * itemsize();
        r.shape[0] = e - b;
        return r;
    }
    DType dtype() const { return dt; }
    int device() const { return -1; }
    bool is_cpu() const { return true; }
};

// dtype dispatch over the floating types the CPU conv path is used with here (fp32, fp64)
template <class... Ts, class F> void dispatch(DType dt, F &&f) {
    if (dt == float32) f(float());
    else if (dt == float64) f(double());
    else throw std::runtime_error("ref_shim: dispatch supports float32 / float64 only");
}

// 1-D loop helper: f(begin, end, step); contiguous OpenMP ranges
template <class F> void kernel_1d(int /*device*/, int64_t n, F &&f) {
#ifdef _OPENMP
    if (n >= 4096) {
#pragma omp parallel
        {
            const int nt = omp_get_num_threads(), t = omp_get_thread_num();
            const int64_t per = (n + nt - 1) / nt;
            const int64_t b = t * per, e = b + per < n ? b + per : n;
            if (b < e) f((int)b, (int)e, 1);
        }
        return;
    }
#endif
    f(0, (int)n, 1);
}
}  // namespace tv

namespace refshim {

template <int N, typename Index> struct TensorGeneric {
    Index strides[N];
    int shape[N];
    static TensorGeneric from_shape(tv::array<int, N> const &s) {
        TensorGeneric l;
        Index acc = 1;
        for (int i = N - 1; i >= 0; --i) { l.shape[i] = s[i]; l.strides[i] = acc; acc *= (Index)s[i]; }
        return l;
    }
    Index operator()(const int *idx) const {
        Index r = 0;
        for (int i = 0; i < N; ++i) r += (Index)idx[i] * strides[i];
        return r;
    }
    Index operator()(tv::array<int, N> const &idx) const { return (*this)(&idx[0]); }
    void inverse(Index index, tv::array<int, N> &out) const {
        for (int i = 0; i < N; ++i) { out[i] = (int)(index / strides[i]); index -= (Index)out[i] * strides[i]; }
    }
};

template <int NDim> struct ConvProblem {
    int N, C, K;
    tv::array<int, NDim> input_dims, output_dims, ksize, padding, stride, dilation;
    ConvProblem(int n, int c, int k, tv::array<int, NDim> in, tv::array<int, NDim> out, tv::array<int, NDim> ks,
                tv::array<int, NDim> pad, tv::array<int, NDim> st, tv::array<int, NDim> dil)
        : N(n), C(c), K(k), input_dims(in), output_dims(out), ksize(ks), padding(pad), stride(st), dilation(dil) {}
    bool check_npq_not_overflow() const {
        int64_t v = N;
        for (int i = 0; i < NDim; ++i) v *= (int64_t)output_dims[i];
        return v < (int64_t)std::numeric_limits<int32_t>::max();
    }
};

}  // namespace refshim
