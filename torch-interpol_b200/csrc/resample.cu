// Separable resampling along ONE axis of a dense tensor viewed as (outer, n_in, inner):
//   out[o, i, j] = sum_k  w_k(c_i) * sign_k * in[o, fold(start(c_i) + k), j]          (c_i = coords[i])
// This is what interpol.resize computes (interpol/resize.py:91-117): its sampling grid is the tensor
// product of one 1-D coordinate vector per axis, and B-spline weights are separable, so grid_pull on the
// dense (B, *out, D) grid equals one such pass per axis -- without ever materialising the grid (12 of the
// 20 B/voxel of a pull) and with 3 * (ORDER+1) taps per voxel instead of (ORDER+1)^3.  Weights, boundary
// maps and the extrapolation mask are those of the gather kernels (support.cuh: nd.py:31-77, bounds.py,
// nd.py:11-27); the mask of an N-D point is the product of the per-axis masks, applied pass by pass.
#include <cstdio>
#include "support.cuh"

namespace ib200 {

template <typename T> struct __align__(4 * sizeof(T)) Vec4 { T v[4]; };

// taps of output index i: folded input indices and weights (sign and extrapolation mask folded in)
template <typename R, typename T>
__device__ __forceinline__ void resample_taps(const KParams &kp, const T *coords, int i, int n_in, int order, int bound,
                                              int extrapolate, int (&idx)[8], R (&w)[8]) {
    const R c = (R)Traits<T>::load(coords + i);
    bool ok = extrapolate == 1 || (c > Thr<R>::lo(kp, 0) && c < Thr<R>::hi(kp, 0));
    Axis<R, 8> ax;
    if (ok) ok = setup_axis<R, -1, 0, 8>(ax, c, order, bound, n_in, 1, kp);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const bool use = ok && k <= order;
        idx[k] = use ? ax.off[k] : 0;
        w[k] = use ? ax.w[k] : R(0);
    }
}

// inner == 1 (the resampled axis is contiguous): a thread owns one output index i, keeps its taps in
// registers and walks over the lines; lanes read neighbouring words, writes are coalesced.
template <typename T>
__global__ void __launch_bounds__(256)
resample_last_kernel(const __grid_constant__ KParams kp, const T *__restrict__ in, T *__restrict__ out,
                     const T *__restrict__ coords, const i64 outer, const int n_in, const int n_out,
                     const int order, const int bound, const int extrapolate) {
    typedef typename Traits<T>::Real R;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_out) return;
    int idx[8]; R w[8];
    resample_taps<R, T>(kp, coords, i, n_in, order, bound, extrapolate, idx, w);
    for (i64 o = blockIdx.y; o < outer; o += gridDim.y) {
        const T *line = in + o * n_in;
        R acc = R(0);
        for (int k = 0; k <= order; ++k) acc = fma(w[k], (R)Traits<T>::load(line + idx[k]), acc);
        Traits<T>::store(out + o * n_out + i, acc);
    }
}

// inner > 1: a CTA owns IT consecutive output indices (their taps in shared memory, read as broadcasts) and
// a slice of the (outer, inner) columns; threads run along `inner`, VEC elements each: every access is coalesced.
constexpr int kResIT = 16;

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
resample_inner_kernel(const __grid_constant__ KParams kp, const T *__restrict__ in, T *__restrict__ out,
                      const T *__restrict__ coords, const i64 outer, const int n_in, const int n_out, const i64 inner_v,
                      const int order, const int bound, const int extrapolate) {
    typedef typename Traits<T>::Real R;
    __shared__ int s_idx[kResIT][8];
    __shared__ R s_w[kResIT][8];
    const int i0 = blockIdx.x * kResIT, ni = min(kResIT, n_out - i0);
    if ((int)threadIdx.x < ni) {
        int idx[8]; R w[8];
        resample_taps<R, T>(kp, coords, i0 + threadIdx.x, n_in, order, bound, extrapolate, idx, w);
#pragma unroll
        for (int k = 0; k < 8; ++k) { s_idx[threadIdx.x][k] = idx[k]; s_w[threadIdx.x][k] = w[k]; }
    }
    __syncthreads();
    const i64 ncol = outer * inner_v;
    for (i64 col = (i64)blockIdx.y * 256 + threadIdx.x; col < ncol; col += (i64)gridDim.y * 256) {
        const i64 o = col / inner_v, j = col - o * inner_v;
        const T *base = in + (o * n_in * inner_v + j) * VEC;
        T *dst = out + ((o * n_out + i0) * inner_v + j) * VEC;
        for (int ii = 0; ii < ni; ++ii) {
            R acc[VEC];
#pragma unroll
            for (int q = 0; q < VEC; ++q) acc[q] = R(0);
            for (int k = 0; k <= order; ++k) {
                const T *p = base + (i64)s_idx[ii][k] * inner_v * VEC;
                const R wk = s_w[ii][k];
                if (VEC == 4) {
                    const Vec4<T> v = *reinterpret_cast<const Vec4<T> *>(p);
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[q] = fma(wk, (R)v.v[q], acc[q]);
                } else {
                    acc[0] = fma(wk, (R)Traits<T>::load(p), acc[0]);
                }
            }
            if (VEC == 4) {
                Vec4<T> v;
#pragma unroll
                for (int q = 0; q < 4; ++q) Traits<T>::store(&v.v[q], acc[q]);
                *reinterpret_cast<Vec4<T> *>(dst + (i64)ii * inner_v * VEC) = v;
            } else {
                Traits<T>::store(dst + (i64)ii * inner_v, acc[0]);
            }
        }
    }
}

template <typename T>
static int launch_resample_t(const KParams &kp, const void *in, void *out, const void *coords, i64 outer, i64 n_in, i64 n_out,
                             i64 inner, int order, int bound, int extrapolate, cudaStream_t stream) {
    if (outer * n_out * inner == 0) return IB200_OK;
    if (inner == 1) {
        const unsigned gx = (unsigned)((n_out + 255) / 256);
        i64 gy = ((i64)kNumSMs * 16 + gx - 1) / gx;
        if (gy > outer) gy = outer;
        if (gy > 65535) gy = 65535;
        resample_last_kernel<T><<<dim3(gx, (unsigned)gy), 256, 0, stream>>>(kp, (const T *)in, (T *)out, (const T *)coords, outer,
                                                                            (int)n_in, (int)n_out, order, bound, extrapolate);
        note_launch("resample_axis_last");
    } else {
        const bool vec = inner % 4 == 0 && (uintptr_t)in % (4 * sizeof(T)) == 0 && (uintptr_t)out % (4 * sizeof(T)) == 0;
        const i64 inner_v = vec ? inner / 4 : inner;
        const unsigned gx = (unsigned)((n_out + kResIT - 1) / kResIT);
        const i64 ncol = outer * inner_v;
        i64 gy = ((i64)kNumSMs * 16 + gx - 1) / gx;
        if (gy > (ncol + 255) / 256) gy = (ncol + 255) / 256;
        if (gy > 65535) gy = 65535;
        if (gy < 1) gy = 1;
        if (vec)
            resample_inner_kernel<T, 4><<<dim3(gx, (unsigned)gy), 256, 0, stream>>>(kp, (const T *)in, (T *)out, (const T *)coords, outer,
                                                                                    (int)n_in, (int)n_out, inner_v, order, bound, extrapolate);
        else
            resample_inner_kernel<T, 1><<<dim3(gx, (unsigned)gy), 256, 0, stream>>>(kp, (const T *)in, (T *)out, (const T *)coords, outer,
                                                                                    (int)n_in, (int)n_out, inner_v, order, bound, extrapolate);
        note_launch(vec ? "resample_axis_v4" : "resample_axis");
    }
    IB200_CUDA_CHECK(cudaGetLastError());
    return IB200_OK;
}

// ---- adjoint pass (restrict, backward of resize) ---------------------------------------------------
// out[o, fold(start(c_i) + k), j] += w_k(c_i) * sign_k * in[o, i, j]: the transpose of the pass above,
// i.e. one axis of grid_push on the tensor-product grid (interpol/restrict.py:86-120).  Global float
// REDs, (order+1) per element and axis instead of (order+1)^dim per voxel; threads run along `inner`
// (or along i when the axis is contiguous), so the REDs of a warp are coalesced.  `out` is zero-filled
// by the launcher.  float32 / float64 storage.
template <typename T>
__global__ void __launch_bounds__(256)
resample_adjoint_last_kernel(const __grid_constant__ KParams kp, const T *__restrict__ in, T *__restrict__ out,
                             const T *__restrict__ coords, const i64 outer, const int n_in, const int n_out,
                             const int order, const int bound, const int extrapolate) {
    typedef typename Traits<T>::Real R;
    const int i = blockIdx.x * 256 + threadIdx.x;      // index of the SOURCE sample (one coordinate each)
    if (i >= n_in) return;
    int idx[8]; R w[8];
    resample_taps<R, T>(kp, coords, i, n_out, order, bound, extrapolate, idx, w);
    for (i64 o = blockIdx.y; o < outer; o += gridDim.y) {
        const R v = (R)Traits<T>::load(in + o * n_in + i);
        T *line = out + o * n_out;
        for (int k = 0; k <= order; ++k)
            if (w[k] != R(0)) atomicAdd(line + idx[k], (T)(w[k] * v));
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
resample_adjoint_inner_kernel(const __grid_constant__ KParams kp, const T *__restrict__ in, T *__restrict__ out,
                              const T *__restrict__ coords, const i64 outer, const int n_in, const int n_out, const i64 inner,
                              const int order, const int bound, const int extrapolate) {
    typedef typename Traits<T>::Real R;
    __shared__ int s_idx[kResIT][8];
    __shared__ R s_w[kResIT][8];
    const int i0 = blockIdx.x * kResIT, ni = min(kResIT, n_in - i0);
    if ((int)threadIdx.x < ni) {
        int idx[8]; R w[8];
        resample_taps<R, T>(kp, coords, i0 + threadIdx.x, n_out, order, bound, extrapolate, idx, w);
#pragma unroll
        for (int k = 0; k < 8; ++k) { s_idx[threadIdx.x][k] = idx[k]; s_w[threadIdx.x][k] = w[k]; }
    }
    __syncthreads();
    const i64 ncol = outer * inner;
    for (i64 col = (i64)blockIdx.y * 256 + threadIdx.x; col < ncol; col += (i64)gridDim.y * 256) {
        const i64 o = col / inner, j = col - o * inner;
        const T *src = in + (o * n_in + i0) * inner + j;
        T *base = out + o * n_out * inner + j;
        for (int ii = 0; ii < ni; ++ii) {
            const R v = (R)Traits<T>::load(src + (i64)ii * inner);
            for (int k = 0; k <= order; ++k) {
                const R wk = s_w[ii][k];
                if (wk != R(0)) atomicAdd(base + (i64)s_idx[ii][k] * inner, (T)(wk * v));
            }
        }
    }
}

template <typename T>
static int launch_resample_adjoint_t(const KParams &kp, const void *in, void *out, const void *coords, i64 outer, i64 n_in,
                                     i64 n_out, i64 inner, int order, int bound, int extrapolate, cudaStream_t stream) {
    IB200_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)(outer * n_out * inner) * sizeof(T), stream));
    if (outer * n_in * inner == 0) return IB200_OK;
    if (inner == 1) {
        const unsigned gx = (unsigned)((n_in + 255) / 256);
        i64 gy = ((i64)kNumSMs * 16 + gx - 1) / gx;
        if (gy > outer) gy = outer;
        if (gy > 65535) gy = 65535;
        resample_adjoint_last_kernel<T><<<dim3(gx, (unsigned)gy), 256, 0, stream>>>(kp, (const T *)in, (T *)out, (const T *)coords, outer,
                                                                                    (int)n_in, (int)n_out, order, bound, extrapolate);
        note_launch("resample_adjoint_last");
    } else {
        const unsigned gx = (unsigned)((n_in + kResIT - 1) / kResIT);
        const i64 ncol = outer * inner;
        i64 gy = ((i64)kNumSMs * 16 + gx - 1) / gx;
        if (gy > (ncol + 255) / 256) gy = (ncol + 255) / 256;
        if (gy > 65535) gy = 65535;
        if (gy < 1) gy = 1;
        resample_adjoint_inner_kernel<T><<<dim3(gx, (unsigned)gy), 256, 0, stream>>>(kp, (const T *)in, (T *)out, (const T *)coords, outer,
                                                                                     (int)n_in, (int)n_out, inner, order, bound, extrapolate);
        note_launch("resample_adjoint");
    }
    IB200_CUDA_CHECK(cudaGetLastError());
    return IB200_OK;
}

int launch_resample_adjoint(const KParams &kp, int dtype, const void *in, void *out, const void *coords, i64 outer, i64 n_in,
                            i64 n_out, i64 inner, int order, int bound, int extrapolate, cudaStream_t stream) {
    switch (dtype) {
    case IB200_F32: return launch_resample_adjoint_t<float>(kp, in, out, coords, outer, n_in, n_out, inner, order, bound, extrapolate, stream);
    case IB200_F64: return launch_resample_adjoint_t<double>(kp, in, out, coords, outer, n_in, n_out, inner, order, bound, extrapolate, stream);
    }
    return IB200_ERR_DTYPE;
}

int launch_resample(const KParams &kp, int dtype, const void *in, void *out, const void *coords, i64 outer, i64 n_in, i64 n_out,
                    i64 inner, int order, int bound, int extrapolate, cudaStream_t stream) {
    switch (dtype) {
    case IB200_F32: return launch_resample_t<float>(kp, in, out, coords, outer, n_in, n_out, inner, order, bound, extrapolate, stream);
    case IB200_F16: return launch_resample_t<__half>(kp, in, out, coords, outer, n_in, n_out, inner, order, bound, extrapolate, stream);
    case IB200_BF16: return launch_resample_t<__nv_bfloat16>(kp, in, out, coords, outer, n_in, n_out, inner, order, bound, extrapolate, stream);
    case IB200_F64: return launch_resample_t<double>(kp, in, out, coords, outer, n_in, n_out, inner, order, bound, extrapolate, stream);
    }
    return IB200_ERR_DTYPE;
}

}  // namespace ib200
