// Shared declarations of the interpol_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/interpol_b200.h"

namespace ib200 {

typedef long long i64;

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---------------------------------------------------------------- dtypes --
// Storage type T -> arithmetic type Real (float for 16/32-bit storage,
// double for double: SURVEY quirk Q8, we never accumulate in half).
template <typename T> struct Traits;
template <> struct Traits<float> {
    typedef float Real;
    static __device__ __forceinline__ float load(const float *p) { return __ldg(p); }
    static __device__ __forceinline__ float load_rw(const float *p) { return *p; }
    static __device__ __forceinline__ void store(float *p, float v) { *p = v; }
};
template <> struct Traits<double> {
    typedef double Real;
    static __device__ __forceinline__ double load(const double *p) { return __ldg(p); }
    static __device__ __forceinline__ double load_rw(const double *p) { return *p; }
    static __device__ __forceinline__ void store(double *p, double v) { *p = v; }
};
template <> struct Traits<__half> {
    typedef float Real;
    static __device__ __forceinline__ float load(const __half *p) { return __half2float(__ldg(p)); }
    static __device__ __forceinline__ float load_rw(const __half *p) { return __half2float(*p); }
    static __device__ __forceinline__ void store(__half *p, float v) { *p = __float2half_rn(v); }
};
template <> struct Traits<__nv_bfloat16> {
    typedef float Real;
    static __device__ __forceinline__ float load(const __nv_bfloat16 *p) { return __bfloat162float(__ldg(p)); }
    static __device__ __forceinline__ float load_rw(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }
};

// ------------------------------------------------------------ parameters --
// Plain-old-data kernel argument shared by the gather and scatter kernels.
struct KParams {
    int dim;
    int bound[3];
    int order[3];
    int extrapolate;
    unsigned flags;
    int round_nearest;   // every order == 0: iso0.py:12 uses round-half-even
    int all_linear;      // every order == 1: iso1 closed forms
    int pts_dense;       // grid + image are dense over the lattice: skip index decomposition
    int vol_n[3];        // extents of the volume (1 for unused axes)
    int pts_n[3];        // extents of the lattice (1 for unused axes)
    i64 pts_total;       // lattice points per batch element
    i64 vol_total;       // voxels per (b, c) volume
    i64 batch, channels;
    i64 vol_sb, vol_sc, vol_s[3];              // volume strides (elements)
    i64 grid_sb, grid_s[3], grid_sd;           // grid strides
    i64 img_sb, img_sc, img_s[3], img_sd;      // lattice image strides
    float thr_lo[3], thr_hi[3];                // extrapolate thresholds, float
    double thr_lo_d[3], thr_hi_d[3];           // and double
};

// ----------------------------------------------------------------- errors --
#define IB200_CUDA_CHECK(expr)                                       \
    do {                                                             \
        cudaError_t _e = (expr);                                     \
        if (_e != cudaSuccess) return IB200_ERR_CUDA - (int)_e;      \
    } while (0)

void note_launch(const char *kernel_name);   // abi.cu: bookkeeping for introspection

struct DeviceGuard {
    int prev;
    bool ok;
    explicit DeviceGuard(int dev) : prev(-1), ok(true) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
        want = dev;
    }
    ~DeviceGuard() { if (prev >= 0 && prev != want) cudaSetDevice(prev); }
    int want = -1;
};

// launch helpers implemented per translation unit
// one translation unit per storage type (gather.cu / scatter.cu compiled with -DIB200_T=...)
#define IB200_DECL_LAUNCHERS(NAME)                                                                  \
    int launch_gather_##NAME(int op, const KParams &kp, const void *vol, const void *grid,          \
                             const void *gout, void *out, cudaStream_t stream);                     \
    int launch_scatter_##NAME(int op, const KParams &kp, const void *img, const void *grid,         \
                              void *out, void *scratch, cudaStream_t stream);
IB200_DECL_LAUNCHERS(f32)
IB200_DECL_LAUNCHERS(f64)
IB200_DECL_LAUNCHERS(f16)
IB200_DECL_LAUNCHERS(bf16)
#undef IB200_DECL_LAUNCHERS
// spline orders that get a fully unrolled, register-resident instantiation for the
// storage type being compiled (everything else runs the runtime-order kernel)
#ifndef IB200_STATIC_ORDERS
#define IB200_STATIC_ORDERS(X)
#endif
int launch_coeff(void *data, int dtype, i64 outer, i64 n, i64 inner, int bound, int order,
                 cudaStream_t stream);
// tiled fast paths: return 1 when they handled the call, 0 when not applicable, <0 on error
// (`gout`: lattice image the fused backward OP_PULL_BWD_GRID multiplies the gradient with, NULL otherwise)
int try_pull_tiled(int op, const KParams &kp, int dtype, const void *vol, const void *grid, const void *gout, void *out,
                   cudaStream_t stream);
int try_pull_pipe(int op, const KParams &kp, int dtype, const void *vol, const void *grid, const void *gout, void *out,
                  cudaStream_t stream);
bool push_tiled_applicable(int op, const KParams &kp, int dtype);
int try_push_tiled(int op, const KParams &kp, int dtype, const void *img, const void *grid,
                   void *acc, cudaStream_t stream);   // acc: zero-filled float32 accumulation volume
int try_push_box(int op, const KParams &kp, int dtype, const void *img, const void *grid,
                 void *acc, cudaStream_t stream);    // acc: zero-filled float32 accumulation volume
int convert_from_f32(int dtype, const void *src, void *dst, i64 n, cudaStream_t stream);

int launch_resample(const KParams &kp, int dtype, const void *in, void *out, const void *coords, i64 outer, i64 n_in,
                    i64 n_out, i64 inner, int order, int bound, int extrapolate, cudaStream_t stream);
int launch_resample_adjoint(const KParams &kp, int dtype, const void *in, void *out, const void *coords, i64 outer, i64 n_in,
                            i64 n_out, i64 inner, int order, int bound, int extrapolate, cudaStream_t stream);
int launch_pull_labels(const KParams &kp, int grid_dtype, int label_type, const void *vol, const void *grid, void *out, cudaStream_t stream);

enum { OP_PULL = 0, OP_GRAD = 1, OP_HESS = 2, OP_PULL_BWD_GRID = 3, OP_GRAD_BWD_GRID = 4 };
enum { OP_PUSH = 0, OP_COUNT = 1, OP_PUSHGRAD = 2 };

}  // namespace ib200
