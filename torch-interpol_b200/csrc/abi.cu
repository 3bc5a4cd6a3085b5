// extern "C" entry points of libinterpol_b200.so: argument validation,
// translation of ib200_problem into kernel parameters, kernel selection.
// See include/interpol_b200.h for the contract of every function.
#include <atomic>
#include <cstdio>
#include <cstring>
#include "common.cuh"

namespace ib200 {

static std::atomic<uint64_t> g_launches{0};
static thread_local char g_last_kernel[96] = "";

void note_launch(const char *kernel_name) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    strncpy(g_last_kernel, kernel_name, sizeof(g_last_kernel) - 1);
    g_last_kernel[sizeof(g_last_kernel) - 1] = 0;
}

static size_t dtype_size(int dtype) {
    switch (dtype) {
    case IB200_F16: case IB200_BF16: return 2;
    case IB200_F32: return 4;
    case IB200_F64: return 8;
    }
    return 0;
}

// round a double threshold the way a comparison against a tensor of `dtype` does
static double round_to_dtype(double v, int dtype) {
    switch (dtype) {
    case IB200_F16: return (double)__half2float(__float2half_rn((float)v));
    case IB200_BF16: return (double)__bfloat162float(__float2bfloat16_rn((float)v));
    case IB200_F32: return (double)(float)v;
    }
    return v;
}

enum { NEED_VOL_IN = 1, NEED_IMG_IN = 2, IMG_HAS_COMP = 4 };

static int build_params(const ib200_problem *p, int need, KParams &kp) {
    if (!p) return IB200_ERR_NULL;
    if (p->dim < 1 || p->dim > 3) return IB200_ERR_DIM;
    if (dtype_size(p->dtype) == 0) return IB200_ERR_DTYPE;
    if (p->extrapolate < 0 || p->extrapolate > 2) return IB200_ERR_EXTRAPOLATE;
    if (p->batch < 0 || p->channels < 0) return IB200_ERR_SHAPE;
    memset(&kp, 0, sizeof(kp));
    kp.dim = p->dim;
    kp.extrapolate = p->extrapolate;
    kp.flags = p->flags;
    kp.batch = p->batch;
    kp.channels = p->channels;
    kp.round_nearest = 1;
    kp.all_linear = 1;
    kp.pts_total = 1;
    kp.vol_total = 1;
    for (int d = 0; d < 3; ++d) {
        kp.vol_n[d] = 1; kp.pts_n[d] = 1; kp.bound[d] = 0; kp.order[d] = 0;
    }
    for (int d = 0; d < p->dim; ++d) {
        if (p->bound[d] < 0 || p->bound[d] > 6) return IB200_ERR_BOUND;
        if (p->order[d] < 0 || p->order[d] > 7) return IB200_ERR_ORDER;
        if (p->vol_shape[d] < 1 || p->pts_shape[d] < 0) return IB200_ERR_SHAPE;
        if (p->vol_shape[d] > 0x7fffffffLL || p->pts_shape[d] > 0x7fffffffLL) return IB200_ERR_TOO_LARGE;
        // the boundary maps fold with 2 * n and 2 * (n + 1) in 32-bit arithmetic (support.cuh)
        if (p->vol_shape[d] > 0x3fffffffLL) return IB200_ERR_TOO_LARGE;
        kp.bound[d] = p->bound[d];
        kp.order[d] = p->order[d];
        kp.vol_n[d] = (int)p->vol_shape[d];
        kp.pts_n[d] = (int)p->pts_shape[d];
        kp.pts_total *= p->pts_shape[d];
        kp.vol_total *= p->vol_shape[d];
        if (p->order[d] != 0) kp.round_nearest = 0;
        if (p->order[d] != 1) kp.all_linear = 0;
        // nd.py:15-26: python-double thresholds cast to the grid dtype by the comparison
        const double thr = p->extrapolate == 2 ? 0.5 + 5e-2 : 5e-2;
        const double lo = round_to_dtype(-thr, p->dtype);
        const double hi = round_to_dtype((double)(p->vol_shape[d] - 1) + thr, p->dtype);
        kp.thr_lo[d] = (float)lo; kp.thr_hi[d] = (float)hi;
        kp.thr_lo_d[d] = lo; kp.thr_hi_d[d] = hi;
    }
    if (kp.vol_total > 0x7fffffffLL) return IB200_ERR_TOO_LARGE;

    // volume strides: honoured for an input volume, dense for an output volume
    i64 dense = 1;
    i64 dense_s[3] = {0, 0, 0};
    for (int d = p->dim - 1; d >= 0; --d) { dense_s[d] = dense; dense *= p->vol_shape[d]; }
    if (need & NEED_VOL_IN) {
        kp.vol_sb = p->vol_stride[0]; kp.vol_sc = p->vol_stride[1];
        i64 reach = 0;
        for (int d = 0; d < p->dim; ++d) {
            kp.vol_s[d] = p->vol_stride[2 + d];
            if (kp.vol_s[d] < 0) return IB200_ERR_SHAPE;
            reach += kp.vol_s[d] * (p->vol_shape[d] - 1);
        }
        if (reach > 0x7fffffffLL) return IB200_ERR_TOO_LARGE;
    } else {
        kp.vol_sc = kp.vol_total; kp.vol_sb = kp.vol_total * kp.channels;
        for (int d = 0; d < p->dim; ++d) kp.vol_s[d] = dense_s[d];
    }
    // grid strides
    kp.grid_sb = p->grid_stride[0];
    for (int d = 0; d < p->dim; ++d) kp.grid_s[d] = p->grid_stride[1 + d];
    kp.grid_sd = p->grid_stride[1 + p->dim];
    bool pts_dense = (kp.grid_sd == 1);
    {
        i64 acc = p->dim;
        for (int d = p->dim - 1; d >= 0; --d) {
            if (p->pts_shape[d] != 1 && kp.grid_s[d] != acc) pts_dense = false;
            acc *= p->pts_shape[d];
        }
    }
    // lattice image strides
    if (need & NEED_IMG_IN) {
        kp.img_sb = p->img_stride[0]; kp.img_sc = p->img_stride[1];
        for (int d = 0; d < p->dim; ++d) kp.img_s[d] = p->img_stride[2 + d];
        kp.img_sd = (need & IMG_HAS_COMP) ? p->img_stride[2 + p->dim] : 0;
        i64 acc = (need & IMG_HAS_COMP) ? p->dim : 1;
        if ((need & IMG_HAS_COMP) && kp.img_sd != 1) pts_dense = false;
        for (int d = p->dim - 1; d >= 0; --d) {
            if (p->pts_shape[d] != 1 && kp.img_s[d] != acc) pts_dense = false;
            acc *= p->pts_shape[d];
        }
    }
    kp.pts_dense = pts_dense ? 1 : 0;
    return IB200_OK;
}

static int run_gather(int op, const ib200_problem *p, const void *vol, const void *grid,
                      const void *gout, void *out, void *stream) {
    KParams kp;
    const bool fused_bwd = op == OP_PULL_BWD_GRID || op == OP_GRAD_BWD_GRID;
    int need = NEED_VOL_IN | (fused_bwd ? NEED_IMG_IN : 0) | (op == OP_GRAD_BWD_GRID ? IMG_HAS_COMP : 0);
    int st = build_params(p, need, kp);
    if (st != IB200_OK) return st;
    if (kp.batch * kp.pts_total == 0 || (kp.channels == 0 && !fused_bwd)) return IB200_OK;
    if (!vol || !grid || !out || (fused_bwd && !gout)) return IB200_ERR_NULL;
    DeviceGuard guard(p->device);
    if (!guard.ok) return IB200_ERR_CUDA - (int)cudaErrorInvalidDevice;
    cudaStream_t s = (cudaStream_t)stream;
    if ((op == OP_PULL || op == OP_GRAD || op == OP_PULL_BWD_GRID) && !(p->flags & IB200_FLAG_NO_TILES)) {
        st = try_pull_pipe(op, kp, p->dtype, vol, grid, gout, out, s);
        if (st != 0) return st < 0 ? st : IB200_OK;
        st = try_pull_tiled(op, kp, p->dtype, vol, grid, gout, out, s);
        if (st != 0) return st < 0 ? st : IB200_OK;
    }
    switch (p->dtype) {
    case IB200_F32: return launch_gather_f32(op, kp, vol, grid, gout, out, s);
    case IB200_F64: return launch_gather_f64(op, kp, vol, grid, gout, out, s);
    case IB200_F16: return launch_gather_f16(op, kp, vol, grid, gout, out, s);
    case IB200_BF16: return launch_gather_bf16(op, kp, vol, grid, gout, out, s);
    }
    return IB200_ERR_DTYPE;
}

static int run_scatter(int op, const ib200_problem *p, const void *img, const void *grid,
                       void *out, void *scratch, void *stream) {
    KParams kp;
    int need = (op == OP_COUNT) ? 0 : (NEED_IMG_IN | (op == OP_PUSHGRAD ? IMG_HAS_COMP : 0));
    ib200_problem q;
    if (!p) return IB200_ERR_NULL;
    q = *p;
    if (op == OP_COUNT) q.channels = 1;
    int st = build_params(&q, need, kp);
    if (st != IB200_OK) return st;
    if (kp.batch * kp.channels * kp.vol_total == 0) return IB200_OK;
    if (!grid || !out || (op != OP_COUNT && !img)) {
        if (kp.batch * kp.pts_total != 0 || !out) return IB200_ERR_NULL;
    }
    DeviceGuard guard(p->device);
    if (!guard.ok) return IB200_ERR_CUDA - (int)cudaErrorInvalidDevice;
    cudaStream_t s = (cudaStream_t)stream;
    if (!(p->flags & IB200_FLAG_NO_TILES) && push_tiled_applicable(op, kp, p->dtype)) {
        const bool half = p->dtype != IB200_F32;
        if (half && !scratch) return IB200_ERR_SCRATCH;
        void *acc = half ? scratch : out;
        const i64 n = kp.batch * kp.channels * kp.vol_total;
        IB200_CUDA_CHECK(cudaMemsetAsync(acc, 0, (size_t)n * sizeof(float), s));
        st = try_push_box(op, kp, p->dtype, img, grid, acc, s);
        if (st < 0) return st;
        if (st == 0) st = try_push_tiled(op, kp, p->dtype, img, grid, acc, s);
        if (st < 0) return st;
        if (st > 0) return half ? convert_from_f32(p->dtype, acc, out, n, s) : IB200_OK;
    }
    switch (p->dtype) {
    case IB200_F32: return launch_scatter_f32(op, kp, img, grid, out, scratch, s);
    case IB200_F64: return launch_scatter_f64(op, kp, img, grid, out, scratch, s);
    case IB200_F16: return launch_scatter_f16(op, kp, img, grid, out, scratch, s);
    case IB200_BF16: return launch_scatter_bf16(op, kp, img, grid, out, scratch, s);
    }
    return IB200_ERR_DTYPE;
}

}  // namespace ib200

using namespace ib200;

extern "C" {

int ib200_pull(const ib200_problem *p, const void *vol, const void *grid, void *out, void *stream) {
    return run_gather(OP_PULL, p, vol, grid, nullptr, out, stream);
}

int ib200_grad(const ib200_problem *p, const void *vol, const void *grid, void *out, void *stream) {
    return run_gather(OP_GRAD, p, vol, grid, nullptr, out, stream);
}

int ib200_hess(const ib200_problem *p, const void *vol, const void *grid, void *out, void *stream) {
    return run_gather(OP_HESS, p, vol, grid, nullptr, out, stream);
}

int ib200_pull_backward_grid(const ib200_problem *p, const void *vol, const void *grid,
                             const void *gout, void *out, void *stream) {
    return run_gather(OP_PULL_BWD_GRID, p, vol, grid, gout, out, stream);
}

int ib200_grad_backward_grid(const ib200_problem *p, const void *vol, const void *grid,
                             const void *gout, void *out, void *stream) {
    return run_gather(OP_GRAD_BWD_GRID, p, vol, grid, gout, out, stream);
}

int ib200_push(const ib200_problem *p, const void *img, const void *grid, void *vol_out,
               void *scratch, void *stream) {
    return run_scatter(OP_PUSH, p, img, grid, vol_out, scratch, stream);
}

int ib200_count(const ib200_problem *p, const void *grid, void *vol_out, void *scratch, void *stream) {
    return run_scatter(OP_COUNT, p, nullptr, grid, vol_out, scratch, stream);
}

int ib200_pushgrad(const ib200_problem *p, const void *img, const void *grid, void *vol_out,
                   void *scratch, void *stream) {
    return run_scatter(OP_PUSHGRAD, p, img, grid, vol_out, scratch, stream);
}

size_t ib200_scratch_bytes(const ib200_problem *p) {
    if (!p || (p->dtype != IB200_F16 && p->dtype != IB200_BF16)) return 0;
    size_t n = (size_t)(p->batch > 0 ? p->batch : 0) * (size_t)(p->channels > 0 ? p->channels : 0);
    for (int d = 0; d < p->dim && d < 3; ++d) n *= (size_t)(p->vol_shape[d] > 0 ? p->vol_shape[d] : 0);
    return n * sizeof(float);
}

int ib200_spline_coeff(void *data, int32_t dtype, int64_t outer, int64_t n, int64_t inner,
                       int32_t bound, int32_t order, int32_t device, void *stream) {
    if (dtype_size(dtype) == 0) return IB200_ERR_DTYPE;
    if (bound < 0 || bound > 6) return IB200_ERR_BOUND;
    if (order < 0 || order > 7) return IB200_ERR_ORDER;
    if (outer < 0 || n < 0 || inner < 0) return IB200_ERR_SHAPE;
    if (n > 0x7fffffffLL) return IB200_ERR_TOO_LARGE;
    if (order <= 1) return IB200_OK;                                  // coeff.py:306-307
    if (bound == IB200_BOUND_DST1 || bound == IB200_BOUND_DST2) return IB200_ERR_BOUND_UNSUPPORTED;
    if (outer * n * inner == 0) return IB200_OK;
    if (!data) return IB200_ERR_NULL;
    DeviceGuard guard(device);
    if (!guard.ok) return IB200_ERR_CUDA - (int)cudaErrorInvalidDevice;
    return launch_coeff(data, dtype, outer, n, inner, bound, order, (cudaStream_t)stream);
}

} // extern "C" (reopened below)

namespace ib200 {
static int run_labels(const ib200_problem *p, const void *vol, const void *grid, void *out, void *stream) {
    KParams kp;
    int st = build_params(p, NEED_VOL_IN, kp);
    if (st != IB200_OK) return st;
    if (kp.batch * kp.pts_total == 0 || kp.channels == 0) return IB200_OK;
    if (!vol || !grid || !out) return IB200_ERR_NULL;
    DeviceGuard guard(p->device);
    if (!guard.ok) return IB200_ERR_CUDA - (int)cudaErrorInvalidDevice;
    return launch_pull_labels(kp, p->dtype, (int)p->reserved, vol, grid, out, (cudaStream_t)stream);
}
}  // namespace ib200

extern "C" {

int ib200_pull_labels(const ib200_problem *p, const void *vol, const void *grid, void *out, void *stream) {
    return run_labels(p, vol, grid, out, stream);
}

static int resample_common(int adjoint, const void *in, void *out, const void *coords, int32_t dtype, int64_t outer, int64_t n_src,
                           int64_t n_dst, int64_t inner, int32_t order, int32_t bound, int32_t extrapolate,
                           int32_t all_nearest, int32_t all_linear, int32_t device, void *stream) {
    // forward: n_src = input extent, n_dst = number of coordinates (outputs); adjoint: n_src = number of
    // coordinates (inputs), n_dst = output extent.  The coordinates always index the axis of extent `n_vol`.
    const int64_t n_vol = adjoint ? n_dst : n_src, n_pts = adjoint ? n_src : n_dst;
    if (dtype_size(dtype) == 0) return IB200_ERR_DTYPE;
    if (adjoint && dtype != IB200_F32 && dtype != IB200_F64) return IB200_ERR_DTYPE;
    if (bound < 0 || bound > 6) return IB200_ERR_BOUND;
    if (order < 0 || order > 7) return IB200_ERR_ORDER;
    if (extrapolate < 0 || extrapolate > 2) return IB200_ERR_EXTRAPOLATE;
    if (outer < 0 || n_vol < 1 || n_pts < 0 || inner < 0) return IB200_ERR_SHAPE;
    if (n_vol > 0x3fffffffLL || n_pts > 0x7fffffffLL || n_vol * inner > 0x7fffffffLL) return IB200_ERR_TOO_LARGE;
    if (outer * n_dst * inner == 0) return IB200_OK;
    if (!out || (outer * n_pts * inner != 0 && (!in || !coords))) return IB200_ERR_NULL;
    KParams kp;
    memset(&kp, 0, sizeof(kp));
    kp.dim = 1;
    kp.extrapolate = extrapolate;
    kp.round_nearest = all_nearest ? 1 : 0;
    kp.all_linear = all_linear ? 1 : 0;
    kp.vol_n[0] = (int)n_vol;
    const double thr = extrapolate == 2 ? 0.5 + 5e-2 : 5e-2;                  // nd.py:15-26
    const double lo = round_to_dtype(-thr, dtype), hi = round_to_dtype((double)(n_vol - 1) + thr, dtype);
    kp.thr_lo[0] = (float)lo; kp.thr_hi[0] = (float)hi; kp.thr_lo_d[0] = lo; kp.thr_hi_d[0] = hi;
    DeviceGuard guard(device);
    if (!guard.ok) return IB200_ERR_CUDA - (int)cudaErrorInvalidDevice;
    if (adjoint)
        return launch_resample_adjoint(kp, dtype, in, out, coords, outer, n_src, n_dst, inner, order, bound, extrapolate, (cudaStream_t)stream);
    return launch_resample(kp, dtype, in, out, coords, outer, n_src, n_dst, inner, order, bound, extrapolate, (cudaStream_t)stream);
}

int ib200_resample_axis(const void *in, void *out, const void *coords, int32_t dtype, int64_t outer, int64_t n_in,
                        int64_t n_out, int64_t inner, int32_t order, int32_t bound, int32_t extrapolate,
                        int32_t all_nearest, int32_t all_linear, int32_t device, void *stream) {
    return resample_common(0, in, out, coords, dtype, outer, n_in, n_out, inner, order, bound, extrapolate, all_nearest, all_linear, device, stream);
}

int ib200_resample_axis_adjoint(const void *in, void *out, const void *coords, int32_t dtype, int64_t outer, int64_t n_in,
                                int64_t n_out, int64_t inner, int32_t order, int32_t bound, int32_t extrapolate,
                                int32_t all_nearest, int32_t all_linear, int32_t device, void *stream) {
    return resample_common(1, in, out, coords, dtype, outer, n_in, n_out, inner, order, bound, extrapolate, all_nearest, all_linear, device, stream);
}

int ib200_abi_version(void) { return IB200_ABI_VERSION; }

const char *ib200_error_string(int status) {
    switch (status) {
    case IB200_OK: return "success";
    case IB200_ERR_NULL: return "null pointer argument";
    case IB200_ERR_DTYPE: return "unknown dtype code";
    case IB200_ERR_DIM: return "only 1, 2 or 3 spatial dimensions are supported";
    case IB200_ERR_BOUND: return "unknown boundary condition";
    case IB200_ERR_ORDER: return "unknown interpolation order";
    case IB200_ERR_SHAPE: return "invalid shape or stride";
    case IB200_ERR_BOUND_UNSUPPORTED: return "boundary condition not implemented for the spline prefilter";
    case IB200_ERR_SCRATCH: return "16-bit scatter needs a float32 scratch volume";
    case IB200_ERR_EXTRAPOLATE: return "extrapolate must be 0, 1 or 2";
    case IB200_ERR_TOO_LARGE: return "a single volume must have fewer than 2^31 voxels (and fewer than 2^30 along one axis)";
    }
    if (status <= IB200_ERR_CUDA) return cudaGetErrorString((cudaError_t)(IB200_ERR_CUDA - status));
    return "unknown error";
}

const char *ib200_last_kernel(void) { return g_last_kernel; }

uint64_t ib200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
