/*
 * interpol_b200.h -- C ABI of the B200-native spline sampling engine.
 *
 * This is the drop-in boundary for ONE hot path of balbasty/torch-interpol:
 * the non-differentiable forward/backward building blocks that
 * interpol/autograd.py calls in interpol/pushpull.py and interpol/coeff.py.
 * Every entry point below names the reference function it replaces
 * (file:line relative to the reference tree).  Plain pointers and sizes only:
 * no torch types cross this boundary.  All pointers are DEVICE pointers on
 * `device` unless stated otherwise; work is enqueued on `stream`
 * (a cudaStream_t passed as void*) and the call returns without
 * synchronising.  The library never allocates or frees device memory: outputs
 * and scratch are owned by the caller.
 *
 * Layouts: volumes are (B, C, X[, Y[, Z]]), grids are
 * (B, X[, Y[, Z]], D) in voxel units with component d indexing spatial axis d,
 * last spatial axis fastest in the dense case.  Inputs may be arbitrarily
 * strided (element strides, 0 allowed for broadcast batch / channel axes);
 * outputs are always dense C-contiguous.
 *
 * Return value: 0 on success, a negative IB200_ERR_* code otherwise
 * (IB200_ERR_CUDA - cudaError_t for CUDA runtime failures).  No exception ever
 * crosses the boundary; the Python host maps codes to the reference's
 * exception classes (ValueError / NotImplementedError / RuntimeError).
 */
#ifndef INTERPOL_B200_H
#define INTERPOL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IB200_ABI_VERSION 1

#if defined(__GNUC__)
#define IB200_API __attribute__((visibility("default")))
#else
#define IB200_API
#endif

/* dtype codes (storage type; arithmetic is f32 for F16/BF16/F32, f64 for F64) */
enum { IB200_F16 = 0, IB200_F32 = 1, IB200_F64 = 2, IB200_BF16 = 3 };

/* boundary codes == interpol/bounds.py:8-15 (BoundType) */
enum {
    IB200_BOUND_ZERO = 0, IB200_BOUND_REPLICATE = 1, IB200_BOUND_DCT1 = 2,
    IB200_BOUND_DCT2 = 3, IB200_BOUND_DST1 = 4, IB200_BOUND_DST2 = 5,
    IB200_BOUND_DFT = 6
};

/* extrapolate codes == interpol/bounds.py:18-21 (ExtrapolateType) */
enum { IB200_EXTRAPOLATE_NO = 0, IB200_EXTRAPOLATE_YES = 1, IB200_EXTRAPOLATE_HIST = 2 };

/* status codes */
enum {
    IB200_OK = 0,
    IB200_ERR_NULL = -1,          /* null pointer argument */
    IB200_ERR_DTYPE = -2,         /* unknown dtype code */
    IB200_ERR_DIM = -3,           /* dim not in {1,2,3} (reference: nd.py handles any dim on CPU) */
    IB200_ERR_BOUND = -4,         /* bound code not in 0..6 (autograd.py:95 ValueError) */
    IB200_ERR_ORDER = -5,         /* order not in 0..7 (autograd.py:145 ValueError) */
    IB200_ERR_SHAPE = -6,         /* non-positive extent / inconsistent shapes */
    IB200_ERR_BOUND_UNSUPPORTED = -7, /* prefilter with dst1/dst2 (coeff.py:237-254 NotImplementedError) */
    IB200_ERR_SCRATCH = -8,       /* scratch buffer required but missing */
    IB200_ERR_EXTRAPOLATE = -9,   /* extrapolate not in {0,1,2} */
    IB200_ERR_TOO_LARGE = -10,    /* a single volume exceeds 2^31-1 voxels */
    IB200_ERR_CUDA = -1000        /* IB200_ERR_CUDA - (int)cudaError_t */
};

/* storage type of the label maps of ib200_pull_labels (ib200_problem.reserved); `vol` and `out` share it */
enum {
    IB200_LABEL_I32 = 0,
    IB200_LABEL_I64 = 1,
    IB200_LABEL_U8 = 2,
    IB200_LABEL_I16 = 3
};

/* behaviour switches (ib200_problem.flags) */
enum {
    IB200_FLAG_NONE = 0,
    /* never take the shared-memory tiled fast paths (A/B testing, profiling) */
    IB200_FLAG_NO_TILES = 1u << 0,
    /* reproduce the reference's sign error for d/dx of the order-1 spline on an
     * axis of a mixed-order call (splines.py:96-97); default is the true derivative */
    IB200_FLAG_REF_LINEAR_GRAD_SIGN = 1u << 1,
    /* never take the persistent warp-specialised pull / grad, the boxed push / count or the
     * channel-interleaved tile kernels (fall back to the round-1 one-tile-per-CTA tiled kernels;
     * A/B testing, profiling) */
    IB200_FLAG_NO_PIPE = 1u << 2,
    /* take the persistent kernels even for problems too small to amortise their ramp-up (tests) */
    IB200_FLAG_FORCE_PIPE = 1u << 3,
    /* `grid` holds DISPLACEMENTS in voxels: the sampling coordinate of lattice point x is x + grid[x]
     * (what interpol.add_identity_grid builds, api.py:482-520, formed in registers instead: the identity
     * grid is never read or written).  Results equal add_identity_grid + the same call. */
    IB200_FLAG_DISPLACEMENT = 1u << 4
};

/*
 * Geometry + options shared by the six push/pull entry points.  It carries
 * what pushpull.py's functions receive as (bound: List[int],
 * interpolation: List[int], extrapolate: int) plus the tensor shapes/strides
 * that the reference reads off the torch tensors.
 */
typedef struct ib200_problem {
    int32_t dim;            /* number of spatial dimensions D: 1, 2 or 3 */
    int32_t dtype;          /* IB200_F16 / F32 / F64 / BF16, same for every tensor */
    int32_t extrapolate;    /* 0 / 1 / 2, nd.py:11-27 */
    int32_t device;         /* CUDA device ordinal that owns every pointer */
    int32_t bound[3];       /* per spatial axis, already padded (jit_utils.py:10-15) */
    int32_t order[3];       /* spline order 0..7 per spatial axis, already padded */
    uint32_t flags;         /* IB200_FLAG_* */
    uint32_t reserved;      /* ib200_pull_labels: storage type of the label maps (IB200_LABEL_*); 0 elsewhere */
    int64_t batch;          /* B = max over operands (operands with B=1 use stride 0) */
    int64_t channels;       /* C */
    int64_t vol_shape[3];   /* spatial shape of the volume that is gathered from
                               (pull/grad/hess) or scattered into (push/count/pushgrad) */
    int64_t pts_shape[3];   /* spatial shape of the grid lattice */
    int64_t vol_stride[5];  /* element strides B, C, X, Y, Z of an INPUT volume (pull/grad/hess) */
    int64_t grid_stride[5]; /* element strides B, X, Y, Z, D of the grid */
    int64_t img_stride[6];  /* element strides B, C, X, Y, Z, D of an INPUT image living on the
                               grid lattice (push / pushgrad); the D entry is used by pushgrad only */
} ib200_problem;

/* out (B, C, *pts_shape) = pull(vol, grid).  Replaces pushpull.grid_pull
 * (interpol/pushpull.py:35-66 -> nd.py:81-143 / iso1.py / iso0.py). */
IB200_API int ib200_pull(const ib200_problem *p, const void *vol, const void *grid,
               void *out, void *stream);

/* out (B, C, *pts_shape, D) = spatial gradient of the interpolated volume.
 * Replaces pushpull.grid_grad (interpol/pushpull.py:146-172 -> nd.py:217-288). */
IB200_API int ib200_grad(const ib200_problem *p, const void *vol, const void *grid,
               void *out, void *stream);

/* out (B, C, *pts_shape, D, D) = Hessian.  Replaces pushpull.grid_hess
 * (interpol/pushpull.py:207-233 -> nd.py:368-464); only used by GridGrad.backward. */
IB200_API int ib200_hess(const ib200_problem *p, const void *vol, const void *grid,
               void *out, void *stream);

/* vol_out (B, C, *vol_shape) = splat of img (B, C, *pts_shape) along grid; the
 * callee zero-fills vol_out.  Replaces pushpull.grid_push
 * (interpol/pushpull.py:70-102 -> nd.py:147-213).
 * `scratch`: required for F16/BF16 (float32 accumulation volume of
 * ib200_scratch_bytes() bytes), ignored (may be NULL) otherwise. */
IB200_API int ib200_push(const ib200_problem *p, const void *img, const void *grid,
               void *vol_out, void *scratch, void *stream);

/* vol_out (B, 1, *vol_shape) = splat of ones.  Replaces pushpull.grid_count
 * (interpol/pushpull.py:106-142); `channels` is ignored (treated as 1). */
IB200_API int ib200_count(const ib200_problem *p, const void *grid, void *vol_out,
                void *scratch, void *stream);

/* vol_out (B, C, *vol_shape) = adjoint of ib200_grad applied to
 * img (B, C, *pts_shape, D).  Replaces pushpull.grid_pushgrad
 * (interpol/pushpull.py:175-204 -> nd.py:292-364); only used by GridGrad.backward. */
IB200_API int ib200_pushgrad(const ib200_problem *p, const void *img, const void *grid,
                   void *vol_out, void *scratch, void *stream);

/* Fused backward of pull w.r.t. the grid: out (B, *pts_shape, D) =
 * sum_c grad(vol, grid)[b,c,...,:] * gout[b,c,...] without materialising the
 * (B, C, *pts, D) temporary of pushpull.grid_pull_backward
 * (interpol/pushpull.py:254-257).  `gout` uses img_stride (B, C, X, Y, Z). */
IB200_API int ib200_pull_backward_grid(const ib200_problem *p, const void *vol, const void *grid,
                             const void *gout, void *out, void *stream);

/* Fused backward of grad w.r.t. the grid: out (B, *pts_shape, D)[d] =
 * sum_c sum_e hess(vol, grid)[b,c,...,d,e] * gout[b,c,...,e] without materialising the
 * (B, C, *pts, D, D) Hessian of pushpull.grid_grad_backward
 * (interpol/pushpull.py:318-324).  `gout` uses img_stride (B, C, X, Y, Z, D). */
IB200_API int ib200_grad_backward_grid(const ib200_problem *p, const void *vol, const void *grid,
                             const void *gout, void *out, void *stream);

/* Number of bytes of scratch the scatter entry points need for this problem
 * (0 for F32/F64). */
IB200_API size_t ib200_scratch_bytes(const ib200_problem *p);

/* In-place spline prefilter along one axis of a dense tensor viewed as
 * (outer, n, inner).  Replaces coeff.spline_coeff (interpol/coeff.py:288-313
 * -> coeff.filter :258-284); orders 0/1 are no-ops, n == 1 is a no-op,
 * bounds zero->dct1 and replicate->dct2 alias like coeff.py:237-252. */
IB200_API int ib200_spline_coeff(void *data, int32_t dtype, int64_t outer, int64_t n,
                       int64_t inner, int32_t bound, int32_t order,
                       int32_t device, void *stream);

/* out (B, C, *pts_shape) = label map `vol` (B, C, *vol_shape) resampled at `grid` (integer storage type
 * p->reserved = IB200_LABEL_*, the same for both: torch's default int64 maps are read and written as they
 * are, no int32 round trip): for every
 * point the label whose soft mask (vol == label) interpolates to the largest value (> 0, ties to the
 * smallest label, 0 when nothing is in bounds).  One pass replaces the loop over `input.unique()` of
 * interpol.grid_pull (interpol/api.py:194-205: one full pull per label).  Orders 0 / 1 per axis
 * (IB200_ERR_ORDER otherwise: higher orders prefilter the masks); p->dtype is the dtype of the GRID
 * (F32 / F64), vol_stride counts label elements. */
IB200_API int ib200_pull_labels(const ib200_problem *p, const void *vol, const void *grid,
                      void *out, void *stream);

/* Separable resampling along one axis of a dense tensor viewed as (outer, n_in, inner):
 * out (outer, n_out, inner)[o, i, j] = sum_k w_k(coords[i]) * in[o, fold(start + k), j].
 * One call per axis replaces the dense-grid construction + grid_pull of interpol.resize
 * (interpol/resize.py:91-117; plugin seam interpol/jitfields.py:95): the sampling grid of a resize is
 * the tensor product of one coordinate vector per axis.  `coords` (n_out values, same dtype as the
 * data, voxel units of the input axis) are those vectors; `all_nearest` / `all_linear` tell whether
 * EVERY axis of the N-D call has order 0 / 1 (iso0.py:12 rounding, iso1.py closed forms). */
IB200_API int ib200_resample_axis(const void *in, void *out, const void *coords, int32_t dtype,
                        int64_t outer, int64_t n_in, int64_t n_out, int64_t inner,
                        int32_t order, int32_t bound, int32_t extrapolate,
                        int32_t all_nearest, int32_t all_linear, int32_t device, void *stream);

/* Adjoint of ib200_resample_axis: out (outer, n_out, inner)[o, fold(start(coords[i]) + k), j] +=
 * w_k(coords[i]) * in (outer, n_in, inner)[o, i, j]; `coords` has n_in values in voxel units of the
 * OUTPUT axis; the callee zero-fills `out`.  One call per axis replaces the dense-grid construction +
 * grid_push of interpol.restrict (interpol/restrict.py:86-120; plugin seam interpol/jitfields.py:106)
 * and is the backward of a resize pass.  F32 / F64 only (IB200_ERR_DTYPE otherwise). */
IB200_API int ib200_resample_axis_adjoint(const void *in, void *out, const void *coords, int32_t dtype,
                                int64_t outer, int64_t n_in, int64_t n_out, int64_t inner,
                                int32_t order, int32_t bound, int32_t extrapolate,
                                int32_t all_nearest, int32_t all_linear, int32_t device, void *stream);

/* Introspection */
IB200_API int ib200_abi_version(void);
IB200_API const char *ib200_error_string(int status);
/* name of the kernel variant the last call on this thread dispatched to */
IB200_API const char *ib200_last_kernel(void);
/* number of kernel launches issued by this library since load (all threads) */
IB200_API uint64_t ib200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* INTERPOL_B200_H */
