// Per-axis spline support: boundary index/sign maps, node weights and memory
// offsets, all in registers.  Semantics follow interpol/bounds.py:30-89
// (Bound.index / Bound.transform), interpol/nd.py:31-77 (get_weights),
// interpol/iso1.py:11-20, interpol/iso0.py:11-15 and interpol/nd.py:11-27
// (inbounds_mask); see SURVEY.md 8.3 for the closed forms.
#pragma once
#include "common.cuh"
#include "splines.cuh"

namespace ib200 {

// ------------------------------------------------------------ boundaries --
// I is int (coordinates known to fit) or long long (wild coordinates).

template <typename I>
__device__ __forceinline__ I pymod(I a, I n) {   // python-style remainder, n > 0
    I r = a % n;
    return r < 0 ? r + n : r;
}

template <typename I>
__device__ __forceinline__ I bound_index(int bound, I i, I n) {
    switch (bound) {
    case IB200_BOUND_ZERO:
    case IB200_BOUND_REPLICATE:                       // bounds.py:31-32
        return i < 0 ? I(0) : (i > n - 1 ? n - 1 : i);
    case IB200_BOUND_DCT2:
    case IB200_BOUND_DST2: {                          // bounds.py:33-38
        const I n2 = 2 * n;
        if (i < 0) i = -i - 1;                        // mirror about -1/2
        if (i >= n2) i = i % n2;
        return i >= n ? n2 - 1 - i : i;
    }
    case IB200_BOUND_DCT1: {                          // bounds.py:39-46
        if (n == 1) return I(0);
        const I n2 = 2 * (n - 1);
        if (i < 0) i = -i;
        if (i >= n2) i = i % n2;
        return i >= n ? n2 - i : i;
    }
    case IB200_BOUND_DST1: {                          // bounds.py:47-56
        const I n2 = 2 * (n + 1);
        if (i < 0) i = -i - 2;
        if (i >= n2 || i < 0) i = pymod(i, n2);       // i == -1 -> n2 - 1
        if (i > n) i = n2 - 2 - i;
        if (i == -1) i = 0;
        if (i == n) i = n - 1;
        return i;
    }
    case IB200_BOUND_DFT:                             // bounds.py:57-58
        if (i >= 0 && i < n) return i;
        return pymod(i, n);
    }
    return i;
}

// -1 / 0 / +1 (reference None == +1)
template <typename I>
__device__ __forceinline__ int bound_sign(int bound, I i, I n) {
    switch (bound) {
    case IB200_BOUND_ZERO:                            // bounds.py:82-87
        return (i < 0 || i >= n) ? 0 : 1;
    case IB200_BOUND_DST2: {                          // bounds.py:76-81
        if (i < 0) i = n - 1 - i;
        return ((i / n) & 1) ? -1 : 1;
    }
    case IB200_BOUND_DST1: {                          // bounds.py:63-75 (Q1: 0 at i == 0 mod 2(n+1))
        if (n == 1) return 1;
        const I n2 = 2 * (n + 1);
        if (i < 0) i = -i + (n - 1);
        if (i >= n2) i = i % n2;
        int x = (i == 0) ? 0 : 1;
        if (i % (n + 1) == n) x = 0;
        return ((i / (n + 1)) & 1) ? -x : x;
    }
    }
    return 1;
}

// ------------------------------------------------------------------ axis --
// NODES = ORDER + 1 for a compile-time order, 8 for a runtime order.
template <typename R, int NODES>
struct Axis {
    R w[NODES];      // weight * sign
    R g[NODES];      // first derivative * sign   (filled when NEED >= 1)
    R h[NODES];      // second derivative * sign  (filled when NEED >= 2)
    int off[NODES];  // index * stride (elements)
    int n;           // number of nodes actually used
};

template <typename R> __device__ __forceinline__ R rint_(R x);
template <> __device__ __forceinline__ float rint_<float>(float x) { return rintf(x); }
template <> __device__ __forceinline__ double rint_<double>(double x) { return rint(x); }

// Fill one axis.  Returns false when the coordinate is not finite (treated as
// out of bounds: contributes nothing; reference behaviour is undefined, Q9).
// ORDER < 0: runtime order `order_rt`.
template <typename R, int ORDER, int NEED, int NODES>
__device__ __forceinline__ bool setup_axis(Axis<R, NODES> &ax, R coord, int order_rt, int bound,
                                           int n, int stride, const KParams &kp) {
    const int order = ORDER >= 0 ? ORDER : order_rt;
    ax.n = order + 1;
    // nd.py:45 / iso1.py:13 / iso0.py:12
    R g0 = kp.round_nearest ? rint_<R>(coord) : floor(coord - R(0.5) * R(order - 1));
    const R t = coord - g0;                                   // nd.py:46
    if (!(fabs(g0) < R(4e18))) {                              // NaN / inf / absurd
#pragma unroll
        for (int k = 0; k < NODES; ++k) { ax.w[k] = R(0); ax.off[k] = 0; if (NEED >= 1) ax.g[k] = R(0); if (NEED >= 2) ax.h[k] = R(0); }
        return false;
    }
    const bool small = fabs(g0) < R(1e9);
    const int i0 = small ? (int)g0 : 0;
    const int lo = (bound == IB200_BOUND_DST1) ? 1 : 0;       // dst1 zeroes voxel 0 (Q1)
    const bool interior = small && i0 >= lo && i0 + order <= n - 1;
#pragma unroll
    for (int k = 0; k < NODES; ++k) {
        if (ORDER < 0 && k > order) break;
        int idx, sgn;
        if (interior) {
            idx = i0 + k; sgn = 1;
        } else if (small) {
            idx = bound_index<int>(bound, i0 + k, n);
            sgn = bound_sign<int>(bound, i0 + k, n);
        } else {
            const i64 ik = (i64)g0 + k;
            idx = (int)bound_index<i64>(bound, ik, (i64)n);
            sgn = bound_sign<i64>(bound, ik, (i64)n);
        }
        ax.off[k] = idx * stride;
        const R x = t - R(k);                                 // nd.py:60
        const R s = R(sgn);
        R w, g = R(0), h = R(0);
        if (kp.round_nearest) {
            w = R(1);
        } else if (order == 1 && !(kp.flags & IB200_FLAG_REF_LINEAR_GRAD_SIGN && !kp.all_linear)) {
            // iso1.py:19 + the closed forms of iso1.grad*: d/dg (1-t) = -1, d/dg t = +1
            w = kp.all_linear ? (k == 0 ? R(1) - t : t) : R(1) - fabs(x);
            g = (k == 0) ? R(-1) : R(1);
        } else if (order == 1) {
            // reference ND path on a linear axis of a mixed-order call (splines.py:96-97)
            w = R(1) - fabs(x);
            g = x > R(0) ? R(1) : (x < R(0) ? R(-1) : R(0));
        } else {
            if (sizeof(R) == 4 && order >= 6) {
                // degree-6/7 Horner forms cancel ~2 digits in float32 (the reference's own
                // float32 noise at order 7 is 1-2e-5): evaluate in double
                w = (R)spline_weight<double>(order, (double)x);
                if (NEED >= 1) g = (R)spline_grad<double>(order, (double)x);
                if (NEED >= 2) h = (R)spline_hess<double>(order, (double)x);
            } else {
                w = spline_weight<R>(order, x);
                if (NEED >= 1) g = spline_grad<R>(order, x);
                if (NEED >= 2) h = spline_hess<R>(order, x);
            }
        }
        ax.w[k] = w * s;
        if (NEED >= 1) ax.g[k] = g * s;
        if (NEED >= 2) ax.h[k] = h * s;
    }
    return true;
}

// unused axis (dim < 3): a single node of weight one
template <typename R, int NODES>
__device__ __forceinline__ void unit_axis(Axis<R, NODES> &ax) {
    ax.n = 1;
    ax.w[0] = R(1); ax.g[0] = R(0); ax.h[0] = R(0); ax.off[0] = 0;
}

// nd.py:11-27 / jit_utils.py:242-285.  extrapolate==1: always true.
template <typename R> struct Thr;
template <> struct Thr<float> {
    static __device__ __forceinline__ float lo(const KParams &kp, int d) { return kp.thr_lo[d]; }
    static __device__ __forceinline__ float hi(const KParams &kp, int d) { return kp.thr_hi[d]; }
};
template <> struct Thr<double> {
    static __device__ __forceinline__ double lo(const KParams &kp, int d) { return kp.thr_lo_d[d]; }
    static __device__ __forceinline__ double hi(const KParams &kp, int d) { return kp.thr_hi_d[d]; }
};

template <typename R, int DIM>
__device__ __forceinline__ bool inbounds(const KParams &kp, const R *coord) {
    if (kp.extrapolate == 1) return true;
    bool ok = true;
#pragma unroll
    for (int d = 0; d < DIM; ++d)
        ok = ok && (coord[d] > Thr<R>::lo(kp, d)) && (coord[d] < Thr<R>::hi(kp, d));
    return ok;
}

}  // namespace ib200
