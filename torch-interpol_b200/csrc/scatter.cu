// Scatter family: push / count / pushgrad -- the adjoints of pull and grad.
//
// One thread owns one lattice point and adds its (order+1)^D weighted
// contributions into the target volume with global reductions (RED.ADD);
// lanes of a warp own consecutive points along the fastest axis so that, for
// coherent deformations, one warp-wide RED touches one or two cache lines.
// 16-bit storage types accumulate in a caller-provided float32 scratch volume
// that is converted once at the end (more accurate than the reference's
// half-precision scatter_add_).
//
// Replaces interpol/nd.py:147-213 (push), :292-364 (pushgrad), the iso0/iso1
// variants and pushpull.grid_count's expanded ones (interpol/pushpull.py:106-142).
#include <cstdio>
#include "support.cuh"

#ifndef IB200_T
#error "compile with -DIB200_T=<storage type> -DIB200_TNAME=<f32|f64|f16|bf16> -DIB200_A=<accumulator type>"
#endif

namespace ib200 {

template <typename A> __device__ __forceinline__ void red_add(A *p, A v) { atomicAdd(p, v); }

// T: storage type of img/grid; A: accumulation type of the target volume
template <typename T, typename A, int DIM, int ORDER, int OP>
__global__ void __launch_bounds__(256)
scatter_kernel(const __grid_constant__ KParams kp, const T *__restrict__ img,
               const T *__restrict__ grid, A *__restrict__ out) {
    typedef typename Traits<T>::Real R;
    constexpr int NODES = ORDER >= 0 ? ORDER + 1 : 8;
    constexpr int NEED = (OP == OP_PUSHGRAD) ? 1 : 0;
    constexpr int UNR = ORDER >= 0 ? NODES : 1;   // never unroll the runtime-order loops

    const i64 total = kp.batch * kp.pts_total;
    for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < total;
         p += (i64)gridDim.x * blockDim.x) {
        i64 b = p / kp.pts_total;
        i64 goff, ioff;
        const bool disp = (kp.flags & IB200_FLAG_DISPLACEMENT) != 0;
        int xyz[3] = {0, 0, 0};
        if (!kp.pts_dense || disp) {
            i64 r = p - (p / kp.pts_total) * kp.pts_total;
            if (DIM == 3) {
                const i64 yz = (i64)kp.pts_n[1] * kp.pts_n[2];
                xyz[0] = (int)(r / yz); r -= (i64)xyz[0] * yz;
                xyz[1] = (int)(r / kp.pts_n[2]); xyz[2] = (int)(r - (i64)xyz[1] * kp.pts_n[2]);
            } else if (DIM == 2) {
                xyz[0] = (int)(r / kp.pts_n[1]); xyz[1] = (int)(r - (i64)xyz[0] * kp.pts_n[1]);
            } else {
                xyz[0] = (int)r;
            }
        }
        if (kp.pts_dense) {
            b = p / kp.pts_total;
            const i64 r = p - b * kp.pts_total;
            goff = b * kp.grid_sb + r * DIM;
            ioff = r * (OP == OP_PUSHGRAD ? DIM : 1);
        } else {
            b = p / kp.pts_total;
            goff = b * kp.grid_sb; ioff = 0;
#pragma unroll
            for (int d = 0; d < DIM; ++d) { goff += xyz[d] * kp.grid_s[d]; ioff += xyz[d] * kp.img_s[d]; }
        }

        R coord[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) coord[d] = Traits<T>::load(grid + goff + d * kp.grid_sd) + (disp ? (R)xyz[d] : R(0));

        bool ok = inbounds<R, DIM>(kp, coord);      // nd.py:201-203: masked sources add nothing
        Axis<R, NODES> ax[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (d < DIM) ok = setup_axis<R, ORDER, NEED, NODES>(ax[d], coord[d], kp.order[d], kp.bound[d], kp.vol_n[d], (int)kp.vol_s[d], kp) && ok;
            else unit_axis(ax[d]);
        }
        if (!ok) continue;

        for (i64 c = 0; c < kp.channels; ++c) {
            A *dst = out + (b * kp.channels + c) * kp.vol_total;
            R val = R(1);
            R vg[DIM];
            if (OP == OP_PUSH) {
                val = Traits<T>::load(img + b * kp.img_sb + c * kp.img_sc + ioff);
            } else if (OP == OP_PUSHGRAD) {
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    vg[d] = Traits<T>::load(img + b * kp.img_sb + c * kp.img_sc + ioff + d * kp.img_sd);
            }
#pragma unroll UNR
            for (int i = 0; i < NODES; ++i) {
                if (ORDER < 0 && i >= ax[0].n) break;
#pragma unroll UNR
                for (int j = 0; j < (DIM >= 2 ? NODES : 1); ++j) {
                    if (ORDER < 0 && j >= ax[1].n) break;
                    // coefficients of this (x, y) row: value = a * w_z + bz * g_z
                    R a, bz = R(0);
                    if (OP == OP_PUSHGRAD) {
                        a = vg[0] * ax[0].g[i] * ax[1].w[j];
                        if (DIM >= 2) a = fma(vg[1 % DIM] * ax[0].w[i], ax[1].g[j], a);
                        if (DIM >= 3) bz = vg[2 % DIM] * ax[0].w[i] * ax[1].w[j];
                    } else {
                        a = val * ax[0].w[i] * ax[1].w[j];
                    }
                    const int row = ax[0].off[i] + ax[1].off[j];
#pragma unroll UNR
                    for (int k = 0; k < (DIM >= 3 ? NODES : 1); ++k) {
                        if (ORDER < 0 && k >= ax[2].n) break;
                        R contrib = a * ax[2].w[k];
                        if (OP == OP_PUSHGRAD && DIM >= 3) contrib = fma(bz, ax[2].g[k], contrib);
                        red_add<A>(dst + row + ax[2].off[k], (A)contrib);
                    }
                }
            }
        }
    }
}

// float32 scratch -> 16-bit output
template <typename T, typename A>
__global__ void __launch_bounds__(256)
convert_kernel(const A *__restrict__ src, T *__restrict__ dst, i64 n) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        Traits<T>::store(dst + i, (typename Traits<T>::Real)src[i]);
}

// ---------------------------------------------------------------- launch --

static const char *kOpName[3] = {"push", "count", "pushgrad"};

template <typename T, typename A, int DIM, int ORDER, int OP>
static int launch_one(const KParams &kp, const void *img, const void *grid, void *out,
                      cudaStream_t stream) {
    const i64 total = kp.batch * kp.pts_total;
    if (total == 0) return IB200_OK;
    const int threads = 256;
    i64 blocks = (total + threads - 1) / threads;
    const i64 cap = (i64)kNumSMs * 64;
    if (blocks > cap) blocks = cap;
    scatter_kernel<T, A, DIM, ORDER, OP><<<(unsigned)blocks, threads, 0, stream>>>(
        kp, (const T *)img, (const T *)grid, (A *)out);
    static thread_local char name[64];
    snprintf(name, sizeof(name), "scatter_%s_%dd_o%d", kOpName[OP], DIM, ORDER);
    note_launch(name);
    IB200_CUDA_CHECK(cudaGetLastError());
    return IB200_OK;
}

template <typename T, typename A, int DIM, int OP>
static int dispatch_order(const KParams &kp, const void *img, const void *grid, void *out,
                          cudaStream_t stream) {
    bool iso = true;
    for (int d = 1; d < DIM; ++d) iso = iso && kp.order[d] == kp.order[0];
    if (iso && OP != OP_PUSHGRAD) {
        switch (kp.order[0]) {
#define IB200_CASE(O) case O: return launch_one<T, A, DIM, O, OP>(kp, img, grid, out, stream);
        IB200_STATIC_ORDERS(IB200_CASE)
#undef IB200_CASE
        default: break;
        }
    }
    return launch_one<T, A, DIM, -1, OP>(kp, img, grid, out, stream);
}

template <typename T, typename A, int OP>
static int dispatch_dim(const KParams &kp, const void *img, const void *grid, void *out,
                        cudaStream_t stream) {
    switch (kp.dim) {
    case 1: return dispatch_order<T, A, 1, OP>(kp, img, grid, out, stream);
    case 2: return dispatch_order<T, A, 2, OP>(kp, img, grid, out, stream);
    case 3: return dispatch_order<T, A, 3, OP>(kp, img, grid, out, stream);
    }
    return IB200_ERR_DIM;
}

template <typename T, typename A>
static int dispatch_op(int op, const KParams &kp, const void *img, const void *grid, void *out,
                       cudaStream_t stream) {
    switch (op) {
    case OP_PUSH: return dispatch_dim<T, A, OP_PUSH>(kp, img, grid, out, stream);
    case OP_COUNT: return dispatch_dim<T, A, OP_COUNT>(kp, img, grid, out, stream);
    case OP_PUSHGRAD: return dispatch_dim<T, A, OP_PUSHGRAD>(kp, img, grid, out, stream);
    }
    return IB200_ERR_NULL;
}

#define IB200_CAT_(a, b) a##b
#define IB200_CAT(a, b) IB200_CAT_(a, b)
int IB200_CAT(launch_scatter_, IB200_TNAME)(int op, const KParams &kp, const void *img, const void *grid,
                                            void *out, void *scratch, cudaStream_t stream) {
    typedef IB200_T T;
    typedef IB200_A A;
    const i64 n = kp.batch * kp.channels * kp.vol_total;
    if (sizeof(T) == sizeof(A)) {   // f32 / f64: accumulate straight into the output
        IB200_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)n * sizeof(A), stream));
        return dispatch_op<T, A>(op, kp, img, grid, out, stream);
    }
    // 16-bit storage: float32 scratch volume, converted once
    if (!scratch) return IB200_ERR_SCRATCH;
    IB200_CUDA_CHECK(cudaMemsetAsync(scratch, 0, (size_t)n * sizeof(A), stream));
    int st = dispatch_op<T, A>(op, kp, img, grid, scratch, stream);
    if (st != IB200_OK || n == 0) return st;
    i64 blocks = (n + 255) / 256;
    if (blocks > (i64)kNumSMs * 32) blocks = (i64)kNumSMs * 32;
    convert_kernel<T, A><<<(unsigned)blocks, 256, 0, stream>>>((const A *)scratch, (T *)out, n);
    note_launch("convert_f32_to_16bit");
    IB200_CUDA_CHECK(cudaGetLastError());
    return IB200_OK;
}

}  // namespace ib200
