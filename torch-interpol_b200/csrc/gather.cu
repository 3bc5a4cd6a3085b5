// Gather family: pull / grad / hess / fused d(pull)/d(grid).
//
// One thread owns one lattice point (b, x, y, z): it reads its D coordinates,
// evaluates the boundary maps and spline weights of the (order+1)^D support in
// registers, and loops over channels re-using them.  Lanes of a warp own
// consecutive points along the fastest axis so grid reads, output writes and
// (for coherent deformations) the taps themselves are coalesced.
//
// Replaces interpol/nd.py:81-143 (pull), :217-288 (grad), :368-464 (hess) and
// the iso0/iso1 specialisations behind interpol/pushpull.py:35-66,146-172,207-233.
#include <cstdio>
#include "support.cuh"

#ifndef IB200_T
#error "compile with -DIB200_T=<storage type> -DIB200_TNAME=<f32|f64|f16|bf16>"
#endif

namespace ib200 {

template <int DIM>
__device__ __forceinline__ void decompose(const KParams &kp, i64 p, i64 &b, int (&xyz)[3]) {
    b = p / kp.pts_total;
    i64 r = p - b * kp.pts_total;
    xyz[0] = xyz[1] = xyz[2] = 0;
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

// OP: OP_PULL / OP_GRAD / OP_HESS / OP_PULL_BWD_GRID / OP_GRAD_BWD_GRID
template <typename T, int DIM, int ORDER, int OP>
__global__ void __launch_bounds__(256)
gather_kernel(const __grid_constant__ KParams kp, const T *__restrict__ vol,
              const T *__restrict__ grid, const T *__restrict__ gout, T *__restrict__ out) {
    typedef typename Traits<T>::Real R;
    constexpr int NODES = ORDER >= 0 ? ORDER + 1 : 8;
    constexpr bool HESS = (OP == OP_HESS || OP == OP_GRAD_BWD_GRID);   // second derivatives needed
    constexpr int NEED = (OP == OP_PULL) ? 0 : (HESS ? 2 : 1);
    constexpr int NH = DIM * (DIM + 1) / 2;
    constexpr int UNR = ORDER >= 0 ? NODES : 1;   // never unroll the runtime-order loops

    const i64 total = kp.batch * kp.pts_total;
    for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < total;
         p += (i64)gridDim.x * blockDim.x) {
        i64 b; int xyz[3];
        i64 goff, ioff;   // offsets of this point in the grid / in a lattice image
        const bool disp = (kp.flags & IB200_FLAG_DISPLACEMENT) != 0;
        if (kp.pts_dense) {
            b = p / kp.pts_total;
            const i64 r = p - b * kp.pts_total;
            goff = b * kp.grid_sb + r * DIM;
            ioff = r;
            if (disp) decompose<DIM>(kp, p, b, xyz);
        } else {
            decompose<DIM>(kp, p, b, xyz);
            goff = b * kp.grid_sb; ioff = 0;
#pragma unroll
            for (int d = 0; d < DIM; ++d) { goff += xyz[d] * kp.grid_s[d]; ioff += xyz[d] * kp.img_s[d]; }
        }
        const i64 r_dense = p - b * kp.pts_total;   // position in the dense output lattice

        R coord[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) coord[d] = Traits<T>::load(grid + goff + d * kp.grid_sd) + (disp ? (R)xyz[d] : R(0));

        bool ok = inbounds<R, DIM>(kp, coord);
        Axis<R, NODES> ax[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (d < DIM) ok = setup_axis<R, ORDER, NEED, NODES>(ax[d], coord[d], kp.order[d], kp.bound[d], kp.vol_n[d], (int)kp.vol_s[d], kp) && ok;
            else unit_axis(ax[d]);
        }

        R bwd[DIM];   // OP_PULL_BWD_GRID / OP_GRAD_BWD_GRID accumulate over channels
#pragma unroll
        for (int d = 0; d < DIM; ++d) bwd[d] = R(0);

        for (i64 c = 0; c < kp.channels; ++c) {
            const T *src = vol + b * kp.vol_sb + c * kp.vol_sc;
            R acc0 = R(0);          // pull
            R accg[DIM];            // grad
            R acch[NH];             // hess, upper triangle row-major
#pragma unroll
            for (int d = 0; d < DIM; ++d) accg[d] = R(0);
#pragma unroll
            for (int q = 0; q < NH; ++q) acch[q] = R(0);

            if (ok) {
#pragma unroll UNR
                for (int i = 0; i < NODES; ++i) {
                    if (ORDER < 0 && i >= ax[0].n) break;
                    // partial sums over the (y, z) plane of this x node
                    R s00 = R(0), s10 = R(0), s01 = R(0), s20 = R(0), s11 = R(0), s02 = R(0);
#pragma unroll UNR
                    for (int j = 0; j < (DIM >= 2 ? NODES : 1); ++j) {
                        if (ORDER < 0 && j >= ax[1].n) break;
                        R t0 = R(0), t1 = R(0), t2 = R(0);   // sums over z with w, g, h
#pragma unroll UNR
                        for (int k = 0; k < (DIM >= 3 ? NODES : 1); ++k) {
                            if (ORDER < 0 && k >= ax[2].n) break;
                            const R v = Traits<T>::load(src + ax[0].off[i] + ax[1].off[j] + ax[2].off[k]);
                            t0 = fma(ax[2].w[k], v, t0);
                            if (NEED >= 1 && DIM >= 3) t1 = fma(ax[2].g[k], v, t1);
                            if (NEED >= 2 && DIM >= 3) t2 = fma(ax[2].h[k], v, t2);
                        }
                        s00 = fma(ax[1].w[j], t0, s00);
                        if (NEED >= 1 && DIM >= 2) s10 = fma(ax[1].g[j], t0, s10);   // d/dy
                        if (NEED >= 1 && DIM >= 3) s01 = fma(ax[1].w[j], t1, s01);   // d/dz
                        if (NEED >= 2 && DIM >= 2) s20 = fma(ax[1].h[j], t0, s20);   // d2/dy2
                        if (NEED >= 2 && DIM >= 3) s11 = fma(ax[1].g[j], t1, s11);   // d2/dydz
                        if (NEED >= 2 && DIM >= 3) s02 = fma(ax[1].w[j], t2, s02);   // d2/dz2
                    }
                    const R wx = ax[0].w[i];
                    if (OP == OP_PULL) acc0 = fma(wx, s00, acc0);
                    if (NEED >= 1) {
                        const R gx = ax[0].g[i];
                        if (!HESS) {
                            accg[0] = fma(gx, s00, accg[0]);
                            if (DIM >= 2) accg[1 % DIM] = fma(wx, s10, accg[1 % DIM]);
                            if (DIM >= 3) accg[2 % DIM] = fma(wx, s01, accg[2 % DIM]);
                        } else {
                            const R hx = ax[0].h[i];
                            // upper triangle, row-major: 1D [xx]; 2D [xx xy yy]; 3D [xx xy xz yy yz zz]
                            acch[0] = fma(hx, s00, acch[0]);
                            if (DIM == 2) {
                                acch[1 % NH] = fma(gx, s10, acch[1 % NH]);
                                acch[2 % NH] = fma(wx, s20, acch[2 % NH]);
                            }
                            if (DIM == 3) {
                                acch[1 % NH] = fma(gx, s10, acch[1 % NH]);
                                acch[2 % NH] = fma(gx, s01, acch[2 % NH]);
                                acch[3 % NH] = fma(wx, s20, acch[3 % NH]);
                                acch[4 % NH] = fma(wx, s11, acch[4 % NH]);
                                acch[5 % NH] = fma(wx, s02, acch[5 % NH]);
                            }
                        }
                    }
                }
            }

            if (OP == OP_PULL) {
                Traits<T>::store(out + (b * kp.channels + c) * kp.pts_total + r_dense, acc0);
            } else if (OP == OP_GRAD) {
                T *o = out + ((b * kp.channels + c) * kp.pts_total + r_dense) * DIM;
#pragma unroll
                for (int d = 0; d < DIM; ++d) Traits<T>::store(o + d, accg[d]);
            } else if (OP == OP_HESS) {
                T *o = out + ((b * kp.channels + c) * kp.pts_total + r_dense) * (DIM * DIM);
                if (DIM == 1) {
                    Traits<T>::store(o, acch[0]);
                } else if (DIM == 2) {
                    Traits<T>::store(o + 0, acch[0]); Traits<T>::store(o + 1, acch[1 % NH]);
                    Traits<T>::store(o + 2, acch[1 % NH]); Traits<T>::store(o + 3, acch[2 % NH]);
                } else {
                    Traits<T>::store(o + 0, acch[0]);      Traits<T>::store(o + 1, acch[1 % NH]); Traits<T>::store(o + 2, acch[2 % NH]);
                    Traits<T>::store(o + 3, acch[1 % NH]); Traits<T>::store(o + 4, acch[3 % NH]); Traits<T>::store(o + 5, acch[4 % NH]);
                    Traits<T>::store(o + 6, acch[2 % NH]); Traits<T>::store(o + 7, acch[4 % NH]); Traits<T>::store(o + 8, acch[5 % NH]);
                }
            } else if (OP == OP_PULL_BWD_GRID) {   // sum_c grad * gout   (pushpull.py:257)
                const R go = Traits<T>::load(gout + b * kp.img_sb + c * kp.img_sc + ioff);
#pragma unroll
                for (int d = 0; d < DIM; ++d) bwd[d] = fma(accg[d], go, bwd[d]);
            } else {   // OP_GRAD_BWD_GRID: sum_c hess . gout  (pushpull.py:321-324), gout (B, C, *pts, D)
                const T *gp = gout + b * kp.img_sb + c * kp.img_sc + (kp.pts_dense ? ioff * DIM : ioff);
                R go[DIM];
#pragma unroll
                for (int e = 0; e < DIM; ++e) go[e] = Traits<T>::load(gp + e * kp.img_sd);
                if (DIM == 1) {
                    bwd[0] = fma(acch[0], go[0], bwd[0]);
                } else if (DIM == 2) {
                    bwd[0] = fma(acch[0], go[0], fma(acch[1 % NH], go[1 % DIM], bwd[0]));
                    bwd[1 % DIM] = fma(acch[1 % NH], go[0], fma(acch[2 % NH], go[1 % DIM], bwd[1 % DIM]));
                } else {
                    bwd[0] = fma(acch[0], go[0], fma(acch[1 % NH], go[1 % DIM], fma(acch[2 % NH], go[2 % DIM], bwd[0])));
                    bwd[1 % DIM] = fma(acch[1 % NH], go[0], fma(acch[3 % NH], go[1 % DIM], fma(acch[4 % NH], go[2 % DIM], bwd[1 % DIM])));
                    bwd[2 % DIM] = fma(acch[2 % NH], go[0], fma(acch[4 % NH], go[1 % DIM], fma(acch[5 % NH], go[2 % DIM], bwd[2 % DIM])));
                }
            }
        }
        if (OP == OP_PULL_BWD_GRID || OP == OP_GRAD_BWD_GRID) {
            T *o = out + (b * kp.pts_total + r_dense) * DIM;
#pragma unroll
            for (int d = 0; d < DIM; ++d) Traits<T>::store(o + d, bwd[d]);
        }
    }
}

// ---------------------------------------------------------------- launch --

static const char *kOpName[5] = {"pull", "grad", "hess", "pull_bwd_grid", "grad_bwd_grid"};

template <typename T, int DIM, int ORDER, int OP>
static int launch_one(const KParams &kp, const void *vol, const void *grid, const void *gout,
                      void *out, cudaStream_t stream) {
    const i64 total = kp.batch * kp.pts_total;
    if (total == 0) return IB200_OK;
    const int threads = 256;
    i64 blocks = (total + threads - 1) / threads;
    const i64 cap = (i64)kNumSMs * 64;            // grid-stride beyond 64 CTAs per SM
    if (blocks > cap) blocks = cap;
    gather_kernel<T, DIM, ORDER, OP><<<(unsigned)blocks, threads, 0, stream>>>(
        kp, (const T *)vol, (const T *)grid, (const T *)gout, (T *)out);
    static thread_local char name[64];
    snprintf(name, sizeof(name), "gather_%s_%dd_o%d", kOpName[OP], DIM, ORDER);
    note_launch(name);
    IB200_CUDA_CHECK(cudaGetLastError());
    return IB200_OK;
}

template <typename T, int DIM, int OP>
static int dispatch_order(const KParams &kp, const void *vol, const void *grid, const void *gout,
                          void *out, cudaStream_t stream) {
    bool iso = true;
    for (int d = 1; d < DIM; ++d) iso = iso && kp.order[d] == kp.order[0];
    // the order-1 closed forms depend on kp flags only; compile-time orders are safe
    if (iso && OP != OP_HESS && OP != OP_GRAD_BWD_GRID) {
        switch (kp.order[0]) {
#define IB200_CASE(O) case O: return launch_one<T, DIM, O, OP>(kp, vol, grid, gout, out, stream);
        IB200_STATIC_ORDERS(IB200_CASE)
#undef IB200_CASE
        default: break;
        }
    }
    return launch_one<T, DIM, -1, OP>(kp, vol, grid, gout, out, stream);
}

template <typename T, int OP>
static int dispatch_dim(const KParams &kp, const void *vol, const void *grid, const void *gout,
                        void *out, cudaStream_t stream) {
    switch (kp.dim) {
    case 1: return dispatch_order<T, 1, OP>(kp, vol, grid, gout, out, stream);
    case 2: return dispatch_order<T, 2, OP>(kp, vol, grid, gout, out, stream);
    case 3: return dispatch_order<T, 3, OP>(kp, vol, grid, gout, out, stream);
    }
    return IB200_ERR_DIM;
}

template <typename T>
static int dispatch_op(int op, const KParams &kp, const void *vol, const void *grid,
                       const void *gout, void *out, cudaStream_t stream) {
    switch (op) {
    case OP_PULL: return dispatch_dim<T, OP_PULL>(kp, vol, grid, gout, out, stream);
    case OP_GRAD: return dispatch_dim<T, OP_GRAD>(kp, vol, grid, gout, out, stream);
    case OP_HESS: return dispatch_dim<T, OP_HESS>(kp, vol, grid, gout, out, stream);
    case OP_PULL_BWD_GRID: return dispatch_dim<T, OP_PULL_BWD_GRID>(kp, vol, grid, gout, out, stream);
    case OP_GRAD_BWD_GRID: return dispatch_dim<T, OP_GRAD_BWD_GRID>(kp, vol, grid, gout, out, stream);
    }
    return IB200_ERR_NULL;
}

#define IB200_CAT_(a, b) a##b
#define IB200_CAT(a, b) IB200_CAT_(a, b)
int IB200_CAT(launch_gather_, IB200_TNAME)(int op, const KParams &kp, const void *vol, const void *grid,
                                           const void *gout, void *out, cudaStream_t stream) {
    return dispatch_op<IB200_T>(op, kp, vol, grid, gout, out, stream);
}

}  // namespace ib200
