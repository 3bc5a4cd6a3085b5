// Tiled pull / grad (3-D, isotropic compile-time order, 16/32-bit storage).
//
// A CTA owns a TX x TY x TZ block of output voxels:
//   1. their grid coordinates are staged in shared memory with cp.async
//      (16 B per request, coalesced);
//   2. the bounding box of all spline supports is reduced per x-plane of the
//      block, and the block is split into 1, 2, 4 ... groups of planes until
//      every group's box fits in shared memory (steeper deformations simply
//      use more, smaller boxes -- there is no cliff);
//   3. each box of the input volume is staged in shared memory: straight 16-byte
//      cp.async copies when the box lies inside the volume, otherwise through
//      per-axis index/sign tables so that every boundary condition (fold +
//      sign) is applied while staging and costs nothing in the tap loop;
//   4. every thread evaluates its voxels with (ORDER+1)^3 LDS taps.  Rows of
//      the box are padded to a multiple of 32 words so the bank of a tap
//      depends only on its z coordinate: lanes on different (x, y) rows never
//      conflict, the only conflicts left are runs longer than 32 words where
//      the deformation locally expands along z.
// Groups whose box cannot fit even for a single plane (incoherent deformation)
// fall back, per group, to direct global gathers with identical arithmetic.
//
// Replaces interpol/nd.py:81-143 / :217-288 (and iso1.py) for the shapes that
// matter for throughput; the generic kernels in gather.cu cover the rest.
#include <cstdio>
#include <cstdlib>
#include "tile_common.cuh"

namespace ib200 {

template <typename T, int ORDER, int OP, int TX, int TY, int TZ, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
pull_tile3d_kernel(const __grid_constant__ KParams kp, const T *__restrict__ vol,
                   const T *__restrict__ grid, const T *__restrict__ gout, T *__restrict__ out, const int cap, const int vec_ok) {
    constexpr int NPT = TX * TY * TZ;            // points per tile
    constexpr int W = ORDER + 1;
    constexpr int NW = NT / 32;
    constexpr bool F32 = sizeof(T) == 4;
    constexpr bool GRAD = (OP == OP_GRAD || OP == OP_PULL_BWD_GRID);
    // fused backward w.r.t. the grid (pushpull.py:254-257): out (B, N, 3) = sum_c grad_c * gout_c; a thread owns the
    // same voxels for every channel, so the sum is a plain read-modify-write of its own output
    constexpr bool BWD = (OP == OP_PULL_BWD_GRID);
    constexpr int UI = ORDER <= 3 ? W : 1, UJ = ORDER <= 5 ? W : 1;   // keep the code of high orders compact
    static_assert(NT == TY * TZ, "one thread per (y, z) column of the tile; x-planes are looped");
    static_assert(TZ % 4 == 0, "z rows are staged 16 bytes at a time");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *tile = reinterpret_cast<float *>(smem_raw);                  // [cap] input box
    T *gtile = reinterpret_cast<T *>(tile + cap);                       // [NPT * 3] grid coordinates
    int *idx_tab = reinterpret_cast<int *>(reinterpret_cast<float *>(gtile) + (NPT * 3 * sizeof(T)) / 4);
    float *sgn_tab = reinterpret_cast<float *>(idx_tab + 3 * kMaxExt);
    int *red = reinterpret_cast<int *>(sgn_tab + 3 * kMaxExt);          // [TX][NW][6]
    PlaneBox *pb = reinterpret_cast<PlaneBox *>(red + TX * NW * 8);     // [TX]
    TileGeom *geoms = reinterpret_cast<TileGeom *>(pb + TX);            // [TX]
    int *nsub_p = reinterpret_cast<int *>(geoms + TX);
    // ---- which tile ------------------------------------------------------
    const int ntx = (kp.pts_n[0] + TX - 1) / TX, nty = (kp.pts_n[1] + TY - 1) / TY, ntz = (kp.pts_n[2] + TZ - 1) / TZ;
    int tid = blockIdx.x;
    const int tz = tid % ntz; tid /= ntz;
    const int ty = tid % nty; tid /= nty;
    const int tx = tid % ntx; tid /= ntx;
    const i64 b = tid;
    const int x0 = tx * TX, y0 = ty * TY, z0 = tz * TZ;
    const int nzv = min(TZ, kp.pts_n[2] - z0);          // valid points along z in this tile
    const int lz = threadIdx.x % TZ, ly = threadIdx.x / TZ;
    const bool col_ok = (y0 + ly < kp.pts_n[1]) && (lz < nzv);

    // ---- 1. + 2. grid coordinates, bounding boxes, plan (tile_common.cuh) ------
    stage_grid_tile<T, TX, TY, TZ, NT>(kp, grid + b * kp.grid_sb, gtile, x0, y0, z0, nzv, vec_ok);
    cp_async_wait_all();
    __syncthreads();
    tile_add_identity<T, TX, NT>(kp, gtile, x0, y0 + ly, z0 + lz);
    plan_from_coords<T, ORDER, TX, NT>(kp, gtile, col_ok, x0, red, pb, geoms, nsub_p, cap, 0.f, TZ == 16 ? 16 : 32);
    const int nsub = *nsub_p;
    const int per = TX / nsub;

    for (int s = 0; s < nsub; ++s) {
        const TileGeom g = geoms[s];
        if (g.ext[0] == 0 && g.fits) {
            // nothing in bounds in this group: zeros
            for (i64 c = 0; c < (BWD ? 1 : kp.channels); ++c)
                for (int p = s * per; p < (s + 1) * per; ++p)
                    if (col_ok && x0 + p < kp.pts_n[0]) {
                        const i64 r = ((i64)(x0 + p) * kp.pts_n[1] + (y0 + ly)) * kp.pts_n[2] + (z0 + lz);
                        T *dst = out + ((b * (BWD ? 1 : kp.channels) + c) * kp.pts_total + r) * (GRAD ? 3 : 1);
                        Traits<T>::store(dst, 0.f);
                        if (GRAD) { Traits<T>::store(dst + 1, 0.f); Traits<T>::store(dst + 2, 0.f); }
                    }
            continue;
        }
        if (g.fits) { __syncthreads(); build_tables<NT>(kp, g, idx_tab, sgn_tab); }
        for (i64 c = 0; c < kp.channels; ++c) {
            const T *src = vol + b * kp.vol_sb + c * kp.vol_sc;
            T *dst = out + (BWD ? b : b * kp.channels + c) * kp.pts_total * (GRAD ? 3 : 1);
            const T *gm = BWD ? gout + b * kp.img_sb + c * kp.img_sc : nullptr;
            if (g.fits) {
                // ---- 3. stage the box, one 4-word vector per thread -------------------
                // x / y folding comes from the per-axis tables (row base, row sign).  A
                // vector whose 4 source voxels lie inside the volume along z on a row of
                // sign +1 is one asynchronous 16-byte copy; the others (volume border,
                // sign -1 / 0) are folded element by element.
                __syncthreads();          // previous taps done, tables visible
                const int nrows = g.ext[0] * g.ext[1];
                const int total = nrows * g.vpr;
                const int zlo = (kp.bound[2] == IB200_BOUND_DST1) ? 1 : 0;
                for (int q = threadIdx.x; q < total; q += NT) {
                    const int r = fast_div(q, g.inv_vpr), v = q - r * g.vpr;
                    const int a = fast_div(r, g.inv_e1), bb = r - a * g.ext[1];
                    const int rowbase = idx_tab[a] + idx_tab[kMaxExt + bb];
                    const float rowsgn = sgn_tab[a] * sgn_tab[kMaxExt + bb];
                    float *tdst = tile + a * g.sxy + bb * g.sz + v * 4;
                    const int zs = g.lo[2] + v * 4;
                    if (F32 && vec_ok && rowsgn == 1.f && zs >= zlo && zs + 3 <= kp.vol_n[2] - 1) {
                        cp_async16(tdst, src + rowbase + zs);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int ee = min(v * 4 + e, g.ext[2] - 1);
                            tdst[e] = rowsgn * sgn_tab[2 * kMaxExt + ee] * Traits<T>::load(src + rowbase + idx_tab[2 * kMaxExt + ee]);
                        }
                    }
                }
                cp_async_wait_all();
                __syncthreads();
                // ---- 4. taps from shared memory ---------------------------------------
#pragma unroll 1
                for (int p = s * per; p < (s + 1) * per; ++p) {
                    if (!(col_ok && x0 + p < kp.pts_n[0])) continue;
                    const T *gp = gtile + (p * NT + threadIdx.x) * 3;
                    const float cc[3] = {(float)gp[0], (float)gp[1], (float)gp[2]};
                    const float f0 = floorf(cc[0] - 0.5f * (ORDER - 1)), f1 = floorf(cc[1] - 0.5f * (ORDER - 1)),
                                f2 = floorf(cc[2] - 0.5f * (ORDER - 1));
                    const bool actp = inbounds<float, 3>(kp, cc) && fabsf(f0) < 4e18f && fabsf(f1) < 4e18f && fabsf(f2) < 4e18f;
                    float acc = 0.f, ax_ = 0.f, ay_ = 0.f, az_ = 0.f;
                    if (actp) {
                        float wx[W], wy[W], wz[W], gx[W], gy[W], gz[W];
                        fast_weights<ORDER>(cc[0] - f0, wx);
                        fast_weights<ORDER>(cc[1] - f1, wy);
                        fast_weights<ORDER>(cc[2] - f2, wz);
                        if (GRAD) {
                            fast_dweights<ORDER>(cc[0] - f0, gx);
                            fast_dweights<ORDER>(cc[1] - f1, gy);
                            fast_dweights<ORDER>(cc[2] - f2, gz);
                        }
                        const float *ri = tile + ((int)f0 - g.lo[0]) * g.sxy + ((int)f1 - g.lo[1]) * g.sz + ((int)f2 - g.lo[2]);
#pragma unroll UI
                        for (int i = 0; i < W; ++i) {
                            const float *rj = ri;
                            float s00 = 0.f, s10 = 0.f, s01 = 0.f;
#pragma unroll UJ
                            for (int j = 0; j < W; ++j) {
                                float t0 = 0.f, t1 = 0.f;
#pragma unroll
                                for (int k = 0; k < W; ++k) {
                                    const float v = rj[k];
                                    t0 = fmaf(wz[k], v, t0);
                                    if (GRAD) t1 = fmaf(gz[k], v, t1);
                                }
                                s00 = fmaf(wy[j], t0, s00);
                                if (GRAD) { s10 = fmaf(gy[j], t0, s10); s01 = fmaf(wy[j], t1, s01); }
                                rj += g.sz;
                            }
                            if (!GRAD) acc = fmaf(wx[i], s00, acc);
                            else { ax_ = fmaf(gx[i], s00, ax_); ay_ = fmaf(wx[i], s10, ay_); az_ = fmaf(wx[i], s01, az_); }
                            ri += g.sxy;
                        }
                    }
                    const i64 r = ((i64)(x0 + p) * kp.pts_n[1] + (y0 + ly)) * kp.pts_n[2] + (z0 + lz);
                    if (BWD) {
                        const float m = Traits<T>::load(gm + r);
                        ax_ *= m; ay_ *= m; az_ *= m;
                        if (c > 0) { ax_ += Traits<T>::load_rw(dst + r * 3); ay_ += Traits<T>::load_rw(dst + r * 3 + 1); az_ += Traits<T>::load_rw(dst + r * 3 + 2); }
                    }
                    if (!GRAD) Traits<T>::store(dst + r, acc);
                    else { Traits<T>::store(dst + r * 3, ax_); Traits<T>::store(dst + r * 3 + 1, ay_); Traits<T>::store(dst + r * 3 + 2, az_); }
                }
            } else {
                // ---- incoherent group: direct global gathers --------------------------
#pragma unroll 1
                for (int p = s * per; p < (s + 1) * per; ++p) {
                    if (!(col_ok && x0 + p < kp.pts_n[0])) continue;
                    const T *gp = gtile + (p * NT + threadIdx.x) * 3;
                    const float cc[3] = {(float)gp[0], (float)gp[1], (float)gp[2]};
                    float acc = 0.f, ag[3] = {0.f, 0.f, 0.f};
                    if (inbounds<float, 3>(kp, cc)) {
                        Axis<float, W> ax[3];
                        bool ok = setup_axis<float, ORDER, GRAD ? 1 : 0, W>(ax[0], cc[0], ORDER, kp.bound[0], kp.vol_n[0], (int)kp.vol_s[0], kp);
                        ok = setup_axis<float, ORDER, GRAD ? 1 : 0, W>(ax[1], cc[1], ORDER, kp.bound[1], kp.vol_n[1], (int)kp.vol_s[1], kp) && ok;
                        ok = setup_axis<float, ORDER, GRAD ? 1 : 0, W>(ax[2], cc[2], ORDER, kp.bound[2], kp.vol_n[2], (int)kp.vol_s[2], kp) && ok;
                        if (ok) {
#pragma unroll
                            for (int i = 0; i < W; ++i) {
                                float s00 = 0.f, s10 = 0.f, s01 = 0.f;
#pragma unroll
                                for (int j = 0; j < W; ++j) {
                                    float t0 = 0.f, t1 = 0.f;
#pragma unroll
                                    for (int k = 0; k < W; ++k) {
                                        const float v = Traits<T>::load(src + ax[0].off[i] + ax[1].off[j] + ax[2].off[k]);
                                        t0 = fmaf(ax[2].w[k], v, t0);
                                        if (GRAD) t1 = fmaf(ax[2].g[k], v, t1);
                                    }
                                    s00 = fmaf(ax[1].w[j], t0, s00);
                                    if (GRAD) { s10 = fmaf(ax[1].g[j], t0, s10); s01 = fmaf(ax[1].w[j], t1, s01); }
                                }
                                if (!GRAD) acc = fmaf(ax[0].w[i], s00, acc);
                                else { ag[0] = fmaf(ax[0].g[i], s00, ag[0]); ag[1] = fmaf(ax[0].w[i], s10, ag[1]); ag[2] = fmaf(ax[0].w[i], s01, ag[2]); }
                            }
                        }
                    }
                    const i64 r = ((i64)(x0 + p) * kp.pts_n[1] + (y0 + ly)) * kp.pts_n[2] + (z0 + lz);
                    if (BWD) {
                        const float m = Traits<T>::load(gm + r);
                        ag[0] *= m; ag[1] *= m; ag[2] *= m;
                        if (c > 0) { ag[0] += Traits<T>::load_rw(dst + r * 3); ag[1] += Traits<T>::load_rw(dst + r * 3 + 1); ag[2] += Traits<T>::load_rw(dst + r * 3 + 2); }
                    }
                    if (!GRAD) Traits<T>::store(dst + r, acc);
                    else { Traits<T>::store(dst + r * 3, ag[0]); Traits<T>::store(dst + r * 3 + 1, ag[1]); Traits<T>::store(dst + r * 3 + 2, ag[2]); }
                }
            }
        }
    }
}

// ---------------------------------------------------------------- launch --

template <typename T, int ORDER, int OP, int TX, int TY, int TZ, int NT, int MINB>
static int launch_pull_tile_cfg(const KParams &kp, const void *vol, const void *grid, const void *gout, void *out, cudaStream_t stream,
                                size_t smem_total) {
    const size_t fixed = (size_t)TX * TY * TZ * 3 * sizeof(T) + 3 * kMaxExt * (sizeof(int) + sizeof(float)) +
                         TX * (NT / 32) * 8 * sizeof(int) + TX * (sizeof(PlaneBox) + sizeof(TileGeom)) + 32;
    const int cap = (int)((smem_total - fixed) / sizeof(float)) & ~31;
    const i64 ntiles = kp.batch * ((kp.pts_n[0] + TX - 1) / TX) * ((kp.pts_n[1] + TY - 1) / TY) * ((kp.pts_n[2] + TZ - 1) / TZ);
    if (ntiles == 0) return 1;
    if (ntiles > 0x7fffffffLL) return 0;
    // 16-byte staging needs 16-byte aligned rows in both the grid and the volume
    const int ev = 16 / (int)sizeof(T);
    const bool vec_ok = ((uintptr_t)vol % 16 == 0) && ((uintptr_t)grid % 16 == 0) &&
                        (kp.vol_s[0] % ev == 0) && (kp.vol_s[1] % ev == 0) && (kp.vol_sb % ev == 0) && (kp.vol_sc % ev == 0) &&
                        ((kp.pts_n[2] * 3) % ev == 0) && (kp.grid_sb % ev == 0);
    auto kern = pull_tile3d_kernel<T, ORDER, OP, TX, TY, TZ, NT, MINB>;
    IB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_total));
    kern<<<(unsigned)ntiles, NT, smem_total, stream>>>(kp, (const T *)vol, (const T *)grid, (const T *)gout, (T *)out, cap, vec_ok ? 1 : 0);
    static thread_local char name[64];
    snprintf(name, sizeof(name), "%s_tile3d_o%d_%dx%dx%d", OP == OP_GRAD ? "grad" : OP == OP_PULL_BWD_GRID ? "pullbwd" : "pull", ORDER, TX, TY, TZ);
    note_launch(name);
    IB200_CUDA_CHECK(cudaGetLastError());
    return 1;
}

template <typename T, int ORDER, int OP>
static int launch_pull_tile(const KParams &kp, const void *vol, const void *grid, const void *gout, void *out, cudaStream_t stream) {
#ifdef IB200_TUNE
    const char *e = getenv("IB200_VARIANT");
    const int v = e ? atoi(e) : 0;
    if (sizeof(T) == 4 && OP == OP_PULL && (ORDER == 3 || ORDER == 1)) {
        switch (v) {
        case 1: return launch_pull_tile_cfg<T, ORDER, OP, 8, 8, 32, 256, 2>(kp, vol, grid, gout, out, stream, 84 * 1024);
        case 2: return launch_pull_tile_cfg<T, ORDER, OP, 8, 4, 32, 128, 4>(kp, vol, grid, gout, out, stream, 56 * 1024);
        case 3: return launch_pull_tile_cfg<T, ORDER, OP, 8, 4, 32, 128, 5>(kp, vol, grid, gout, out, stream, 44 * 1024);
        case 4: return launch_pull_tile_cfg<T, ORDER, OP, 4, 8, 32, 256, 3>(kp, vol, grid, gout, out, stream, 56 * 1024);
        case 5: return launch_pull_tile_cfg<T, ORDER, OP, 4, 4, 32, 128, 6>(kp, vol, grid, gout, out, stream, 36 * 1024);
        }
    }
#endif
    return launch_pull_tile_cfg<T, ORDER, OP, 8, 8, 32, 256, 2>(kp, vol, grid, gout, out, stream, 110 * 1024);
}

template <typename T, int OP>
static int dispatch_pull_tile(const KParams &kp, const void *vol, const void *grid, const void *gout, void *out, cudaStream_t stream) {
    switch (kp.order[0]) {
    case 1: return launch_pull_tile<T, 1, OP>(kp, vol, grid, gout, out, stream);
    case 2: return launch_pull_tile<T, 2, OP>(kp, vol, grid, gout, out, stream);
    case 3: return launch_pull_tile<T, 3, OP>(kp, vol, grid, gout, out, stream);
    case 4: return launch_pull_tile<T, 4, OP>(kp, vol, grid, gout, out, stream);
    case 5: return launch_pull_tile<T, 5, OP>(kp, vol, grid, gout, out, stream);
    case 6: return launch_pull_tile<T, 6, OP>(kp, vol, grid, gout, out, stream);
    case 7: return launch_pull_tile<T, 7, OP>(kp, vol, grid, gout, out, stream);
    }
    return 0;
}

int try_pull_tiled(int op, const KParams &kp, int dtype, const void *vol, const void *grid, const void *gout, void *out,
                   cudaStream_t stream) {
    if (op != OP_PULL && op != OP_GRAD && op != OP_PULL_BWD_GRID) return 0;
    if (kp.dim != 3 || !kp.pts_dense) return 0;
    if (kp.order[0] != kp.order[1] || kp.order[0] != kp.order[2]) return 0;
    if (kp.pts_total < 32768) return 0;                        // small problems: one generic launch
    if (kp.pts_total * 3 > 0x7fffffffLL) return 0;             // 32-bit offsets inside one batch element
    if (kp.vol_s[2] != 1) return 0;                            // staging wants a unit innermost stride
    if (kp.flags & IB200_FLAG_REF_LINEAR_GRAD_SIGN) return 0;
    // displacement fields in 16-bit storage: the generic kernels add the lattice index in float32
    if ((kp.flags & IB200_FLAG_DISPLACEMENT) && dtype != IB200_F32) return 0;
    if (op == OP_PULL) {
        switch (dtype) {
        case IB200_F32: return dispatch_pull_tile<float, OP_PULL>(kp, vol, grid, nullptr, out, stream);
        case IB200_F16: return dispatch_pull_tile<__half, OP_PULL>(kp, vol, grid, nullptr, out, stream);
        }
    } else if (op == OP_GRAD) {
        switch (dtype) {
        case IB200_F32: return dispatch_pull_tile<float, OP_GRAD>(kp, vol, grid, nullptr, out, stream);
        case IB200_F16: return dispatch_pull_tile<__half, OP_GRAD>(kp, vol, grid, nullptr, out, stream);
        }
    } else if (dtype == IB200_F32 && gout && kp.channels >= 1) {
        // (16-bit storage would round the running sum once per channel: the generic kernel sums in registers)
        return dispatch_pull_tile<float, OP_PULL_BWD_GRID>(kp, vol, grid, gout, out, stream);
    }
    return 0;
}

}  // namespace ib200
