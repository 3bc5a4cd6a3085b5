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

// IL = 4: channel-interleaved variant for float32 volumes with a multiple of 4 channels.  The box holds one float4 per
// voxel (channels c .. c+3, gathered from the four channel planes while staging), so a tap is ONE LDS.128 for four
// channels and coordinates, weights and row addresses are evaluated once per point instead of once per point and
// channel: 64 + 4 * 44 instructions per point for the taps of four channels instead of 4 * (64 + 84).  A quarter warp
// (8 lanes, consecutive z) reads 8 consecutive float4 = all 32 banks, so rows need no padding and lanes on different
// rows never collide.
template <typename T, int ORDER, int OP, int TX, int TY, int TZ, int NT, int MINB, int IL>
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
    static_assert(IL == 1 || (IL == 4 && sizeof(T) == 4 && OP != OP_PULL_BWD_GRID), "interleaved boxes: float32 pull / grad");
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
    plan_from_coords<T, ORDER, TX, NT>(kp, gtile, col_ok, x0, red, pb, geoms, nsub_p, IL == 4 ? cap / 4 : cap, 0.f,
                                       IL == 4 ? 4 : (TZ == 16 ? 16 : 32));
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
        for (i64 c = 0; c < kp.channels; c += IL) {
            const T *src = vol + b * kp.vol_sb + c * kp.vol_sc;
            T *dst = out + (BWD ? b : b * kp.channels + c) * kp.pts_total * (GRAD ? 3 : 1);
            const T *gm = BWD ? gout + b * kp.img_sb + c * kp.img_sc : nullptr;
            if constexpr (IL == 4) {
              if (g.fits) {
                // ---- 3'. stage the box: one voxel (float4 = four channels) per thread and step ----
                __syncthreads();          // previous taps done, tables visible
                float4 *tile4 = reinterpret_cast<float4 *>(tile);
                const int e2 = g.ext[2];
                const int total = g.ext[0] * g.ext[1] * e2;
                const unsigned inv_e2 = make_inv(e2);
                const i64 sc = kp.vol_sc;
                const int rot = (threadIdx.x >> 3) & 3;
                // (a warp per box row -- row base and sign once per row -- was measured slower: fewer copies in flight)
#pragma unroll 4
                for (int q = threadIdx.x; q < total; q += NT) {
                    const int r = fast_div(q, inv_e2), zz = q - r * e2;
                    const int a = fast_div(r, g.inv_e1), bb = r - a * g.ext[1];
                    const float sg = sgn_tab[a] * sgn_tab[kMaxExt + bb] * sgn_tab[2 * kMaxExt + zz];
                    const float *sp = src + idx_tab[a] + idx_tab[kMaxExt + bb] + idx_tab[2 * kMaxExt + zz];
                    float4 *td = tile4 + a * g.sxy + bb * g.sz + zz;
                    if (sg == 1.f) {
                        // The common case travels asynchronously, one 4-byte copy per channel plane.  Lane v writes word
                        // 4 v + ch of 128 consecutive words: with the same ch for all lanes the 32 words sit in 8 banks
                        // (4-way conflict, measured: 49 M wavefronts for 19 M ideal).  Copy m of lane v takes channel
                        // (m + v / 8) mod 4 instead: banks 4 (v mod 8) + ch are all distinct, and every group of 8 lanes
                        // still reads 32 contiguous bytes of one channel plane (32 M wavefronts).
                        float *tf = reinterpret_cast<float *>(td);
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            const int ch = (m + rot) & 3;
                            cp_async4(tf + ch, sp + ch * sc);
                        }
                    } else {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (sg != 0.f) v = make_float4(sg * __ldg(sp), sg * __ldg(sp + sc), sg * __ldg(sp + 2 * sc), sg * __ldg(sp + 3 * sc));
                        *td = v;
                    }
                }
                cp_async_wait_all();
                __syncthreads();
                // ---- 4'. taps: one LDS.128 per node, packed FFMA2 over the channel pairs ----
#pragma unroll 1
                for (int p = s * per; p < (s + 1) * per; ++p) {
                    if (!(col_ok && x0 + p < kp.pts_n[0])) continue;
                    const T *gp = gtile + (p * NT + threadIdx.x) * 3;
                    const float cc[3] = {(float)gp[0], (float)gp[1], (float)gp[2]};
                    const float f0 = floorf(cc[0] - 0.5f * (ORDER - 1)), f1 = floorf(cc[1] - 0.5f * (ORDER - 1)),
                                f2 = floorf(cc[2] - 0.5f * (ORDER - 1));
                    const bool actp = inbounds<float, 3>(kp, cc) && fabsf(f0) < 4e18f && fabsf(f1) < 4e18f && fabsf(f2) < 4e18f;
                    // (lo, hi) = channels (c, c+1), (c+2, c+3); value / d-dx / d-dy / d-dz accumulators
                    float2 a_lo = make_float2(0.f, 0.f), a_hi = a_lo, x_lo = a_lo, x_hi = a_lo, y_lo = a_lo, y_hi = a_lo, z_lo = a_lo, z_hi = a_lo;
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
                        const float4 *ri = tile4 + ((int)f0 - g.lo[0]) * g.sxy + ((int)f1 - g.lo[1]) * g.sz + ((int)f2 - g.lo[2]);
#pragma unroll
                        for (int i = 0; i < W; ++i) {
                            const float4 *rj = ri;
                            float2 s00l = make_float2(0.f, 0.f), s00h = s00l, s10l = s00l, s10h = s00l, s01l = s00l, s01h = s00l;
#pragma unroll
                            for (int j = 0; j < W; ++j) {
                                float2 t0l = make_float2(0.f, 0.f), t0h = t0l, t1l = t0l, t1h = t0l;
#pragma unroll
                                for (int k = 0; k < W; ++k) {
                                    const float4 v = rj[k];
                                    const float2 wk = make_float2(wz[k], wz[k]);
                                    t0l = __ffma2_rn(wk, make_float2(v.x, v.y), t0l);
                                    t0h = __ffma2_rn(wk, make_float2(v.z, v.w), t0h);
                                    if (GRAD) {
                                        const float2 gk = make_float2(gz[k], gz[k]);
                                        t1l = __ffma2_rn(gk, make_float2(v.x, v.y), t1l);
                                        t1h = __ffma2_rn(gk, make_float2(v.z, v.w), t1h);
                                    }
                                }
                                const float2 wj = make_float2(wy[j], wy[j]);
                                s00l = __ffma2_rn(wj, t0l, s00l); s00h = __ffma2_rn(wj, t0h, s00h);
                                if (GRAD) {
                                    const float2 gj = make_float2(gy[j], gy[j]);
                                    s10l = __ffma2_rn(gj, t0l, s10l); s10h = __ffma2_rn(gj, t0h, s10h);
                                    s01l = __ffma2_rn(wj, t1l, s01l); s01h = __ffma2_rn(wj, t1h, s01h);
                                }
                                rj += g.sz;
                            }
                            const float2 wi = make_float2(wx[i], wx[i]);
                            if (!GRAD) { a_lo = __ffma2_rn(wi, s00l, a_lo); a_hi = __ffma2_rn(wi, s00h, a_hi); }
                            else {
                                const float2 gi = make_float2(gx[i], gx[i]);
                                x_lo = __ffma2_rn(gi, s00l, x_lo); x_hi = __ffma2_rn(gi, s00h, x_hi);
                                y_lo = __ffma2_rn(wi, s10l, y_lo); y_hi = __ffma2_rn(wi, s10h, y_hi);
                                z_lo = __ffma2_rn(wi, s01l, z_lo); z_hi = __ffma2_rn(wi, s01h, z_hi);
                            }
                            ri += g.sxy;
                        }
                    }
                    const i64 r = ((i64)(x0 + p) * kp.pts_n[1] + (y0 + ly)) * kp.pts_n[2] + (z0 + lz);
                    const i64 cs = kp.pts_total * (GRAD ? 3 : 1);
                    if (!GRAD) {
                        dst[r] = a_lo.x; dst[cs + r] = a_lo.y; dst[2 * cs + r] = a_hi.x; dst[3 * cs + r] = a_hi.y;
                    } else {
                        const float gx4[4] = {x_lo.x, x_lo.y, x_hi.x, x_hi.y}, gy4[4] = {y_lo.x, y_lo.y, y_hi.x, y_hi.y},
                                    gz4[4] = {z_lo.x, z_lo.y, z_hi.x, z_hi.y};
#pragma unroll
                        for (int cg = 0; cg < 4; ++cg) {
                            T *d = dst + cg * cs + r * 3;
                            d[0] = gx4[cg]; d[1] = gy4[cg]; d[2] = gz4[cg];
                        }
                    }
                }
                continue;                 // next group of four channels
              }
            }
            if (IL == 1 && g.fits) {
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
                // ---- incoherent group: direct global gathers (channel by channel) ------
#pragma unroll 1
              for (int cg = 0; cg < IL; ++cg) {
                const T *src = vol + b * kp.vol_sb + (c + cg) * kp.vol_sc;
                T *dst = out + (BWD ? b : b * kp.channels + c + cg) * kp.pts_total * (GRAD ? 3 : 1);
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
}

// ---------------------------------------------------------------- launch --

template <typename T, int ORDER, int OP, int TX, int TY, int TZ, int NT, int MINB, int IL = 1>
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
    auto kern = pull_tile3d_kernel<T, ORDER, OP, TX, TY, TZ, NT, MINB, IL>;
    IB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_total));
    kern<<<(unsigned)ntiles, NT, smem_total, stream>>>(kp, (const T *)vol, (const T *)grid, (const T *)gout, (T *)out, cap, vec_ok ? 1 : 0);
    static thread_local char name[64];
    snprintf(name, sizeof(name), "%s_tile3d_o%d_%dx%dx%d%s", OP == OP_GRAD ? "grad" : OP == OP_PULL_BWD_GRID ? "pullbwd" : "pull", ORDER, TX, TY, TZ,
             IL == 4 ? "_c4" : "");
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
    // float32 volumes with a multiple of 4 channels: channel-interleaved boxes (see the kernel's header)
    if constexpr (sizeof(T) == 4 && OP != OP_PULL_BWD_GRID && ORDER <= 3) {
        if (kp.channels >= 4 && kp.channels % 4 == 0 && !(kp.flags & IB200_FLAG_NO_PIPE))
            return launch_pull_tile_cfg<T, ORDER, OP, 8, 8, 16, 128, 3, 4>(kp, vol, grid, gout, out, stream, 74 * 1024);   // (2 CTAs x 110 KB: 5-20 % slower)
    }
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
