// Helpers shared by the shared-memory tiled pull / push kernels (3-D, compile-
// time isotropic order): fixed node->piece weight evaluation, tile geometry,
// CTA-wide bounding-box reduction, boundary tables.
#pragma once
#include "support.cuh"

namespace ib200 {

// Weights of the ORDER+1 nodes from t = g - i0, with the polynomial piece of
// every node fixed at compile time: node k sits at signed distance x = t - k
// and t lies in [(ORDER-1)/2, (ORDER+1)/2] so |x| falls in a known knot
// interval (B-splines are continuous, so the closed end does not matter).
// Same polynomials as splines.cuh / interpol/splines.py:30-80.
template <int ORDER>
__device__ __forceinline__ void fast_weights(float t, float (&w)[ORDER + 1]) {
    if constexpr (ORDER == 0) {
        w[0] = 1.f;
    } else if constexpr (ORDER == 1) {
        w[0] = 1.f - t; w[1] = t;
    } else if constexpr (ORDER == 2) {
        // t in [0.5, 1.5]: nodes at |x| = t (outer), |t-1| (inner), 2-t (outer)
        const float a = 1.5f - t, c = t - 0.5f, x1 = t - 1.f;
        w[0] = 0.5f * a * a;
        w[1] = 0.75f - x1 * x1;
        w[2] = 0.5f * c * c;
    } else if constexpr (ORDER == 3) {
        // t in [1, 2]: |x| = t (outer), t-1 (inner), 2-t (inner), 3-t (outer)
        // u = t - 1 in [0, 1], a = 1 - u:  u^3/6, (3u^3 - 6u^2 + 4)/6 and their mirror images, in FMA form
        const float u = t - 1.f, a = 2.f - t, u2 = u * u, a2 = a * a;
        w[0] = a2 * (a * (1.f / 6.f));
        w[1] = fmaf(u2, fmaf(u, 0.5f, -1.f), 2.f / 3.f);
        w[2] = fmaf(a2, fmaf(a, 0.5f, -1.f), 2.f / 3.f);
        w[3] = u2 * (u * (1.f / 6.f));
    } else {
#pragma unroll
        for (int k = 0; k <= ORDER; ++k) {
            if (ORDER >= 6) w[k] = (float)spline_weight<double>(ORDER, (double)(t - (float)k));
            else w[k] = spline_weight<float>(ORDER, t - (float)k);
        }
    }
}

// largest value a single node weight can take (bounds the splat contributions)
__host__ __device__ constexpr float max_weight(int order) {
    return order <= 1 ? 1.f : order == 2 ? 0.75f : order == 3 ? (2.f / 3.f) : order == 4 ? (115.f / 192.f)
         : order == 5 ? 0.55f : order == 6 ? (5887.f / 11520.f) : (151.f / 315.f);
}

// first derivative of the node weights, same fixed node->piece assignment
// (interpol/splines.py:90-139; order 1 uses the iso1 closed form -1/+1)
template <int ORDER>
__device__ __forceinline__ void fast_dweights(float t, float (&g)[ORDER + 1]) {
    if constexpr (ORDER == 0) {
        g[0] = 0.f;
    } else if constexpr (ORDER == 1) {
        g[0] = -1.f; g[1] = 1.f;
    } else if constexpr (ORDER == 2) {
        g[0] = t - 1.5f; g[1] = -2.f * (t - 1.f); g[2] = t - 0.5f;
    } else if constexpr (ORDER == 3) {
        const float a = 2.f - t, x1 = t - 1.f;
        g[0] = -0.5f * a * a;
        g[1] = x1 * (1.5f * x1 - 2.f);
        g[2] = -a * (1.5f * a - 2.f);
        g[3] = 0.5f * x1 * x1;
    } else {
#pragma unroll
        for (int k = 0; k <= ORDER; ++k) {
            if (ORDER >= 6) g[k] = (float)spline_grad<double>(ORDER, (double)(t - (float)k));
            else g[k] = spline_grad<float>(ORDER, t - (float)k);
        }
    }
}

struct TileGeom {
    int lo[3];      // unfolded source coordinate of tile element (0,0,0)
    int ext[3];     // extents
    int sz, sxy;    // strides (elements) of the y and x axes inside the tile
    int fits;       // the box fits in shared memory
    int plain;      // the box lies strictly inside the volume: no fold, no sign
    int vpr;        // 16-byte vectors per staged row
    unsigned inv_vpr, inv_e1, inv_cpr;   // ceil(2^32 / d) for d = vpr, ext[1], sz/32
};

// exact q / d for q * d < 2^32 with inv = floor(2^32 / d) + 1
__device__ __forceinline__ int fast_div(int q, unsigned inv) { return inv ? (int)__umulhi((unsigned)q, inv) : q; }
// (32-bit form: 0xffffffff / d + 1 == floor(2^32 / d) + 1 unless d is a power of two, where it is 2^32 / d: exact too)
__host__ __device__ __forceinline__ unsigned make_inv(int d) { return d <= 1 ? 0u : 0xffffffffu / (unsigned)d + 1u; }

constexpr int kMaxExt = 160;      // longest tile edge the boundary tables hold
constexpr int kIntMax = 0x7fffffff;
constexpr int kIntMin = -0x7fffffff - 1;

struct PlaneBox { int mn[3], mx[3]; };   // support starts of one x-plane of the tile

// Geometry of the union of planes [p0, p1).  Rows are padded to a multiple of 32
// words: the bank of a tap then depends on its z only, so lanes whose supports
// sit on different (x, y) rows never collide; the z origin is aligned to 4 words
// so rows can be staged / flushed 16 bytes at a time.
template <int ORDER>
__device__ __forceinline__ TileGeom make_geom(const KParams &kp, const PlaneBox *pb, int p0, int p1, int cap, int zround = 32,
                                                int maxext = kMaxExt) {
    TileGeom g;
    bool any = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        int a = kIntMax, b = kIntMin;
        for (int p = p0; p < p1; ++p) { a = min(a, pb[p].mn[d]); b = max(b, pb[p].mx[d]); }
        if (a > b) { any = false; a = 0; b = 0; }
        if (d == 2) a &= ~3;
        g.lo[d] = a;
        const long long e = (long long)b - a + 1 + ORDER;
        g.ext[d] = (int)(e > 0x3fffffff ? 0x3fffffff : e);
    }
    g.sz = (int)(((long long)g.ext[2] + zround - 1) / zround * zround);
    // zround == 16: lanes are tiled 2 rows x 16 z; an ODD multiple of 16 puts neighbouring rows 16 banks apart
    if (zround == 16 && (g.sz & 16) == 0) g.sz += 16;
    g.sxy = g.sz * (g.ext[1] < 4096 ? g.ext[1] : 4096);
    const long long vol = (long long)g.sz * g.ext[1] * g.ext[0];
    g.fits = any && g.ext[0] <= maxext && g.ext[1] <= maxext && g.ext[2] <= maxext && vol <= cap;
    if (!any) { g.fits = 1; g.ext[0] = g.ext[1] = g.ext[2] = 0; g.sz = 32; g.sxy = 0; }
    bool plain = g.fits && any;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int lo_ok = (kp.bound[d] == IB200_BOUND_DST1) ? 1 : 0;     // dst1 zeroes voxel 0 (Q1)
        const int hi = g.lo[d] + (d == 2 ? ((g.ext[2] + 3) & ~3) : g.ext[d]) - 1;
        plain = plain && g.lo[d] >= lo_ok && hi <= kp.vol_n[d] - 1;
    }
    g.plain = plain;
    g.vpr = (g.ext[2] + 3) >> 2;
    g.inv_vpr = make_inv(g.vpr);
    g.inv_e1 = make_inv(g.ext[1]);
    g.inv_cpr = make_inv(g.sz >> 5 > 0 ? g.sz >> 5 : 1);
    return g;
}

// idx_tab[d][e] = bound_index(lo_d + e) * stride_d ; sgn_tab[d][e] = bound_sign(lo_d + e)
template <int NT>
__device__ __forceinline__ void build_tables(const KParams &kp, const TileGeom &g, int *idx_tab, float *sgn_tab) {
    for (int q = threadIdx.x; q < 3 * kMaxExt; q += NT) {
        const int d = q / kMaxExt, e = q - d * kMaxExt;
        if (e < g.ext[d]) {
            const int src = g.lo[d] + e;
            idx_tab[q] = bound_index<int>(kp.bound[d], src, kp.vol_n[d]) * (int)kp.vol_s[d];
            sgn_tab[q] = (float)bound_sign<int>(kp.bound[d], src, kp.vol_n[d]);
        }
    }
}

// ---- cp.async helpers (LDGSTS: global -> shared without register staging) ----
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

}  // namespace ib200

namespace ib200 {

// ---- shared phases of the tiled kernels -----------------------------------

// Phase 1: stage the grid coordinates of a TX x TY x TZ tile (TX*TY rows of TZ*3
// values) into shared memory; a warp copies one row per pass, 16 bytes per lane.
template <typename T, int TX, int TY, int TZ, int NT>
__device__ __forceinline__ void stage_grid_tile(const KParams &kp, const T *gridb, T *gtile, int x0, int y0, int z0,
                                                int nzv, int vec_ok) {
    constexpr int ROWV = TZ * 3 * (int)sizeof(T) / 16;               // 16-byte vectors per row (<= 32)
    constexpr int EPV = 16 / (int)sizeof(T);
    constexpr int NW = NT / 32;
    static_assert(ROWV <= 32, "one warp stages one row of grid coordinates per pass");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool vec = vec_ok && nzv == TZ;
#pragma unroll
    for (int rw = warp; rw < TX * TY; rw += NW) {                    // warp-uniform row
        const int lx = rw / TY, lyy = rw - lx * TY;
        if (lane < ROWV && x0 + lx < kp.pts_n[0] && y0 + lyy < kp.pts_n[1]) {
            const int off = (((x0 + lx) * kp.pts_n[1] + (y0 + lyy)) * kp.pts_n[2] + z0) * 3 + lane * EPV;
            T *sdst = gtile + rw * (TZ * 3) + lane * EPV;
            if (vec) {
                cp_async16(sdst, gridb + off);
            } else {
                for (int e = 0; e < EPV; ++e)
                    if (lane * EPV + e < nzv * 3) sdst[e] = gridb[off + e];
            }
        }
    }
}

// Displacement fields (IB200_FLAG_DISPLACEMENT): the staged tile holds displacements; every thread turns ITS OWN
// points (plane p, column threadIdx.x -- the only ones it ever reads) into coordinates, in place.  (float32
// tiles only: 16-bit displacement fields take the generic kernels, which add the index in float32.)
template <typename T, int TX, int NT>
__device__ __forceinline__ void tile_add_identity(const KParams &kp, T *gtile, int x0, int y, int z) {
    if (!(kp.flags & IB200_FLAG_DISPLACEMENT)) return;
#pragma unroll
    for (int p = 0; p < TX; ++p) {
        T *g = gtile + (p * NT + threadIdx.x) * 3;
        g[0] = (T)((float)g[0] + (float)(x0 + p));
        g[1] = (T)((float)g[1] + (float)y);
        g[2] = (T)((float)g[2] + (float)z);
    }
}

// support start of the point of this thread in plane p: 0 inactive, 1 ok, 2 absurd
template <typename T, int ORDER, int NT>
__device__ __forceinline__ int support_start(const KParams &kp, const T *gtile, int p, bool in_tile, int (&i0)[3]) {
    if (!in_tile) return 0;
    const T *g = gtile + (p * NT + threadIdx.x) * 3;
    const float c[3] = {(float)g[0], (float)g[1], (float)g[2]};
    // nd.py:45: support start floor(g - (order-1)/2)
    const float f0 = floorf(c[0] - 0.5f * (ORDER - 1)), f1 = floorf(c[1] - 0.5f * (ORDER - 1)),
                f2 = floorf(c[2] - 0.5f * (ORDER - 1));
    if (!(inbounds<float, 3>(kp, c) && fabsf(f0) < 4e18f && fabsf(f1) < 4e18f && fabsf(f2) < 4e18f)) return 0;
    if (!(fabsf(f0) < 1e9f && fabsf(f1) < 1e9f && fabsf(f2) < 1e9f)) return 2;
    i0[0] = (int)f0; i0[1] = (int)f1; i0[2] = (int)f2;
    return 1;
}

// order-preserving float <-> int key (so that integer REDUX min/max works on floats)
__device__ __forceinline__ int fkey(float f) { const int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }
__device__ __forceinline__ float fkey_inv(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }
// floor(c - (ORDER-1)/2) as a saturated int
template <int ORDER>
__device__ __forceinline__ int start_of(float c) {
    const float f = floorf(c - 0.5f * (ORDER - 1));
    return (int)fminf(fmaxf(f, -1.5e9f), 1.5e9f);
}

// Phase 2: bounding boxes -> plan (whole tile first, per-plane split when it does not fit).
// The box is reduced on the raw coordinates (floor is monotone, so the floor of the
// min / max coordinate is the min / max support start); NaNs drop out of fminf/fmaxf,
// infinities make the box infinite and send the tile to the global fallback.
// `extra` is an additional non-negative per-thread value reduced with max (|value| for push);
// its CTA-wide maximum is left in red[kExtraSlot] as raw float bits.
constexpr int kExtraSlot = 6;
template <typename T, int ORDER, int TX, int NT>
__device__ __forceinline__ void plan_from_coords(const KParams &kp, const T *gtile, bool col_ok, int x0, int *red,
                                                 PlaneBox *pb, TileGeom *geoms, int *nsub, int cap, float extra = 0.f,
                                                 int zround = 32, int maxext = kMaxExt) {
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool masked = kp.extrapolate != 1;
    auto reduce_planes = [&](int p0, int p1, int slot) {
        float mn[3] = {3e38f, 3e38f, 3e38f}, mx[3] = {-3e38f, -3e38f, -3e38f};
        for (int p = p0; p < p1; ++p) {
            if (col_ok && x0 + p < kp.pts_n[0]) {
                const T *g = gtile + (p * NT + threadIdx.x) * 3;
                const float c[3] = {(float)g[0], (float)g[1], (float)g[2]};
                if (!masked || inbounds<float, 3>(kp, c)) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) { mn[d] = fminf(mn[d], c[d]); mx[d] = fmaxf(mx[d], c[d]); }
                }
            }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int a = __reduce_min_sync(0xffffffffu, fkey(mn[d]));
            const int c = __reduce_max_sync(0xffffffffu, fkey(mx[d]));
            if (lane == 0) { red[(slot * NW + warp) * 8 + 2 * d] = a; red[(slot * NW + warp) * 8 + 2 * d + 1] = c; }
        }
    };
    auto combine = [&](int slot) -> PlaneBox {
        PlaneBox b;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            int a = kIntMax, c = kIntMin;
            for (int w = 0; w < NW; ++w) { a = min(a, red[(slot * NW + w) * 8 + 2 * d]); c = max(c, red[(slot * NW + w) * 8 + 2 * d + 1]); }
            const float fa = fkey_inv(a), fc = fkey_inv(c);
            if (fa > fc) { b.mn[d] = kIntMax; b.mx[d] = kIntMin; }        // no active point
            else { b.mn[d] = start_of<ORDER>(fa); b.mx[d] = start_of<ORDER>(fc); }
        }
        return b;
    };
    // whole tile (slot 0) + the extra maximum
#pragma unroll
    for (int dummy = 0; dummy < 1; ++dummy) {
        float mn[3] = {3e38f, 3e38f, 3e38f}, mx[3] = {-3e38f, -3e38f, -3e38f};
#pragma unroll
        for (int p = 0; p < TX; ++p) {
            if (col_ok && x0 + p < kp.pts_n[0]) {
                const T *g = gtile + (p * NT + threadIdx.x) * 3;
                const float c[3] = {(float)g[0], (float)g[1], (float)g[2]};
                if (!masked || inbounds<float, 3>(kp, c)) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) { mn[d] = fminf(mn[d], c[d]); mx[d] = fmaxf(mx[d], c[d]); }
                }
            }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int a = __reduce_min_sync(0xffffffffu, fkey(mn[d]));
            const int c = __reduce_max_sync(0xffffffffu, fkey(mx[d]));
            if (lane == 0) { red[warp * 8 + 2 * d] = a; red[warp * 8 + 2 * d + 1] = c; }
        }
        const unsigned e = __reduce_max_sync(0xffffffffu, __float_as_uint(extra));
        if (lane == 0) red[warp * 8 + kExtraSlot] = (int)e;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        pb[0] = combine(0);
        geoms[0] = make_geom<ORDER>(kp, pb, 0, 1, cap, zround, maxext);
        *nsub = geoms[0].fits ? 1 : 0;
        unsigned e = 0;
        for (int w = 0; w < NW; ++w) e = max(e, (unsigned)red[w * 8 + kExtraSlot]);
        red[kExtraSlot] = (int)e;           // slot of warp 0: read by everyone after the barrier
    }
    __syncthreads();
    if (*nsub == 1) return;
    // ---- rare: the whole-tile box does not fit -> per-plane boxes, groups of planes ----
    const int extra_bits = red[kExtraSlot];
    __syncthreads();
    for (int p = 0; p < TX; ++p) reduce_planes(p, p + 1, p);
    __syncthreads();
    if (threadIdx.x < TX) pb[threadIdx.x] = combine(threadIdx.x);
    __syncthreads();
    if (threadIdx.x == 0) {
        int ns = 2;
        for (; ns <= TX; ns *= 2) {
            bool ok = true;
            const int per = TX / ns;
            for (int s = 0; s < ns; ++s) {
                geoms[s] = make_geom<ORDER>(kp, pb, s * per, (s + 1) * per, cap, zround, maxext);
                ok = ok && geoms[s].fits;
            }
            if (ok || ns == TX) break;
        }
        *nsub = ns;
        red[kExtraSlot] = extra_bits;
    }
    __syncthreads();
}

}  // namespace ib200
