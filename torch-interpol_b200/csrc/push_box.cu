// Boxed push / count (3-D, isotropic compile-time order, 16 / 32-bit storage): the default scatter kernel of
// round 2.  Same arithmetic as push_tile.cu -- a CTA owns a block of SOURCE voxels and privatises the scatter
// in a shared-memory box of 32-bit fixed-point accumulators (native ATOMS.ADD), flushed once with vector REDs --
// restructured around what the round-2 micro-benchmarks (profiles/micro/smem_micro.cu, tap_lab.cu) and the bank
// model (profiles/sim/bank_*.py, validated against ncu) showed:
//
//   * ATOMS.ADD without a return value issues at the LDS rate (1 wavefront / clk).  On a smooth deformation a warp
//     needs 1.9 wavefronts per atomic at best, whatever the lane <-> voxel tiling -- provided box rows that
//     neighbouring lanes hit sit in disjoint banks; rows padded to anything else cost 2.4-2.8 (ncu confirmed both).
//     Floor of the atomics: 120 clk per 32 sources per SM, 0.22 ms for 256^3.
//   * The round-1 kernel was NOT at that floor (0.61 ms, LSU data pipe 54 % busy): the other phases of a tile
//     (staging through LDG + STS, the |value| histogram behind the overflow bound, flush) ran at 2 CTAs per SM with
//     little to overlap them.  What fills the pipe is MANY SMALL INDEPENDENT CTAs (5 per SM here; persistent CTAs
//     with prefetch, fewer barriers, or 3 larger CTAs all measured slower -- profiles/README.md r2b).
//   * Tile = 8 x 8 x 16 sources, 128 threads, a warp = two z-rows of 16.  Coordinates and values arrive through
//     two TMA tile copies (no LSU instructions, no registers).  Box rows are 32 words; rows of even and odd y live
//     in two half-planes 16 banks apart ("split-parity" layout), which is what the 2 x 16 lane tiling needs to
//     reach 1.95 wavefronts per atomic, at 2/3 of the shared memory of 48-word rows.
//   * The overflow bound is the sum of |value| over the tile (every accumulator is a sub-sum of it, weights are
//     <= w3): one block reduction fused with the bounding-box pass instead of a histogram pass; the binding
//     constraint on the scale is the 21-bit single contribution of the float -> fixed conversion anyway.
//   * Values that do not fit the tile's scale -- NaN / Inf, and outliers above 16 x the (trimmed) mean |value| of
//     the tile -- bypass the box and go to the output with float atomics, like scatter.cu: non-finite values
//     propagate as in the reference, and one hot voxel no longer sets the quantisation step of its whole tile.
//     Resolution of the box: 2^-21 of min(max |value|, 16 x trimmed mean |value|) x max weight, per tile.
//
// Replaces interpol/nd.py:147-213 (and iso1.py push); push_tile.cu remains for problems whose rows are not
// 16-byte aligned (TMA), scatter.cu for everything else.
#include <cstdio>
#include <cstdlib>
#include "pipe_common.cuh"

namespace ib200 {

namespace {

constexpr float kBoxMagic = 12582912.f;   // 1.5 * 2^23: float -> int by mantissa alignment
constexpr int kBoxMagicBits = 0x4B400000;
constexpr int kBoxMaxExt = 64;            // longest box edge the boundary tables hold
constexpr int kBoxRow = 32;               // words per box row
constexpr float kOutlier = 16.f;          // outliers: |value| > kOutlier * trimmed mean |value| of the tile

// Split-parity box: element (a, r, z) of a box of ext[0] planes x ext[1] rows x ext[2] <= 32 words lives at
//     a * ps + (r >> 1) * 32 + (r & 1) * hs + z,     hs = ceil(ext[1] / 2) * 32 + 16,  ps = 2 * hs
// so rows of equal parity are 32 words apart (same banks), rows of different parity 16 banks apart, planes a
// multiple of 32 words apart.
struct BoxGeom {
    int lo[3], ext[3];
    int hs, ps;
    int fits, plain, vpr;
    unsigned inv_vpr, inv_e1;
};

__device__ __forceinline__ BoxGeom to_box(const TileGeom &g, int cap) {
    BoxGeom q;
#pragma unroll
    for (int d = 0; d < 3; ++d) { q.lo[d] = g.lo[d]; q.ext[d] = g.ext[d]; }
    q.hs = ((g.ext[1] + 1) >> 1) * kBoxRow + 16;
    q.ps = 2 * q.hs;
    const bool empty = g.ext[0] == 0;
    q.fits = g.fits && (empty || (g.ext[0] <= kBoxMaxExt && g.ext[1] <= kBoxMaxExt && g.ext[2] <= kBoxRow &&
                                  (long long)q.ps * g.ext[0] <= cap));
    q.plain = g.plain;
    q.vpr = (g.ext[2] + 3) >> 2;
    q.inv_vpr = g.inv_vpr; q.inv_e1 = g.inv_e1;
    return q;
}

// One source straight to the output volume (float REDs): incoherent groups, outliers, non-finite values.
template <int ORDER>
__device__ __noinline__ void push_point_global(const KParams &kp, float *dst, float c0, float c1, float c2, float v) {
    constexpr int W = ORDER + 1;
    const float cc[3] = {c0, c1, c2};
    if (!inbounds<float, 3>(kp, cc)) return;
    Axis<float, W> ax[3];
    bool ok = setup_axis<float, ORDER, 0, W>(ax[0], cc[0], ORDER, kp.bound[0], kp.vol_n[0], (int)kp.vol_s[0], kp);
    ok = setup_axis<float, ORDER, 0, W>(ax[1], cc[1], ORDER, kp.bound[1], kp.vol_n[1], (int)kp.vol_s[1], kp) && ok;
    ok = setup_axis<float, ORDER, 0, W>(ax[2], cc[2], ORDER, kp.bound[2], kp.vol_n[2], (int)kp.vol_s[2], kp) && ok;
    if (!ok) return;
#pragma unroll 1
    for (int i = 0; i < W; ++i)
#pragma unroll 1
        for (int j = 0; j < W; ++j) {
            const float vij = v * ax[0].w[i] * ax[1].w[j];
#pragma unroll
            for (int k = 0; k < W; ++k) atomicAdd(dst + ax[0].off[i] + ax[1].off[j] + ax[2].off[k], vij * ax[2].w[k]);
        }
}

template <typename T, int ORDER, int OP, int MINB>
__global__ void __launch_bounds__(128, MINB)
push_box3d_kernel(const __grid_constant__ KParams kp, const __grid_constant__ CUtensorMap tm_grid,
                  const __grid_constant__ CUtensorMap tm_img, float *__restrict__ out, const int cap, const int vec_ok,
                  const int gbmul, const int ibmul, const int icmul) {
    constexpr int TX = 8, TY = 8, TZ = 16, NT = TY * TZ, NPT = TX * NT;
    constexpr int W = ORDER + 1;
    constexpr int NW = NT / 32;
    constexpr bool COUNT = (OP == OP_COUNT);
    constexpr int UI = ORDER <= 3 ? W : 1, UJ = ORDER <= 5 ? W : 1;   // keep the code of high orders compact
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    T *gtile = reinterpret_cast<T *>(smem_raw);                          // [TX][TY][TZ * 3] grid coordinates (TMA)
    T *vals = gtile + NPT * 3;                                            // [TX][TY][TZ] source values (TMA)
    int *acc = reinterpret_cast<int *>(vals + NPT);                       // [cap] fixed-point accumulators
    int *idx_tab = acc + cap;                                             // [3][kBoxMaxExt]
    float *sgn_tab = reinterpret_cast<float *>(idx_tab + 3 * kBoxMaxExt);
    int *red = reinterpret_cast<int *>(sgn_tab + 3 * kBoxMaxExt);         // [TX][NW][8] general planner
    PlaneBox *pb = reinterpret_cast<PlaneBox *>(red + TX * NW * 8);       // [TX]
    TileGeom *geoms = reinterpret_cast<TileGeom *>(pb + TX);              // [TX]
    int *nsub_p = reinterpret_cast<int *>(geoms + TX);                    // [4]
    float *stat = reinterpret_cast<float *>(nsub_p + 4);                  // [NW][4] partial sums
    int *slots = reinterpret_cast<int *>(stat + NW * 4);                  // [NW][12] front end: box keys, sum / count / max |v|
    float *plan = reinterpret_cast<float *>(slots + NW * 12);             // [4] fast flag, thr, scale, 1 / scale
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(plan + 4);   // coordinates, values
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float w3 = max_weight(ORDER) * max_weight(ORDER) * max_weight(ORDER);

    // ---- which tile ------------------------------------------------------
    const int ntx = (kp.pts_n[0] + TX - 1) / TX, nty = (kp.pts_n[1] + TY - 1) / TY, ntz = (kp.pts_n[2] + TZ - 1) / TZ;
    int tid = blockIdx.x;
    const int tz = tid % ntz; tid /= ntz;
    const int ty = tid % nty; tid /= nty;
    const int tx = tid % ntx; tid /= ntx;
    const int b = tid;
    const int x0 = tx * TX, y0 = ty * TY, z0 = tz * TZ;
    const int nzv = min(TZ, kp.pts_n[2] - z0);
    const int lz = threadIdx.x % TZ, ly = threadIdx.x / TZ;
    const bool col_ok = (y0 + ly < kp.pts_n[1]) && (lz < nzv);
    const bool masked = kp.extrapolate != 1;

    // ---- 1. coordinates + values of channel 0: two TMA tile copies --------------
    auto request_values = [&](int c) {     // thread 0, once nobody reads `vals` any more
        mbar_expect_tx(bar + 1, NPT * (unsigned)sizeof(T));
        tma_load_5d(vals, &tm_img, z0, y0, x0, c * icmul, b * ibmul, bar + 1);
        mbar_arrive(bar + 1);
    };
    if (threadIdx.x == 0) {
        mbar_init(bar, 1); mbar_init(bar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        mbar_expect_tx(bar, NPT * 3 * (unsigned)sizeof(T));
        tma_load_4d(gtile, &tm_grid, z0 * 3, y0, x0, b * gbmul, bar);
        mbar_arrive(bar);
        if (!COUNT) request_values(0);
    }
    __syncthreads();                       // barrier initialisation visible
    // the accumulators are zeroed while the copies are in flight (the whole capacity: the box is not known yet)
    {
        int4 *a4 = reinterpret_cast<int4 *>(acc);
        for (int q = threadIdx.x; q < (cap >> 2); q += NT) a4[q] = make_int4(0, 0, 0, 0);
    }
    unsigned vphase = 0;                   // parity of the value barrier
    bool dirty = false;                    // (block-uniform) the box holds flushed sums / a flush may still be running
    mbar_wait(bar, 0);
    tile_add_identity<T, TX, NT>(kp, gtile, x0, y0 + ly, z0 + lz);
    if (!COUNT) { mbar_wait(bar + 1, 0); ++vphase; }

    // ---- 2. front end: ONE pass over the tile's coordinates and values -------------------
    // bounding box of all supports (floor is monotone: the floor of the min / max coordinate is the min / max
    // support start; NaNs drop out of fminf / fmaxf, infinities make the box infinite) together with sum / count /
    // max of |value|; warp 0 turns them into the plan of the common case -- the whole tile in one box, no
    // outliers -- and everything else goes through the general planner below.
    {
        float mn[3] = {3e38f, 3e38f, 3e38f}, mx[3] = {-3e38f, -3e38f, -3e38f};
        float ssum = 0.f, smax = 0.f;
        int scnt = 0;
#pragma unroll
        for (int p = 0; p < TX; ++p) {
            if (col_ok && x0 + p < kp.pts_n[0]) {
                const T *gq = gtile + (p * NT + threadIdx.x) * 3;
                const float cq[3] = {(float)gq[0], (float)gq[1], (float)gq[2]};
                if (!masked || inbounds<float, 3>(kp, cq)) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) { mn[d] = fminf(mn[d], cq[d]); mx[d] = fmaxf(mx[d], cq[d]); }
                }
                if (!COUNT) {
                    const float a = fabsf((float)vals[p * NT + threadIdx.x]);
                    const bool use = a <= 3e38f && a > 0.f;            // finite, non-zero
                    ssum += use ? a : 0.f; scnt += use ? 1 : 0; smax = fmaxf(smax, use ? a : 0.f);
                    if (!(a <= 3e38f)) smax = 3.4e38f;                 // NaN / Inf: not the common case
                }
            }
        }
        int key[6];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            key[2 * d] = __reduce_min_sync(0xffffffffu, fkey(mn[d]));
            key[2 * d + 1] = __reduce_max_sync(0xffffffffu, fkey(mx[d]));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
        scnt = __reduce_add_sync(0xffffffffu, scnt);
        const unsigned mbits = __reduce_max_sync(0xffffffffu, __float_as_uint(smax));
        if (lane == 0) {
#pragma unroll
            for (int d = 0; d < 6; ++d) slots[warp * 12 + d] = key[d];
            slots[warp * 12 + 6] = __float_as_int(ssum); slots[warp * 12 + 7] = scnt; slots[warp * 12 + 8] = (int)mbits;
        }
    }
    __syncthreads();
    if (warp == 0) {
        const int w = lane < NW ? lane : 0;
        PlaneBox box0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const int a = __reduce_min_sync(0xffffffffu, slots[w * 12 + 2 * d]);
            const int c2 = __reduce_max_sync(0xffffffffu, slots[w * 12 + 2 * d + 1]);
            const float fa = fkey_inv(a), fc = fkey_inv(c2);
            if (fa > fc) { box0.mn[d] = kIntMax; box0.mx[d] = kIntMin; }
            else { box0.mn[d] = start_of<ORDER>(fa); box0.mx[d] = start_of<ORDER>(fc); }
        }
        float ssum = lane < NW ? __int_as_float(slots[w * 12 + 6]) : 0.f;
#pragma unroll
        for (int o = NW / 2; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
        const int scnt = __reduce_add_sync(0xffffffffu, lane < NW ? slots[w * 12 + 7] : 0);
        const float smax = __uint_as_float(__reduce_max_sync(0xffffffffu, (unsigned)slots[w * 12 + 8]));
        if (lane == 0) {
            pb[0] = box0;
            const TileGeom g0 = make_geom<ORDER>(kp, pb, 0, 1, 0x3fffffff, 4, kBoxMaxExt);
            geoms[0] = g0;
            bool fast = to_box(g0, cap).fits != 0;
            float thr = 1.f, sum = (float)NPT;                         // COUNT: every value is 1
            if (!COUNT) {
                thr = smax; sum = ssum;
                fast = fast && smax < 3e38f && (scnt == 0 || smax <= kOutlier * ssum / (float)scnt);
            }
            float scale = 0.f, inv = 0.f;
            if (sum > 0.f) {
                int e1, e2;
                frexpf(sum * w3, &e1);           // any accumulator < 2^e1
                frexpf(thr * w3, &e2);           // single contribution < 2^e2
                int k = min(30 - e1, 21 - e2);
                k = max(-120, min(120, k));
                scale = ldexpf(1.f, k); inv = ldexpf(1.f, -k);
            }
            plan[0] = fast ? 1.f : 0.f; plan[1] = thr; plan[2] = scale; plan[3] = inv;
            *nsub_p = 1;
        }
    }
    __syncthreads();
    const bool fast = plan[0] != 0.f;
    if (!fast)       // the box does not fit, or the values need trimming: general planner (block-uniform branch;
                     // 48-word rows there are an upper bound of what the split-parity layout needs)
        plan_from_coords<T, ORDER, TX, NT>(kp, gtile, col_ok, x0, red, pb, geoms, nsub_p, cap, 0.f, 16, kBoxMaxExt);
    const int nsub = *nsub_p;
    const int per = TX / nsub;

    for (i64 c = 0; c < kp.channels; ++c) {
        float *dst = out + ((i64)b * kp.channels + c) * kp.vol_total;
        // ---- 3. scale of this channel's values -----------------------------------------------
        float thr = 3e38f, scale = 0.f, inv = 0.f;
        bool next_requested = COUNT || c + 1 >= kp.channels;
        if (c == 0 && (fast || COUNT)) {
            thr = plan[1]; scale = plan[2]; inv = plan[3];
        } else if (!COUNT) {
            if (c > 0) { mbar_wait(bar + 1, vphase & 1); ++vphase; }
            float av[TX];
#pragma unroll
            for (int p = 0; p < TX; ++p)
                av[p] = (col_ok && x0 + p < kp.pts_n[0]) ? fabsf((float)vals[p * NT + threadIdx.x]) : 0.f;
            // (trimmed) statistics of the finite values not above `thr`: sum, count, max
            float ssum = 0.f, smax = 0.f;
            for (int pass = 0; pass < 6; ++pass) {
                float s = 0.f, n = 0.f, m = 0.f;
#pragma unroll
                for (int p = 0; p < TX; ++p) {
                    const bool use = av[p] <= thr && av[p] > 0.f;      // NaN fails both
                    s += use ? av[p] : 0.f; n += use ? 1.f : 0.f; m = fmaxf(m, use ? av[p] : 0.f);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s += __shfl_xor_sync(0xffffffffu, s, o); n += __shfl_xor_sync(0xffffffffu, n, o);
                    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                }
                __syncthreads();                                        // previous pass read
                if (lane == 0) *reinterpret_cast<float4 *>(stat + warp * 4) = make_float4(s, n, m, 0.f);
                __syncthreads();
                s = 0.f; n = 0.f; m = 0.f;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const float4 q = *reinterpret_cast<const float4 *>(stat + w * 4);
                    s += q.x; n += q.y; m = fmaxf(m, q.z);
                }
                ssum = s; smax = m;
                const float cut = n > 0.f ? kOutlier * s / n : 0.f;
                if (!(m > cut)) break;                                  // every value kept is a regular value
                thr = cut;                                              // (block-uniform decision)
            }
            thr = fminf(thr, smax);
            if (ssum > 0.f) {
                int e1, e2;
                frexpf(ssum * w3, &e1);
                frexpf(thr * w3, &e2);
                int k = min(30 - e1, 21 - e2);
                k = max(-120, min(120, k));
                scale = ldexpf(1.f, k); inv = ldexpf(1.f, -k);
            }
        }

        for (int s = 0; s < nsub; ++s) {
            const BoxGeom g = to_box(geoms[s], cap);
            if (g.ext[0] == 0 && g.fits) continue;                      // nothing in bounds
            if (g.fits) {
                // ---- a. boundary tables (boxes that reach a face of the volume); the box is clean for the first
                //         group of a tile, zeroed again for the others ----
                if (dirty) {
                    __syncthreads();                                    // previous flush done
                    const int n4 = (g.ps * g.ext[0]) >> 2;
                    int4 *a4 = reinterpret_cast<int4 *>(acc);
                    for (int q = threadIdx.x; q < n4; q += NT) a4[q] = make_int4(0, 0, 0, 0);
                }
                if (!g.plain) {
                    for (int q = threadIdx.x; q < 3 * kBoxMaxExt; q += NT) {
                        const int d = q / kBoxMaxExt, e = q - d * kBoxMaxExt;
                        if (e < g.ext[d]) {
                            const int src = g.lo[d] + e;
                            idx_tab[q] = bound_index<int>(kp.bound[d], src, kp.vol_n[d]) * (int)kp.vol_s[d];
                            sgn_tab[q] = (float)bound_sign<int>(kp.bound[d], src, kp.vol_n[d]);
                        }
                    }
                }
                if (dirty || !g.plain) __syncthreads();
                dirty = true;
                // ---- b. integer shared-memory atomics ---------------------------------------
#pragma unroll 1
                for (int p = s * per; p < (s + 1) * per; ++p) {
                    int i0[3];
                    if (support_start<T, ORDER, NT>(kp, gtile, p, col_ok && x0 + p < kp.pts_n[0], i0) != 1) continue;
                    const T *gp = gtile + (p * NT + threadIdx.x) * 3;
                    const float c0 = (float)gp[0], c1 = (float)gp[1], c2 = (float)gp[2];
                    float v = 1.f;
                    if (!COUNT) {
                        v = (float)vals[p * NT + threadIdx.x];
                        if (!(fabsf(v) <= thr)) {                       // outlier / NaN / Inf: float REDs to the output
                            push_point_global<ORDER>(kp, dst, c0, c1, c2, v);
                            continue;
                        }
                    }
                    float wx[W], wy[W], wz[W];
                    fast_weights<ORDER>(c0 - (float)i0[0], wx);
                    fast_weights<ORDER>(c1 - (float)i0[1], wy);
                    fast_weights<ORDER>(c2 - (float)i0[2], wz);
                    v *= scale;
                    // rows r, r + 2, ... (one parity) are 32 words apart; r + 1, r + 3, ... start `hs` words away
                    const int r = i0[1] - g.lo[1];
                    const int base = (i0[0] - g.lo[0]) * g.ps + (i0[2] - g.lo[2]);
                    int *pa = acc + base + (r >> 1) * kBoxRow + (r & 1) * g.hs;
                    int *pbb = acc + base + ((r + 1) >> 1) * kBoxRow + ((r + 1) & 1) * g.hs;
#pragma unroll UI
                    for (int i = 0; i < W; ++i) {
                        const float vi = v * wx[i];
#pragma unroll UJ
                        for (int j = 0; j < W; ++j) {
                            int *rj = ((j & 1) ? pbb : pa) + (j >> 1) * kBoxRow;
                            const float vij = vi * wy[j];
                            const float2 v2 = make_float2(vij, vij), m2 = make_float2(kBoxMagic, kBoxMagic);
#pragma unroll
                            for (int k = 0; k + 1 < W; k += 2) {
                                const float2 q = __ffma2_rn(v2, make_float2(wz[k], wz[k + 1]), m2);
                                atomicAdd(rj + k, __float_as_int(q.x) - kBoxMagicBits);
                                atomicAdd(rj + k + 1, __float_as_int(q.y) - kBoxMagicBits);
                            }
                            if (W & 1) atomicAdd(rj + W - 1, __float_as_int(fmaf(vij, wz[W - 1], kBoxMagic)) - kBoxMagicBits);
                        }
                        pa += g.ps; pbb += g.ps;
                    }
                }
                __syncthreads();
                // values of the next channel travel while this one is flushed
                if (!next_requested && s == nsub - 1) {
                    if (threadIdx.x == 0) request_values((int)c + 1);
                    next_requested = true;
                }
                // ---- c. flush the box: fixed -> float, fold + sign, global REDs -------------
                // (the geometry lives on the stack -- its arrays are indexed in loops elsewhere -- and the flush loop
                // used to re-read seven of its words from local memory per vector: copies in registers)
                const int g_vpr = g.vpr, g_e1 = g.ext[1], g_e2 = g.ext[2], g_ps = g.ps, g_hs = g.hs;
                const int g_lo0 = g.lo[0], g_lo1 = g.lo[1], g_lo2 = g.lo[2];
                const bool g_plain = g.plain != 0;
                const int vs0 = (int)kp.vol_s[0], vs1 = (int)kp.vol_s[1], vs2 = (int)kp.vol_s[2], nz1 = kp.vol_n[2] - 1;
                const int total = g.ext[0] * g_e1 * g_vpr;
                const int zlo = (kp.bound[2] == IB200_BOUND_DST1) ? 1 : 0;
                const unsigned inv_vpr = g.inv_vpr, inv_e1 = g.inv_e1;
                for (int q = threadIdx.x; q < total; q += NT) {
                    const int rr = fast_div(q, inv_vpr), v4 = q - rr * g_vpr;
                    const int a = fast_div(rr, inv_e1), bb = rr - a * g_e1;
                    const int4 iv = *reinterpret_cast<const int4 *>(acc + a * g_ps + (bb >> 1) * kBoxRow + (bb & 1) * g_hs + v4 * 4);
                    if ((iv.x | iv.y | iv.z | iv.w) == 0) continue;
                    float rowsgn = 1.f;
                    int rowbase;
                    if (g_plain) {
                        rowbase = (g_lo0 + a) * vs0 + (g_lo1 + bb) * vs1;
                    } else {
                        rowsgn = sgn_tab[a] * sgn_tab[kBoxMaxExt + bb];
                        if (rowsgn == 0.f) continue;
                        rowbase = idx_tab[a] + idx_tab[kBoxMaxExt + bb];
                    }
                    const float f = inv * rowsgn;
                    const int zs = g_lo2 + v4 * 4;
                    if (vec_ok && zs >= zlo && zs + 3 <= nz1) {
                        atomicAdd(reinterpret_cast<float4 *>(dst + rowbase + zs),
                                  make_float4(f * (float)iv.x, f * (float)iv.y, f * (float)iv.z, f * (float)iv.w));
                    } else {
                        const int ivs[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int ee = v4 * 4 + e;
                            if (ivs[e] != 0 && ee < g_e2) {
                                const float sg = g_plain ? 1.f : sgn_tab[2 * kBoxMaxExt + ee];
                                const int zi = g_plain ? (g_lo2 + ee) * vs2 : idx_tab[2 * kBoxMaxExt + ee];
                                if (sg != 0.f) atomicAdd(dst + rowbase + zi, f * sg * (float)ivs[e]);
                            }
                        }
                    }
                }
            } else {
                // ---- incoherent group: direct global REDs (same arithmetic as scatter.cu) ---
#pragma unroll 1
                for (int p = s * per; p < (s + 1) * per; ++p) {
                    if (!(col_ok && x0 + p < kp.pts_n[0])) continue;
                    const T *gp = gtile + (p * NT + threadIdx.x) * 3;
                    const float v = COUNT ? 1.f : (float)vals[p * NT + threadIdx.x];
                    push_point_global<ORDER>(kp, dst, (float)gp[0], (float)gp[1], (float)gp[2], v);
                }
            }
        }
        if (!next_requested) {             // (block-uniform) last group was empty or incoherent
            __syncthreads();
            if (threadIdx.x == 0) request_values((int)c + 1);
        }
    }
}

template <typename T> struct TmaType;
template <> struct TmaType<float> { static constexpr CUtensorMapDataType value = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; };
template <> struct TmaType<__half> { static constexpr CUtensorMapDataType value = CU_TENSOR_MAP_DATA_TYPE_FLOAT16; };

template <typename T, int ORDER, int OP, int MINB>
int launch_push_box(const KParams &kp, const void *img, const void *grid, float *out, cudaStream_t stream) {
    constexpr int TX = 8, TY = 8, TZ = 16, NT = TY * TZ, NPT = TX * NT, NW = NT / 32;
    const i64 ntiles = kp.batch * ((kp.pts_n[0] + TX - 1) / TX) * ((kp.pts_n[1] + TY - 1) / TY) * ((kp.pts_n[2] + TZ - 1) / TZ);
    if (ntiles == 0) return 1;
    if (ntiles > 0x7fffffffLL) return 0;
    // shared memory: as much box as MINB CTAs per SM leave room for (227 KB per SM, 1 KB reserved per CTA)
    const size_t fixed = (size_t)NPT * 4 * sizeof(T) + 3 * kBoxMaxExt * 8 + (size_t)TX * NW * 8 * 4 +
                         TX * (sizeof(PlaneBox) + sizeof(TileGeom)) + 16 + NW * 4 * 4 + NW * 12 * 4 + 16 + 16 + 64;
    const size_t budget = (size_t)(227 * 1024) / MINB - 1024;
    if (budget < fixed + 4096 * 4) return 0;
    const int cap = (int)((budget - fixed) / sizeof(int)) & ~31;
    const size_t smem_total = fixed + (size_t)cap * sizeof(int);
    const int gbmul = (kp.grid_sb != 0 && kp.batch > 1) ? 1 : 0;
    const int ibmul = (kp.img_sb != 0 && kp.batch > 1) ? 1 : 0, icmul = (kp.img_sc != 0 && kp.channels > 1) ? 1 : 0;
    CUtensorMap tm_grid, tm_img;
    {
        const long long row = (long long)kp.pts_n[2] * 3;
        const long long dim[4] = {row, kp.pts_n[1], kp.pts_n[0], gbmul ? kp.batch : 1};
        const long long str[4] = {1, row, row * kp.pts_n[1], gbmul ? kp.grid_sb : row * kp.pts_n[1] * kp.pts_n[0]};
        const int box[4] = {TZ * 3, TY, TX, 1};
        if (!make_tensor_map_t(&tm_grid, grid, 4, dim, str, box, TmaType<T>::value, (int)sizeof(T))) return 0;
    }
    if (OP != OP_COUNT) {
        const long long nz = kp.pts_n[2], ny = kp.pts_n[1], nx = kp.pts_n[0];
        const long long vol = ((nz * ny * nx) + 15) & ~15LL;
        const long long dim[5] = {nz, ny, nx, icmul ? kp.channels : 1, ibmul ? kp.batch : 1};
        const long long str[5] = {1, nz, nz * ny, icmul ? kp.img_sc : vol, ibmul ? kp.img_sb : vol};
        const int box[5] = {TZ, TY, TX, 1, 1};
        if (!make_tensor_map_t(&tm_img, img, 5, dim, str, box, TmaType<T>::value, (int)sizeof(T))) return 0;
    } else {
        tm_img = tm_grid;
    }
    const bool vec_ok = ((uintptr_t)out % 16 == 0) && (kp.vol_n[2] % 4 == 0);
    auto kern = push_box3d_kernel<T, ORDER, OP, MINB>;
    IB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_total));
    kern<<<(unsigned)ntiles, NT, smem_total, stream>>>(kp, tm_grid, tm_img, out, cap, vec_ok ? 1 : 0, gbmul, ibmul, icmul);
    static thread_local char name[64];
    snprintf(name, sizeof(name), "%s_box3d_o%d_8x8x16", OP == OP_COUNT ? "count" : "push", ORDER);
    note_launch(name);
    IB200_CUDA_CHECK(cudaGetLastError());
    return 1;
}

template <typename T, int OP>
int dispatch_push_box(const KParams &kp, const void *img, const void *grid, float *out, cudaStream_t stream) {
    const char *e = getenv("IB200_PUSH_MINB");      // tuning aid: CTAs per SM the shared-memory budget is cut for
    const int minb = e ? atoi(e) : 5;
#define IB200_BOX_CASE(O)                                                                                        \
    case O:                                                                                                      \
        if constexpr (O == 3 && sizeof(T) == 4) {                                                                \
            if (minb == 4) return launch_push_box<T, O, OP, 4>(kp, img, grid, out, stream);                      \
            if (minb == 6) return launch_push_box<T, O, OP, 6>(kp, img, grid, out, stream);                      \
        }                                                                                                        \
        return launch_push_box<T, O, OP, 5>(kp, img, grid, out, stream);
    switch (kp.order[0]) {
        IB200_BOX_CASE(1) IB200_BOX_CASE(2) IB200_BOX_CASE(3) IB200_BOX_CASE(4)
        IB200_BOX_CASE(5) IB200_BOX_CASE(6) IB200_BOX_CASE(7)
    }
#undef IB200_BOX_CASE
    return 0;
}

}  // namespace

// `acc` is the float32 accumulation target (the output itself for F32, the scratch volume for 16-bit storage),
// already zero-filled by the caller.  Returns 1 when handled, 0 when not applicable (the caller falls back to
// push_tile.cu), < 0 on error.
int try_push_box(int op, const KParams &kp, int dtype, const void *img, const void *grid, void *acc, cudaStream_t stream) {
    if (!push_tiled_applicable(op, kp, dtype)) return 0;
    if (kp.flags & IB200_FLAG_NO_PIPE) return 0;         // A/B switch: the round-1 tile kernel
    const int es = dtype == IB200_F32 ? 4 : 2;
    // TMA: 16-byte aligned bases and row strides, dense lattice image
    if ((uintptr_t)grid % 16 || (op == OP_PUSH && (uintptr_t)img % 16)) return 0;
    if ((kp.pts_n[2] * 3 * es) % 16 || (kp.pts_n[2] * es) % 16) return 0;
    if ((kp.grid_sb * es) % 16 || (kp.img_sb * es) % 16 || (kp.img_sc * es) % 16) return 0;
    if (kp.grid_sb < 0 || kp.img_sb < 0 || kp.img_sc < 0) return 0;
    float *out = (float *)acc;
    if (op == OP_PUSH) {
        if (dtype == IB200_F32) return dispatch_push_box<float, OP_PUSH>(kp, img, grid, out, stream);
        return dispatch_push_box<__half, OP_PUSH>(kp, img, grid, out, stream);
    }
    if (dtype == IB200_F32) return dispatch_push_box<float, OP_COUNT>(kp, img, grid, out, stream);
    return dispatch_push_box<__half, OP_COUNT>(kp, img, grid, out, stream);
}

}  // namespace ib200
