// Persistent warp-specialised pull / grad (3-D, isotropic compile-time order,
// float32 storage): the tiled gather of pull_tile.cu restructured so that the tap
// loop never waits for memory.
//
//   grid = one CTA per SM, each looping over tiles of 8 x 8 x 32 output voxels; tiles beyond a CTA's first
//   are claimed from a global counter (a CTA that draws expensive tiles draws fewer).
//   warps 0 .. NCW-1 (consumers)
//     - wait on the box's mbarrier, (rarely) fix up folded elements, then evaluate
//       (ORDER+1)^3 LDS taps per voxel and store the result; rows are 64 words so
//       that the bank of a tap depends on z only; the 64 z-rows of a tile are dealt to
//       the warps statically, rotated from box to box.
//   warp NCW (producer)
//     - claims tiles two ahead of the taps and streams their grid coordinates into a
//       4-deep ring, one TMA tile copy (cp.async.bulk.tensor, box 8 x 8 x 96 floats) per tile;
//     - turns the bounding box of a tile (from the scout) into a plan: whole tile, z halves or
//       z quarters, plain / folded box, or the global fallback;
//     - issues one TMA tile copy per x-plane of the box of the input volume (box
//       16 rows x 64 words) into a 2-deep ring of boxes; the TMA unit zero-fills
//       whatever lies outside the volume, which IS the `zero` boundary condition;
//       for the other bounds a fix-up pass rewrites the out-of-volume elements.
//   warp NCW+1 (scout)
//     - reduces the bounding box of all spline supports of every tile as soon as its
//       coordinates have landed (REDUX + shared atomics).
//   Boxes that do not fit (incoherent deformation) are gathered from global
//   memory by the consumers with identical arithmetic.
//
// Replaces interpol/nd.py:81-143 / :217-288 (and iso1.py) for the shapes that
// matter for throughput; pull_tile.cu / gather.cu cover the rest.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <map>
#include <mutex>
#include <utility>
#include "pipe_common.cuh"

namespace ib200 {

// Phase timers (clock64 ticks, accumulated in registers by the producer warp and consumer warp 0 of
// CTA 0, written once at kernel exit) -- filled only when IB200_PIPE_DEBUG is set in the environment.
__device__ long long g_pipe_dbg[16];
__device__ long long g_pipe_cta[256 * 4];   // per CTA: start / end of consumer warp 0 (globaltimer ns), its loop ticks, its tiles
// (compiled in with -DIB200_PIPE_TIMERS only -- IB200_TUNE builds, profiles/pipe_debug.py: the clock reads and
// their predicated adds were ~30 instructions per warp and tile on the consumers' path)
#if defined(IB200_TUNE) && !defined(IB200_PIPE_TIMERS)
#define IB200_PIPE_TIMERS
#endif
#ifdef IB200_PIPE_TIMERS
#define TICK(var) const long long var = dbg ? clock64() : 0
#define TOCK(acc, var) do { if (dbg) acc += clock64() - (var); } while (0)
#else
#define TICK(var) do { } while (0)
#define TOCK(acc, var) do { } while (0)
#endif

__device__ __forceinline__ void st_global(float *p, float v) {
    asm volatile("st.global.f32 [%0], %1;\n" ::"l"(p), "f"(v) : "memory");
}

constexpr int kNG = 4;     // ring of grid-coordinate tiles
constexpr int kNB = 2;     // ring of boxes
constexpr int kTileCtrSlots = 1024;

template <int ORDER, int OP, int W>
__device__ __noinline__ float3 pull_point_global(const KParams &kp, const float *src, float c0, float c1, float c2) {
    constexpr bool GRAD = (OP == OP_GRAD || OP == OP_PULL_BWD_GRID);
    const float cc[3] = {c0, c1, c2};
    float acc = 0.f, ag[3] = {0.f, 0.f, 0.f};
    if (inbounds<float, 3>(kp, cc)) {
        Axis<float, W> ax[3];
        bool ok = setup_axis<float, ORDER, GRAD ? 1 : 0, W>(ax[0], cc[0], ORDER, kp.bound[0], kp.vol_n[0], (int)kp.vol_s[0], kp);
        ok = setup_axis<float, ORDER, GRAD ? 1 : 0, W>(ax[1], cc[1], ORDER, kp.bound[1], kp.vol_n[1], (int)kp.vol_s[1], kp) && ok;
        ok = setup_axis<float, ORDER, GRAD ? 1 : 0, W>(ax[2], cc[2], ORDER, kp.bound[2], kp.vol_n[2], (int)kp.vol_s[2], kp) && ok;
        if (ok) {
#pragma unroll 1
            for (int i = 0; i < W; ++i) {
                float s00 = 0.f, s10 = 0.f, s01 = 0.f;
#pragma unroll 1
                for (int j = 0; j < W; ++j) {
                    float t0 = 0.f, t1 = 0.f;
#pragma unroll
                    for (int k = 0; k < W; ++k) {
                        const float v = __ldg(src + ax[0].off[i] + ax[1].off[j] + ax[2].off[k]);
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
    return GRAD ? make_float3(ag[0], ag[1], ag[2]) : make_float3(acc, 0.f, 0.f);
}

template <int ORDER, int OP, int NCW>
__global__ void __launch_bounds__(32 * (NCW + 2), 1)
pull_pipe3d_kernel(const __grid_constant__ KParams kp, const __grid_constant__ CUtensorMap tm_vol,
                   const __grid_constant__ CUtensorMap tm_grid, const float *__restrict__ vol,
                   const float *__restrict__ gout, float *__restrict__ out, const int ntiles, const int cmul, const int bmul, const int gbmul,
                   const unsigned inv_ntz, const unsigned inv_nty, const unsigned inv_ntx, int *tile_ctr, long long *dbg) {
    constexpr int TX = 8, TY = 8, TZ = 32;
    constexpr int NPT = TX * TY * TZ;
    constexpr int NROWS = TX * TY;
    constexpr int W = ORDER + 1;
    constexpr bool GRAD = (OP == OP_GRAD || OP == OP_PULL_BWD_GRID);
    constexpr bool BWD = (OP == OP_PULL_BWD_GRID);     // fused backward w.r.t. the grid: grad * gout (one channel)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float *box = reinterpret_cast<float *>(smem_raw);                                   // [kNB][kBoxWords]
    float *gtile = box + (size_t)kNB * kBoxWords;                                       // [kNG][NPT * 3]
    PipeGeom *geoms = reinterpret_cast<PipeGeom *>(gtile + (size_t)kNG * NPT * 3);      // [kNB]
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(geoms + kNB);
    unsigned long long *gfull = bars, *gempty = bars + kNG, *bfull = bars + 2 * kNG, *bempty = bars + 2 * kNG + kNB;
    unsigned long long *kfull = bars + 2 * kNG + 2 * kNB;                               // [2] box of a tile reduced
    int *fixctr = reinterpret_cast<int *>(bars + 2 * kNG + 2 * kNB + 2) + kNB;                                                                 // [kNB] next unclaimed fix-up plane
    int *fixdone = fixctr + kNB;                                                        // [kNB] fixed planes
    int *keys_base = fixdone + kNB;                                                      // [2][6] boxes of tiles j, j+1 (+2 pad each)
    int *qkeys = keys_base + 16;                                                        // [24] quarter boxes (rare)
    PipeGeom *planned = reinterpret_cast<PipeGeom *>(qkeys + 24);                       // [4] parts of the planned tile
    int *zoff = reinterpret_cast<int *>(planned + 4);                                   // [kNB][kZLut] folded z word of an unfolded index
    float *zsgn = reinterpret_cast<float *>(zoff + kNB * kZLut);                        // [kNB][kZLut] its sign
    int *tile_ring = reinterpret_cast<int *>(zsgn + kNB * kZLut);                       // [kNG] tile held by a coordinate slot (-1: none left)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kNG; ++i) { mbar_init(gfull + i, 1); mbar_init(gempty + i, NCW); }
        for (int i = 0; i < kNB; ++i) { mbar_init(bfull + i, 1); mbar_init(bempty + i, NCW); }
        mbar_init(kfull, 1); mbar_init(kfull + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        tma_prefetch_desc(&tm_vol);
        tma_prefetch_desc(&tm_grid);
    }
    if (threadIdx.x < 16) keys_base[threadIdx.x] = pipe_key_init(threadIdx.x);
    __syncthreads();

    const int ntx = (kp.pts_n[0] + TX - 1) / TX, nty = (kp.pts_n[1] + TY - 1) / TY, ntz = (kp.pts_n[2] + TZ - 1) / TZ;
    const int C = (int)kp.channels;
    // Tiles are handed out in ascending order: the first one is blockIdx.x, every further one is claimed from a
    // global counter by the producer when it requests the tile's coordinates (two tiles ahead of the taps), so
    // a CTA that drew expensive tiles (stretched or boundary regions cost up to 2x) simply draws fewer of them
    // -- with the static round-robin the slowest CTA ran 25 % longer than the average one.  Neighbouring CTAs
    // still work on neighbouring tiles at the same time (halos shared through L2).  tile_ctr == nullptr (stream
    // capture): static round-robin.  The tile of sequence number q travels with its coordinate slot
    // (tile_ring[q % kNG], published by the gfull barrier); -1 ends the sequence for every role.
    auto decode = [&](int t, int &b, int &x0, int &y0, int &z0, int &nzv) {
        const int t1 = fast_div(t, inv_ntz), t2 = fast_div(t1, inv_nty), t3 = fast_div(t2, inv_ntx);
        b = t3; x0 = (t2 - t3 * ntx) * TX; y0 = (t1 - t2 * nty) * TY; z0 = (t - t1 * ntz) * TZ;
        nzv = min(TZ, kp.pts_n[2] - z0);
    };

    // Schedule.  The box of tile j is reduced by the consumers while they process tile j-2 (its
    // coordinates arrive one tile earlier still), turned into a plan + TMA requests by the producer
    // when tile j-2 releases its buffer, and lands while the consumers process tile j-1: nothing on
    // the consumers' path ever waits for memory or for the (slow, single-warp) producer.
    if (warp == NCW) {
        // ================================ producer ================================
        bool ended = false;                          // the sentinel has been published
        // claim(q): lane 0 draws the tile of sequence number q (the global atomic is issued early, its latency
        // hides behind the planning of tile q - 2); publish(q, t): coordinates requested, tile id in the ring
        auto claim = [&](int q) {
            int t = -1;
            if (lane == 0 && !ended) {
                if (q == 0) t = (int)blockIdx.x;
                else if (tile_ctr) t = (int)gridDim.x + atomicAdd(tile_ctr, 1);
                else t = (int)blockIdx.x + q * (int)gridDim.x;
                if (t >= ntiles) t = -1;
            }
            return t;
        };
        auto publish = [&](int q, int t) {
            if (ended) return;
            const int s = q % kNG, u = q / kNG;
            t = __shfl_sync(0xffffffffu, t, 0);
            if (u > 0) mbar_wait(gempty + s, (u - 1) & 1);
            if (lane == 0) {
                tile_ring[s] = t;
                if (t >= 0) {
                    int b, x0, y0, z0, nzv;
                    decode(t, b, x0, y0, z0, nzv);
                    mbar_expect_tx(gfull + s, NPT * 3 * 4);
                    tma_load_4d(gtile + (size_t)s * NPT * 3, &tm_grid, z0 * 3, y0, x0, b * gbmul, gfull + s);
                }
                mbar_arrive(gfull + s);
            }
            __syncwarp();
            ended = t < 0;
        };
        publish(0, claim(0));
        publish(1, claim(1));
        int n = 0;                                   // box sequence number
#ifdef IB200_PIPE_TIMERS
        long long a_kfull = 0, a_plan = 0, a_grid = 0, a_bempty = 0, a_issue = 0;
#endif
        for (int j = 0;; ++j) {
            const int t = tile_ring[j % kNG];        // written by lane 0 of this warp (request_grid ends with __syncwarp)
            if (t < 0) break;
            int b, x0, y0, z0, nzv;
            decode(t, b, x0, y0, z0, nzv);
            const int t_ahead = claim(j + 2);
            // ---- plan: whole tile, z halves or z quarters ----
            TICK(t_k);
            mbar_wait(kfull + (j & 1), (j >> 1) & 1);
            TOCK(a_kfull, t_k);
            TICK(t_p);
            int nparts = 1;
            {
                PipeGeom g0 = pipe_geom<ORDER>(kp, keys_base + (j & 1) * 8, 0, 1);
                g0.zlo = 0; g0.zhi = TZ; g0.last = 1;
                if (g0.mode != PIPE_GLOBAL) {
                    if (lane == 0) planned[0] = g0;
                } else {
                    const int nxv = min(TX, kp.pts_n[0] - x0), nyv = min(TY, kp.pts_n[1] - y0);
                    pipe_quarter_boxes<TX, TY, TZ>(kp, gtile + (size_t)(j % kNG) * NPT * 3, nxv, nyv, nzv, qkeys);
                    g0 = pipe_geom<ORDER>(kp, qkeys, 0, 2);
                    const PipeGeom g1 = pipe_geom<ORDER>(kp, qkeys, 2, 4);
                    if (g0.mode != PIPE_GLOBAL && g1.mode != PIPE_GLOBAL) {
                        nparts = 2;
                        if (lane == 0) { planned[0] = g0; planned[1] = g1; }
                    } else {
                        nparts = 4;
                        for (int p = 0; p < 4; ++p) {
                            const PipeGeom gq = pipe_geom<ORDER>(kp, qkeys, p, p + 1);
                            if (lane == 0) planned[p] = gq;
                        }
                    }
                }
                __syncwarp();
                if (lane < 6) keys_base[(j & 1) * 8 + lane] = pipe_key_init(lane);      // ready for tile j + 2
                __syncwarp();
            }
            TOCK(a_plan, t_p);
            // ---- one box per (channel, part) ----
            for (int c = 0; c < C; ++c) {
                for (int part = 0; part < nparts; ++part, ++n) {
                    const PipeGeom g = planned[part];
                    const int s = n % kNB, u = n / kNB;
                    TICK(t_b);
                    if (u > 0) mbar_wait(bempty + s, (u - 1) & 1);
                    TOCK(a_bempty, t_b);
                    TICK(t_i);
                    if (lane == 0) {
                        geoms[s] = g;
                        fixctr[s] = 0; fixdone[s] = 0;
                    }
                    if (g.zfold) {
                        for (int e = lane; e < g.zn; e += 32) {
                            zoff[s * kZLut + e] = bound_index<int>(kp.bound[2], g.za + e, kp.vol_n[2]) - g.lo[2];
                            zsgn[s * kZLut + e] = (float)bound_sign<int>(kp.bound[2], g.za + e, kp.vol_n[2]);
                        }
                    }
                    if (lane == 0) {
                        if (g.mode == PIPE_PLAIN || g.mode == PIPE_FOLD) mbar_expect_tx(bfull + s, (unsigned)g.ext[0] * kBoxPlane * 4u);
                    }
                    __syncwarp();
                    if ((g.mode == PIPE_PLAIN || g.mode == PIPE_FOLD) && lane < g.ext[0]) {
                        const int px = g.xfold ? bound_index<int>(kp.bound[0], g.lo[0] + lane, kp.vol_n[0]) : g.lo[0] + lane;
                        tma_load_5d(box + (size_t)s * kBoxWords + lane * kBoxPlane, &tm_vol, g.lo[2], g.lo[1], px,
                                    c * cmul, b * bmul, bfull + s);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bfull + s);
                    TOCK(a_issue, t_i);
                }
            }
            // coordinates of tile j + 2 (their buffer was released with tile j - 2, like the box just refilled):
            // after the boxes, which are what the consumers are about to wait for
            TICK(t_g);
            publish(j + 2, t_ahead);
            TOCK(a_grid, t_g);
        }
#ifdef IB200_PIPE_TIMERS
        if (dbg && blockIdx.x == 0 && lane == 0) { dbg[0] = a_kfull; dbg[1] = a_plan; dbg[2] = a_grid; dbg[3] = a_bempty; dbg[4] = a_issue; }
#endif
    } else if (warp == NCW + 1) {
        // ================================= scout ==================================
        // reduces the bounding box of the support starts of every tile as soon as its coordinates have
        // landed (two tiles ahead of the consumers), so the tap loop carries no look-ahead state
        for (int q = 0;; ++q) {
            mbar_wait(gfull + q % kNG, (q / kNG) & 1);
            const int t = tile_ring[q % kNG];
            if (t < 0) break;
            int b, x0, y0, z0, nzv;
            decode(t, b, x0, y0, z0, nzv);
            pipe_tile_box<TX, TY, TZ, 1>(kp, gtile + (size_t)(q % kNG) * NPT * 3, min(TX, kp.pts_n[0] - x0), min(TY, kp.pts_n[1] - y0),
                                         nzv, keys_base + (q & 1) * 8, 0);
            __syncwarp();
            if (lane == 0) mbar_arrive(kfull + (q & 1));
        }
    } else {
        // ================================ consumers ===============================
        int n = 0;
#ifdef IB200_PIPE_TIMERS
        long long a_gfull = 0, a_bfull = 0, a_fix = 0, a_rows = 0, a_rel = 0, a_items = 0;
#endif
        TICK(t_all);
#ifdef IB200_PIPE_TIMERS
        long long gt0 = 0;
        int ntl = 0;
        if (dbg) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt0));
#endif
        for (int q = 0;; ++q) {
            mbar_wait(gfull + q % kNG, (q / kNG) & 1);      // landed long ago (the scout has been through it): acquire only
            const int t = tile_ring[q % kNG];
            if (t < 0) break;
            int b, x0, y0, z0, nzv;
            decode(t, b, x0, y0, z0, nzv);
#ifdef IB200_PIPE_TIMERS
            ++ntl;
#endif
            const int nxv = min(TX, kp.pts_n[0] - x0), nyv = min(TY, kp.pts_n[1] - y0);
            const float *gt = gtile + (size_t)(q % kNG) * NPT * 3;
            for (int c = 0; c < C; ++c) {
                const float *src = vol + (i64)b * kp.vol_sb + (i64)c * kp.vol_sc;
                // lattice offset of this lane's voxel in row 0 of the tile (32-bit: pts_total * 3 < 2^31)
                const int o0 = (x0 * kp.pts_n[1] + y0) * kp.pts_n[2] + z0 + lane;
                float *dst = out + ((i64)b * kp.channels + c) * kp.pts_total * (GRAD ? 3 : 1) + (GRAD ? 3 * o0 : o0);
                const float *gm = BWD ? gout + (i64)b * kp.img_sb + (i64)c * kp.img_sc + o0 : nullptr;
                int ostride_y = kp.pts_n[2], ostride_x = kp.pts_n[1] * kp.pts_n[2];
                // kept in registers: the compiler otherwise re-reads them from the constant bank on every row's store
                // path and the shared-window base through S2R at every row start (2-3 % of the stall samples each)
                unsigned gt_s = smem_u32(gt);
                asm volatile("" : "+r"(ostride_y), "+r"(ostride_x), "+r"(gt_s), "+l"(dst));
                bool last;
                do {
                    const int s = n % kNB;
                    TICK(t_b);
                    mbar_wait(bfull + s, (n / kNB) & 1);
                    TOCK(a_bfull, t_b);
                    TICK(t_f);
                    const PipeGeom g = geoms[s];
                    last = g.last != 0;
                    float *bx = box + (size_t)s * kBoxWords;
                    if (g.mode == PIPE_FOLD) {
                        // x-planes are fixed by whichever warps get here first; nobody waits for a late warp,
                        // only for the planes still being fixed
                        for (;;) {
                            int a = 0;
                            if (lane == 0) a = atomicAdd(fixctr + s, 1);
                            a = __shfl_sync(0xffffffffu, a, 0);
                            if (a >= g.ext[0]) break;
                            pipe_fixup_plane(kp, g, bx, src, a);
                            __syncwarp();
                            if (lane == 0) { __threadfence_block(); atomicAdd(fixdone + s, 1); }
                        }
                        while (*reinterpret_cast<volatile int *>(fixdone + s) < g.ext[0]) {}
                        __threadfence_block();
                    }
                    TOCK(a_fix, t_f);
                    TICK(t_r);
                    // ---- taps ----
                    const bool lane_ok = lane < nzv && lane >= g.zlo && lane < g.zhi;
                    // z-rows are dealt statically, rotated from box to box so that the warps that get the extra
                    // row (NROWS % NCW of them) change every time: claiming rows from a shared counter (round 1)
                    // balanced perfectly but cost an ATOMS round trip and a SHFL per row -- measured 5 % slower (profiles/r2zi)
                    int r = (warp + (n % NCW) * (NCW - NROWS % NCW)) % NCW;
                    while (r < NROWS) {
                        const int rn = r + NCW;
                        const int p = r / TY, ly = r - p * TY;
                        if (p < nxv && ly < nyv && lane_ok) {
                            float cc[3];
                            {
                                const unsigned ga = gt_s + (unsigned)((r * TZ + lane) * 12);
                                asm volatile("ld.shared.f32 %0, [%3];\n\tld.shared.f32 %1, [%3+4];\n\tld.shared.f32 %2, [%3+8];\n"
                                             : "=f"(cc[0]), "=f"(cc[1]), "=f"(cc[2]) : "r"(ga));
                            }
                            float res[3] = {0.f, 0.f, 0.f};
                            if (g.mode == PIPE_GLOBAL) {
                                const float3 r3 = pull_point_global<ORDER, OP, W>(kp, src, cc[0], cc[1], cc[2]);
                                res[0] = r3.x; res[1] = r3.y; res[2] = r3.z;
                            } else if (g.mode != PIPE_EMPTY) {
                                const float f0 = floorf(cc[0] - 0.5f * (ORDER - 1)), f1 = floorf(cc[1] - 0.5f * (ORDER - 1)),
                                            f2 = floorf(cc[2] - 0.5f * (ORDER - 1));
                                const bool actp = inbounds<float, 3>(kp, cc) && fabsf(f0) < 4e18f && fabsf(f1) < 4e18f && fabsf(f2) < 4e18f;
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
                                    const float *rxy = bx + ((int)f0 - g.lo[0]) * kBoxPlane + ((int)f1 - g.lo[1]) * kBoxZ;
                                    const float *rk[W];
                                    if (g.zfold) {
                                        const int e0 = s * kZLut + (int)f2 - g.za;
#pragma unroll
                                        for (int k = 0; k < W; ++k) {
                                            rk[k] = rxy + zoff[e0 + k];
                                            const float sg = zsgn[e0 + k];
                                            wz[k] *= sg;
                                            if (GRAD) gz[k] *= sg;
                                        }
                                    } else {
#pragma unroll
                                        for (int k = 0; k < W; ++k) rk[k] = rxy + ((int)f2 - g.lo[2]) + k;
                                    }
                                    float acc = 0.f, ax_ = 0.f, ay_ = 0.f, az_ = 0.f;
                                    if constexpr (!GRAD && W % 2 == 0) {
                                        // packed FFMA2 (sm_100): rows j, j+1 of a plane share the z weights, so each
                                        // instruction advances two row sums; 11 instructions per plane instead of 21
                                        float2 wz2[W], acc2 = make_float2(0.f, 0.f);
#pragma unroll
                                        for (int k = 0; k < W; ++k) wz2[k] = make_float2(wz[k], wz[k]);
#pragma unroll
                                        for (int i = 0; i < W; ++i) {
                                            float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
                                            for (int jj = 0; jj < W; jj += 2) {
                                                float2 t2 = make_float2(0.f, 0.f);
#pragma unroll
                                                for (int k = 0; k < W; ++k)
                                                    t2 = __ffma2_rn(wz2[k], make_float2(rk[k][i * kBoxPlane + jj * kBoxZ], rk[k][i * kBoxPlane + (jj + 1) * kBoxZ]), t2);
                                                s2 = __ffma2_rn(make_float2(wy[jj], wy[jj + 1]), t2, s2);
                                            }
                                            acc2 = __ffma2_rn(make_float2(wx[i], wx[i]), s2, acc2);
                                        }
                                        acc = acc2.x + acc2.y;
                                    } else if constexpr (GRAD) {
                                        // packed FFMA2: (value, d/dz) share the tap, (value, d/dy) share the row sum
                                        float2 wgz[W];
#pragma unroll
                                        for (int k = 0; k < W; ++k) wgz[k] = make_float2(wz[k], gz[k]);
#pragma unroll
                                        for (int i = 0; i < W; ++i) {
                                            float2 s2 = make_float2(0.f, 0.f);      // (s00, s10)
                                            float s01 = 0.f;
#pragma unroll
                                            for (int jj = 0; jj < W; ++jj) {
                                                float2 t2 = make_float2(0.f, 0.f);  // (t0, t1)
#pragma unroll
                                                for (int k = 0; k < W; ++k) {
                                                    const float v = rk[k][i * kBoxPlane + jj * kBoxZ];
                                                    t2 = __ffma2_rn(wgz[k], make_float2(v, v), t2);
                                                }
                                                s2 = __ffma2_rn(make_float2(wy[jj], gy[jj]), make_float2(t2.x, t2.x), s2);
                                                s01 = fmaf(wy[jj], t2.y, s01);
                                            }
                                            ax_ = fmaf(gx[i], s2.x, ax_); ay_ = fmaf(wx[i], s2.y, ay_); az_ = fmaf(wx[i], s01, az_);
                                        }
                                    } else {
#pragma unroll
                                    for (int i = 0; i < W; ++i) {
                                        float s00 = 0.f, s10 = 0.f, s01 = 0.f;
#pragma unroll
                                        for (int jj = 0; jj < W; ++jj) {
                                            float t0 = 0.f, t1 = 0.f;
#pragma unroll
                                            for (int k = 0; k < W; ++k) {
                                                const float v = rk[k][i * kBoxPlane + jj * kBoxZ];
                                                t0 = fmaf(wz[k], v, t0);
                                                if (GRAD) t1 = fmaf(gz[k], v, t1);
                                            }
                                            s00 = fmaf(wy[jj], t0, s00);
                                            if (GRAD) { s10 = fmaf(gy[jj], t0, s10); s01 = fmaf(wy[jj], t1, s01); }
                                        }
                                        if (!GRAD) acc = fmaf(wx[i], s00, acc);
                                        else { ax_ = fmaf(gx[i], s00, ax_); ay_ = fmaf(wx[i], s10, ay_); az_ = fmaf(wx[i], s01, az_); }
                                    }
                                    }
                                    if (!GRAD) res[0] = acc;
                                    else { res[0] = ax_; res[1] = ay_; res[2] = az_; }
                                }
                            }
                            const int o = p * ostride_x + ly * ostride_y;
                            if (BWD) { const float m = gm[o]; res[0] *= m; res[1] *= m; res[2] *= m; }
                            // (explicit global stores: `dst` went through an opaque register copy above)
                            if (!GRAD) st_global(dst + o, res[0]);
                            else { st_global(dst + o * 3, res[0]); st_global(dst + o * 3 + 1, res[1]); st_global(dst + o * 3 + 2, res[2]); }
                        }
                        r = rn;
                    }
                    TOCK(a_rows, t_r);
                    TICK(t_e);
                    // ---- release the box (and, after the last part of the last channel, the coordinates) ----
                    if (g.mode == PIPE_FOLD) fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(bempty + s);
                        if (last && c == C - 1) mbar_arrive(gempty + q % kNG);
                    }
                    ++n;
                    TOCK(a_rel, t_e);
#ifdef IB200_PIPE_TIMERS
                    a_items += 1;
#endif
                } while (!last);
            }
        }
#ifdef IB200_PIPE_TIMERS
        if (dbg && warp == 0 && lane == 0 && blockIdx.x < 256) {
            long long gt1;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt1));
            g_pipe_cta[blockIdx.x * 4] = gt0; g_pipe_cta[blockIdx.x * 4 + 1] = gt1;
            g_pipe_cta[blockIdx.x * 4 + 2] = clock64() - t_all; g_pipe_cta[blockIdx.x * 4 + 3] = ntl;
        }
        if (dbg && blockIdx.x == 0 && warp == 0 && lane == 0) {
            dbg[5] = a_gfull; dbg[6] = a_bfull; dbg[7] = a_fix; dbg[8] = a_rows; dbg[9] = a_rel; dbg[10] = a_items;
            dbg[11] = clock64() - t_all;
        }
#endif
    }
}

// ---------------------------------------------------------------- launch --

// Tile counters of the dynamic schedule: one word per stream in flight (launches on one stream are ordered, so
// they can share a word; it is zeroed in-stream before every launch).  A stream under capture gets none (static
// round-robin): a captured node would carry its word into replays on other streams.
__device__ int g_tile_ctr[kTileCtrSlots];

static int *g_last_ctl = nullptr;    // counter of the last launch (debugging aid, ib200_debug_pipe_control)

static int *pipe_control_for(cudaStream_t stream) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) { cudaGetLastError(); return nullptr; }
    if (getenv("IB200_STATIC_TILES")) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, int> slots;
    static std::map<int, int *> base;
    std::lock_guard<std::mutex> lock(mu);
    int *b = nullptr;
    auto ib = base.find(dev);
    if (ib == base.end()) {
        if (cudaGetSymbolAddress((void **)&b, g_tile_ctr) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        base[dev] = b;
    } else b = ib->second;
    auto key = std::make_pair(dev, stream);
    auto it = slots.find(key);
    if (it == slots.end()) {
        if ((int)slots.size() >= kTileCtrSlots) return nullptr;      // more live streams than words: static schedule
        it = slots.emplace(key, (int)slots.size()).first;
    }
    int *ctr = b + it->second;
    if (cudaMemsetAsync(ctr, 0, sizeof(int), stream) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return ctr;
}

template <int ORDER, int OP, int NCW>
static int launch_pull_pipe(const KParams &kp, const float *vol, const float *grid, const float *gout, float *out, cudaStream_t stream) {
    constexpr int NPT = 8 * 8 * 32;
    const size_t smem_total = (size_t)kNB * kBoxWords * 4 + (size_t)kNG * NPT * 3 * 4 + kNB * sizeof(PipeGeom) +
                              (2 * kNG + 2 * kNB + 2) * sizeof(unsigned long long) + (3 * kNB + 40 + kNG) * sizeof(int) + 4 * sizeof(PipeGeom) + (size_t)kNB * kZLut * 8 + 64;
    const i64 ntiles = kp.batch * ((kp.pts_n[0] + 7) / 8) * ((kp.pts_n[1] + 7) / 8) * ((kp.pts_n[2] + 31) / 32);
    if (ntiles == 0) return 1;
    if (ntiles * kp.channels > 0x3fffffffLL) return 0;
    // the tile decode divides by multiply-high: exact only while tile * divisor < 2^32
    {
        const i64 dmax = std::max<i64>((kp.pts_n[2] + 31) / 32, std::max<i64>((kp.pts_n[1] + 7) / 8, (kp.pts_n[0] + 7) / 8));
        if (ntiles * dmax >= (1LL << 32)) return 0;
    }
    // tensor maps: volume (z, y, x, c, b), box {64, 16, 1, 1, 1}; grid (z*3, y, x, b), box {96, 8, 8, 1}
    CUtensorMap tm_vol, tm_grid;
    const long long vbytes = (((long long)kp.vol_n[0] * kp.vol_s[0]) + 3) & ~3LL;
    const int cmul = (kp.vol_sc != 0 && kp.channels > 1) ? 1 : 0, bmul = (kp.vol_sb != 0 && kp.batch > 1) ? 1 : 0;
    const int gbmul = (kp.grid_sb != 0 && kp.batch > 1) ? 1 : 0;
    {
        const long long dim[5] = {kp.vol_n[2], kp.vol_n[1], kp.vol_n[0], cmul ? kp.channels : 1, bmul ? kp.batch : 1};
        const long long str[5] = {1, kp.vol_n[1] > 1 ? kp.vol_s[1] : kp.vol_n[2], kp.vol_n[0] > 1 ? kp.vol_s[0] : vbytes,
                                  cmul ? kp.vol_sc : vbytes, bmul ? kp.vol_sb : vbytes};
        const int box[5] = {kBoxZ, kBoxY, 1, 1, 1};
        if (!make_tensor_map(&tm_vol, vol, 5, dim, str, box)) return 0;
    }
    {
        const long long row = (long long)kp.pts_n[2] * 3;
        const long long dim[4] = {row, kp.pts_n[1], kp.pts_n[0], gbmul ? kp.batch : 1};
        const long long str[4] = {1, row, row * kp.pts_n[1], gbmul ? kp.grid_sb : row * kp.pts_n[1] * kp.pts_n[0]};
        const int box[4] = {96, 8, 8, 1};
        if (!make_tensor_map(&tm_grid, grid, 4, dim, str, box)) return 0;
    }
    auto kern = pull_pipe3d_kernel<ORDER, OP, NCW>;
    IB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_total));
    const int nblocks = (int)(ntiles < pipe_sm_count() ? ntiles : pipe_sm_count());
    long long *dbg = nullptr;
    if (getenv("IB200_PIPE_DEBUG")) IB200_CUDA_CHECK(cudaGetSymbolAddress((void **)&dbg, g_pipe_dbg));
    int *ctl = ntiles > nblocks ? pipe_control_for(stream) : nullptr;
    g_last_ctl = ctl;
    kern<<<nblocks, 32 * (NCW + 2), smem_total, stream>>>(kp, tm_vol, tm_grid, vol, gout, out, (int)ntiles, cmul, bmul, gbmul,
        make_inv((kp.pts_n[2] + 31) / 32), make_inv((kp.pts_n[1] + 7) / 8), make_inv((kp.pts_n[0] + 7) / 8), ctl, dbg);
    static thread_local char name[64];
    snprintf(name, sizeof(name), "%s_pipe3d_o%d", OP == OP_GRAD ? "grad" : OP == OP_PULL_BWD_GRID ? "pullbwd" : "pull", ORDER);
    note_launch(name);
    IB200_CUDA_CHECK(cudaGetLastError());
    return 1;
}

int try_pull_pipe(int op, const KParams &kp, int dtype, const void *vol, const void *grid, const void *gout_, void *out, cudaStream_t stream) {
    if (op != OP_PULL && op != OP_GRAD && op != OP_PULL_BWD_GRID) return 0;
    if (op == OP_PULL_BWD_GRID && (kp.channels != 1 || !gout_)) return 0;        // several channels: the tile kernel sums them
    if (dtype != IB200_F32) return 0;
    if (kp.dim != 3 || !kp.pts_dense) return 0;
    if (kp.order[0] != kp.order[1] || kp.order[0] != kp.order[2]) return 0;
    if (kp.order[0] < 1 || kp.order[0] > 3) return 0;
    if (kp.pts_total < 32768) return 0;
    if (kp.pts_total * 3 > 0x7fffffffLL) return 0;
    if (kp.flags & (IB200_FLAG_REF_LINEAR_GRAD_SIGN | IB200_FLAG_NO_PIPE | IB200_FLAG_DISPLACEMENT)) return 0;   // (displacement fields: tile kernel)
    // The pipeline needs a dozen or two tiles per CTA to amortise its ramp-up and its three-tile tail
    // (profiles/r2zl, benchmark deformation: 160^3 = 13.5 tiles per SM, cubic 0.195 ms against 0.138 ms for the
    // one-tile-per-CTA kernel, linear 0.058 against 0.065; 192^3 = 23 per SM: 0.151 / 0.170 and 0.077 / 0.106), and
    // with several channels the tile kernel, which shares one plan and one staged grid tile between the channels
    // of a tile, is still ahead (256^3 C=4: pull 1.05 vs 1.17 ms).  IB200_FLAG_FORCE_PIPE overrides (tests, profiling).
    if (!(kp.flags & IB200_FLAG_FORCE_PIPE)) {
        const i64 tiles = kp.batch * ((kp.pts_n[0] + 7) / 8) * ((kp.pts_n[1] + 7) / 8) * ((kp.pts_n[2] + 31) / 32);
        const i64 per_sm = kp.order[0] == 1 ? 10 : 18;
        if (tiles < per_sm * (i64)pipe_sm_count() || kp.channels > 1) return 0;
    }
    // TMA: unit innermost stride, 16-byte aligned bases and strides
    if (kp.vol_s[2] != 1) return 0;
    if ((uintptr_t)vol % 16 || (uintptr_t)grid % 16) return 0;
    if (kp.vol_s[0] % 4 || kp.vol_s[1] % 4 || kp.vol_sb % 4 || kp.vol_sc % 4) return 0;
    if (kp.pts_n[2] % 4 || kp.grid_sb % 4) return 0;
    const float *v = (const float *)vol, *g = (const float *)grid, *go = (const float *)gout_;
    float *o = (float *)out;
    constexpr int NCW = 14;     // (16 warps of 4 rows each need <= 96 registers: measured 4 % slower, profiles/r2zj)
    if (op == OP_PULL) {
        switch (kp.order[0]) {
        case 1: return launch_pull_pipe<1, OP_PULL, NCW>(kp, v, g, nullptr, o, stream);
        case 2: return launch_pull_pipe<2, OP_PULL, NCW>(kp, v, g, nullptr, o, stream);
        case 3: return launch_pull_pipe<3, OP_PULL, NCW>(kp, v, g, nullptr, o, stream);
        }
    } else if (op == OP_PULL_BWD_GRID) {
        switch (kp.order[0]) {
        case 1: return launch_pull_pipe<1, OP_PULL_BWD_GRID, NCW>(kp, v, g, go, o, stream);
        case 2: return launch_pull_pipe<2, OP_PULL_BWD_GRID, NCW>(kp, v, g, go, o, stream);
        case 3: return launch_pull_pipe<3, OP_PULL_BWD_GRID, NCW>(kp, v, g, go, o, stream);
        }
    } else {
        switch (kp.order[0]) {
        case 1: return launch_pull_pipe<1, OP_GRAD, NCW>(kp, v, g, nullptr, o, stream);
        case 2: return launch_pull_pipe<2, OP_GRAD, NCW>(kp, v, g, nullptr, o, stream);
        case 3: return launch_pull_pipe<3, OP_GRAD, NCW>(kp, v, g, nullptr, o, stream);
        }
    }
    return 0;
}

}  // namespace ib200

// debugging aid (not part of the public ABI): phase timers of the last pipe launch
// debugging aid (not part of the public ABI): tiles claimed from the counter by the last persistent pull / grad
// launch (including the failed claim that ends every CTA); -1 when that launch had no counter.  Synchronises.
extern "C" __attribute__((visibility("default"))) int ib200_debug_pipe_control(int *out1) {
    if (!ib200::g_last_ctl) return -1;
    if (cudaDeviceSynchronize() != cudaSuccess) return -2;
    return cudaMemcpy(out1, ib200::g_last_ctl, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
extern "C" __attribute__((visibility("default"))) int ib200_debug_pipe_cta(long long *out1024) {
    return cudaMemcpyFromSymbol(out1024, ib200::g_pipe_cta, sizeof(long long) * 1024) == cudaSuccess ? 0 : -1;
}
extern "C" __attribute__((visibility("default"))) int ib200_debug_pipe_counters(long long *out16) {
    return cudaMemcpyFromSymbol(out16, ib200::g_pipe_dbg, sizeof(long long) * 16) == cudaSuccess ? 0 : -1;
}
