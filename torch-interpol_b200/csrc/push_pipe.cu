// Persistent warp-specialised push / count (3-D, isotropic compile-time order,
// float32 storage): the adjoint of pull_pipe.cu, with the scatter privatised in a
// shared-memory box of 32-bit fixed-point accumulators (native integer ATOMS, see
// push_tile.cu for why) that is flushed by the TMA unit.
//
//   grid = one CTA per SM, each looping over tiles of 8 x 8 x 32 SOURCE voxels.
//   producer warp
//     - streams grid coordinates through a 4-deep TMA ring; turns the bounding box
//       the consumers reduced two tiles ahead into a plan (whole tile / z halves /
//       z quarters) and publishes the geometry of each (channel, part) item;
//     - reopens the box as soon as every consumer warp has flushed and re-zeroed its planes
//       (two CTAs per SM with one box each: the serial phases of one CTA hide behind the
//       atomics of the other).
//   consumer warps, per item (every step is claimed dynamically and closed by a
//   counter, never by a barrier: a late warp delays nobody)
//     Z. load the values of the warp's rows (kept in registers), reduce max |value|;
//     H. sum of |value| (quantised upwards) per cell of 4^3 support starts;
//     S. rigorous bound on any accumulator (a target voxel receives from 2x2x2 cells)
//        -> power-of-two fixed-point scale (every warp recomputes it: no extra hand-shake);
//     A. (ORDER+1)^3 integer atomics per source, z folded through a table for the
//        mirror-type bounds; on the side, the box of tile q+2 is reduced;
//     F. fold what sits outside the volume back inside (bounds the TMA clip / the
//        x-plane coordinate / the z table do not already cover);
//     C. per x-plane, by the warp that claims it: fixed -> float in place, ONE TMA float reduction into
//        the volume, wait until the plane has been read, zero it for the next item.
//
// Replaces interpol/nd.py:147-213 (and iso1.py push) for the shapes that matter
// for throughput; push_tile.cu / scatter.cu cover the rest.
#include <cstdio>
#include <cstdlib>
#include "pipe_common.cuh"

namespace ib200 {

__device__ long long g_push_dbg[16];   // phase timers, see pull_pipe.cu (IB200_PIPE_DEBUG)
#define PTICK(var) const long long var = dbg ? clock64() : 0
#define PTOCK(acc, var) do { if (dbg) acc += clock64() - (var); } while (0)

constexpr float kPMagic = 12582912.f;   // 1.5 * 2^23: float -> int by mantissa alignment
constexpr int kPMagicBits = 0x4B400000;
constexpr int kPBoxX = 15;              // x-planes per accumulator box (60 KB: leaves room for the cell counters)
constexpr int kPBoxWords = kPBoxX * kBoxPlane;
constexpr int kCell = 4;                // edge of a bound cell (>= ORDER + 1 for the orders served here)
constexpr int kCells = (kPBoxX / kCell + 2) * (kBoxY / kCell + 2) * (kBoxZ / kCell + 2);

__device__ __forceinline__ void tma_reduce_add_5d(const CUtensorMap *tm, const void *src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.reduce.async.bulk.tensor.5d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];\n"
                 ::"l"(tm), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}

// per-item control block shared by the consumers (reset by the producer when it opens the item)
struct PushCtl {
    int zero_next, zero_done;    // Z: planes
    int hist_next, hist_done;    // H: rows
    int bound_done;              // S: warps that contributed
    int cells_max;               //    max over cells of the 2x2x2 neighbourhood count
    unsigned vmax_bits;          //    max |value| over the tile (float bits)
    int row_next, row_done;      // A: rows
    int fold_next, fold_done;    // F: planes
    int conv_next;               // C: planes
};

template <int ORDER, int OP, int NCW, int kPNG, int kPNB, int CTAS>
__global__ void __launch_bounds__(32 * (NCW + 1), CTAS)
push_pipe3d_kernel(const __grid_constant__ KParams kp, const __grid_constant__ CUtensorMap tm_out,
                   const __grid_constant__ CUtensorMap tm_grid, const float *__restrict__ img,
                   float *__restrict__ out, const int ntiles, const int gbmul,
                   const unsigned inv_ntz, const unsigned inv_nty, const unsigned inv_ntx, long long *dbg) {
    constexpr int TX = 8, TY = 8, TZ = 32;
    constexpr int NPT = TX * TY * TZ;
    constexpr int NROWS = TX * TY;
    constexpr int W = ORDER + 1;
    constexpr bool COUNT = (OP == OP_COUNT);
    constexpr int URW = (NROWS + NCW - 1) / NCW;   // rows per consumer warp
    constexpr int LA = kPNG / 2;                   // tiles of look-ahead for the bounding boxes
    constexpr int QBITS = 18;                      // |value| quantisation of the bound: 2048 * 2^18 < 2^31
    static_assert(W <= kCell, "a target voxel must receive from at most 2 cells per axis");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    int *box = reinterpret_cast<int *>(smem_raw);                                       // [kPNB][kBoxWords] fixed point, then float
    float *gtile = reinterpret_cast<float *>(box + (size_t)kPNB * kPBoxWords);           // [kPNG][NPT * 3]
    PipeGeom *geoms = reinterpret_cast<PipeGeom *>(gtile + (size_t)kPNG * NPT * 3);     // [kPNB]
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(geoms + kPNB);
    unsigned long long *gfull = bars, *gempty = bars + kPNG, *bfull = bars + 2 * kPNG, *bdone = bars + 2 * kPNG + kPNB;
    unsigned long long *kfull = bars + 2 * kPNG + 2 * kPNB;                             // [2] box of a tile reduced
    PushCtl *ctl = reinterpret_cast<PushCtl *>(bars + 2 * kPNG + 2 * kPNB + 2);         // [kPNB]
    int *keys_base = reinterpret_cast<int *>(ctl + kPNB);                               // [2][8]
    int *qkeys = keys_base + 16;                                                        // [24]
    PipeGeom *planned = reinterpret_cast<PipeGeom *>(qkeys + 24);                       // [4]
    int *zoff = reinterpret_cast<int *>(planned + 4);                                   // [kPNB][kZLut]
    float *zsgn = reinterpret_cast<float *>(zoff + kPNB * kZLut);                       // [kPNB][kZLut]
    int *cells = reinterpret_cast<int *>(zsgn + kPNB * kZLut);                          // [kCells] sources per cell (one item at a time)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kPNG; ++i) { mbar_init(gfull + i, 1); mbar_init(gempty + i, NCW); }
        for (int i = 0; i < kPNB; ++i) { mbar_init(bfull + i, 1); mbar_init(bdone + i, NCW); }
        mbar_init(kfull, NCW); mbar_init(kfull + 1, NCW);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        tma_prefetch_desc(&tm_out);
        tma_prefetch_desc(&tm_grid);
    }
    if (threadIdx.x < 16) keys_base[threadIdx.x] = pipe_key_init(threadIdx.x);
    // boxes and cell sums start clean; afterwards whoever flushes a plane re-zeroes it
    for (int e = threadIdx.x; e < kPNB * kPBoxWords / 4; e += blockDim.x) reinterpret_cast<int4 *>(box)[e] = make_int4(0, 0, 0, 0);
    for (int e = threadIdx.x; e < kCells; e += blockDim.x) cells[e] = 0;
    __syncthreads();

    const int ntx = (kp.pts_n[0] + TX - 1) / TX, nty = (kp.pts_n[1] + TY - 1) / TY, ntz = (kp.pts_n[2] + TZ - 1) / TZ;
    const int C = (int)kp.channels;
    const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    auto decode = [&](int q, int &b, int &x0, int &y0, int &z0) {
        const int t = (int)blockIdx.x + q * (int)gridDim.x;
        const int t1 = fast_div(t, inv_ntz), t2 = fast_div(t1, inv_nty), t3 = fast_div(t2, inv_ntx);
        b = t3; x0 = (t2 - t3 * ntx) * TX; y0 = (t1 - t2 * nty) * TY; z0 = (t - t1 * ntz) * TZ;
    };

    if (warp == NCW) {
        // ================================ producer ================================
        auto request_grid = [&](int q) {
            int b, x0, y0, z0;
            decode(q, b, x0, y0, z0);
            const int s = q % kPNG, u = q / kPNG;
            if (u > 0) mbar_wait(gempty + s, (u - 1) & 1);
            if (lane == 0) {
                mbar_expect_tx(gfull + s, NPT * 3 * 4);
                tma_load_4d(gtile + (size_t)s * NPT * 3, &tm_grid, z0 * 3, y0, x0, b * gbmul, gfull + s);
                mbar_arrive(gfull + s);
            }
        };
        // item m is finished when every consumer warp has converted, flushed (and re-zeroed) its planes of the box
        long long p_wait = 0, p_issue = 0, p_read = 0, p_kfull = 0;
        auto flush = [&](int m, int fb, int fc) {
            const int s = m % kPNB;
            PTICK(t_w);
            mbar_wait(bdone + s, (m / kPNB) & 1);
            PTOCK(p_wait, t_w);
            (void)fb; (void)fc; (void)p_issue; (void)p_read;
        };
        for (int q = 0; q < LA && q < my_tiles; ++q) request_grid(q);
        int n = 0;                                   // item sequence number
        int pb = 0, pc = 0, ppb = 0, ppc = 0;        // (batch, channel) of items n-1 and n-2
        for (int j = 0; j < my_tiles; ++j) {
            int b, x0, y0, z0;
            decode(j, b, x0, y0, z0);
            PTICK(t_k);
            mbar_wait(kfull + (j & 1), (j >> 1) & 1);
            PTOCK(p_kfull, t_k);
            int nparts = 1;
            {
                PipeGeom g0 = pipe_geom<ORDER>(kp, keys_base + (j & 1) * 8, 0, 1, kPBoxX);
                g0.zlo = 0; g0.zhi = TZ; g0.last = 1;
                if (g0.mode != PIPE_GLOBAL) {
                    if (lane == 0) planned[0] = g0;
                } else {
                    const int nxv = min(TX, kp.pts_n[0] - x0), nyv = min(TY, kp.pts_n[1] - y0), nzv = min(TZ, kp.pts_n[2] - z0);
                    pipe_quarter_boxes<TX, TY, TZ>(kp, gtile + (size_t)(j % kPNG) * NPT * 3, nxv, nyv, nzv, qkeys);
                    g0 = pipe_geom<ORDER>(kp, qkeys, 0, 2, kPBoxX);
                    const PipeGeom g1 = pipe_geom<ORDER>(kp, qkeys, 2, 4, kPBoxX);
                    if (g0.mode != PIPE_GLOBAL && g1.mode != PIPE_GLOBAL) {
                        nparts = 2;
                        if (lane == 0) { planned[0] = g0; planned[1] = g1; }
                    } else {
                        nparts = 4;
                        for (int p = 0; p < 4; ++p) {
                            const PipeGeom gq = pipe_geom<ORDER>(kp, qkeys, p, p + 1, kPBoxX);
                            if (lane == 0) planned[p] = gq;
                        }
                    }
                }
                __syncwarp();
                if (lane < 6) keys_base[(j & 1) * 8 + lane] = pipe_key_init(lane);
                __syncwarp();
            }
            if (j + LA < my_tiles) request_grid(j + LA);
            for (int c = 0; c < C; ++c) {
                for (int part = 0; part < nparts; ++part, ++n) {
                    if (n >= kPNB) flush(n - kPNB, kPNB >= 2 ? ppb : pb, kPNB >= 2 ? ppc : pc);       // frees the slot of item n
                    const PipeGeom g = planned[part];
                    const int s = n % kPNB;
                    if (g.zfold) {
                        for (int e = lane; e < g.zn; e += 32) {
                            zoff[s * kZLut + e] = bound_index<int>(kp.bound[2], g.za + e, kp.vol_n[2]) - g.lo[2];
                            zsgn[s * kZLut + e] = (float)bound_sign<int>(kp.bound[2], g.za + e, kp.vol_n[2]);
                        }
                    }
                    if (lane == 0) {
                        geoms[s] = g;
                        PushCtl z;
                        z.zero_next = z.zero_done = z.hist_next = z.hist_done = z.bound_done = z.cells_max = 0;
                        z.vmax_bits = 0u; z.row_next = z.row_done = z.fold_next = z.fold_done = z.conv_next = 0;
                        ctl[s] = z;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bfull + s);
                    ppb = pb; ppc = pc; pb = b; pc = c;
                }
            }
        }
        // drain: the last items
        if (kPNB >= 2 && n >= 2) flush(n - 2, ppb, ppc);
        if (n >= 1) flush(n - 1, pb, pc);
        if (dbg && blockIdx.x == 0 && lane == 0) { dbg[0] = p_wait; dbg[1] = p_issue; dbg[2] = p_read; dbg[3] = p_kfull; }
    } else {
        // ================================ consumers ===============================
        const bool masked = kp.extrapolate != 1;
        const float w3 = max_weight(ORDER) * max_weight(ORDER) * max_weight(ORDER);
        for (int q = 0; q < LA && q < my_tiles; ++q) {      // boxes of the first tile(s), cooperatively
            int b, x0, y0, z0;
            decode(q, b, x0, y0, z0);
            mbar_wait(gfull + q % kPNG, (q / kPNG) & 1);
            pipe_tile_box<TX, TY, TZ, NCW>(kp, gtile + (size_t)(q % kPNG) * NPT * 3, min(TX, kp.pts_n[0] - x0), min(TY, kp.pts_n[1] - y0),
                                           min(TZ, kp.pts_n[2] - z0), keys_base + (q & 1) * 8, warp);
            __syncwarp();
            if (lane == 0) mbar_arrive(kfull + (q & 1));
        }
        auto spin_until = [&](const int *p, int target) {
            while (*reinterpret_cast<const volatile int *>(p) < target) {}
            __threadfence_block();
        };
        auto claim = [&](int *p) -> int {
            int v = 0;
            if (lane == 0) v = atomicAdd(p, 1);
            return __shfl_sync(0xffffffffu, v, 0);
        };
        auto finish = [&](int *p) {
            __syncwarp();
            if (lane == 0) { __threadfence_block(); atomicAdd(p, 1); }
        };
        int n = 0;
        long long c_bfull = 0, c_z = 0, c_h = 0, c_s = 0, c_a = 0, c_aw = 0, c_f = 0, c_c = 0, c_items = 0;
        PTICK(t_all);
        for (int q = 0; q < my_tiles; ++q) {
            int b, x0, y0, z0;
            decode(q, b, x0, y0, z0);
            const int nxv = min(TX, kp.pts_n[0] - x0), nyv = min(TY, kp.pts_n[1] - y0), nzv = min(TZ, kp.pts_n[2] - z0);
            const float *gt = gtile + (size_t)(q % kPNG) * NPT * 3;
            bool look = q + LA < my_tiles;
            int nxv2 = 0, nyv2 = 0, nzv2 = 0;
            const float *gt2 = gtile + (size_t)((q + LA) % kPNG) * NPT * 3 + lane * 3;
            float mn[3] = {3e38f, 3e38f, 3e38f}, mx[3] = {-3e38f, -3e38f, -3e38f};
            if (look) {
                int b2, x2, y2, z2;
                decode(q + LA, b2, x2, y2, z2);
                nxv2 = min(TX, kp.pts_n[0] - x2); nyv2 = min(TY, kp.pts_n[1] - y2); nzv2 = min(TZ, kp.pts_n[2] - z2);
            }
            bool look_waited = false;
            const int tile_off = (x0 * kp.pts_n[1] + y0) * kp.pts_n[2] + z0;
            for (int c = 0; c < C; ++c) {
                const float *src = COUNT ? nullptr : img + (i64)b * kp.img_sb + (i64)c * kp.img_sc + tile_off;
                float *dst = out + ((i64)b * kp.channels + c) * kp.vol_total;
                bool last;
                do {
                    const int s = n % kPNB;
                    PTICK(t0);
                    mbar_wait(bfull + s, (n / kPNB) & 1);
                    PTOCK(c_bfull, t0);
                    const PipeGeom g = geoms[s];
                    last = g.last != 0;
                    int *bx = box + (size_t)s * kPBoxWords;
                    PushCtl *ct = ctl + s;
                    const bool lane_ok = lane < nzv && lane >= g.zlo && lane < g.zhi;
                    const bool boxed = g.mode == PIPE_PLAIN || g.mode == PIPE_FOLD;
                    const int nc0 = g.ext[0] / kCell + 1, nc1 = g.ext[1] / kCell + 1, nc2 = g.ext[2] / kCell + 1;
                    auto row_of = [&](int r, float (&cc)[3], bool &act) {
                        const int p = r / TY, ly = r - p * TY;
                        act = p < nxv && ly < nyv && lane_ok;
                        const float *gp = gt + (r * TZ + lane) * 3;
                        cc[0] = gp[0]; cc[1] = gp[1]; cc[2] = gp[2];
                        const float f0 = floorf(cc[0] - 0.5f * (ORDER - 1)), f1 = floorf(cc[1] - 0.5f * (ORDER - 1)),
                                    f2 = floorf(cc[2] - 0.5f * (ORDER - 1));
                        act = act && inbounds<float, 3>(kp, cc) && fabsf(f0) < 4e18f && fabsf(f1) < 4e18f && fabsf(f2) < 4e18f;
                    };
                    float scale = 0.f;
                    float rval[URW];
#pragma unroll
                    for (int u = 0; u < URW; ++u) rval[u] = COUNT ? 1.f : 0.f;
                    PTICK(t1);
                    if (boxed) {
                        // ---- Z. values of this warp's rows (kept in registers), max |value| ----
                        float vm = 0.f;
#pragma unroll
                        for (int u = 0; u < URW; ++u) {
                            const int r = warp + u * NCW;
                            const int p = r / TY, ly = r - p * TY;
                            rval[u] = COUNT ? 1.f : 0.f;
                            if (!COUNT && r < NROWS && p < nxv && ly < nyv && lane_ok) rval[u] = __ldg(src + (p * kp.pts_n[1] + ly) * kp.pts_n[2] + lane);
                        }
#pragma unroll
                        for (int u = 0; u < URW; ++u) { const float av = fabsf(rval[u]); vm = fmaxf(vm, av < 3e38f ? av : 3e38f); }
                        {
                            const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(vm));
                            if (lane == 0 && m != 0u) atomicMax(&ct->vmax_bits, m);
                        }
                        finish(&ct->bound_done);
                        spin_until(&ct->bound_done, NCW);
                        const float vmax = __uint_as_float(*reinterpret_cast<volatile unsigned *>(&ct->vmax_bits));
                        PTOCK(c_z, t1);
                        PTICK(t2);
                        // ---- H. sum of |value| (quantised upwards) per cell of kCell^3 support starts ----
                        if (vmax > 0.f) {
                            const float qs = (float)(1 << QBITS) / vmax;
#pragma unroll
                            for (int u = 0; u < URW; ++u) {
                                const int r = warp + u * NCW;
                                if (r < NROWS) {
                                    float cc[3]; bool act;
                                    row_of(r, cc, act);
                                    if (act) {
                                        const int i0 = (int)floorf(cc[0] - 0.5f * (ORDER - 1)) - g.lo[0], i1 = (int)floorf(cc[1] - 0.5f * (ORDER - 1)) - g.lo[1];
                                        int i2 = (int)floorf(cc[2] - 0.5f * (ORDER - 1));
                                        if (g.zfold) {
                                            // folded z: the taps of one source stay within W consecutive folded words; cell of the lowest
                                            const int e0 = s * kZLut + i2 - g.za;
                                            int m = zoff[e0];
#pragma unroll
                                            for (int k = 1; k < W; ++k) m = min(m, zoff[e0 + k]);
                                            i2 = m;
                                        } else {
                                            i2 -= g.lo[2];
                                        }
                                        const int qv = min(1 << QBITS, (int)ceilf(fabsf(rval[u]) * qs));
                                        atomicAdd(&cells[((i0 / kCell) * nc1 + i1 / kCell) * nc2 + i2 / kCell], qv);
                                    }
                                }
                            }
                        }
                        finish(&ct->hist_done);
                        spin_until(&ct->hist_done, NCW);
                        PTOCK(c_h, t2);
                        PTICK(t3);
                        // ---- S. bound -> scale ----
                        {
                            // lanes own (c0, c1) columns of cells and walk along c2 with a sliding pair sum
                            int wmax = 0;
                            if (lane < nc0 * nc1) {
                                const int c0 = lane / nc1, c1 = lane - c0 * nc1;
                                const bool h0 = c0 + 1 < nc0, h1 = c1 + 1 < nc1;
                                const int *cp = cells + (c0 * nc1 + c1) * nc2;
                                int prev = 0;
                                for (int c2 = nc2 - 1; c2 >= 0; --c2) {
                                    int t = cp[c2];
                                    if (h1) t += cp[nc2 + c2];
                                    if (h0) t += cp[nc1 * nc2 + c2];
                                    if (h0 && h1) t += cp[(nc1 + 1) * nc2 + c2];
                                    wmax = max(wmax, t + prev);
                                    prev = t;
                                }
                            }
                            const int m = __reduce_max_sync(0xffffffffu, wmax);     // every warp: same cells, same bound
                            if (vmax > 0.f && m > 0) {
                                // |any accumulator| <= M = m * 2^-QBITS * vmax * w3 ; single contribution <= vmax * w3
                                const float M = (float)m * (1.f / (float)(1 << QBITS)) * vmax * w3;
                                int e1, e2;
                                frexpf(M, &e1);
                                frexpf(vmax * w3, &e2);
                                int k = min(30 - e1, 21 - e2);
                                k = max(-120, min(120, k));
                                scale = ldexpf(1.f, k);
                            }
                        }
                        PTOCK(c_s, t3);
                    }
                    PTICK(t4);
                    // ---- A. atomics; on the side, the box of tile q + LA (its coordinates land during Z / H / S) ----
                    if (look && !look_waited) { mbar_wait(gfull + (q + LA) % kPNG, ((q + LA) / kPNG) & 1); look_waited = true; }
#pragma unroll 1
                    for (int u = 0; u < URW; ++u) {
                        const int r = warp + u * NCW;
                        if (r >= NROWS) break;
                        const int p = r / TY, ly = r - p * TY;
                        if (look) {
                            const float c2[3] = {gt2[r * (TZ * 3)], gt2[r * (TZ * 3) + 1], gt2[r * (TZ * 3) + 2]};
                            const bool use = p < nxv2 && ly < nyv2 && lane < nzv2 && (!masked || inbounds<float, 3>(kp, c2));
#pragma unroll
                            for (int d = 0; d < 3; ++d) {
                                mn[d] = fminf(mn[d], use ? c2[d] : 3e38f);
                                mx[d] = fmaxf(mx[d], use ? c2[d] : -3e38f);
                            }
                        }
                        float cc[3]; bool act;
                        row_of(r, cc, act);
                        float val = rval[0];
#pragma unroll
                        for (int uu = 1; uu < URW; ++uu) val = (u == uu) ? rval[uu] : val;
                        if (!boxed && !COUNT && act) val = __ldg(src + (p * kp.pts_n[1] + ly) * kp.pts_n[2] + lane);
                        if (act && g.mode == PIPE_GLOBAL) {
                            // box does not fit: direct global REDs (same arithmetic as scatter.cu)
                            Axis<float, W> ax[3];
                            bool ok = setup_axis<float, ORDER, 0, W>(ax[0], cc[0], ORDER, kp.bound[0], kp.vol_n[0], (int)kp.vol_s[0], kp);
                            ok = setup_axis<float, ORDER, 0, W>(ax[1], cc[1], ORDER, kp.bound[1], kp.vol_n[1], (int)kp.vol_s[1], kp) && ok;
                            ok = setup_axis<float, ORDER, 0, W>(ax[2], cc[2], ORDER, kp.bound[2], kp.vol_n[2], (int)kp.vol_s[2], kp) && ok;
                            if (ok) {
#pragma unroll 1
                                for (int i = 0; i < W; ++i)
#pragma unroll 1
                                    for (int jj = 0; jj < W; ++jj) {
                                        const float vij = val * ax[0].w[i] * ax[1].w[jj];
#pragma unroll
                                        for (int k = 0; k < W; ++k)
                                            atomicAdd(dst + ax[0].off[i] + ax[1].off[jj] + ax[2].off[k], vij * ax[2].w[k]);
                                    }
                            }
                        } else if (act && boxed && scale != 0.f) {
                            const float f0 = floorf(cc[0] - 0.5f * (ORDER - 1)), f1 = floorf(cc[1] - 0.5f * (ORDER - 1)),
                                        f2 = floorf(cc[2] - 0.5f * (ORDER - 1));
                            float wx[W], wy[W], wz[W];
                            fast_weights<ORDER>(cc[0] - f0, wx);
                            fast_weights<ORDER>(cc[1] - f1, wy);
                            fast_weights<ORDER>(cc[2] - f2, wz);
                            int *rxy = bx + ((int)f0 - g.lo[0]) * kBoxPlane + ((int)f1 - g.lo[1]) * kBoxZ;
                            int *rk[W];
                            if (g.zfold) {
                                const int e0 = s * kZLut + (int)f2 - g.za;
#pragma unroll
                                for (int k = 0; k < W; ++k) { rk[k] = rxy + zoff[e0 + k]; wz[k] *= zsgn[e0 + k]; }
                            } else {
#pragma unroll
                                for (int k = 0; k < W; ++k) rk[k] = rxy + ((int)f2 - g.lo[2]) + k;
                            }
                            const float v = val * scale;
#pragma unroll
                            for (int i = 0; i < W; ++i) {
                                const float vi = v * wx[i];
#pragma unroll
                                for (int jj = 0; jj < W; ++jj) {
                                    const float vij = vi * wy[jj];
#pragma unroll
                                    for (int k = 0; k < W; ++k)
                                        atomicAdd(rk[k] + i * kBoxPlane + jj * kBoxZ, __float_as_int(fmaf(vij, wz[k], kPMagic)) - kPMagicBits);
                                }
                            }
                        }
                    }
                    finish(&ct->row_done);
                    if (look) {
                        pipe_merge_box(mn, mx, keys_base + ((q + LA) & 1) * 8);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(kfull + ((q + LA) & 1));
                        look = false;
                    }
                    PTOCK(c_a, t4);
                    PTICK(t5);
                    if (boxed) {
                        spin_until(&ct->row_done, NCW);
                        PTOCK(c_aw, t5);
                        PTICK(t6);
                        const float inv = scale != 0.f ? 1.f / scale : 0.f;
                        // ---- F. fold what sits outside the volume back inside ----
                        if (g.mode == PIPE_FOLD) {
                            for (;;) {
                                const int a = claim(&ct->fold_next);
                                if (a >= g.ext[0]) break;
                                push_fold_plane(kp, g, bx, dst, a, inv);
                                finish(&ct->fold_done);
                            }
                            spin_until(&ct->fold_done, g.ext[0]);
                        }
                        PTOCK(c_f, t6);
                        PTICK(t7);
                        // ---- C. per x-plane, by the warp that claims it: fixed -> float in place, flush through the
                        //         TMA unit (float add; folds along x go into the plane coordinate, what overhangs an
                        //         upper face is clipped), wait until the plane has been read, zero it for the next item ----
                        const bool by_tensor = g.lo[1] >= 0 && g.lo[2] >= 0;     // the TMA unit TRAPS on negative start coordinates
                        const int zs = max(g.lo[2], 0), ze = min(g.lo[2] + kBoxZ, kp.vol_n[2]);
                        for (;;) {
                            const int a = claim(&ct->conv_next);
                            if (a >= g.ext[0]) break;
                            int4 *p4 = reinterpret_cast<int4 *>(bx + a * kBoxPlane);
                            const int px = g.xfold ? bound_index<int>(kp.bound[0], g.lo[0] + a, kp.vol_n[0]) : g.lo[0] + a;
                            const bool px_ok = px >= 0 && px < kp.vol_n[0];
                            if (scale != 0.f && px_ok) {
                                for (int e = lane; e < kBoxPlane / 4; e += 32) {
                                    const int4 iv = p4[e];
                                    *reinterpret_cast<float4 *>(p4 + e) = make_float4(inv * (float)iv.x, inv * (float)iv.y, inv * (float)iv.z, inv * (float)iv.w);
                                }
                                fence_async_smem();
                                __syncwarp();
                                if (by_tensor) {
                                    if (lane == 0) tma_reduce_add_5d(&tm_out, p4, g.lo[2], g.lo[1], px, c, b);
                                } else if (lane < g.ext[1] && ze > zs) {
                                    // box hanging over a lower face (bound `zero` / `dft`, or y): one 1-D bulk reduction per row, clipped by hand
                                    const int sy = g.lo[1] + lane;
                                    if (sy >= 0 && sy < kp.vol_n[1])
                                        bulk_red_add_f32(dst + ((i64)px * kp.vol_n[1] + sy) * kp.vol_n[2] + zs,
                                                         bx + a * kBoxPlane + lane * kBoxZ + (zs - g.lo[2]), (unsigned)(ze - zs) * 4u);
                                }
                                bulk_commit();
                                bulk_wait_read<0>();
                                __syncwarp();
                            }
                            if (scale != 0.f)
                                for (int e = lane; e < kBoxPlane / 4; e += 32) p4[e] = make_int4(0, 0, 0, 0);
                            if (a == 0)
                                for (int e = lane; e < nc0 * nc1 * nc2; e += 32) cells[e] = 0;
                        }
                        PTOCK(c_c, t7);
                    }
                    c_items += 1;
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(bdone + s);
                        if (last && c == C - 1) mbar_arrive(gempty + q % kPNG);
                    }
                    ++n;
                } while (!last);
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");       // every reduction this thread issued has landed
        if (dbg && blockIdx.x == 0 && warp == 0 && lane == 0) {
            dbg[4] = c_bfull; dbg[5] = c_z; dbg[6] = c_h; dbg[7] = c_s; dbg[8] = c_a; dbg[9] = c_aw; dbg[10] = c_f; dbg[11] = c_c;
            dbg[12] = c_items; dbg[13] = clock64() - t_all;
        }
    }
}

// ---------------------------------------------------------------- launch --

template <int ORDER, int OP, int NCW, int kPNG, int kPNB, int CTAS>
static int launch_push_pipe(const KParams &kp, const float *img, const float *grid, float *out, cudaStream_t stream) {
    constexpr int NPT = 8 * 8 * 32;
    const size_t smem_total = (size_t)kPNB * kPBoxWords * 4 + (size_t)kPNG * NPT * 3 * 4 + kPNB * sizeof(PipeGeom) +
                              (2 * kPNG + 2 * kPNB + 2) * sizeof(unsigned long long) + kPNB * sizeof(PushCtl) + 40 * sizeof(int) +
                              4 * sizeof(PipeGeom) + (size_t)kPNB * kZLut * 8 + (size_t)kCells * sizeof(int) + 64;
    if (smem_total > (CTAS == 1 ? 227 * 1024 : 113 * 1024)) return 0;
    const i64 ntiles = kp.batch * ((kp.pts_n[0] + 7) / 8) * ((kp.pts_n[1] + 7) / 8) * ((kp.pts_n[2] + 31) / 32);
    if (ntiles == 0) return 1;
    if (ntiles * kp.channels > 0x3fffffffLL) return 0;
    CUtensorMap tm_out, tm_grid;
    const int gbmul = (kp.grid_sb != 0 && kp.batch > 1) ? 1 : 0;
    {
        // dense (B, C, X, Y, Z) float32 accumulation volume
        const long long dim[5] = {kp.vol_n[2], kp.vol_n[1], kp.vol_n[0], kp.channels, kp.batch};
        const long long str[5] = {1, kp.vol_n[2], (long long)kp.vol_n[2] * kp.vol_n[1], kp.vol_total, kp.vol_total * kp.channels};
        const int box[5] = {kBoxZ, kBoxY, 1, 1, 1};
        if (!make_tensor_map(&tm_out, out, 5, dim, str, box)) return 0;
    }
    {
        const long long row = (long long)kp.pts_n[2] * 3;
        const long long dim[4] = {row, kp.pts_n[1], kp.pts_n[0], gbmul ? kp.batch : 1};
        const long long str[4] = {1, row, row * kp.pts_n[1], gbmul ? kp.grid_sb : row * kp.pts_n[1] * kp.pts_n[0]};
        const int box[4] = {96, 8, 8, 1};
        if (!make_tensor_map(&tm_grid, grid, 4, dim, str, box)) return 0;
    }
    auto kern = push_pipe3d_kernel<ORDER, OP, NCW, kPNG, kPNB, CTAS>;
    IB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_total));
    const int nblocks = (int)(ntiles < CTAS * pipe_sm_count() ? ntiles : CTAS * pipe_sm_count());
    long long *dbg = nullptr;
    if (getenv("IB200_PIPE_DEBUG")) IB200_CUDA_CHECK(cudaGetSymbolAddress((void **)&dbg, g_push_dbg));
    kern<<<nblocks, 32 * (NCW + 1), smem_total, stream>>>(kp, tm_out, tm_grid, img, out, (int)ntiles, gbmul,
        make_inv((kp.pts_n[2] + 31) / 32), make_inv((kp.pts_n[1] + 7) / 8), make_inv((kp.pts_n[0] + 7) / 8), dbg);
    static thread_local char name[64];
    snprintf(name, sizeof(name), "%s_pipe3d_o%d", OP == OP_COUNT ? "count" : "push", ORDER);
    note_launch(name);
    IB200_CUDA_CHECK(cudaGetLastError());
    return 1;
}

// `acc` is the zero-filled float32 accumulation volume (the output itself for F32)
int try_push_pipe(int op, const KParams &kp, int dtype, const void *img, const void *grid, void *acc, cudaStream_t stream) {
    if (op != OP_PUSH && op != OP_COUNT) return 0;
    if (dtype != IB200_F32) return 0;
    if (kp.dim != 3 || !kp.pts_dense) return 0;
    if (kp.order[0] != kp.order[1] || kp.order[0] != kp.order[2]) return 0;
    if (kp.order[0] < 1 || kp.order[0] > 3) return 0;
    if (kp.pts_total < 32768) return 0;
    if (kp.pts_total * 3 > 0x7fffffffLL) return 0;
    if (kp.flags & (IB200_FLAG_NO_PIPE | IB200_FLAG_DISPLACEMENT)) return 0;
    // Not the default: both scatter kernels are bound by the shared-memory atomic pipe (64 ATOMS per source at
    // ~2.7 clk each once same-bank / same-address lanes serialise), and on the 256^3 cubic workload the
    // one-tile-per-CTA kernel (0.67 ms) still beats this one (0.73 ms) -- see DESIGN.md.  IB200_FLAG_FORCE_PIPE selects it.
    if (!(kp.flags & IB200_FLAG_FORCE_PIPE)) return 0;
    // TMA: 16-byte aligned bases and strides of the accumulation volume and of the grid
    if ((uintptr_t)acc % 16 || (uintptr_t)grid % 16) return 0;
    if (kp.vol_n[2] % 4 || kp.pts_n[2] % 4 || kp.grid_sb % 4) return 0;
    const float *v = (const float *)img, *g = (const float *)grid;
    float *o = (float *)acc;
    // two CTAs per SM, each with one box: the fill / bound / flush phases of one hide behind the atomics of the other
    constexpr int NCW = 7;
#define IB200_PP NCW, 2, 1, 2
    if (op == OP_PUSH) {
        switch (kp.order[0]) {
        case 1: return launch_push_pipe<1, OP_PUSH, IB200_PP>(kp, v, g, o, stream);
        case 2: return launch_push_pipe<2, OP_PUSH, IB200_PP>(kp, v, g, o, stream);
        case 3: return launch_push_pipe<3, OP_PUSH, IB200_PP>(kp, v, g, o, stream);
        }
    } else {
        switch (kp.order[0]) {
        case 1: return launch_push_pipe<1, OP_COUNT, IB200_PP>(kp, v, g, o, stream);
        case 2: return launch_push_pipe<2, OP_COUNT, IB200_PP>(kp, v, g, o, stream);
        case 3: return launch_push_pipe<3, OP_COUNT, IB200_PP>(kp, v, g, o, stream);
        }
    }
    return 0;
}

}  // namespace ib200

extern "C" __attribute__((visibility("default"))) int ib200_debug_push_counters(long long *out16) {
    return cudaMemcpyFromSymbol(out16, ib200::g_push_dbg, sizeof(long long) * 16) == cudaSuccess ? 0 : -1;
}
