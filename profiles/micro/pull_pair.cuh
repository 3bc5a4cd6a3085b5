// EXPERIMENT (round 2, not adopted -- profiles/README.md r2f): paired tap loop for the persistent pull kernel: one lane evaluates TWO output voxels that are neighbours
// along z, A = (x, y, z) and B = (x, y, z + 1), against a box of the input volume staged in shared memory.
//
// Why (profiles/micro/smem_micro.cu, pair_lab.cu): with one voxel per lane a z-row of 32
// voxels costs (ORDER+1)^3 LDS.32 per lane, each 1 wavefront when the 32 supports span <= 32 banks and 2 as
// soon as the deformation stretches along z -- 69 / 132 clk per row in isolation, 64 taps * 4 B = 256 B of
// crossbar traffic per voxel.  Neighbours along z share all but one or two of their z taps:  a lane that owns A
// and B reads, per (x, y) row of the support, ONE aligned window of NWIN = 6 words (three LDS.64, order 3) that
// holds the taps of both -- 3 words per voxel and row instead of 4, in 64-bit accesses whose bank conflicts
// are decided per half warp (16 lanes, span 16 * 2 * stretch words) rather than per warp.  Each voxel applies
// its ORDER+1 z weights through a zero-padded vector of NWIN weights (packed FFMA2, sm_100: the even / odd
// words of the window accumulate in the two halves of a register pair and are added once at the very end).
//
// B rides along with A only when its support starts within one cell of A's along x and y and fits A's z window:
//   * same (x, y) cell (4 of 5 pairs on the benchmark deformation): nothing else to do;
//   * one cell off along x and / or y: the rows B shares with A are taken in the main loop with B's x / y
//     weights shifted by one (zero-padded), the plane / row it does not share in two short predicated passes;
//   * anything else (folds, steep shear, A masked): B is evaluated on its own in a second trip of the pass loop.
// Zero-padded weights multiply words OUTSIDE a voxel's support: harmless unless such a word is NaN / Inf
// (0 * NaN), so any non-finite result is recomputed from the voxel's own taps only (`exact` below) -- the
// reference's NaN semantics (interpol/nd.py:118-136 touches the support and nothing else) are kept exactly.
#pragma once
#include "tile_common.cuh"   // -I torch-interpol_b200/csrc

namespace ib200 {

// P[j] = w[j - o] for 0 <= j - o < W, 0 elsewhere; o in [0, NWIN - W]
template <int W, int NWIN>
__device__ __forceinline__ void pad_shift(const float (&w)[W], int o, float (&P)[NWIN]) {
#pragma unroll
    for (int j = 0; j < NWIN; ++j) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < W; ++k) {
            const int oo = j - k;
            if (oo >= 0 && oo <= NWIN - W) v = (o == oo) ? w[k] : v;
        }
        P[j] = v;
    }
}

// S[i] = w[i - d] for 0 <= i - d < W (d in {-1, 0, 1}), 0 elsewhere; everything 0 when !on
template <int W>
__device__ __forceinline__ void shift1(const float (&w)[W], int d, bool on, float (&S)[W]) {
#pragma unroll
    for (int i = 0; i < W; ++i) {
        float v = w[i];                                     // d == 0
        v = (d > 0) ? (i >= 1 ? w[i >= 1 ? i - 1 : 0] : 0.f) : v;
        v = (d < 0) ? (i + 1 < W ? w[i + 1 < W ? i + 1 : 0] : 0.f) : v;
        S[i] = on ? v : 0.f;
    }
}

template <int NWIN>
__device__ __forceinline__ void load_window(const float *p, float2 (&q)[NWIN / 2]) {
    if constexpr (NWIN % 4 == 0) {
#pragma unroll
        for (int m = 0; m < NWIN / 4; ++m) {
            const float4 v = *reinterpret_cast<const float4 *>(p + 4 * m);
            q[2 * m] = make_float2(v.x, v.y); q[2 * m + 1] = make_float2(v.z, v.w);
        }
    } else {
#pragma unroll
        for (int m = 0; m < NWIN / 2; ++m) q[m] = *reinterpret_cast<const float2 *>(p + 2 * m);
    }
}

template <int NWIN>
__device__ __forceinline__ float2 dot_window(const float2 (&W2)[NWIN / 2], const float2 (&q)[NWIN / 2]) {
    float2 t = __fmul2_rn(W2[0], q[0]);
#pragma unroll
    for (int m = 1; m < NWIN / 2; ++m) t = __ffma2_rn(W2[m], q[m], t);
    return t;
}

// one voxel from its own taps only (rolled loops: rare path)
template <int ORDER, int PLANE, int ROW>
__device__ __noinline__ float pull_point_box(const float *bx, int lo0, int lo1, int lo2, float c0, float c1, float c2) {
    constexpr int W = ORDER + 1;
    const float f0 = floorf(c0 - 0.5f * (ORDER - 1)), f1 = floorf(c1 - 0.5f * (ORDER - 1)), f2 = floorf(c2 - 0.5f * (ORDER - 1));
    float wx[W], wy[W], wz[W];
    fast_weights<ORDER>(c0 - f0, wx); fast_weights<ORDER>(c1 - f1, wy); fast_weights<ORDER>(c2 - f2, wz);
    const float *r = bx + ((int)f0 - lo0) * PLANE + ((int)f1 - lo1) * ROW + ((int)f2 - lo2);
    float acc = 0.f;
#pragma unroll 1
    for (int i = 0; i < W; ++i) {
        float s = 0.f;
#pragma unroll 1
        for (int j = 0; j < W; ++j) {
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < W; ++k) t = fmaf(wz[k], r[i * PLANE + j * ROW + k], t);
            s = fmaf(wy[j], t, s);
        }
        acc = fmaf(wx[i], s, acc);
    }
    return acc;
}

// Evaluates voxels A (coordinates a*) and B (b*) of this lane; act* = the voxel exists and is not masked.
// Must be called by all 32 lanes of the warp.  bx = box element (0, 0, 0) = source voxel (lo0, lo1, lo2),
// lo2 % 4 == 0, rows of ROW words, planes of PLANE words; every tap of an active voxel lies inside the box.
template <int ORDER, int NWIN, int PLANE, int ROW>
__device__ __forceinline__ void pull_pair_eval(const float *bx, const int lo0, const int lo1, const int lo2,
                                               const float a0, const float a1, const float a2,
                                               const float b0, const float b1, const float b2,
                                               const bool actA, const bool actB, float &resA, float &resB) {
    constexpr int W = ORDER + 1;
    constexpr int VEC = (NWIN % 4 == 0) ? 4 : 2;          // words per load: the window starts on a VEC boundary
    constexpr int OMAX = NWIN - W;
    static_assert(VEC - 1 <= OMAX, "the primary voxel must fit its own window");
    constexpr float kHalf = 0.5f * (ORDER - 1);
    resA = 0.f; resB = 0.f;
    float p0 = a0, p1 = a1, p2 = a2;                      // primary voxel of this trip
    bool actP = actA, actS = actB;
    // (no lane-dependent branch around the warp votes: a lane without work runs on zero weights at the box origin)
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const float fp0 = floorf(p0 - kHalf), fp1 = floorf(p1 - kHalf), fp2 = floorf(p2 - kHalf);
        const float fs0 = floorf(b0 - kHalf), fs1 = floorf(b1 - kHalf), fs2 = floorf(b2 - kHalf);
        const bool okP = actP && fabsf(fp0) < 4e18f && fabsf(fp1) < 4e18f && fabsf(fp2) < 4e18f;
        const bool okS = actS && fabsf(fs0) < 4e18f && fabsf(fs1) < 4e18f && fabsf(fs2) < 4e18f;
        const int ip0 = okP ? (int)fp0 - lo0 : 0, ip1 = okP ? (int)fp1 - lo1 : 0, zP = okP ? (int)fp2 - lo2 : 0;
        const int base = zP & ~(VEC - 1), oP = zP & (VEC - 1);
        // (differences of floats that are integers: exact, and huge ones stay huge)
        const float ddx = fs0 - fp0, ddy = fs1 - fp1, ddz = fs2 - fp2;
        const bool paired = okP && okS && fabsf(ddx) <= 1.f && fabsf(ddy) <= 1.f && ddz >= (float)(-oP) && ddz <= (float)(OMAX - oP);
        const int dx = paired ? (int)ddx : 0, dy = paired ? (int)ddy : 0, oS = paired ? oP + (int)ddz : 0;
        const bool redo = okS && !paired;                 // B on its own next trip
        float wxP[W], wyP[W], wzP[W], wxS[W], wyS[W], wzS[W];
        fast_weights<ORDER>(p0 - fp0, wxP); fast_weights<ORDER>(p1 - fp1, wyP); fast_weights<ORDER>(p2 - fp2, wzP);
        fast_weights<ORDER>(b0 - fs0, wxS); fast_weights<ORDER>(b1 - fs1, wyS); fast_weights<ORDER>(b2 - fs2, wzS);
        float WP[NWIN], WS[NWIN], sxS[W], syS[W];
        pad_shift<W, NWIN>(wzP, oP, WP);
        pad_shift<W, NWIN>(wzS, oS, WS);
        shift1<W>(wxS, dx, paired, sxS);
        shift1<W>(wyS, dy, paired, syS);
        float2 WP2[NWIN / 2], WS2[NWIN / 2];
#pragma unroll
        for (int m = 0; m < NWIN / 2; ++m) { WP2[m] = make_float2(WP[2 * m], WP[2 * m + 1]); WS2[m] = make_float2(WS[2 * m], WS[2 * m + 1]); }
        const float *rb = bx + ip0 * PLANE + ip1 * ROW + base;
        float2 accP = make_float2(0.f, 0.f), accS = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < W; ++i) {
            float2 sP = make_float2(0.f, 0.f), sS = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < W; ++j) {
                float2 q[NWIN / 2];
                load_window<NWIN>(rb + i * PLANE + j * ROW, q);
                sP = __ffma2_rn(make_float2(wyP[j], wyP[j]), dot_window<NWIN>(WP2, q), sP);
                sS = __ffma2_rn(make_float2(syS[j], syS[j]), dot_window<NWIN>(WS2, q), sS);
            }
            accP = __ffma2_rn(make_float2(wxP[i], wxP[i]), sP, accP);
            accS = __ffma2_rn(make_float2(sxS[i], sxS[i]), sS, accS);
        }
        // the plane / the row B does not share with A
        const bool ex = paired && dx != 0, ey = paired && dy != 0;
        if (__any_sync(0xffffffffu, ex)) {
            if (ex) {
                const float *pe = rb + (dx > 0 ? W : -1) * PLANE + dy * ROW;
                const float we = dx > 0 ? wxS[W - 1] : wxS[0];
                float2 s = make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    float2 q[NWIN / 2];
                    load_window<NWIN>(pe + j * ROW, q);
                    s = __ffma2_rn(make_float2(wyS[j], wyS[j]), dot_window<NWIN>(WS2, q), s);
                }
                accS = __ffma2_rn(make_float2(we, we), s, accS);
            }
        }
        if (__any_sync(0xffffffffu, ey)) {
            if (ey) {
                const float *pe = rb + (dy > 0 ? W : -1) * ROW;
                const float we = dy > 0 ? wyS[W - 1] : wyS[0];
                float2 s = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < W; ++i) {
                    float2 q[NWIN / 2];
                    load_window<NWIN>(pe + i * PLANE, q);
                    s = __ffma2_rn(make_float2(sxS[i], sxS[i]), dot_window<NWIN>(WS2, q), s);
                }
                accS = __ffma2_rn(make_float2(we, we), s, accS);
            }
        }
        float rP = accP.x + accP.y, rS = accS.x + accS.y;
        // zero-padded weights met a non-finite word: redo from the voxel's own taps (exact NaN semantics)
        if (okP && !(fabsf(rP) <= 3.4e38f)) rP = pull_point_box<ORDER, PLANE, ROW>(bx, lo0, lo1, lo2, p0, p1, p2);
        if (paired && !(fabsf(rS) <= 3.4e38f)) rS = pull_point_box<ORDER, PLANE, ROW>(bx, lo0, lo1, lo2, b0, b1, b2);
        if (pass == 0) { if (okP) resA = rP; if (paired) resB = rS; }
        else if (okP) resB = rP;
        if (pass == 1 || !__any_sync(0xffffffffu, redo)) break;
        p0 = b0; p1 = b1; p2 = b2; actP = redo; actS = false;
    }
}

}  // namespace ib200
