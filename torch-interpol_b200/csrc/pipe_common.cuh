// Building blocks of the persistent, warp-specialised ("pipe") kernels: mbarrier and TMA tile-copy
// wrappers, the geometry of a staged box (planning from a bounding box of support
// starts, boundary handling modes), the fix-up pass for what the TMA unit cannot express,
// and the host-side tensor-map encoder.
//
// One CTA per SM loops over tiles of TX x TY x TZ lattice points.  The last warp
// is the producer: it streams the grid coordinates of upcoming tiles into a ring of shared-memory
// buffers (one cp.async.bulk.tensor per tile), turns the bounding box the consumers reduced for a
// tile into a plan (whole tile / z halves / z quarters), and requests one TMA box per x-plane of the
// input volume.  (push_box.cu shares the TMA wrappers and the tensor-map encoder.)  All other warps are
// consumers: they only ever wait on mbarriers and counters, so the tap loop of tile n overlaps with
// every memory phase of tiles n+1, n+2.
#pragma once
#include <cuda.h>
#include "tile_common.cuh"

namespace ib200 {

// ---- mbarrier / bulk copy PTX ------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// tiled tensor copies (TMA): coordinates innermost first, out-of-bounds elements are zero-filled
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3,
                                            unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_5d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, int c3, int c4,
                                            unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tm) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(tm) : "memory");
}
// order generic-proxy shared-memory accesses before later async-proxy accesses
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// ---- geometry of one staged box ---------------------------------------------
// A box is [<= kBoxX planes][kBoxY rows][kBoxZ words], loaded plane by plane with one
// TMA tile copy each (box {kBoxZ, kBoxY, 1, 1, 1}); elements outside the volume are
// zero-filled by the TMA unit.  Rows are 64 words so that the bank of a tap depends
// on z only, and all strides are compile-time constants (tap offsets are immediates).
constexpr int kBoxZ = 64, kBoxY = 16, kBoxX = 16;
constexpr int kBoxPlane = kBoxZ * kBoxY;          // words per x-plane
constexpr int kBoxWords = kBoxPlane * kBoxX;      // 64 KB

enum { PIPE_EMPTY = 0, PIPE_PLAIN = 1, PIPE_FOLD = 2, PIPE_GLOBAL = 3 };

struct PipeGeom {
    int lo[3];       // unfolded source coordinate of box element (0, 0, 0); lo[2] % 4 == 0
    int ext[3];      // extents (<= kBoxX, kBoxY, kBoxZ)
    int mode;        // PIPE_*
    int r0[3], r1[3];   // box indices [r0, r1) along x / y (elements) and z (4-word vectors) whose source
                        // lies inside the volume (or whose bound is `zero`: the TMA fill is the answer)
    int zlo, zhi;       // lanes (z offsets inside the tile) this box serves
    int last;           // last part of its tile
    int xfold;          // x-planes are requested through the boundary map of x
    int zfold;          // z is stored FOLDED: lo[2] / ext[2] span the folded indices, taps go through a table
    int za, zn;         // the table covers the unfolded indices [za, za + zn)
};
constexpr int kZLut = 128;   // longest unfolded z range a folded box may serve

__device__ __forceinline__ int floor_div4(int a) { return a >> 2; }   // arithmetic shift == floor

// Bounding box of the raw coordinates of a TX x TY x TZ tile whose coordinates sit in `gt`
// ([TX*TY rows][TZ * 3]).  Cooperative: consumer warp `cw` of NCW reduces the rows cw,
// cw + NCW, ... and merges its result into keys[2 * d + {0, 1}] (order-preserving integer
// keys of the min / max coordinate, fkey) with shared atomics.  The single producer warp
// cannot afford this loop: it only gets its scheduler's leftovers.
__device__ __forceinline__ int pipe_key_init(int slot) { return (slot & 1) ? fkey(-3e38f) : fkey(3e38f); }

// warp-reduce a per-lane box and merge it into keys[2 * d + {0, 1}]
__device__ __forceinline__ void pipe_merge_box(const float (&mn)[3], const float (&mx)[3], int *keys) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int a = __reduce_min_sync(0xffffffffu, fkey(mn[d]));
        const int b = __reduce_max_sync(0xffffffffu, fkey(mx[d]));
        if (lane == 0) { atomicMin(keys + 2 * d, a); atomicMax(keys + 2 * d + 1, b); }
    }
}

template <int TX, int TY, int TZ, int NCW>
__device__ __forceinline__ void pipe_tile_box(const KParams &kp, const float *gt, int nxv, int nyv, int nzv, int *keys, int cw) {
    static_assert(TZ == 32, "one lane per z");
    const int lane = threadIdx.x & 31;
    const bool masked = kp.extrapolate != 1;
    float mn[3] = {3e38f, 3e38f, 3e38f}, mx[3] = {-3e38f, -3e38f, -3e38f};
    const float *gl = gt + lane * 3;
#pragma unroll 2
    for (int r = cw; r < TX * TY; r += NCW) {
        const int lx = r / TY, ly = r % TY;
        const float c[3] = {gl[r * (TZ * 3)], gl[r * (TZ * 3) + 1], gl[r * (TZ * 3) + 2]};
        const bool use = lane < nzv && lx < nxv && ly < nyv && (!masked || inbounds<float, 3>(kp, c));
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            mn[d] = fminf(mn[d], use ? c[d] : 3e38f);
            mx[d] = fmaxf(mx[d], use ? c[d] : -3e38f);
        }
    }
    pipe_merge_box(mn, mx, keys);
}

// One box per quarter of the z range (8 lanes each), for tiles whose whole box does not fit:
// the tile is then planned in z halves or z quarters, whichever is the coarsest split whose
// boxes all fit (displacements that shear x / y along z make the box of 32 z-neighbours much
// wider than the box of 8).  Rare, so the producer warp does it alone:
// keys[quarter * 6 + 2 * d + {0, 1}].
template <int TX, int TY, int TZ>
__device__ __forceinline__ void pipe_quarter_boxes(const KParams &kp, const float *gt, int nxv, int nyv, int nzv, int *keys) {
    const int lane = threadIdx.x & 31;
    const bool masked = kp.extrapolate != 1;
    float mn[3] = {3e38f, 3e38f, 3e38f}, mx[3] = {-3e38f, -3e38f, -3e38f};
    if (lane < nzv) {
#pragma unroll 4
        for (int r = 0; r < TX * TY; ++r) {
            const int lx = r / TY, ly = r % TY;
            if (lx < nxv && ly < nyv) {
                const float *g = gt + (r * TZ + lane) * 3;
                const float c[3] = {g[0], g[1], g[2]};
                if (!masked || inbounds<float, 3>(kp, c)) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) { mn[d] = fminf(mn[d], c[d]); mx[d] = fmaxf(mx[d], c[d]); }
                }
            }
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        // butterfly inside each group of 8 lanes
        int a = fkey(mn[d]), b = fkey(mx[d]);
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            a = min(a, __shfl_xor_sync(0xffffffffu, a, o));
            b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
        }
        if ((lane & 7) == 0) { keys[(lane >> 3) * 6 + 2 * d] = a; keys[(lane >> 3) * 6 + 2 * d + 1] = b; }
    }
    __syncwarp();
}

// geometry of the box of quarters [q0, q1) from the keys above
template <int ORDER>
__device__ __forceinline__ PipeGeom pipe_geom(const KParams &kp, const int *keys, int q0, int q1, int max_x = kBoxX) {
    PipeGeom g;
    g.xfold = 0; g.zfold = 0; g.za = 0; g.zn = 0;
    bool any = true;
    bool zfit = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        int ka = kIntMax, kb = kIntMin;
        for (int q = q0; q < q1; ++q) { ka = min(ka, keys[q * 6 + 2 * d]); kb = max(kb, keys[q * 6 + 2 * d + 1]); }
        const float fa = fkey_inv(ka), fc = fkey_inv(kb);
        int a = 0, b = 0;
        if (fa > fc) any = false;
        else { a = start_of<ORDER>(fa); b = start_of<ORDER>(fc); }
        if (d == 2 && any) {
            // Mirror-type bounds along z: instead of staging out-of-volume words and rewriting them,
            // stage the FOLDED index range (a sub-range of the volume, usually smaller) and let the taps
            // look their word up in a small table (the producer fills it, pull_pipe.cu).
            const int bz = kp.bound[2];
            const bool mirror = bz == IB200_BOUND_REPLICATE || bz == IB200_BOUND_DCT1 || bz == IB200_BOUND_DCT2 ||
                                bz == IB200_BOUND_DST1 || bz == IB200_BOUND_DST2;
            const int lo_ok = bz == IB200_BOUND_DST1 ? 1 : 0;
            const long long top = (long long)b + ORDER;
            if (mirror && (a < lo_ok || top > kp.vol_n[2] - 1)) {
                const long long len = top - a + 1;
                if (len > kZLut) {
                    zfit = false;
                } else {
                    const int lane = threadIdx.x & 31;
                    int fmin = kIntMax, fmax = kIntMin;
                    for (int e = lane; e < (int)len; e += 32) {
                        const int f = bound_index<int>(bz, a + e, kp.vol_n[2]);
                        fmin = min(fmin, f); fmax = max(fmax, f);
                    }
                    fmin = __reduce_min_sync(0xffffffffu, fmin);
                    fmax = __reduce_max_sync(0xffffffffu, fmax);
                    g.zfold = 1; g.za = a; g.zn = (int)len;
                    a = fmin; b = fmax - ORDER;        // so that ext = fmax - (fmin & ~3) + 1 below
                }
            }
        }
        if (d == 2) a &= ~3;
        g.lo[d] = a;
        const long long e = (long long)b - a + 1 + ORDER;
        g.ext[d] = (int)(e > 0x3fffffff ? 0x3fffffff : e);
    }
    const bool fits = zfit && g.ext[0] <= max_x && g.ext[1] <= kBoxY && g.ext[2] <= kBoxZ;
    if (!any) {
        g.mode = PIPE_EMPTY; g.ext[0] = g.ext[1] = g.ext[2] = 0; g.lo[0] = g.lo[1] = g.lo[2] = 0;
    } else if (!fits) {
        g.mode = PIPE_GLOBAL;
    } else {
        g.mode = PIPE_PLAIN;
    }
    const int vpr = (min(g.ext[2], kBoxZ) + 3) >> 2;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int e = d == 2 ? vpr : min(g.ext[d], d == 0 ? kBoxX : kBoxY);
        int r0 = 0, r1 = e;
        // x: one TMA request per plane, so the producer simply requests the folded plane (bounds of sign +1)
        const bool by_tma = d == 0 && (kp.bound[0] == IB200_BOUND_REPLICATE || kp.bound[0] == IB200_BOUND_DCT1 ||
                                       kp.bound[0] == IB200_BOUND_DCT2 || kp.bound[0] == IB200_BOUND_DFT);
        if (by_tma) {
            g.xfold = (g.lo[0] < 0 || g.lo[0] + e > kp.vol_n[0]) ? 1 : 0;
        } else if (d == 2 && g.zfold) {
            // folded range: inside the volume by construction
        } else if (kp.bound[d] != IB200_BOUND_ZERO) {
            const int lo_ok = (kp.bound[d] == IB200_BOUND_DST1) ? 1 : 0;     // dst1 zeroes voxel 0 (Q1)
            if (d == 2) {
                // vectors whose four source voxels lie in [lo_ok, nz - 1]
                r0 = floor_div4(lo_ok - g.lo[2] + 3);
                r1 = floor_div4(kp.vol_n[2] - 4 - g.lo[2]) + 1;
            } else {
                r0 = lo_ok - g.lo[d];
                r1 = kp.vol_n[d] - g.lo[d];
            }
            r0 = max(0, min(r0, e)); r1 = max(r0, min(r1, e));
        }
        g.r0[d] = r0; g.r1[d] = r1;
        if (g.mode == PIPE_PLAIN && (r0 != 0 || r1 != e)) g.mode = PIPE_FOLD;
    }
    g.zlo = q0 * 8; g.zhi = q1 * 8; g.last = (q1 == 4) ? 1 : 0;
    return g;
}

// Fix-up of x-plane `a` of a PIPE_FOLD box by one warp: every vector with a source outside the
// volume along an axis whose fold the TMA requests did not already apply is recomputed through
// the boundary maps.  The folded source usually lies inside the box itself (mirror / replicate
// bounds next to a face): it is then read back from shared memory, otherwise from global memory.
// Only in-range elements are ever read, only out-of-range elements are written: no hazard
// between warps fixing different planes.
__device__ __forceinline__ void pipe_fixup_plane(const KParams &kp, const PipeGeom &g, float *bx, const float *src, int a) {
    const int lane = threadIdx.x & 31;
    const int vpr = (g.ext[2] + 3) >> 2;
    const int sx = g.lo[0] + a;
    const int fx = bound_index<int>(kp.bound[0], sx, kp.vol_n[0]);
    const int sgx = bound_sign<int>(kp.bound[0], sx, kp.vol_n[0]);
    const bool x_in = a >= g.r0[0] && a < g.r1[0];
    const int ax = x_in ? a : fx - g.lo[0];
    const bool x_src = x_in || (ax >= g.r0[0] && ax < g.r1[0]);
    // (row, 4-word vector) items of the plane are dealt to the lanes round-robin: what needs rewriting is usually
    // a few whole rows next to a face, and with one row per lane pair only a handful of lanes had any work
    const int nitems = g.ext[1] * vpr;
    for (int it = lane; it < nitems; it += 32) {
        const int bb = it / vpr, v = it - bb * vpr;
        const bool y_in = bb >= g.r0[1] && bb < g.r1[1];
        const bool v_in = v >= g.r0[2] && v < g.r1[2];
        if (x_in && y_in && v_in) continue;
        const int sy = g.lo[1] + bb;
        // (axes of bound `zero` count as in range -- the TMA fill is their answer -- but their sign is 0 outside)
        const int fy = bound_index<int>(kp.bound[1], sy, kp.vol_n[1]);
        const int sgxy = sgx * bound_sign<int>(kp.bound[1], sy, kp.vol_n[1]);
        const int by = fy - g.lo[1];
        const bool xy_src = x_src && by >= g.r0[1] && by < g.r1[1];
        const float *brow = bx + ax * kBoxPlane + by * kBoxZ;
        const float *grow = src + fx * (int)kp.vol_s[0] + fy * (int)kp.vol_s[1];
        float *dstrow = bx + a * kBoxPlane + bb * kBoxZ;
        float val[4];
        // a folded-z box holds RAW volume words (the taps apply the z map, dst1's zero at voxel 0 included)
        const int nzvol = kp.vol_n[2], sz0 = g.lo[2] + 4 * v;
        const bool z_inside = sz0 >= 0 && sz0 + 3 <= nzvol - 1;
        int vec = 0;            // 1: signed copy of a vector of the box, 2: of the volume (16-byte aligned: lo[2] % 4 == 0)
        int gz0 = sz0;
        if (xy_src && v_in && (!g.zfold || sz0 + 3 <= nzvol - 1)) vec = 1;      // a row / plane mirrored at an x / y face
        else if (!xy_src && z_inside && (g.zfold || kp.bound[2] != IB200_BOUND_DST1)) vec = 2;   // its source lies outside the box
        else if (!g.zfold && !z_inside && kp.bound[2] == IB200_BOUND_DFT && (nzvol & 3) == 0 && nzvol >= 4) {
            gz0 = bound_index<int>(IB200_BOUND_DFT, sz0, nzvol);                  // wraps as a whole: nz % 4 == 0
            vec = 2;
        }
        if (sgxy == 0) {
            val[0] = 0.f; val[1] = 0.f; val[2] = 0.f; val[3] = 0.f;              // (nothing is read: the folded indices may mean nothing)
        } else if (vec != 0) {
            const float4 t = vec == 1 ? *reinterpret_cast<const float4 *>(brow + 4 * v) : __ldg(reinterpret_cast<const float4 *>(grow + gz0));
            const float sg = (float)sgxy;
            val[0] = sg * t.x; val[1] = sg * t.y; val[2] = sg * t.z; val[3] = sg * t.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int sz = sz0 + k;
                const int fz = g.zfold ? min(sz, nzvol - 1) : bound_index<int>(kp.bound[2], sz, nzvol);
                const int sg = g.zfold ? sgxy : sgxy * bound_sign<int>(kp.bound[2], sz, nzvol);
                const int zz = fz - g.lo[2];
                float t = 0.f;
                if (sg != 0) t = (xy_src && zz >= 4 * g.r0[2] && zz < 4 * g.r1[2]) ? brow[zz] : __ldg(grow + fz);
                val[k] = (float)sg * t;
            }
        }
        *reinterpret_cast<float4 *>(dstrow + 4 * v) = make_float4(val[0], val[1], val[2], val[3]);
    }
}

static inline int pipe_sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) n = v;
        else n = kNumSMs;
    }
    return n;
}

// ---- host side: tensor maps ---------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        tried = true;
    }
    return fn;
}
// tensor of `rank` dims (innermost first) of `esize`-byte elements, element strides `stride[1..rank-1]`
static inline bool make_tensor_map_t(CUtensorMap *tm, const void *base, int rank, const long long *dim,
                                     const long long *stride, const int *box, CUtensorMapDataType dt, int esize) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return false;
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        if (dim[i] < 1 || dim[i] > 0xffffffffLL) return false;
        gdim[i] = (cuuint64_t)dim[i]; bdim[i] = (cuuint32_t)box[i]; estr[i] = 1;
        if (i > 0) {
            if (stride[i] <= 0 || (stride[i] * esize) % 16 != 0 || stride[i] * esize >= (1LL << 40)) return false;
            gstr[i - 1] = (cuuint64_t)stride[i] * esize;
        }
    }
    return enc(tm, dt, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr, bdim, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static inline bool make_tensor_map(CUtensorMap *tm, const void *base, int rank, const long long *dim,
                                   const long long *stride, const int *box) {
    return make_tensor_map_t(tm, base, rank, dim, stride, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4);
}

}  // namespace ib200
