// Tiled push / count (3-D, isotropic compile-time order, 16/32-bit storage):
// the adjoint of the tiled pull, with the scatter privatised in shared memory.
//
// A CTA owns a TX x TY x TZ block of SOURCE voxels.  As in pull_tile.cu it stages
// their grid coordinates, reduces the bounding box of all spline supports and
// plans 1, 2, 4 ... groups of x-planes whose boxes fit.  For every group:
//   a. a box of 32-bit FIXED-POINT accumulators is zeroed in shared memory;
//   b. a rigorous bound on the largest sum any accumulator can reach is derived
//      from a coarse histogram of |value| over cells of (ORDER+1)^3 support
//      starts (a target voxel only receives from 2x2x2 such cells) -- it fixes a
//      power-of-two scale such that no accumulator can overflow;
//   c. every source adds its (ORDER+1)^3 weighted contributions with NATIVE
//      integer shared-memory atomics (ATOMS.ADD runs at the LDS rate, 5x faster
//      than the CAS loop a float atomicAdd compiles to, and 6x faster than L2
//      REDs -- profiles/micro/atomics_micro.cu); float -> fixed conversion is one
//      FFMA (magic-number rounding) + one IADD;
//   d. the box is flushed once to the output volume with (vector) global REDs,
//      boundary conditions (index fold + sign) applied through per-axis tables.
// Integer accumulation is exact and order independent, so the shared-memory
// stage is deterministic; resolution is <= 2^-21 of the largest |value| * weight
// bound, i.e. comparable to float32 atomics.
//
// Replaces interpol/nd.py:147-213 (and iso1.py push) for the shapes that matter
// for throughput; the generic kernel in scatter.cu covers the rest.
#include <cstdio>
#include <cstdlib>
#include "tile_common.cuh"

namespace ib200 {

constexpr int kHist = 1024;            // coarse cells available for the overflow bound
constexpr float kMagic = 12582912.f;   // 1.5 * 2^23: float -> int by mantissa alignment
constexpr int kMagicBits = 0x4B400000;

template <typename T, int ORDER, int OP, int TX, int TY, int TZ, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
push_tile3d_kernel(const __grid_constant__ KParams kp, const T *__restrict__ img,
                   const T *__restrict__ grid, float *__restrict__ out, const int cap, const int vec_ok) {
    constexpr int NPT = TX * TY * TZ;
    constexpr int W = ORDER + 1;
    constexpr int NW = NT / 32;
    constexpr bool COUNT = (OP == OP_COUNT);
    constexpr int UI = ORDER <= 3 ? W : 1, UJ = ORDER <= 5 ? W : 1;   // keep the code of high orders compact
    constexpr int QBITS = 18;              // |value| quantisation for the bound: NPT * 2^18 < 2^31
    static_assert(NT == TY * TZ, "one thread per (y, z) column of the tile; x-planes are looped");
    static_assert(NPT <= 4096, "bound histogram would overflow");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *acc = reinterpret_cast<int *>(smem_raw);                       // [cap] fixed-point accumulators
    T *gtile = reinterpret_cast<T *>(acc + cap);                        // [NPT * 3] grid coordinates
    int *idx_tab = reinterpret_cast<int *>(reinterpret_cast<float *>(gtile) + (NPT * 3 * sizeof(T)) / 4);
    float *sgn_tab = reinterpret_cast<float *>(idx_tab + 3 * kMaxExt);
    int *red = reinterpret_cast<int *>(sgn_tab + 3 * kMaxExt);          // [TX][NW][6]
    PlaneBox *pb = reinterpret_cast<PlaneBox *>(red + TX * NW * 8);     // [TX]
    TileGeom *geoms = reinterpret_cast<TileGeom *>(pb + TX);            // [TX]
    int *nsub_p = reinterpret_cast<int *>(geoms + TX);
    int *hist = nsub_p + 4;                                             // [kHist]
    float *scal = reinterpret_cast<float *>(hist + kHist);              // [4] vmax, scale, 1/scale
    float *vals = scal + 4;                                             // [NPT] source values of the current channel
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- which tile ------------------------------------------------------
    const int ntx = (kp.pts_n[0] + TX - 1) / TX, nty = (kp.pts_n[1] + TY - 1) / TY, ntz = (kp.pts_n[2] + TZ - 1) / TZ;
    int tid = blockIdx.x;
    const int tz = tid % ntz; tid /= ntz;
    const int ty = tid % nty; tid /= nty;
    const int tx = tid % ntx; tid /= ntx;
    const i64 b = tid;
    const int x0 = tx * TX, y0 = ty * TY, z0 = tz * TZ;
    const int nzv = min(TZ, kp.pts_n[2] - z0);
    const int lz = threadIdx.x % TZ, ly = threadIdx.x / TZ;
    const bool col_ok = (y0 + ly < kp.pts_n[1]) && (lz < nzv);

    // ---- 1. + 2. grid coordinates (+ values of channel 0), bounding boxes, plan ----
    stage_grid_tile<T, TX, TY, TZ, NT>(kp, grid + b * kp.grid_sb, gtile, x0, y0, z0, nzv, vec_ok);
    auto stage_values = [&](i64 c) {       // vals[p * NT + tid] = img[b, c, x0 + p, y0 + ly, z0 + lz]
        if (COUNT) return;
        const T *src = img + b * kp.img_sb + c * kp.img_sc;
#pragma unroll
        for (int p = 0; p < TX; ++p) {
            float v = 0.f;
            if (col_ok && x0 + p < kp.pts_n[0])
                v = Traits<T>::load(src + (((x0 + p) * kp.pts_n[1] + (y0 + ly)) * kp.pts_n[2] + (z0 + lz)));
            vals[p * NT + threadIdx.x] = v;
        }
    };
    auto local_vmax = [&]() -> float {     // max |value| over this thread's points (own slots: no barrier needed)
        if (COUNT) return 1.f;
        float m = 0.f;
#pragma unroll
        for (int p = 0; p < TX; ++p) m = fmaxf(m, fabsf(vals[p * NT + threadIdx.x]));
        return m < 3e38f ? m : 3e38f;      // inf / NaN values: garbage in, garbage out
    };
    stage_values(0);
    cp_async_wait_all();
    __syncthreads();
    tile_add_identity<T, TX, NT>(kp, gtile, x0, y0 + ly, z0 + lz);
    plan_from_coords<T, ORDER, TX, NT>(kp, gtile, col_ok, x0, red, pb, geoms, nsub_p, cap, local_vmax());
    float vmax_tile = __int_as_float(red[kExtraSlot]);
    i64 cur_c = 0;                         // channel whose values are staged in `vals`
    const int nsub = *nsub_p;
    const int per = TX / nsub;
    const float w3 = max_weight(ORDER) * max_weight(ORDER) * max_weight(ORDER);

    for (int s = 0; s < nsub; ++s) {
        const TileGeom g = geoms[s];
        if (g.ext[0] == 0 && g.fits) continue;                          // nothing in bounds
        // coarse cells of W^3 support starts
        const int nc0 = (g.ext[0] - W) / W + 1, nc1 = (g.ext[1] - W) / W + 1, nc2 = (g.ext[2] - W) / W + 1;
        const bool tiled = g.fits && (i64)nc0 * nc1 * nc2 <= kHist;
        if (tiled) { __syncthreads(); build_tables<NT>(kp, g, idx_tab, sgn_tab); }

        for (i64 c = 0; c < kp.channels; ++c) {
            const T *src = COUNT ? nullptr : img + b * kp.img_sb + c * kp.img_sc;
            float *dst = out + (b * kp.channels + c) * kp.vol_total;
            auto value = [&](int p) -> float { return COUNT ? 1.f : vals[p * NT + threadIdx.x]; };
            if (c != cur_c) {              // another channel: restage the values, redo the maximum
                cur_c = c;
                __syncthreads();
                stage_values(c);
                const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(local_vmax()));
                if (lane == 0) red[warp] = (int)m;
                __syncthreads();
                unsigned mm = 0;
#pragma unroll
                for (int w = 0; w < NW; ++w) mm = max(mm, (unsigned)red[w]);
                vmax_tile = __uint_as_float(mm);
            }
            if (tiled) {
                // ---- a. zero the accumulators and the histogram, find max |value| -----
                __syncthreads();                                        // previous flush done
                {
                    const int n4 = (g.sxy * g.ext[0]) >> 2;
                    int4 *a4 = reinterpret_cast<int4 *>(acc);
                    for (int q = threadIdx.x; q < n4; q += NT) a4[q] = make_int4(0, 0, 0, 0);
                    for (int q = threadIdx.x; q < nc0 * nc1 * nc2; q += NT) hist[q] = 0;
                }
                const float vmax = vmax_tile;   // max |value| over the tile (>= the group's: still a bound)
                __syncthreads();                                        // zeros visible, red reusable
                if (vmax > 0.f) {
                    // ---- b. rigorous bound on any accumulator -> power-of-two scale --------
                    const float qs = (float)(1 << QBITS) / vmax;
#pragma unroll 1
                    for (int p = s * per; p < (s + 1) * per; ++p) {
                        int i0[3];
                        if (support_start<T, ORDER, NT>(kp, gtile, p, col_ok && x0 + p < kp.pts_n[0], i0) == 1) {
                            const int cell = (((i0[0] - g.lo[0]) / W) * nc1 + (i0[1] - g.lo[1]) / W) * nc2 + (i0[2] - g.lo[2]) / W;
                            const int qv = min(1 << QBITS, (int)ceilf(fabsf(value(p)) * qs));
                            atomicAdd(&hist[cell], qv);
                        }
                    }
                    __syncthreads();
                    int wmax = 0;
                    for (int q = threadIdx.x; q < nc0 * nc1 * nc2; q += NT) {
                        const int c2 = q % nc2, c1 = (q / nc2) % nc1, c0 = q / (nc2 * nc1);
                        int sum = 0;
#pragma unroll
                        for (int d0 = 0; d0 < 2; ++d0)
#pragma unroll
                            for (int d1 = 0; d1 < 2; ++d1)
#pragma unroll
                                for (int d2 = 0; d2 < 2; ++d2)
                                    if (c0 + d0 < nc0 && c1 + d1 < nc1 && c2 + d2 < nc2)
                                        sum += hist[((c0 + d0) * nc1 + c1 + d1) * nc2 + c2 + d2];
                        wmax = max(wmax, sum);
                    }
                    wmax = __reduce_max_sync(0xffffffffu, wmax);
                    if (lane == 0) red[warp] = wmax;
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        int m = 1;
                        for (int w = 0; w < NW; ++w) m = max(m, red[w]);
                        // |any accumulator| <= M = m * 2^-QBITS * vmax * w3  (in value units)
                        const float M = (float)m * (1.f / (float)(1 << QBITS)) * vmax * w3;
                        int e1, e2;
                        frexpf(M, &e1);                 // M < 2^e1
                        frexpf(vmax * w3, &e2);         // single contribution < 2^e2
                        int k = min(30 - e1, 21 - e2);          // sums < 2^30, contributions < 2^21
                        k = max(-120, min(120, k));
                        scal[0] = vmax; scal[1] = ldexpf(1.f, k); scal[2] = ldexpf(1.f, -k);
                    }
                    __syncthreads();
                    const float scale = scal[1];
                    // ---- c. integer shared-memory atomics -----------------------------------
#pragma unroll 1
                    for (int p = s * per; p < (s + 1) * per; ++p) {
                        int i0[3];
                        if (support_start<T, ORDER, NT>(kp, gtile, p, col_ok && x0 + p < kp.pts_n[0], i0) != 1) continue;
                        const T *gp = gtile + (p * NT + threadIdx.x) * 3;
                        float wx[W], wy[W], wz[W];
                        fast_weights<ORDER>((float)gp[0] - (float)i0[0], wx);
                        fast_weights<ORDER>((float)gp[1] - (float)i0[1], wy);
                        fast_weights<ORDER>((float)gp[2] - (float)i0[2], wz);
                        const float v = value(p) * scale;
                        int *ri = acc + (i0[0] - g.lo[0]) * g.sxy + (i0[1] - g.lo[1]) * g.sz + (i0[2] - g.lo[2]);
#pragma unroll UI
                        for (int i = 0; i < W; ++i) {
                            int *rj = ri;
                            const float vi = v * wx[i];
#pragma unroll UJ
                            for (int j = 0; j < W; ++j) {
                                const float vij = vi * wy[j];
#pragma unroll
                                for (int k = 0; k < W; ++k)
                                    atomicAdd(rj + k, __float_as_int(fmaf(vij, wz[k], kMagic)) - kMagicBits);
                                rj += g.sz;
                            }
                            ri += g.sxy;
                        }
                    }
                    __syncthreads();
                    // ---- d. flush the box: fixed -> float, fold + sign, global REDs -----------
                    const float inv = scal[2];
                    const int nrows = g.ext[0] * g.ext[1];
                    const int total = nrows * g.vpr;
                    const int zlo = (kp.bound[2] == IB200_BOUND_DST1) ? 1 : 0;
                    for (int q = threadIdx.x; q < total; q += NT) {
                        const int r = fast_div(q, g.inv_vpr), v4 = q - r * g.vpr;
                        const int a = fast_div(r, g.inv_e1), bb = r - a * g.ext[1];
                        const int4 iv = *reinterpret_cast<const int4 *>(acc + a * g.sxy + bb * g.sz + v4 * 4);
                        if ((iv.x | iv.y | iv.z | iv.w) == 0) continue;
                        const float rowsgn = sgn_tab[a] * sgn_tab[kMaxExt + bb];
                        if (rowsgn == 0.f) continue;
                        const int rowbase = idx_tab[a] + idx_tab[kMaxExt + bb];
                        const float f = inv * rowsgn;
                        const int zs = g.lo[2] + v4 * 4;
                        if (vec_ok && zs >= zlo && zs + 3 <= kp.vol_n[2] - 1) {
                            atomicAdd(reinterpret_cast<float4 *>(dst + rowbase + zs),
                                      make_float4(f * (float)iv.x, f * (float)iv.y, f * (float)iv.z, f * (float)iv.w));
                        } else {
                            const int ivs[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int ee = v4 * 4 + e;
                                if (ivs[e] != 0 && ee < g.ext[2]) {
                                    const float sg = sgn_tab[2 * kMaxExt + ee];
                                    if (sg != 0.f) atomicAdd(dst + rowbase + idx_tab[2 * kMaxExt + ee], f * sg * (float)ivs[e]);
                                }
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
                    const float cc[3] = {(float)gp[0], (float)gp[1], (float)gp[2]};
                    if (!inbounds<float, 3>(kp, cc)) continue;
                    Axis<float, W> ax[3];
                    bool ok = setup_axis<float, ORDER, 0, W>(ax[0], cc[0], ORDER, kp.bound[0], kp.vol_n[0], (int)kp.vol_s[0], kp);
                    ok = setup_axis<float, ORDER, 0, W>(ax[1], cc[1], ORDER, kp.bound[1], kp.vol_n[1], (int)kp.vol_s[1], kp) && ok;
                    ok = setup_axis<float, ORDER, 0, W>(ax[2], cc[2], ORDER, kp.bound[2], kp.vol_n[2], (int)kp.vol_s[2], kp) && ok;
                    if (!ok) continue;
                    const float v = value(p);
#pragma unroll
                    for (int i = 0; i < W; ++i)
#pragma unroll
                        for (int j = 0; j < W; ++j) {
                            const float vij = v * ax[0].w[i] * ax[1].w[j];
#pragma unroll
                            for (int k = 0; k < W; ++k)
                                atomicAdd(dst + ax[0].off[i] + ax[1].off[j] + ax[2].off[k], vij * ax[2].w[k]);
                        }
                }
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
convert_from_f32_kernel(const float *__restrict__ src, T *__restrict__ dst, i64 n) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
        Traits<T>::store(dst + i, src[i]);
}

// ---------------------------------------------------------------- launch --

template <typename T, int ORDER, int OP, int TX, int TY, int TZ, int NT, int MINB>
static int launch_push_tile_cfg(const KParams &kp, const void *img, const void *grid, float *out, cudaStream_t stream,
                                size_t smem_total) {
    const size_t fixed = (size_t)TX * TY * TZ * 3 * sizeof(T) + 3 * kMaxExt * (sizeof(int) + sizeof(float)) +
                         TX * (NT / 32) * 8 * sizeof(int) + TX * (sizeof(PlaneBox) + sizeof(TileGeom)) +
                         (kHist + 8) * sizeof(int) + (size_t)TX * TY * TZ * sizeof(float) + 64;
    const int cap = (int)((smem_total - fixed) / sizeof(int)) & ~31;
    const i64 ntiles = kp.batch * ((kp.pts_n[0] + TX - 1) / TX) * ((kp.pts_n[1] + TY - 1) / TY) * ((kp.pts_n[2] + TZ - 1) / TZ);
    if (ntiles == 0) return 1;
    if (ntiles > 0x7fffffffLL) return 0;
    const int ev = 16 / (int)sizeof(T);
    // 16-byte staging of the grid, 16-byte vector REDs into the (dense, float32) target volume
    const bool vec_ok = ((uintptr_t)grid % 16 == 0) && ((uintptr_t)out % 16 == 0) && ((kp.pts_n[2] * 3) % ev == 0) &&
                        (kp.grid_sb % ev == 0) && (kp.vol_n[2] % 4 == 0);
    auto kern = push_tile3d_kernel<T, ORDER, OP, TX, TY, TZ, NT, MINB>;
    IB200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_total));
    kern<<<(unsigned)ntiles, NT, smem_total, stream>>>(kp, (const T *)img, (const T *)grid, out, cap, vec_ok ? 1 : 0);
    static thread_local char name[64];
    snprintf(name, sizeof(name), "%s_tile3d_o%d_%dx%dx%d", OP == OP_COUNT ? "count" : "push", ORDER, TX, TY, TZ);
    note_launch(name);
    IB200_CUDA_CHECK(cudaGetLastError());
    return 1;
}

template <typename T, int ORDER, int OP>
static int launch_push_tile(const KParams &kp, const void *img, const void *grid, float *out, cudaStream_t stream) {
    return launch_push_tile_cfg<T, ORDER, OP, 8, 8, 32, 256, 2>(kp, img, grid, out, stream, 110 * 1024);
}

template <typename T, int OP>
static int dispatch_push_tile(const KParams &kp, const void *img, const void *grid, float *out, cudaStream_t stream) {
    switch (kp.order[0]) {
    case 1: return launch_push_tile<T, 1, OP>(kp, img, grid, out, stream);
    case 2: return launch_push_tile<T, 2, OP>(kp, img, grid, out, stream);
    case 3: return launch_push_tile<T, 3, OP>(kp, img, grid, out, stream);
    case 4: return launch_push_tile<T, 4, OP>(kp, img, grid, out, stream);
    case 5: return launch_push_tile<T, 5, OP>(kp, img, grid, out, stream);
    case 6: return launch_push_tile<T, 6, OP>(kp, img, grid, out, stream);
    case 7: return launch_push_tile<T, 7, OP>(kp, img, grid, out, stream);
    }
    return 0;
}

// `acc` is the float32 accumulation target (the output itself for F32, the scratch
// volume for 16-bit storage), already zero-filled by the caller.
bool push_tiled_applicable(int op, const KParams &kp, int dtype) {
    if (op != OP_PUSH && op != OP_COUNT) return false;
    if (dtype != IB200_F32 && dtype != IB200_F16) return false;
    if (kp.dim != 3 || !kp.pts_dense) return false;
    if (kp.order[0] != kp.order[1] || kp.order[0] != kp.order[2]) return false;
    if (kp.order[0] < 1 || kp.order[0] > 7) return false;
    if (kp.pts_total < 32768) return false;
    if (kp.pts_total * 3 > 0x7fffffffLL) return false;
    // displacement fields in 16-bit storage: the generic kernels add the lattice index in float32
    if ((kp.flags & IB200_FLAG_DISPLACEMENT) && dtype != IB200_F32) return false;
    return true;
}

int try_push_tiled(int op, const KParams &kp, int dtype, const void *img, const void *grid, void *acc,
                   cudaStream_t stream) {
    if (!push_tiled_applicable(op, kp, dtype)) return 0;
    float *out = (float *)acc;
    if (op == OP_PUSH) {
        switch (dtype) {
        case IB200_F32: return dispatch_push_tile<float, OP_PUSH>(kp, img, grid, out, stream);
        case IB200_F16: return dispatch_push_tile<__half, OP_PUSH>(kp, img, grid, out, stream);
        }
    } else {
        switch (dtype) {
        case IB200_F32: return dispatch_push_tile<float, OP_COUNT>(kp, img, grid, out, stream);
        case IB200_F16: return dispatch_push_tile<__half, OP_COUNT>(kp, img, grid, out, stream);
        }
    }
    return 0;
}

int convert_from_f32(int dtype, const void *src, void *dst, i64 n, cudaStream_t stream) {
    if (n == 0) return IB200_OK;
    i64 blocks = (n + 255) / 256;
    if (blocks > (i64)kNumSMs * 32) blocks = (i64)kNumSMs * 32;
    if (dtype == IB200_F16) convert_from_f32_kernel<__half><<<(unsigned)blocks, 256, 0, stream>>>((const float *)src, (__half *)dst, n);
    else if (dtype == IB200_BF16) convert_from_f32_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, stream>>>((const float *)src, (__nv_bfloat16 *)dst, n);
    else return IB200_ERR_DTYPE;
    note_launch("convert_f32_to_16bit");
    IB200_CUDA_CHECK(cudaGetLastError());
    return IB200_OK;
}

}  // namespace ib200
