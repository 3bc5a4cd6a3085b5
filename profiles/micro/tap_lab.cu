// Tap-loop laboratory (B200, via gpurun): the inner loops of the persistent pull / push kernels in isolation
// -- box and coordinates resident in shared memory, no staging, no pipeline -- to measure what a lane <-> voxel
// mapping costs in cycles per row of 32 voxels per SM, as a function of the local stretch of the deformation
// along z (stretch > 1: the 32 supports of a z-row span more than 32 banks; stretch < 1: neighbouring sources
// collide on one accumulator word).
//
//   MODE 0  one lane per voxel, 32 voxels per warp instruction ((ORDER+1)^3 taps per lane)        [round 1]
//   MODE 1  two lanes per voxel ("k-split"): lane 2m takes the z taps {0, 1}, lane 2m+1 the taps {2, 3} of
//           voxel m; 16 consecutive voxels per warp instruction, span 15 s + 4 <= 32 words up to s = 1.87
//   MODE 2  (push only) k-split with the 16 voxels of an instruction taken at stride 2 along z (even voxels,
//           then odd voxels): no two lanes share an accumulator word down to s = 0.5
//   MODE 3  (push only) one lane per source, warp-aggregated: a source in the same cell as its lower z-neighbour
//           rides along with it (4 SHFL, second weight set in the leader) instead of colliding on its 64 words
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int NT = 512, NW = NT / 32;
constexpr int BZ = 64, BY = 16, BX = 16, PLANE = BZ * BY, BOXW = PLANE * BX;
constexpr int TX = 8, TY = 8, TZ = 32, NROWS = TX * TY;
constexpr float kMagic = 12582912.f;
constexpr int kMagicBits = 0x4B400000;

__device__ __forceinline__ void w3(float t, float (&w)[4]) {       // cubic, t in [1, 2)
    const float u = t - 1.f, a = 2.f - t, u2 = u * u, a2 = a * a;
    w[0] = a2 * (a * (1.f / 6.f));
    w[1] = fmaf(u2, fmaf(u, 0.5f, -1.f), 2.f / 3.f);
    w[2] = fmaf(a2, fmaf(a, 0.5f, -1.f), 2.f / 3.f);
    w[3] = u2 * (u * (1.f / 6.f));
}

__device__ void fill_coords(float *gt, float stretch, float shear) {
    // smooth synthetic deformation in box coordinates: x / y drift slowly along z (cell crossings every
    // few voxels), z advances by `stretch` per voxel
    for (int i = threadIdx.x; i < NROWS * TZ; i += NT) {
        const int r = i / TZ, z = i % TZ, p = r / TY, ly = r % TY;
        gt[i * 3 + 0] = 1.3f + p * 1.02f + shear * z + 0.07f * ly;
        gt[i * 3 + 1] = 1.7f + ly * 0.97f + 0.8f * shear * z + 0.05f * p;
        gt[i * 3 + 2] = 1.2f + stretch * z + 0.11f * ly + 0.06f * p;
    }
}

template <int MODE>
__global__ void __launch_bounds__(NT, 1) pull_lab(float stretch, float shear, int ntiles, float *out, long long *cyc) {
    extern __shared__ __align__(16) float smem[];
    float *box = smem, *gt = smem + BOXW;
    __shared__ long long t0s, t1s;
    for (int i = threadIdx.x; i < BOXW; i += NT) box[i] = (float)((i * 2654435761u) >> 20) * (1.f / 4096.f) - 0.5f;
    fill_coords(gt, stretch, shear);
    if (threadIdx.x == 0) { t0s = 0x7fffffffffffffffLL; t1s = 0; }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t0 = clock64();
    for (int t = 0; t < ntiles; ++t) {
        float *dst = out + ((size_t)blockIdx.x * 2 + (t & 1)) * NROWS * TZ;
        for (int r = warp; r < NROWS; r += NW) {
            if (MODE == 0) {
                const float *gp = gt + (r * TZ + lane) * 3;
                const float c0 = gp[0], c1 = gp[1], c2 = gp[2];
                const float f0 = floorf(c0 - 1.f), f1 = floorf(c1 - 1.f), f2 = floorf(c2 - 1.f);
                float wx[4], wy[4], wz[4];
                w3(c0 - f0, wx); w3(c1 - f1, wy); w3(c2 - f2, wz);
                const float *rk = box + (int)f0 * PLANE + (int)f1 * BZ + (int)f2;
                float2 wz2[4], acc2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 4; ++k) wz2[k] = make_float2(wz[k], wz[k]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int jj = 0; jj < 4; jj += 2) {
                        float2 t2 = make_float2(0.f, 0.f);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            t2 = __ffma2_rn(wz2[k], make_float2(rk[i * PLANE + jj * BZ + k], rk[i * PLANE + (jj + 1) * BZ + k]), t2);
                        s2 = __ffma2_rn(make_float2(wy[jj], wy[jj + 1]), t2, s2);
                    }
                    acc2 = __ffma2_rn(make_float2(wx[i], wx[i]), s2, acc2);
                }
                dst[r * TZ + lane] = acc2.x + acc2.y;
            } else {
                const int v = lane >> 1, h = lane & 1;
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
                    const int z = pass * 16 + v;
                    const float *gp = gt + (r * TZ + z) * 3;
                    const float c0 = gp[0], c1 = gp[1], c2 = gp[2];
                    const float f0 = floorf(c0 - 1.f), f1 = floorf(c1 - 1.f), f2 = floorf(c2 - 1.f);
                    float wx[4], wy[4], wz[4];
                    w3(c0 - f0, wx); w3(c1 - f1, wy); w3(c2 - f2, wz);
                    const float wa = h ? wz[2] : wz[0], wb = h ? wz[3] : wz[1];
                    const float *rk = box + (int)f0 * PLANE + (int)f1 * BZ + (int)f2 + 2 * h;
                    const float2 wa2 = make_float2(wa, wa), wb2 = make_float2(wb, wb);
                    float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
                        for (int jj = 0; jj < 4; jj += 2) {
                            float2 t2 = make_float2(0.f, 0.f);
                            t2 = __ffma2_rn(wa2, make_float2(rk[i * PLANE + jj * BZ], rk[i * PLANE + (jj + 1) * BZ]), t2);
                            t2 = __ffma2_rn(wb2, make_float2(rk[i * PLANE + jj * BZ + 1], rk[i * PLANE + (jj + 1) * BZ + 1]), t2);
                            s2 = __ffma2_rn(make_float2(wy[jj], wy[jj + 1]), t2, s2);
                        }
                        acc2 = __ffma2_rn(make_float2(wx[i], wx[i]), s2, acc2);
                    }
                    float acc = acc2.x + acc2.y;
                    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                    if (h == 0) dst[r * TZ + z] = acc;
                }
            }
        }
    }
    const long long t1 = clock64();
    atomicMin(&t0s, t0); atomicMax(&t1s, t1);
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1s - t0s;
}

template <int MODE>
__global__ void __launch_bounds__(NT, 1) push_lab(float stretch, float shear, int ntiles, float *out, long long *cyc) {
    extern __shared__ __align__(16) float smem[];
    int *box = reinterpret_cast<int *>(smem);
    float *gt = smem + BOXW, *vals = gt + NROWS * TZ * 3;
    __shared__ long long t0s, t1s;
    for (int i = threadIdx.x; i < BOXW; i += NT) box[i] = 0;
    for (int i = threadIdx.x; i < NROWS * TZ; i += NT) vals[i] = (float)((i * 2654435761u) >> 20) * (1.f / 4096.f) - 0.5f;
    fill_coords(gt, stretch, shear);
    if (threadIdx.x == 0) { t0s = 0x7fffffffffffffffLL; t1s = 0; }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float scale = 1048576.f;
    const long long t0 = clock64();
    for (int t = 0; t < ntiles; ++t) {
        for (int r = warp; r < NROWS; r += NW) {
            if (MODE == 0) {
                const float *gp = gt + (r * TZ + lane) * 3;
                const float c0 = gp[0], c1 = gp[1], c2 = gp[2];
                const float f0 = floorf(c0 - 1.f), f1 = floorf(c1 - 1.f), f2 = floorf(c2 - 1.f);
                float wx[4], wy[4], wz[4];
                w3(c0 - f0, wx); w3(c1 - f1, wy); w3(c2 - f2, wz);
                const float val = vals[r * TZ + lane] * scale;
                int *rk = box + (int)f0 * PLANE + (int)f1 * BZ + (int)f2;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float vi = val * wx[i];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float vij = vi * wy[j];
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            atomicAdd(rk + i * PLANE + j * BZ + k, __float_as_int(fmaf(vij, wz[k], kMagic)) - kMagicBits);
                    }
                }
            } else if (MODE == 3) {
                // warp-aggregated: a source whose support starts in the same cell as its lower z-neighbour's hands its
                // value and coordinates to that neighbour (4 SHFL) and issues no atomics; the neighbour adds both
                // contributions before the float -> fixed conversion
                const float *gp = gt + (r * TZ + lane) * 3;
                const float c0 = gp[0], c1 = gp[1], c2 = gp[2];
                const float f0 = floorf(c0 - 1.f), f1 = floorf(c1 - 1.f), f2 = floorf(c2 - 1.f);
                const int cell = (int)f0 * PLANE + (int)f1 * BZ + (int)f2;
                const int pcell = __shfl_up_sync(0xffffffffu, cell, 1);
                const bool sp = lane > 0 && pcell == cell;
                const bool spp = __shfl_up_sync(0xffffffffu, (int)sp, 1) != 0;
                const bool follower = sp && !(lane > 1 && spp);
                const bool lead2 = __shfl_down_sync(0xffffffffu, (int)follower, 1) != 0 && lane < 31;
                const float val = vals[r * TZ + lane] * scale;
                const float d0 = __shfl_down_sync(0xffffffffu, c0, 1), d1 = __shfl_down_sync(0xffffffffu, c1, 1),
                            d2 = __shfl_down_sync(0xffffffffu, c2, 1);
                const float vb = lead2 ? __shfl_down_sync(0xffffffffu, val, 1) : 0.f * __shfl_down_sync(0xffffffffu, val, 1);
                if (!follower) {
                    float wx[4], wy[4], wz[4], ux[4], uy[4], uz[4];
                    w3(c0 - f0, wx); w3(c1 - f1, wy); w3(c2 - f2, wz);
                    w3(lead2 ? d0 - f0 : 1.5f, ux); w3(lead2 ? d1 - f1 : 1.5f, uy); w3(lead2 ? d2 - f2 : 1.5f, uz);
                    int *rk = box + cell;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float vi = val * wx[i], ui = vb * ux[i];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float vij = vi * wy[j], uij = ui * uy[j];
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                atomicAdd(rk + i * PLANE + j * BZ + k, __float_as_int(fmaf(vij, wz[k], fmaf(uij, uz[k], kMagic))) - kMagicBits);
                        }
                    }
                }
            } else {
                const int v = lane >> 1, h = lane & 1;
#pragma unroll 1
                for (int pass = 0; pass < 2; ++pass) {
                    const int z = MODE == 1 ? pass * 16 + v : 2 * v + pass;
                    const float *gp = gt + (r * TZ + z) * 3;
                    const float c0 = gp[0], c1 = gp[1], c2 = gp[2];
                    const float f0 = floorf(c0 - 1.f), f1 = floorf(c1 - 1.f), f2 = floorf(c2 - 1.f);
                    float wx[4], wy[4], wz[4];
                    w3(c0 - f0, wx); w3(c1 - f1, wy); w3(c2 - f2, wz);
                    const float wa = h ? wz[2] : wz[0], wb = h ? wz[3] : wz[1];
                    const float val = vals[r * TZ + z] * scale;
                    int *rk = box + (int)f0 * PLANE + (int)f1 * BZ + (int)f2 + 2 * h;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float vi = val * wx[i];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float vij = vi * wy[j];
                            atomicAdd(rk + i * PLANE + j * BZ, __float_as_int(fmaf(vij, wa, kMagic)) - kMagicBits);
                            atomicAdd(rk + i * PLANE + j * BZ + 1, __float_as_int(fmaf(vij, wb, kMagic)) - kMagicBits);
                        }
                    }
                }
            }
        }
    }
    const long long t1 = clock64();
    atomicMin(&t0s, t0); atomicMax(&t1s, t1);
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1s - t0s;
    __syncthreads();
    float a = 0.f;
    for (int i = threadIdx.x; i < BOXW; i += NT) a += (float)box[i];
    out[blockIdx.x * NT + threadIdx.x] = a;
}

int main() {
    const int blocks = 148, ntiles = 64;
    const size_t smem = (size_t)(BOXW + NROWS * TZ * 4) * 4 + 64;
    float *d_out; long long *d_cyc;
    CK(cudaMalloc(&d_out, (size_t)blocks * 2 * NROWS * TZ * 4)); CK(cudaMalloc(&d_cyc, blocks * 8));
    CK(cudaFuncSetAttribute(pull_lab<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(pull_lab<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(push_lab<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(push_lab<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(push_lab<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(push_lab<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto report = [&](const char *name, float s, float sh) {
        CK(cudaDeviceSynchronize());
        long long h[blocks]; CK(cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost));
        double c = 0; for (int i = 0; i < blocks; ++i) c += (double)h[i];
        c /= blocks;
        const double per_row = c / ((double)ntiles * NROWS);
        printf("%-34s stretch %.2f shear %.2f : %7.1f clk / row of 32 voxels / SM  (%.2f clk/voxel; 256^3 at 1.9 GHz: %.0f us)\n",
               name, s, sh, per_row, per_row / 32, per_row / 32 * 16777216.0 / 148 / 1.9e9 * 1e6);
    };
    const float stretches[] = {0.8f, 0.95f, 1.0f, 1.05f, 1.2f, 1.4f};
    const float shears[] = {0.0f, 0.12f};
    for (float sh : shears)
        for (float s : stretches) {
            for (int rep = 0; rep < 2; ++rep) pull_lab<0><<<blocks, NT, smem>>>(s, sh, ntiles, d_out, d_cyc);
            report("pull  one lane per voxel", s, sh);
            for (int rep = 0; rep < 2; ++rep) pull_lab<1><<<blocks, NT, smem>>>(s, sh, ntiles, d_out, d_cyc);
            report("pull  k-split (2 lanes per voxel)", s, sh);
            for (int rep = 0; rep < 2; ++rep) push_lab<0><<<blocks, NT, smem>>>(s, sh, ntiles, d_out, d_cyc);
            report("push  one lane per voxel", s, sh);
            for (int rep = 0; rep < 2; ++rep) push_lab<1><<<blocks, NT, smem>>>(s, sh, ntiles, d_out, d_cyc);
            report("push  k-split adjacent", s, sh);
            for (int rep = 0; rep < 2; ++rep) push_lab<2><<<blocks, NT, smem>>>(s, sh, ntiles, d_out, d_cyc);
            report("push  k-split stride 2", s, sh);
            for (int rep = 0; rep < 2; ++rep) push_lab<3><<<blocks, NT, smem>>>(s, sh, ntiles, d_out, d_cyc);
            report("push  warp-aggregated neighbours", s, sh);
        }
    return 0;
}
