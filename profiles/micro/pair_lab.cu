// Pair laboratory (B200, via gpurun): the paired tap loop of csrc/pull_pair.cuh in isolation -- box and
// coordinates resident in shared memory, no staging, no pipeline -- against the one-voxel-per-lane loop of
// round 1, in cycles per row of 32 voxels per SM as a function of the local stretch (along z) and shear
// (x / y drift along z: how often the two voxels of a pair sit in different x / y cells) of the deformation.
//
//   MODE 0  one lane per voxel, 64 LDS.32 per voxel                                   [round 1 tap loop]
//   MODE 6  two z-neighbours per lane, window of 6 words = 3 LDS.64 per (x, y) row of the support
//   MODE 8  two z-neighbours per lane, window of 8 words = 2 LDS.128 per (x, y) row
// Every mode writes its result; the host compares modes 6 / 8 with mode 0 (max abs difference).
//
// nvcc -O3 -std=c++17 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -I../../torch-interpol_b200/csrc pair_lab.cu -o pair_lab
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "pull_pair.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

using namespace ib200;

constexpr int NT = 512, NW = NT / 32;
constexpr int BZ = 64, BY = 16, BX = 20, PLANE = BZ * BY, BOXW = PLANE * BX;
constexpr int TX = 8, TY = 8, TZ = 32, NROWS = TX * TY;

__device__ void fill_coords(float *gt, float stretch, float shear, int variant) {
    for (int i = threadIdx.x; i < NROWS * TZ; i += NT) {
        const int r = i / TZ, z = i % TZ, p = r / TY, ly = r % TY;
        const float wob = variant ? 0.35f * __sinf(0.7f * z + 0.9f * ly + 1.3f * p) : 0.f;   // local wiggle (variant 1)
        gt[i * 3 + 0] = 1.3f + p * 1.02f + shear * z + 0.07f * ly + 0.5f * wob;
        gt[i * 3 + 1] = 1.7f + ly * 0.97f + 0.8f * shear * z + 0.05f * p - 0.4f * wob;
        gt[i * 3 + 2] = 1.2f + stretch * z + 0.11f * ly + 0.06f * p + wob;
    }
}

template <int MODE>
__global__ void __launch_bounds__(NT, 1) pull_lab(float stretch, float shear, int variant, int ntiles, float *out, long long *cyc) {
    extern __shared__ __align__(16) float smem[];
    float *box = smem, *gt = smem + BOXW;
    __shared__ long long t0s, t1s;
    for (int i = threadIdx.x; i < BOXW; i += NT) box[i] = (float)((i * 2654435761u) >> 20) * (1.f / 4096.f) - 0.5f;
    fill_coords(gt, stretch, shear, variant);
    if (threadIdx.x == 0) { t0s = 0x7fffffffffffffffLL; t1s = 0; }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t0 = clock64();
    for (int t = 0; t < ntiles; ++t) {
        float *dst = out + ((size_t)blockIdx.x * 2 + (t & 1)) * NROWS * TZ;
        if constexpr (MODE == 0) {
            for (int r = warp; r < NROWS; r += NW) {
                const float *gp = gt + (r * TZ + lane) * 3;
                const float c0 = gp[0], c1 = gp[1], c2 = gp[2];
                const float f0 = floorf(c0 - 1.f), f1 = floorf(c1 - 1.f), f2 = floorf(c2 - 1.f);
                float wx[4], wy[4], wz[4];
                fast_weights<3>(c0 - f0, wx); fast_weights<3>(c1 - f1, wy); fast_weights<3>(c2 - f2, wz);
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
            }
        } else {
            // a warp takes two z-rows: half warp h -> row 2 * rp + h, lane m of the half -> voxels 2m, 2m + 1
            const int h = lane >> 4, m = lane & 15;
            for (int rp = warp; rp < NROWS / 2; rp += NW) {
                const int r = 2 * rp + h;
                const float2 *gp = reinterpret_cast<const float2 *>(gt + (r * TZ + 2 * m) * 3);
                const float2 g0 = gp[0], g1 = gp[1], g2 = gp[2];
                float ra, rb;
                pull_pair_eval<3, MODE, PLANE, BZ>(box, 0, 0, 0, g0.x, g0.y, g1.x, g1.y, g2.x, g2.y, true, true, ra, rb);
                *reinterpret_cast<float2 *>(dst + r * TZ + 2 * m) = make_float2(ra, rb);
            }
        }
    }
    const long long t1 = clock64();
    atomicMin(&t0s, t0); atomicMax(&t1s, t1);
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1s - t0s;
}

int main() {
    const int blocks = 148, ntiles = 64;
    const size_t smem = (size_t)(BOXW + NROWS * TZ * 3) * 4 + 64;
    const size_t nout = (size_t)blocks * 2 * NROWS * TZ;
    float *d_out[3]; long long *d_cyc;
    for (int i = 0; i < 3; ++i) CK(cudaMalloc(&d_out[i], nout * 4));
    CK(cudaMalloc(&d_cyc, blocks * 8));
    CK(cudaFuncSetAttribute(pull_lab<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(pull_lab<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(pull_lab<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<float> h0(nout), h1(nout);
    auto clk = [&]() {
        CK(cudaDeviceSynchronize());
        long long h[blocks]; CK(cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost));
        double c = 0; for (int i = 0; i < blocks; ++i) c += (double)h[i];
        return c / blocks / ((double)ntiles * NROWS);
    };
    auto diff = [&](int which) {
        CK(cudaMemcpy(h0.data(), d_out[0], nout * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h1.data(), d_out[which], nout * 4, cudaMemcpyDeviceToHost));
        double d = 0; for (size_t i = 0; i < nout; ++i) d = fmax(d, fabs((double)h0[i] - (double)h1[i]));
        return d;
    };
    const float stretches[] = {0.8f, 0.95f, 1.0f, 1.05f, 1.2f, 1.4f};
    const float shears[] = {0.0f, 0.05f, 0.12f, 0.25f};
    printf("clk per row of 32 voxels per SM (256^3 at 1.9 GHz: clk * 1.865 us)\n");
    printf("%7s %7s %7s | %10s %10s %10s | %9s %9s\n", "variant", "stretch", "shear", "1 vox/lane", "pair LDS64", "pair LDS128", "err(64)", "err(128)");
    for (int variant = 0; variant < 2; ++variant)
        for (float sh : shears)
            for (float s : stretches) {
                for (int rep = 0; rep < 2; ++rep) pull_lab<0><<<blocks, NT, smem>>>(s, sh, variant, ntiles, d_out[0], d_cyc);
                const double c0 = clk();
                for (int rep = 0; rep < 2; ++rep) pull_lab<6><<<blocks, NT, smem>>>(s, sh, variant, ntiles, d_out[1], d_cyc);
                const double c6 = clk();
                for (int rep = 0; rep < 2; ++rep) pull_lab<8><<<blocks, NT, smem>>>(s, sh, variant, ntiles, d_out[2], d_cyc);
                const double c8 = clk();
                printf("%7d %7.2f %7.2f | %10.1f %10.1f %10.1f | %9.2e %9.2e\n", variant, s, sh, c0, c6, c8, diff(1), diff(2)); fflush(stdout);
            }
    return 0;
}
