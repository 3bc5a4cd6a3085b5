// Micro-benchmarks that drove the push design (run on the B200 via gpurun):
// throughput of shared-memory float / int / int64 atomics, global RED (scalar
// and vector) and plain LDS in a splat-like access pattern: every lane adds
// NT taps to consecutive addresses starting at a per-lane base (lanes own
// consecutive bases, like consecutive voxels along the fastest axis).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int TILE = 16384;   // floats of shared memory tile
constexpr int ITERS = 64;     // points per thread
constexpr int ROWS = 16, TAPS = 4;

__device__ __forceinline__ int lane_base(int it, int tid) {
    // pseudo "deformation": rows advance with the iteration, lanes consecutive
    return ((it * 37 + (tid >> 5) * 211) % 9000) + (tid & 31);
}

__global__ void k_smem_float(float *out, float v) {
    extern __shared__ float s[];
    for (int i = threadIdx.x; i < TILE; i += blockDim.x) s[i] = 0.f;
    __syncthreads();
    for (int it = 0; it < ITERS; ++it) {
        int b = lane_base(it, threadIdx.x);
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int k = 0; k < TAPS; ++k) atomicAdd(&s[b + r * 45 + k], v * (r + k));
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = s[17];
}

__global__ void k_smem_int(float *out, float v) {
    extern __shared__ float s[];
    int *si = (int *)s;
    for (int i = threadIdx.x; i < TILE; i += blockDim.x) si[i] = 0;
    __syncthreads();
    for (int it = 0; it < ITERS; ++it) {
        int b = lane_base(it, threadIdx.x);
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int k = 0; k < TAPS; ++k) atomicAdd(&si[b + r * 45 + k], __float2int_rn(v * (r + k) * 1048576.f));
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = si[17];
}

__global__ void k_smem_i64(float *out, float v) {
    extern __shared__ float s[];
    unsigned long long *si = (unsigned long long *)s;
    for (int i = threadIdx.x; i < TILE / 2; i += blockDim.x) si[i] = 0;
    __syncthreads();
    for (int it = 0; it < ITERS; ++it) {
        int b = lane_base(it, threadIdx.x) % 7000;
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int k = 0; k < TAPS; ++k) atomicAdd(&si[b + r * 45 + k], (unsigned long long)__float2ll_rn(v * (r + k) * 1048576.f));
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = (float)si[17];
}

// non-atomic read-modify-write (upper bound for an ownership-based scheme)
__global__ void k_smem_rmw(float *out, float v) {
    extern __shared__ float s[];
    for (int i = threadIdx.x; i < TILE; i += blockDim.x) s[i] = 0.f;
    __syncthreads();
    for (int it = 0; it < ITERS; ++it) {
        int b = lane_base(it, threadIdx.x);
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int k = 0; k < TAPS; ++k) s[b + r * 45 + k] += v * (r + k);
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = s[17];
}

// LDS + FMA (pull-like)
__global__ void k_smem_lds(float *out, float v) {
    extern __shared__ float s[];
    for (int i = threadIdx.x; i < TILE; i += blockDim.x) s[i] = i * v;
    __syncthreads();
    float acc = 0.f;
    for (int it = 0; it < ITERS; ++it) {
        int b = lane_base(it, threadIdx.x);
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int k = 0; k < TAPS; ++k) acc = fmaf(s[b + r * 45 + k], v + k, acc);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// global RED, volume-like addressing: row stride 256, plane stride 65536
__global__ void k_glob_red(float *vol, float v, int nvox) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
        long long p = ((long long)t + (long long)it * gridDim.x * blockDim.x) % (nvox - 4 * 65536);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int k = 0; k < TAPS; ++k) atomicAdd(vol + p + i * 65536 + j * 256 + k, v * (i + j + k));
    }
}

// same with one 16-byte vector RED per row (requires 16 B alignment: p multiple of 4)
__global__ void k_glob_red_v4(float *vol, float v, int nvox) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
        long long p = (((long long)t + (long long)it * gridDim.x * blockDim.x) * 4) % (nvox - 4 * 65536);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 c = make_float4(v * (i + j), v * (i + j + 1), v * (i + j + 2), v * (i + j + 3));
                atomicAdd((float4 *)(vol + p + i * 65536 + j * 256), c);
            }
    }
}

template <typename F>
float time_it(F launch, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a); launch(); cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const int blocks = 148 * 4, threads = 256;
    const size_t smem = TILE * sizeof(float);
    float *out; CK(cudaMalloc(&out, blocks * threads * sizeof(float)));
    const int nvox = 256 * 256 * 256;
    float *vol; CK(cudaMalloc(&vol, (size_t)nvox * sizeof(float))); CK(cudaMemset(vol, 0, (size_t)nvox * sizeof(float)));
    CK(cudaFuncSetAttribute(k_smem_float, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_smem_int, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_smem_i64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_smem_rmw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_smem_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const double ops = (double)blocks * threads * ITERS * ROWS * TAPS;
    struct { const char *name; float ms; } res[8];
    int n = 0;
    res[n++] = {"smem float atomicAdd (CAS loop)", time_it([&] { k_smem_float<<<blocks, threads, smem>>>(out, 1e-3f); })};
    res[n++] = {"smem int32 atomicAdd (native)", time_it([&] { k_smem_int<<<blocks, threads, smem>>>(out, 1e-3f); })};
    res[n++] = {"smem int64 atomicAdd", time_it([&] { k_smem_i64<<<blocks, threads, smem>>>(out, 1e-3f); })};
    res[n++] = {"smem non-atomic RMW (LDS+FADD+STS)", time_it([&] { k_smem_rmw<<<blocks, threads, smem>>>(out, 1e-3f); })};
    res[n++] = {"smem LDS+FFMA (pull-like)", time_it([&] { k_smem_lds<<<blocks, threads, smem>>>(out, 1e-3f); })};
    res[n++] = {"global RED.ADD.F32 (L2)", time_it([&] { k_glob_red<<<blocks, threads>>>(vol, 1e-3f, nvox); })};
    res[n++] = {"global RED.ADD.F32x4 (L2)", time_it([&] { k_glob_red_v4<<<blocks, threads>>>(vol, 1e-3f, nvox); })};
    printf("%-40s %10s %14s %16s\n", "pattern", "ms", "Gtap/s", "clk/voxel/SM@1.9");
    for (int i = 0; i < n; ++i) {
        double gt = ops / (res[i].ms * 1e-3) / 1e9;
        // a "voxel" = 64 taps; cycles per voxel per SM at 1.9 GHz
        double clk = 1.9e9 * 148 / (gt * 1e9 / 64);
        printf("%-40s %10.3f %14.1f %16.2f\n", res[i].name, res[i].ms, gt, clk);
    }
    return 0;
}
