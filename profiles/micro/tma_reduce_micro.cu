// Which TMA store / reduce forms does this part accept on float32 tensors?
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -I torch-interpol_b200/csrc profiles/micro/tma_reduce_micro.cu -o profiles/micro/tma_reduce_micro.bin
//   for m in 0 1 2 3 4 5 6; do ./tma_reduce_micro.bin $m; done     (one process per variant: a fault kills the context)
#include <cstdio>
#include <cstdlib>
#include "pipe_common.cuh"
using namespace ib200;

__global__ void k(const __grid_constant__ CUtensorMap tm, int mode, int c0, int c1, float *gdst) {
    extern __shared__ __align__(1024) float sm[];
    for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) sm[i] = 1.f;
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (mode == 0)        // plain tensor store
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tm), "r"(smem_u32(sm)), "r"(c0), "r"(c1) : "memory");
        else if (mode == 1 || mode == 2 || mode == 5 || mode == 6)   // tensor reduce add
            asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tm), "r"(smem_u32(sm)), "r"(c0), "r"(c1) : "memory");
        else if (mode == 3)   // 1-D bulk reduce add f32
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(sm)), "r"(4096) : "memory");
        else if (mode == 4)   // 1-D bulk store
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(sm)), "r"(4096) : "memory");
        bulk_commit();
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main(int argc, char **argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    float *d;
    cudaMalloc(&d, 64 * 64 * 4 * sizeof(float));
    cudaMemset(d, 0, 64 * 64 * 4 * sizeof(float));
    CUtensorMap tm;
    const long long dim[5] = {64, 64, 4, 1, 1}, str[5] = {1, 64, 4096, 16384, 16384};
    const int box[5] = {64, 16, 1, 1, 1};
    if (!make_tensor_map(&tm, d, 2, dim, str, box)) { printf("encode failed\n"); return 2; }
    k<<<1, 32, 4096>>>(tm, mode, mode == 6 ? 32 : 0, mode == 2 ? -4 : (mode == 5 ? 56 : 8), d);
    cudaError_t e = cudaDeviceSynchronize();
    static float h[64 * 64];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0; for (float v : h) s += v;
    const char *names[] = {"tensor store", "tensor reduce.add (inside)", "tensor reduce.add (clipped)", "bulk reduce.add.f32", "bulk store", "tensor reduce.add (clipped at the top, expect 512)", "tensor reduce.add (clipped inner, expect 512)"};
    printf("mode %d %-28s: %s, sum = %.0f\n", mode, names[mode], cudaGetErrorString(e), s);
    return e != cudaSuccess;
}
