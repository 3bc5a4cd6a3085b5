// Shared-memory pipe micro-benchmarks behind the round-2 kernel designs (B200, via gpurun):
//   * LDS.32 / LDS.64 / LDS.128 throughput per warp instruction for lane -> address patterns with
//     overlap between lanes (does the crossbar merge identical 8 / 16-byte chunks across the whole warp,
//     or only inside half / quarter warps?);
//   * native ATOMS.ADD (32-bit) throughput: conflict-free, partially active warps, bank conflicts,
//     same-address collisions; with and without a returned value.
// One 512-thread CTA per SM; every warp runs ITERS x UNROLL independent instructions; cycles are taken with
// clock64() around the loop (block-wide min start / max end), reported as cycles per warp instruction per SM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int NT = 512, ITERS = 512, UNROLL = 8;
constexpr int SWORDS = 8192;   // 32 KB window the patterns live in

__device__ __forceinline__ unsigned s_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int WIDTH>   // 1, 2, 4 words per lane
__global__ void __launch_bounds__(NT, 1) k_lds(const int *__restrict__ lane_off, float *out, long long *cyc) {
    extern __shared__ __align__(16) float s[];
    __shared__ long long t0s, t1s;
    for (int i = threadIdx.x; i < SWORDS + 4096; i += NT) s[i] = (float)(i & 255);
    if (threadIdx.x == 0) { t0s = 0x7fffffffffffffffLL; t1s = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per-warp base so that warps do not all hammer the same words; keeps 16-byte alignment
    const unsigned base = s_u32(s) + (unsigned)(lane_off[lane] * 4) + (unsigned)(warp * 64 * 4);
    float acc = 0.f;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        const unsigned a = base + (unsigned)((it & 15) * 128 * 4);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (WIDTH == 1) {
                float x;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a + u * 512));
                acc += x;
            } else if (WIDTH == 2) {
                float x, y;
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(a + u * 512));
                acc += x + y;
            } else {
                float x, y, z, w;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a + u * 512));
                acc += x + y + z + w;
            }
        }
    }
    const long long t1 = clock64();
    atomicMin(&t0s, t0); atomicMax(&t1s, t1);
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1s - t0s;
    out[blockIdx.x * NT + threadIdx.x] = acc;
}

template <int RET>   // 0: red (no return), 1: atom (returned value consumed)
__global__ void __launch_bounds__(NT, 1) k_atoms(const int *__restrict__ lane_off, const int *__restrict__ lane_on, int *out, long long *cyc) {
    extern __shared__ __align__(16) int si[];
    __shared__ long long t0s, t1s;
    for (int i = threadIdx.x; i < SWORDS + 4096; i += NT) si[i] = 0;
    if (threadIdx.x == 0) { t0s = 0x7fffffffffffffffLL; t1s = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned base = s_u32(si) + (unsigned)(lane_off[lane] * 4) + (unsigned)(warp * 64 * 4);
    const bool on = lane_on[lane] != 0;
    int acc = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        const unsigned a = base + (unsigned)((it & 15) * 128 * 4);
        if (on) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (RET) {
                    int r;
                    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(a + u * 512), "r"(it + u) : "memory");
                    acc += r;
                } else {
                    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a + u * 512), "r"(it + u) : "memory");
                }
            }
        }
    }
    const long long t1 = clock64();
    atomicMin(&t0s, t0); atomicMax(&t1s, t1);
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1s - t0s;
    out[blockIdx.x * NT + threadIdx.x] = acc + si[threadIdx.x];
}

// plain LDS + STS read-modify-write (what an ownership scheme would issue), and SHFL for comparison
__global__ void __launch_bounds__(NT, 1) k_shfl(float *out, long long *cyc) {
    __shared__ long long t0s, t1s;
    if (threadIdx.x == 0) { t0s = 0x7fffffffffffffffLL; t1s = 0; }
    __syncthreads();
    float v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = (float)(threadIdx.x + u);
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] += __shfl_down_sync(0xffffffffu, v[u], 1 + (it & 3));
    }
    const long long t1 = clock64();
    atomicMin(&t0s, t0); atomicMax(&t1s, t1);
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1s - t0s;
    float a = 0.f;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) a += v[u];
    out[blockIdx.x * NT + threadIdx.x] = a;
}

struct Pat { const char *name; int off[32]; int on[32]; };

static void fill(Pat &p, const char *name, int (*f)(int), int (*g)(int) = nullptr) {
    p.name = name;
    for (int l = 0; l < 32; ++l) { p.off[l] = f(l); p.on[l] = g ? g(l) : 1; }
}

int main() {
    const int blocks = 148;
    const size_t smem = (SWORDS + 4096) * 4;
    int *d_off, *d_on; float *d_out; long long *d_cyc;
    CK(cudaMalloc(&d_off, 128)); CK(cudaMalloc(&d_on, 128));
    CK(cudaMalloc(&d_out, blocks * NT * 4)); CK(cudaMalloc(&d_cyc, blocks * 8));
    CK(cudaFuncSetAttribute(k_lds<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_lds<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_lds<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_atoms<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_atoms<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto report = [&](const char *kind, const char *name) {
        CK(cudaDeviceSynchronize());
        long long h[blocks]; CK(cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost));
        double s = 0; for (int i = 0; i < blocks; ++i) s += (double)h[i];
        s /= blocks;
        printf("%-10s %-58s %8.3f clk / warp instr / SM\n", kind, name, s / ((double)ITERS * UNROLL * (NT / 32)));
    };
    auto run_lds = [&](int width, Pat &p) {
        CK(cudaMemcpy(d_off, p.off, 128, cudaMemcpyHostToDevice));
        for (int r = 0; r < 2; ++r) {
            if (width == 1) k_lds<1><<<blocks, NT, smem>>>(d_off, d_out, d_cyc);
            else if (width == 2) k_lds<2><<<blocks, NT, smem>>>(d_off, d_out, d_cyc);
            else k_lds<4><<<blocks, NT, smem>>>(d_off, d_out, d_cyc);
        }
        report(width == 1 ? "LDS.32" : width == 2 ? "LDS.64" : "LDS.128", p.name);
    };
    auto run_atoms = [&](int ret, Pat &p) {
        CK(cudaMemcpy(d_off, p.off, 128, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_on, p.on, 128, cudaMemcpyHostToDevice));
        for (int r = 0; r < 2; ++r) {
            if (ret) k_atoms<1><<<blocks, NT, smem>>>(d_off, d_on, (int *)d_out, d_cyc);
            else k_atoms<0><<<blocks, NT, smem>>>(d_off, d_on, (int *)d_out, d_cyc);
        }
        report(ret ? "ATOMS.ret" : "ATOMS", p.name);
    };
    Pat p;
    // ---- LDS.32 ----
    fill(p, "word l (conflict free)", [](int l) { return l; }); run_lds(1, p);
    fill(p, "word 2l (2-way bank conflict)", [](int l) { return 2 * l; }); run_lds(1, p);
    fill(p, "word l/2 (pairs broadcast)", [](int l) { return l / 2; }); run_lds(1, p);
    fill(p, "word floor(1.3 l) (span 41)", [](int l) { return (int)(1.3 * l); }); run_lds(1, p);
    // ---- LDS.64 ----
    fill(p, "pair l (256 B distinct)", [](int l) { return 2 * l; }); run_lds(2, p);
    fill(p, "pair l/2 (16 distinct pairs, 128 B)", [](int l) { return 2 * (l / 2); }); run_lds(2, p);
    fill(p, "pair l%16 (halves identical, 128 B)", [](int l) { return 2 * (l % 16); }); run_lds(2, p);
    fill(p, "pair floor(1.2 l / 2) (z-row, stretch 1.2)", [](int l) { return 2 * ((int)(1.2 * l) / 2); }); run_lds(2, p);
    fill(p, "all lanes one pair", [](int) { return 0; }); run_lds(2, p);
    // ---- LDS.128 ----
    fill(p, "chunk l (512 B distinct)", [](int l) { return 4 * l; }); run_lds(4, p);
    fill(p, "chunk l/2 (16 distinct, 256 B)", [](int l) { return 4 * (l / 2); }); run_lds(4, p);
    fill(p, "chunk l/4 (8 distinct, 128 B; 2 per quarter warp)", [](int l) { return 4 * (l / 4); }); run_lds(4, p);
    fill(p, "chunk l%8 (quarters identical, 128 B)", [](int l) { return 4 * (l % 8); }); run_lds(4, p);
    fill(p, "chunk l%16 (halves identical, 256 B)", [](int l) { return 4 * (l % 16); }); run_lds(4, p);
    fill(p, "all lanes one chunk", [](int) { return 0; }); run_lds(4, p);
    fill(p, "chunk floor(l/4) z-row stretch 1.0 (1 voxel per lane)", [](int l) { return 4 * (l / 4); }); run_lds(4, p);
    fill(p, "chunk floor(1.25 l / 4) z-row stretch 1.25 (10 chunks)", [](int l) { return 4 * ((int)(1.25 * l) / 4); }); run_lds(4, p);
    fill(p, "chunk floor(1.5 l / 4) z-row stretch 1.5 (12 chunks)", [](int l) { return 4 * ((int)(1.5 * l) / 4); }); run_lds(4, p);
    fill(p, "16 voxels x 2 halves: chunk floor((l%16)/4)+l/16 (5 chunks)", [](int l) { return 4 * ((l % 16) / 4 + l / 16); }); run_lds(4, p);
    fill(p, "16 voxels x 2 halves stretch 1.4: floor(1.4(l%16)/4)+l/16", [](int l) { return 4 * ((int)(1.4 * (l % 16)) / 4 + l / 16); }); run_lds(4, p);
    fill(p, "8 voxels x 4 rows(64w apart): chunk (l%8)/4 + 16*(l/8)", [](int l) { return 4 * ((l % 8) / 4 + 16 * (l / 8)); }); run_lds(4, p);
    fill(p, "8 voxels x 4 rows skewed 8w: chunk (l%8)/4 + 18*(l/8)", [](int l) { return 4 * ((l % 8) / 4 + 18 * (l / 8)); }); run_lds(4, p);
    // ---- ATOMS ----
    for (int ret = 0; ret < 2; ++ret) {
        fill(p, "word l (conflict free)", [](int l) { return l; }); run_atoms(ret, p);
        fill(p, "word l, 16 lanes active", [](int l) { return l; }, [](int l) { return l < 16 ? 1 : 0; }); run_atoms(ret, p);
        fill(p, "word l, 8 lanes active", [](int l) { return l; }, [](int l) { return l < 8 ? 1 : 0; }); run_atoms(ret, p);
        fill(p, "word l, every 4th lane active", [](int l) { return l; }, [](int l) { return l % 4 == 0 ? 1 : 0; }); run_atoms(ret, p);
        fill(p, "word l, 1 lane active", [](int l) { return l; }, [](int l) { return l == 0 ? 1 : 0; }); run_atoms(ret, p);
        fill(p, "word 2l (2-way bank conflict)", [](int l) { return 2 * l; }); run_atoms(ret, p);
        fill(p, "word 4l (4-way bank conflict)", [](int l) { return 4 * l; }); run_atoms(ret, p);
        fill(p, "word l/2 (pairs collide on one address)", [](int l) { return l / 2; }); run_atoms(ret, p);
        fill(p, "word l/4 (4 lanes per address)", [](int l) { return l / 4; }); run_atoms(ret, p);
        fill(p, "all lanes one address", [](int) { return 0; }); run_atoms(ret, p);
        fill(p, "word floor(0.8 l) (stretch 0.8: some collisions)", [](int l) { return (int)(0.8 * l); }); run_atoms(ret, p);
        fill(p, "word floor(1.2 l) (stretch 1.2: span 38)", [](int l) { return (int)(1.2 * l); }); run_atoms(ret, p);
    }
    // ---- SHFL ----
    for (int r = 0; r < 2; ++r) k_shfl<<<blocks, NT>>>(d_out, d_cyc);
    report("SHFL", "shfl_down + FADD");
    return 0;
}
