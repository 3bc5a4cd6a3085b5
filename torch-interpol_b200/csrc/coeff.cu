// Spline prefilter: separable causal / anti-causal recursive (IIR) filter, one
// axis per launch, lines batched over threads.
//
// The tensor is viewed as (outer, n, inner) and filtered along n, in place.
//  * inner > 1 : a warp owns 32 adjacent lines (consecutive `inner` index); the
//    [n x 32] tile is staged in shared memory with cp.async (16 B per lane when
//    alignment allows), each lane then runs all poles of its line out of its
//    own bank-conflict-free column, and the tile is written back coalesced:
//    one HBM read + one HBM write per axis whatever the number of poles.
//  * inner == 1: the lines themselves are contiguous; a CTA stages L whole
//    lines (a contiguous block of L*n elements) with a padded, odd row stride.
//  * lines too long for shared memory fall back to in-place global recursion.
//
// Replaces interpol/coeff.py:258-284 (filter) with the boundary conditions of
// coeff.py:82-227 (dct1/dct2/dft initial+final; zero->dct1, replicate->dct2).
#include <cstdio>
#include <cmath>
#include "common.cuh"

namespace ib200 {

struct CoeffParams {
    i64 outer, n, inner;
    int kind;            // 1: dct1, 2: dct2, 6: dft
    int npoles;
    double gain;
    double pole[3];      // exact poles (recursions, finals)                coeff.py:276,281
    double pp[3];        // pole rounded through float32 (power vectors)    SURVEY Q11
    double c1[3];        // dct1: pole^(n-1); dct2: pole^n; dft: pole^K
    int K[3];            // truncation length ceil(-30/log|pole|) (dft: min(K, n))
    int K2[3];           // length after which pp^k is < 1e-40 (dct2 exact sum cut-off)
};

// ------------------------------------------------------------------------
// All poles of one line held in shared (or global) memory with element stride st.
template <typename R, typename P>
__device__ __forceinline__ void filter_line(P s, const int st, const int n, const CoeffParams &cp, const R gain0 = R(1)) {
// the line holds raw samples; `gain0` (coeff.py:271 `inp *= gain`) is folded into the first pole's pass
#define S(i) (gn * (R)s[(i64)(i) * st])
#define SW(i) s[(i64)(i) * st]
    for (int p = 0; p < cp.npoles; ++p) {
        const R gn = p == 0 ? gain0 : R(1);
        const R pole = (R)cp.pole[p];
        const R pp = (R)cp.pp[p];
        R init;
        if (cp.kind == 1) {                      // dct1_initial, coeff.py:109-149
            if (cp.K[p] < n) {
                R acc = R(0), a = R(1);
                for (int k = 0; k < cp.K[p]; ++k) { acc = fma((R)S(k), a, acc); a *= pp; }
                init = acc;
            } else {
                const R pn = (R)cp.c1[p];
                const R pn2 = (R)(cp.c1[p] * cp.c1[p]);
                R out = (R)S(0) + pn * (R)S(n - 1);
                if (n > 2) {
                    R acc = R(0), a = pp;
                    for (int k = 1; k < n - 1; ++k) { acc = fma((R)S(k), a + pn2 / a, acc); a *= pp; }
                    out += acc;
                }
                init = out / (R)(1. - cp.c1[p] * cp.c1[p]);
            }
        } else if (cp.kind == 2) {               // dct2_initial, coeff.py:153-179
            const R pn = (R)cp.c1[p];
            R acc = R(0), a = R(1);
            const int kmax = n < cp.K2[p] ? n : cp.K2[p];
            for (int k = 0; k < kmax; ++k) { acc = fma((R)S(k), a, acc); a *= pp; }
            if (pn != R(0)) {
                R acc2 = R(0), b = R(1);
                for (int k = n - 1; k >= n - kmax; --k) { acc2 = fma((R)S(k), b, acc2); b *= pp; }
                acc = fma(pn, acc2, acc);
            }
            init = fma(acc, (R)(cp.pole[p] / (1. - cp.c1[p] * cp.c1[p])), (R)S(0));
        } else {                                 // dft_initial, coeff.py:82-105
            R acc = R(0), a = pp;
            for (int j = 1; j < cp.K[p]; ++j) { acc = fma((R)S(n - j), a, acc); a *= pp; }
            init = (acc + (R)S(0)) / (R)(1. - cp.c1[p]);
        }
        // causal recursion, coeff.py:275-276
        R prev = init;
        SW(0) = prev;
#pragma unroll 8
        for (int i = 1; i < n; ++i) { prev = fma(pole, prev, (R)S(i)); SW(i) = prev; }
        // final condition (reads the causally filtered line)
        R fin;
        if (cp.kind == 1) {                      // dct1_final, coeff.py:210-216
            fin = (pole * (R)SW(n - 2) + (R)SW(n - 1)) * (R)(cp.pole[p] / (cp.pole[p] * cp.pole[p] - 1.));
        } else if (cp.kind == 2) {               // dct2_final, coeff.py:220-227
            fin = (R)SW(n - 1) * (R)(cp.pole[p] / (cp.pole[p] - 1.));
        } else {                                 // dft_final, coeff.py:183-206
            R acc = R(0), a = pp * pp;
            for (int k = 0; k < cp.K[p] - 1; ++k) { acc = fma((R)SW(k), a, acc); a *= pp; }
            fin = fma(pole, (R)SW(n - 1), acc) / (R)(cp.c1[p] - 1.);
        }
        // anti-causal recursion, coeff.py:280-281
        prev = fin;
        SW(n - 1) = prev;
#pragma unroll 8
        for (int i = n - 2; i >= 0; --i) { prev = pole * (prev - (R)SW(i)); SW(i) = prev; }
    }
#undef S
#undef SW
}

// Same filter for a float32 line that is contiguous in shared memory, 16-byte aligned, n % 4 == 0: the
// recursions (and the dct2 boundary sums) move four samples per LDS.128 / STS.128.  With rows padded to n + 4
// words the 8 lanes of a quarter warp hit 8 distinct 16-byte bank groups, so the vector accesses are conflict
// free where scalar accesses at that stride would serialise four ways.  Arithmetic (order of every fma) is that
// of filter_line.
__device__ __forceinline__ void filter_line_vec4(float *s, const int n, const CoeffParams &cp, const float gain0) {
    typedef float R;
    float4 *s4 = reinterpret_cast<float4 *>(s);
    const int nv = n >> 2;
    for (int p = 0; p < cp.npoles; ++p) {
        const R gn = p == 0 ? gain0 : R(1);
        const R pole = (R)cp.pole[p];
        const R pp = (R)cp.pp[p];
        R init;
        if (cp.kind == 2) {                      // dct2_initial, coeff.py:153-179
            const R pn = (R)cp.c1[p];
            R acc = R(0), a = R(1);
            const int kmax = n < cp.K2[p] ? n : cp.K2[p];
            int k = 0;
            for (; k + 4 <= kmax; k += 4) {
                const float4 v = s4[k >> 2];
                acc = fmaf(gn * v.x, a, acc); a *= pp;
                acc = fmaf(gn * v.y, a, acc); a *= pp;
                acc = fmaf(gn * v.z, a, acc); a *= pp;
                acc = fmaf(gn * v.w, a, acc); a *= pp;
            }
            for (; k < kmax; ++k) { acc = fmaf(gn * s[k], a, acc); a *= pp; }
            if (pn != R(0)) {
                R acc2 = R(0), b = R(1);
                int q = n - 1;
                for (; q - 3 >= n - kmax; q -= 4) {
                    const float4 v = s4[q >> 2];
                    acc2 = fmaf(gn * v.w, b, acc2); b *= pp;
                    acc2 = fmaf(gn * v.z, b, acc2); b *= pp;
                    acc2 = fmaf(gn * v.y, b, acc2); b *= pp;
                    acc2 = fmaf(gn * v.x, b, acc2); b *= pp;
                }
                for (; q >= n - kmax; --q) { acc2 = fmaf(gn * s[q], b, acc2); b *= pp; }
                acc = fmaf(pn, acc2, acc);
            }
            init = fmaf(acc, (R)(cp.pole[p] / (1. - cp.c1[p] * cp.c1[p])), gn * s[0]);
        } else if (cp.kind == 1) {               // dct1_initial, coeff.py:109-149
            if (cp.K[p] < n) {
                R acc = R(0), a = R(1);
                for (int k = 0; k < cp.K[p]; ++k) { acc = fmaf(gn * s[k], a, acc); a *= pp; }
                init = acc;
            } else {
                const R pn = (R)cp.c1[p];
                const R pn2 = (R)(cp.c1[p] * cp.c1[p]);
                R out = gn * s[0] + pn * (gn * s[n - 1]);
                if (n > 2) {
                    R acc = R(0), a = pp;
                    for (int k = 1; k < n - 1; ++k) { acc = fmaf(gn * s[k], a + pn2 / a, acc); a *= pp; }
                    out += acc;
                }
                init = out / (R)(1. - cp.c1[p] * cp.c1[p]);
            }
        } else {                                 // dft_initial, coeff.py:82-105
            R acc = R(0), a = pp;
            for (int j = 1; j < cp.K[p]; ++j) { acc = fmaf(gn * s[n - j], a, acc); a *= pp; }
            init = (acc + gn * s[0]) / (R)(1. - cp.c1[p]);
        }
        // causal recursion, coeff.py:275-276
        R prev = init;
#pragma unroll 2
        for (int c = 0; c < nv; ++c) {
            float4 v = s4[c];
            v.x = c == 0 ? prev : fmaf(pole, prev, gn * v.x);
            v.y = fmaf(pole, v.x, gn * v.y);
            v.z = fmaf(pole, v.y, gn * v.z);
            v.w = fmaf(pole, v.z, gn * v.w);
            prev = v.w;
            s4[c] = v;
        }
        // final condition (reads the causally filtered line)
        R fin;
        if (cp.kind == 1) {                      // dct1_final, coeff.py:210-216
            fin = (pole * s[n - 2] + s[n - 1]) * (R)(cp.pole[p] / (cp.pole[p] * cp.pole[p] - 1.));
        } else if (cp.kind == 2) {               // dct2_final, coeff.py:220-227
            fin = s[n - 1] * (R)(cp.pole[p] / (cp.pole[p] - 1.));
        } else {                                 // dft_final, coeff.py:183-206
            R acc = R(0), a = pp * pp;
            for (int k = 0; k < cp.K[p] - 1; ++k) { acc = fmaf(s[k], a, acc); a *= pp; }
            fin = fmaf(pole, s[n - 1], acc) / (R)(cp.c1[p] - 1.);
        }
        // anti-causal recursion, coeff.py:280-281
        prev = fin;
#pragma unroll 2
        for (int c = nv - 1; c >= 0; --c) {
            float4 v = s4[c];
            v.w = c == nv - 1 ? prev : pole * (prev - v.w);
            v.z = pole * (v.w - v.z);
            v.y = pole * (v.z - v.y);
            v.x = pole * (v.y - v.x);
            prev = v.x;
            s4[c] = v;
        }
    }
}

// proxy so that filter_line can run directly on 16-bit global storage
template <typename T>
struct GlobalRef {
    T *p;
    __device__ __forceinline__ operator float() const { return Traits<T>::load_rw(p); }
    __device__ __forceinline__ void operator=(float v) { Traits<T>::store(p, v); }
};
template <typename T>
struct GlobalLine {
    T *base;
    __device__ __forceinline__ GlobalRef<T> operator[](i64 i) const { return GlobalRef<T>{base + i}; }
};

// ---------------------------------------------------------------- kernels --

__device__ __forceinline__ void cpa16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cpa_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// inner > 1.  One warp per CTA owns a tile of 32 adjacent lines ([n][32] in shared
// memory, lane l <-> column l: conflict free).  float32 tiles are staged with
// 16-byte cp.async (a quarter-warp moves one 128-byte row; all n rows are in
// flight at once, so the read runs at HBM rate instead of one row per latency)
// and written back with 16-byte stores.
template <typename T>
__global__ void __launch_bounds__(32)
coeff_strided_kernel(const __grid_constant__ CoeffParams cp, T *__restrict__ data, const int vec_ok) {
    typedef typename Traits<T>::Real R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const int n = (int)cp.n;
    R *tile = reinterpret_cast<R *>(smem_raw);
    const i64 tiles_per_outer = (cp.inner + 31) / 32;
    const i64 ntiles = cp.outer * tiles_per_outer;
    const R gain = (R)cp.gain;
    for (i64 tidx = blockIdx.x; tidx < ntiles; tidx += gridDim.x) {
        const i64 o = tidx / tiles_per_outer;
        const i64 k0 = (tidx - o * tiles_per_outer) * 32;
        const bool full = k0 + 32 <= cp.inner;
        T *g0 = data + o * cp.n * cp.inner + k0;
        if (sizeof(T) == 4 && sizeof(R) == 4 && vec_ok && full) {
            const int rsub = lane >> 3, ch = (lane & 7) * 4;
            for (int i = rsub; i < n; i += 4) cpa16(tile + i * 32 + ch, g0 + (i64)i * cp.inner + ch);
            cpa_wait_all();
            __syncwarp();
            filter_line<R>(tile + lane, 32, n, cp, gain);
            __syncwarp();
            for (int i = rsub; i < n; i += 4)
                *reinterpret_cast<float4 *>(g0 + (i64)i * cp.inner + ch) = *reinterpret_cast<const float4 *>(tile + i * 32 + ch);
        } else if (k0 + lane < cp.inner) {
            T *g = g0 + lane;
#pragma unroll 16
            for (int i = 0; i < n; ++i) tile[i * 32 + lane] = Traits<T>::load_rw(g + (i64)i * cp.inner);
            filter_line<R>(tile + lane, 32, n, cp, gain);
#pragma unroll 16
            for (int i = 0; i < n; ++i) Traits<T>::store(g + (i64)i * cp.inner, tile[i * 32 + lane]);
        }
        __syncwarp();
    }
}

// inner == 1.  The lines themselves are contiguous.  One warp per CTA owns 32 consecutive lines (a contiguous block
// of 32 * n elements): float32 blocks are staged with 16-byte cp.async into rows padded to n + 4 words (16-byte
// aligned rows), every lane filters its own line four samples per LDS.128 / STS.128 (filter_line_vec4: conflict
// free at that stride), and the block goes back with 16-byte stores.  Several CTAs per SM overlap each other's
// load / filter / store phases.  Other types (or unaligned blocks) stage element by element with an odd row stride.
template <typename T>
__global__ void __launch_bounds__(32)
coeff_contig_kernel(const __grid_constant__ CoeffParams cp, T *__restrict__ data, const int row_stride, const int vec_ok) {
    typedef typename Traits<T>::Real R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    R *tile = reinterpret_cast<R *>(smem_raw);
    const int n = (int)cp.n, lane = threadIdx.x;
    const i64 nblk = (cp.outer + 31) / 32;
    const R gain = (R)cp.gain;
    for (i64 blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const i64 l0 = blk * 32;
        const int nl = (int)((cp.outer - l0) < 32 ? (cp.outer - l0) : 32);
        T *g = data + l0 * n;
        if (sizeof(T) == 4 && sizeof(R) == 4 && vec_ok) {
            const int vpl = n >> 2;                       // 16-byte vectors per line
            const int total = nl * vpl;
            int l = lane / vpl, v = lane - l * vpl;       // vector e = lane + 32 k  <->  (line l, vector v)
            const int dl = 32 / vpl, dv = 32 - dl * vpl;
            for (int e = lane; e < total; e += 32) {
                cpa16(tile + l * row_stride + 4 * v, g + 4 * (i64)e);
                l += dl; v += dv;
                if (v >= vpl) { v -= vpl; ++l; }
            }
            cpa_wait_all();
            __syncwarp();
            if (lane < nl) filter_line_vec4(reinterpret_cast<float *>(tile) + lane * row_stride, n, cp, (float)gain);
            __syncwarp();
            l = lane / vpl; v = lane - l * vpl;
            for (int e = lane; e < total; e += 32) {
                *reinterpret_cast<float4 *>(g + 4 * (i64)e) = *reinterpret_cast<const float4 *>(tile + l * row_stride + 4 * v);
                l += dl; v += dv;
                if (v >= vpl) { v -= vpl; ++l; }
            }
        } else {
            const int count = nl * n;
            int l = lane / n, i = lane - l * n;
            const int dl = 32 / n, di = 32 - dl * n;
            for (int e = lane; e < count; e += 32) {
                tile[l * row_stride + i] = (R)Traits<T>::load_rw(g + e);
                l += dl; i += di;
                if (i >= n) { i -= n; ++l; }
            }
            __syncwarp();
            if (lane < nl) filter_line<R>(tile + lane * row_stride, 1, n, cp, gain);
            __syncwarp();
            l = lane / n; i = lane - l * n;
            for (int e = lane; e < count; e += 32) {
                Traits<T>::store(g + e, tile[l * row_stride + i]);
                l += dl; i += di;
                if (i >= n) { i -= n; ++l; }
            }
        }
        __syncwarp();
    }
}

// fallback for lines that do not fit in shared memory: in-place on global.
template <typename T>
__global__ void __launch_bounds__(128)
coeff_global_kernel(const __grid_constant__ CoeffParams cp, T *__restrict__ data) {
    typedef typename Traits<T>::Real R;
    const i64 nlines = cp.outer * cp.inner;
    const R gain = (R)cp.gain;
    for (i64 l = (i64)blockIdx.x * blockDim.x + threadIdx.x; l < nlines; l += (i64)gridDim.x * blockDim.x) {
        const i64 o = l / cp.inner, k = l - o * cp.inner;
        T *g = data + o * cp.n * cp.inner + k;
        if (sizeof(T) == sizeof(R)) filter_line<R>(reinterpret_cast<R *>(g), (int)cp.inner, (int)cp.n, cp, gain);
        else filter_line<R>(GlobalLine<T>{g}, (int)cp.inner, (int)cp.n, cp, gain);
    }
}
// ---------------------------------------------------------------- launch --

static int get_poles(int order, double *poles) {   // coeff.py:35-65
    switch (order) {
    case 2: poles[0] = sqrt(8.) - 3.; return 1;
    case 3: poles[0] = sqrt(3.) - 2.; return 1;
    case 4:
        poles[0] = sqrt(664. - sqrt(438976.)) + sqrt(304.) - 19.;
        poles[1] = sqrt(664. + sqrt(438976.)) - sqrt(304.) - 19.;
        return 2;
    case 5:
        poles[0] = sqrt(67.5 - sqrt(4436.25)) + sqrt(26.25) - 6.5;
        poles[1] = sqrt(67.5 + sqrt(4436.25)) - sqrt(26.25) - 6.5;
        return 2;
    case 6:
        poles[0] = -0.488294589303044755130118038883789062112279161239377608394;
        poles[1] = -0.081679271076237512597937765737059080653379610398148178525368;
        poles[2] = -0.00141415180832581775108724397655859252786416905534669851652709;
        return 3;
    case 7:
        poles[0] = -0.5352804307964381655424037816816460718339231523426924148812;
        poles[1] = -0.122554615192326690515272264359357343605486549427295558490763;
        poles[2] = -0.0091486948096082769285930216516478534156925639545994482648003;
        return 3;
    }
    return 0;
}

template <typename T>
static int launch_typed(const CoeffParams &cp, void *data, cudaStream_t stream) {
    typedef typename Traits<T>::Real R;
    const size_t kSmemCap = 200 * 1024;
    T *d = (T *)data;
    if (cp.inner > 1) {
        const size_t smem = (size_t)cp.n * 32 * sizeof(R);
        if (smem <= kSmemCap) {
            const i64 ntiles = cp.outer * ((cp.inner + 31) / 32);
            i64 blocks = ntiles;
            const i64 cap = (i64)kNumSMs * 32;
            if (blocks > cap) blocks = cap;
            const int vec_ok = ((uintptr_t)data % 16 == 0) && (cp.inner % 4 == 0);
            IB200_CUDA_CHECK(cudaFuncSetAttribute(coeff_strided_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            coeff_strided_kernel<T><<<(unsigned)blocks, 32, smem, stream>>>(cp, d, vec_ok);
            note_launch("coeff_strided");
            IB200_CUDA_CHECK(cudaGetLastError());
            return IB200_OK;
        }
    } else {
        const int vec_ok = sizeof(T) == 4 && sizeof(R) == 4 && ((uintptr_t)data % 16 == 0) && (cp.n % 4 == 0);
        const int row_stride = vec_ok ? (int)cp.n + 4 : ((int)cp.n | 1);   // 16-byte aligned rows / odd stride
        const size_t smem = (size_t)32 * row_stride * sizeof(R);
        if (smem <= kSmemCap) {
            i64 blocks = (cp.outer + 31) / 32;
            const i64 cap = (i64)kNumSMs * 32;
            if (blocks > cap) blocks = cap;
            IB200_CUDA_CHECK(cudaFuncSetAttribute(coeff_contig_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            coeff_contig_kernel<T><<<(unsigned)blocks, 32, smem, stream>>>(cp, d, row_stride, vec_ok);
            note_launch("coeff_contig");
            IB200_CUDA_CHECK(cudaGetLastError());
            return IB200_OK;
        }
    }
    const i64 nlines = cp.outer * cp.inner;
    i64 blocks = (nlines + 127) / 128;
    if (blocks > (i64)kNumSMs * 16) blocks = (i64)kNumSMs * 16;
    coeff_global_kernel<T><<<(unsigned)blocks, 128, 0, stream>>>(cp, d);
    note_launch("coeff_global");
    IB200_CUDA_CHECK(cudaGetLastError());
    return IB200_OK;
}

int launch_coeff(void *data, int dtype, i64 outer, i64 n, i64 inner, int bound, int order,
                 cudaStream_t stream) {
    if (order <= 1 || n == 1 || outer * inner == 0) return IB200_OK;   // coeff.py:263-264, 306-307
    CoeffParams cp;
    cp.outer = outer; cp.n = n; cp.inner = inner;
    if (bound == IB200_BOUND_ZERO || bound == IB200_BOUND_DCT1) cp.kind = 1;          // coeff.py:237-252
    else if (bound == IB200_BOUND_REPLICATE || bound == IB200_BOUND_DCT2) cp.kind = 2;
    else if (bound == IB200_BOUND_DFT) cp.kind = 6;
    else return IB200_ERR_BOUND_UNSUPPORTED;
    cp.npoles = get_poles(order, cp.pole);
    cp.gain = 1.;
    for (int p = 0; p < 3; ++p) { cp.pp[p] = 0; cp.c1[p] = 0; cp.K[p] = 0; cp.K2[p] = 0; if (p >= cp.npoles) cp.pole[p] = 0; }
    for (int p = 0; p < cp.npoles; ++p) {
        const double pole = cp.pole[p];
        cp.gain *= (1. - pole) * (1. - 1. / pole);                                    // coeff.py:69-73
        cp.pp[p] = (double)(float)pole;
        i64 K = (i64)ceil(-30. / log(fabs(pole)));
        cp.K2[p] = (int)ceil(log(1e-40) / log(fabs(pole)));
        if (cp.kind == 1) { cp.c1[p] = pow(pole, (double)(n - 1)); }
        else if (cp.kind == 2) { cp.c1[p] = pow(pole, (double)n); }
        else { if (K > n) K = n; cp.c1[p] = pow(pole, (double)K); }
        cp.K[p] = (int)(K > 0x7fffffff ? 0x7fffffff : K);
    }
    switch (dtype) {
    case IB200_F32: return launch_typed<float>(cp, data, stream);
    case IB200_F64: return launch_typed<double>(cp, data, stream);
    case IB200_F16: return launch_typed<__half>(cp, data, stream);
    case IB200_BF16: return launch_typed<__nv_bfloat16>(cp, data, stream);
    }
    return IB200_ERR_DTYPE;
}

}  // namespace ib200
