// Label-map pull in ONE pass (orders 0 / 1 per axis).
//
// The reference resamples an integer label map by looping over `input.unique()`: one full grid_pull of the
// soft mask (input == label) per label, keeping for every voxel the label whose interpolated mask is the
// largest (interpol/api.py:194-205: ascending labels, strict `>`, starting from out = 0 / pmax = 0).  A
// point only sees the labels of its own (order+1)^dim nodes, so the same arg-max is found per point among
// at most 8 candidates, whatever the number of labels in the volume: candidates are visited in ascending
// order, each one's mask value is accumulated with the node weights / signs / extrapolation mask of the
// gather kernels in the same nested order (z, then y, then x), and the first largest value > 0 wins.
#include <cstdio>
#include "support.cuh"

namespace ib200 {

// (L: storage type of the labels, read and written as they are; candidates are compared as 64-bit integers)
template <typename G, typename L, int DIM>
__global__ void __launch_bounds__(256)
pull_labels_kernel(const __grid_constant__ KParams kp, const L *__restrict__ vol, const G *__restrict__ grid,
                   L *__restrict__ out) {
    typedef typename Traits<G>::Real R;
    const i64 total = kp.batch * kp.pts_total;
    for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (i64)gridDim.x * blockDim.x) {
        i64 b; int xyz[3];
        i64 goff;
        const bool disp = (kp.flags & IB200_FLAG_DISPLACEMENT) != 0;
        if (kp.pts_dense) {
            b = p / kp.pts_total;
            goff = b * kp.grid_sb + (p - b * kp.pts_total) * DIM;
            if (disp) {
                i64 r = p;
#pragma unroll
                for (int d = DIM - 1; d >= 0; --d) { xyz[d] = (int)(r % kp.pts_n[d]); r /= kp.pts_n[d]; }
            }
        } else {
            i64 r = p;
            goff = 0;
#pragma unroll
            for (int d = DIM - 1; d >= 0; --d) { xyz[d] = (int)(r % kp.pts_n[d]); r /= kp.pts_n[d]; goff += xyz[d] * kp.grid_s[d]; }
            b = r;
            goff += b * kp.grid_sb;
        }
        const i64 r_dense = p - b * kp.pts_total;
        R coord[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) coord[d] = Traits<G>::load(grid + goff + d * kp.grid_sd) + (disp ? (R)xyz[d] : R(0));
        bool ok = inbounds<R, DIM>(kp, coord);
        Axis<R, 8> ax[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (d < DIM) ok = setup_axis<R, -1, 0, 8>(ax[d], coord[d], kp.order[d], kp.bound[d], kp.vol_n[d], (int)kp.vol_s[d], kp) && ok;
            else unit_axis(ax[d]);
        }
        const int n0 = ax[0].n, n1 = DIM >= 2 ? ax[1].n : 1, n2 = DIM >= 3 ? ax[2].n : 1;     // 1 or 2 nodes per axis
        for (i64 c = 0; c < kp.channels; ++c) {
            const L *src = vol + b * kp.vol_sb + c * kp.vol_sc;
            long long best = 0;
            if (ok) {
                long long lab[2][2][2];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int k = 0; k < 2; ++k)
                            lab[i][j][k] = (i < n0 && j < n1 && k < n2) ? (long long)__ldg(src + ax[0].off[i] + ax[1].off[j] + ax[2].off[k]) : 0;
                R pmax = R(0);
                constexpr long long kNone = 0x7fffffffffffffffLL;
                bool first = true;                       // (no label is "below every int64 label": a flag instead)
                long long prev = 0;
                for (int cand = 0; cand < 8; ++cand) {
                    // next candidate: the smallest node label above the previous one
                    long long cur = kNone;
                    bool found = false;
#pragma unroll
                    for (int i = 0; i < 2; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j)
#pragma unroll
                            for (int k = 0; k < 2; ++k)
                                if (i < n0 && j < n1 && k < n2 && (first || lab[i][j][k] > prev) && (!found || lab[i][j][k] < cur)) {
                                    cur = lab[i][j][k]; found = true;
                                }
                    if (!found) break;
                    prev = cur; first = false;
                    // interpolated value of the mask (label == cur), nested like gather_kernel
                    R acc = R(0);
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        if (i >= n0) break;
                        R s = R(0);
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            if (j >= n1) break;
                            R t = R(0);
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                if (k >= n2) break;
                                t = fma(ax[2].w[k], lab[i][j][k] == cur ? R(1) : R(0), t);
                            }
                            s = fma(ax[1].w[j], t, s);
                        }
                        acc = fma(ax[0].w[i], s, acc);
                    }
                    if (acc > pmax) { pmax = acc; best = cur; }
                }
            }
            out[(b * kp.channels + c) * kp.pts_total + r_dense] = (L)best;
        }
    }
}

template <typename G, typename L>
static int launch_labels_t(const KParams &kp, const void *vol_, const void *grid, void *out_, cudaStream_t stream) {
    const L *vol = (const L *)vol_;
    L *out = (L *)out_;
    const i64 total = kp.batch * kp.pts_total;
    if (total == 0 || kp.channels == 0) return IB200_OK;
    i64 blocks = (total + 255) / 256;
    if (blocks > (i64)kNumSMs * 32) blocks = (i64)kNumSMs * 32;
    switch (kp.dim) {
    case 1: pull_labels_kernel<G, L, 1><<<(unsigned)blocks, 256, 0, stream>>>(kp, vol, (const G *)grid, out); break;
    case 2: pull_labels_kernel<G, L, 2><<<(unsigned)blocks, 256, 0, stream>>>(kp, vol, (const G *)grid, out); break;
    case 3: pull_labels_kernel<G, L, 3><<<(unsigned)blocks, 256, 0, stream>>>(kp, vol, (const G *)grid, out); break;
    default: return IB200_ERR_DIM;
    }
    note_launch("pull_labels");
    IB200_CUDA_CHECK(cudaGetLastError());
    return IB200_OK;
}

template <typename G>
static int launch_labels_g(const KParams &kp, int label_type, const void *vol, const void *grid, void *out, cudaStream_t stream) {
    switch (label_type) {
    case IB200_LABEL_I32: return launch_labels_t<G, int>(kp, vol, grid, out, stream);
    case IB200_LABEL_I64: return launch_labels_t<G, long long>(kp, vol, grid, out, stream);
    case IB200_LABEL_U8: return launch_labels_t<G, unsigned char>(kp, vol, grid, out, stream);
    case IB200_LABEL_I16: return launch_labels_t<G, short>(kp, vol, grid, out, stream);
    }
    return IB200_ERR_DTYPE;
}

int launch_pull_labels(const KParams &kp, int grid_dtype, int label_type, const void *vol, const void *grid, void *out, cudaStream_t stream) {
    for (int d = 0; d < kp.dim; ++d)
        if (kp.order[d] > 1) return IB200_ERR_ORDER;
    switch (grid_dtype) {
    case IB200_F32: return launch_labels_g<float>(kp, label_type, vol, grid, out, stream);
    case IB200_F64: return launch_labels_g<double>(kp, label_type, vol, grid, out, stream);
    }
    return IB200_ERR_DTYPE;
}

}  // namespace ib200
