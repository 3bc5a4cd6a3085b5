"""GPU parity of the TILED kernels (the default path for dense 3-D problems with >= 32768 lattice points)
against the float64 oracle: pull / grad / push / count x orders 1-7 x {float32, float16} x mixed per-dim
bounds x extrapolate 0 / 1 / 2, on shapes above the dispatch threshold with partial tiles, asserting that
the tile (or pipe) kernel -- not the generic one-thread-per-point kernel -- produced the result.
Full-size BASELINE configs 3, 4 and 5 (one shard) against the OpenMP oracle.

Reference: interpol/nd.py:81-288 (pull / push / grad), interpol/iso1.py (order 1)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from test_gpu_ops import smooth_grid, to_np

pytestmark = pytest.mark.gpu

SHAPE = (40, 44, 36)          # 63 360 points: above the 32 768 threshold, partial tiles along x, y and z
VSHAPE = (36, 50, 41)         # the volume has its own (odd) shape


def _case(order, seed, dtype=torch.float32, B=2, C=2, amp=3.0):
    gen = torch.Generator().manual_seed(seed)
    vol = torch.randn([B, C, *VSHAPE], generator=gen)
    img = torch.randn([B, C, *SHAPE], generator=gen)
    grid = smooth_grid(SHAPE, gen, amp=amp, batch=B)
    scale = torch.tensor([VSHAPE[d] / SHAPE[d] for d in range(3)])
    grid = (grid * scale - 1.25).contiguous()                 # leaves the field of view on the low side
    if dtype != torch.float32:
        vol, img, grid = vol.to(dtype), img.to(dtype), grid.to(dtype)
    return vol, img, grid


def _tiled(name, op):
    return any(name.startswith(op + sfx) for sfx in ('_tile3d', '_pipe3d', '_box3d'))


@pytest.mark.parametrize('extrapolate', [1, 0, 2])
@pytest.mark.parametrize('order', [1, 2, 3, 4, 5, 6, 7])
def test_tile_kernels_vs_oracle_f32(order, extrapolate):
    import oracle
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    vol, img, grid = _case(order, 500 + 10 * order + extrapolate)
    grid = _off_threshold(grid, VSHAPE)
    bound = [(order + extrapolate) % 7, (order + 2) % 7, (order + 4 + extrapolate) % 7]
    o = [order]
    v64, i64, g64 = vol.double().numpy(), img.double().numpy(), grid.double().numpy()
    tol = 1e-5
    got = pp.grid_pull(vol.cuda(), grid.cuda(), bound, o, extrapolate)
    assert _tiled(ib.last_kernel(), 'pull'), ib.last_kernel()
    assert rel_err(to_np(got), oracle.grid_pull(v64, g64, bound, o, extrapolate)) <= tol
    got = pp.grid_grad(vol.cuda(), grid.cuda(), bound, o, extrapolate)
    assert _tiled(ib.last_kernel(), 'grad'), ib.last_kernel()
    assert rel_err(to_np(got), oracle.grid_grad(v64, g64, bound, o, extrapolate)) <= tol
    got = pp.grid_push(img.cuda(), grid.cuda(), list(VSHAPE), bound, o, extrapolate)
    assert _tiled(ib.last_kernel(), 'push'), ib.last_kernel()
    assert rel_err(to_np(got), oracle.grid_push(i64, g64, VSHAPE, bound, o, extrapolate)) <= tol
    got = pp.grid_count(grid.cuda(), list(VSHAPE), bound, o, extrapolate)
    assert _tiled(ib.last_kernel(), 'count'), ib.last_kernel()
    assert rel_err(to_np(got), oracle.grid_count(g64, VSHAPE, bound, o, extrapolate)) <= tol


def _off_threshold(grid, vshape):
    """extrapolate 0 / 2 keep a point iff -t < g < n - 1 + t with the thresholds rounded to the GRID's dtype
    (nd.py:11-27): a float32 coordinate that equals the rounded threshold is masked in float32 and kept in
    float64.  Move such points (one in ~1e6) so that the float64 oracle answers the same question."""
    grid = grid.clone()
    for d in range(3):
        for t in (0.05, 0.55):
            for thr in (-t, vshape[d] - 1 + t):
                near = (grid[..., d] - thr).abs() < 1e-4
                grid[..., d][near] += 3e-4
    return grid


def _splat_tol(fn32, fn64):
    """Tolerance for push / count: 1e-5, or the reference's own float32 noise floor where that is larger --
    bounds that pile everything outside the field of view onto the border voxels (replicate) accumulate
    hundreds of float32 partial sums per voxel and the float32 reference itself is 3-4e-5 away from float64."""
    return max(1e-5, 1.5 * rel_err(np.asarray(fn32, dtype=np.float64), fn64))


@pytest.mark.parametrize('bound', range(7))
@pytest.mark.parametrize('order', [1, 3, 5])
def test_tile_kernels_every_bound_f32(order, bound):
    """isotropic bound, all seven, deformation reaching well outside the volume on both sides"""
    import oracle
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    vol, img, grid = _case(order, 900 + 10 * order + bound, amp=6.0)
    grid = _off_threshold((grid * 1.2 - 3.0).contiguous(), VSHAPE)
    b, o = [bound], [order]
    v64, i64, g64 = vol.double().numpy(), img.double().numpy(), grid.double().numpy()
    for ex in (1, 0):
        got = pp.grid_pull(vol.cuda(), grid.cuda(), b, o, ex)
        assert _tiled(ib.last_kernel(), 'pull'), ib.last_kernel()
        assert rel_err(to_np(got), oracle.grid_pull(v64, g64, b, o, ex)) <= 1e-5
        got = pp.grid_grad(vol.cuda(), grid.cuda(), b, o, ex)
        assert rel_err(to_np(got), oracle.grid_grad(v64, g64, b, o, ex)) <= 1e-5
        got = pp.grid_push(img.cuda(), grid.cuda(), list(VSHAPE), b, o, ex)
        assert _tiled(ib.last_kernel(), 'push'), ib.last_kernel()
        want = oracle.grid_push(i64, g64, VSHAPE, b, o, ex)
        assert rel_err(to_np(got), want) <= _splat_tol(oracle.grid_push(img.numpy(), grid.numpy(), VSHAPE, b, o, ex), want)
        got = pp.grid_count(grid.cuda(), list(VSHAPE), b, o, ex)
        want = oracle.grid_count(g64, VSHAPE, b, o, ex)
        assert rel_err(to_np(got), want) <= _splat_tol(oracle.grid_count(grid.numpy(), VSHAPE, b, o, ex), want)


@pytest.mark.parametrize('extrapolate', [1, 0])
@pytest.mark.parametrize('order', [1, 2, 3, 4, 5, 6, 7])
def test_tile_kernels_vs_oracle_f16(order, extrapolate):
    """float16 storage (volume, image and grid), float32 arithmetic: 1e-2 against the float64 oracle on the
    float16-rounded inputs (north-star tolerance; SURVEY 8d parity gates)"""
    import oracle
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    vol, img, grid = _case(order, 700 + 10 * order + extrapolate, dtype=torch.float16)
    bound = [(order + 1) % 7, 6, (order + 3) % 7]
    o = [order]
    v64, i64, g64 = vol.double().numpy(), img.double().numpy(), grid.double().numpy()
    got = pp.grid_pull(vol.cuda(), grid.cuda(), bound, o, extrapolate)
    assert got.dtype == torch.float16
    assert _tiled(ib.last_kernel(), 'pull'), ib.last_kernel()
    assert rel_err(to_np(got), oracle.grid_pull(v64, g64, bound, o, extrapolate)) <= 1e-2
    got = pp.grid_grad(vol.cuda(), grid.cuda(), bound, o, extrapolate)
    assert _tiled(ib.last_kernel(), 'grad'), ib.last_kernel()
    assert rel_err(to_np(got), oracle.grid_grad(v64, g64, bound, o, extrapolate)) <= 1e-2
    got = pp.grid_push(img.cuda(), grid.cuda(), list(VSHAPE), bound, o, extrapolate)
    assert got.dtype == torch.float16
    assert rel_err(to_np(got), oracle.grid_push(i64, g64, VSHAPE, bound, o, extrapolate)) <= 1e-2
    got = pp.grid_count(grid.cuda(), list(VSHAPE), bound, o, extrapolate)
    assert rel_err(to_np(got), oracle.grid_count(g64, VSHAPE, bound, o, extrapolate)) <= 1e-2


def test_push_tile_nonfinite_values_propagate():
    """NaN / Inf in the pushed image must reach the touched voxels (as in the reference's
    scatter_add_ and in the generic kernel), not turn into finite garbage."""
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    vol, img, grid = _case(3, 4242, B=1, C=1)
    img = img.clone()
    img[0, 0, 10, 12, 9] = float('nan')
    img[0, 0, 30, 5, 20] = float('inf')
    img[0, 0, 31, 40, 3] = -float('inf')
    for order in (1, 3):
        got = pp.grid_push(img.cuda(), grid.cuda(), list(VSHAPE), [3], [order], 1)
        assert _tiled(ib.last_kernel(), 'push'), ib.last_kernel()
        old = pp.flags
        pp.flags = 1            # NO_TILES: the generic scatter kernel
        try:
            want = pp.grid_push(img.cuda(), grid.cuda(), list(VSHAPE), [3], [order], 1)
        finally:
            pp.flags = old
        assert torch.equal(torch.isnan(got), torch.isnan(want))
        assert torch.equal(torch.isposinf(got), torch.isposinf(want))
        assert torch.equal(torch.isneginf(got), torch.isneginf(want))
        ok = torch.isfinite(want)
        assert rel_err(to_np(got[ok]), to_np(want[ok])) <= 1e-5


@pytest.mark.parametrize('ratio', [1e3, 1e5, 1e8])
def test_push_tile_high_dynamic_range(ratio):
    """one large outlier among unit-scale values must not wipe out the precision of its
    neighbours (float64 oracle, error measured on the voxels the outlier does not touch, relative to THEIR
    magnitude)."""
    import oracle
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    vol, img, grid = _case(3, 777, B=1, C=1)
    img = img.clone()
    img[0, 0, 20, 22, 18] = ratio
    got = pp.grid_push(img.cuda(), grid.cuda(), list(VSHAPE), [3], [3], 1)
    assert _tiled(ib.last_kernel(), 'push'), ib.last_kernel()
    want = oracle.grid_push(img.double().numpy(), grid.double().numpy(), VSHAPE, [3], [3], 1)
    base = img.clone(); base[0, 0, 20, 22, 18] = 0
    want0 = oracle.grid_push(base.double().numpy(), grid.double().numpy(), VSHAPE, [3], [3], 1)
    far = np.abs(want - want0) == 0                  # voxels the outlier does not reach
    assert far.sum() > 0.9 * far.size
    assert rel_err(to_np(got)[far], want[far]) <= 1e-5
    assert rel_err(to_np(got), want) <= 1e-5


# ------------------------------------------------ BASELINE configs at full size --

def _bench():
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    return bench


def test_cfg3_full_size_vs_oracle():
    """BASELINE config 3: 256^3 fp32, 4 channels, cubic, dct2: spline_coeff_nd -> grid_pull, grid_grad."""
    import oracle
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    bench = _bench()
    vol, grid = bench.make_workload(256, 'cuda', channels=4)
    coeff = ib.spline_coeff_nd(vol, interpolation=3, bound='dct2', dim=3)
    v64, g64 = to_np(vol), to_np(grid)
    c64 = oracle.spline_coeff_nd(v64, [3], [3], 3)
    assert rel_err(to_np(coeff), c64) <= 1e-5
    del v64
    pull = pp.grid_pull(coeff, grid, [3], [3], 1)
    k_pull = ib.last_kernel()
    grad = pp.grid_grad(coeff, grid, [3], [3], 1)
    k_grad = ib.last_kernel()
    assert _tiled(k_pull, 'pull') and _tiled(k_grad, 'grad'), (k_pull, k_grad)
    cc = to_np(coeff)
    assert rel_err(to_np(pull), oracle.grid_pull(cc, g64, [3], [3], 1)) <= 1e-5
    assert rel_err(to_np(grad), oracle.grid_grad(cc, g64, [3], [3], 1)) <= 1e-5


@pytest.mark.parametrize('kind', ['smooth', 'incoherent'])
def test_cfg4_full_size_vs_oracle(kind):
    """BASELINE config 4: 256^3 fp16 (volume and grid), order 5, dft, grid_push + grid_count; smooth
    deformation and identity + randn * 20 (scatter stress).  1e-2 vs the float64 oracle on the rounded inputs."""
    import oracle
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    bench = _bench()
    vol, grid = bench.make_workload(256, 'cuda', dtype=torch.float16, incoherent=(kind == 'incoherent'))
    push = pp.grid_push(vol, grid, [256] * 3, [6], [5], 1)
    count = pp.grid_count(grid, [256] * 3, [6], [5], 1)
    v64, g64 = to_np(vol), to_np(grid)
    assert rel_err(to_np(push), oracle.grid_push(v64, g64, [256] * 3, [6], [5], 1, nthreads=8)) <= 1e-2
    assert rel_err(to_np(count), oracle.grid_count(g64, [256] * 3, [6], [5], 1, nthreads=8)) <= 1e-2


def test_cfg5_shard_vs_oracle():
    """BASELINE config 5, one GPU's shard: batch 8 x 192^3 fp32, cubic, bounds (dct2, dft, zero), pull + push."""
    import oracle
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    bench = _bench()
    vol, grid = bench.make_workload(192, 'cuda', batch=8)
    bound = [3, 6, 0]
    pull = pp.grid_pull(vol, grid, bound, [3], 1)
    assert _tiled(ib.last_kernel(), 'pull'), ib.last_kernel()
    push = pp.grid_push(pull, grid, [192] * 3, bound, [3], 1)
    assert _tiled(ib.last_kernel(), 'push'), ib.last_kernel()
    v64, g64 = to_np(vol), to_np(grid)
    assert rel_err(to_np(pull), oracle.grid_pull(v64, g64, bound, [3], 1)) <= 1e-5
    assert rel_err(to_np(push), oracle.grid_push(to_np(pull), g64, [192] * 3, bound, [3], 1, nthreads=8)) <= 1e-5


def test_headline_256_vs_oracle():
    """The north-star configuration itself (256^3 fp32 cubic dct2, C = 1) against the OpenMP oracle."""
    import oracle
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    bench = _bench()
    vol, grid = bench.make_workload(256, 'cuda')
    pull = pp.grid_pull(vol, grid, [3], [3], 1)
    k = ib.last_kernel()
    push = pp.grid_push(pull, grid, [256] * 3, [3], [3], 1)
    v64, g64 = to_np(vol), to_np(grid)
    assert rel_err(to_np(pull), oracle.grid_pull(v64, g64, [3], [3], 1)) <= 1e-5, k
    assert rel_err(to_np(push), oracle.grid_push(to_np(pull), g64, [256] * 3, [3], [3], 1, nthreads=8)) <= 1e-5


# ------------------------------------------------------------ fused backward --

@pytest.mark.parametrize('channels', [1, 3])
@pytest.mark.parametrize('order', [1, 3, 5])
def test_fused_backward_vs_oracle_composition(order, channels):
    """GridPull / GridPush / GridCount / GridGrad backward at tiled sizes against the reference's algebra
    (pushpull.py:237-325) evaluated with the float64 oracle: push(grad) + sum_c grad(vol) * grad_out, etc.
    The grid branch runs the fused kernels (pullbwd_pipe3d / pullbwd_tile3d; no (B,C,N,D) temporary)."""
    import oracle
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    gen = torch.Generator().manual_seed(1300 + 10 * order + channels)
    B = 2
    vol = torch.randn([B, channels, *VSHAPE], generator=gen)
    grid = smooth_grid(SHAPE, gen, amp=3.0, batch=B)
    grid = (grid * torch.tensor([VSHAPE[d] / SHAPE[d] for d in range(3)]) - 1.25).contiguous()
    gout = torch.randn([B, channels, *SHAPE], generator=gen)
    bound, o, ex = [3, 6, 1], [order], 1
    v64, g64, o64 = vol.double().numpy(), grid.double().numpy(), gout.double().numpy()
    # ---- pull
    v = vol.cuda().requires_grad_(); g = grid.cuda().requires_grad_()
    gi, gg = pp.grid_pull_backward(gout.cuda(), v, g, bound, o, ex)
    assert ib.last_kernel().startswith(('pullbwd_tile3d', 'pullbwd_pipe3d')), ib.last_kernel()
    want_gi = oracle.grid_push(o64, g64, VSHAPE, bound, o, ex)
    want_gg = (oracle.grid_grad(v64, g64, bound, o, ex) * o64[..., None]).sum(1)
    assert rel_err(to_np(gi), want_gi) <= 1e-5
    assert rel_err(to_np(gg), want_gg) <= 1e-5
    # ---- push (roles swapped: inp lives on the lattice, grad on the volume)
    img = gout.cuda().requires_grad_()
    gvol = vol.cuda()
    gi, gg = pp.grid_push_backward(gvol, img, g, bound, o, ex)
    assert ib.last_kernel().startswith(('pullbwd_tile3d', 'pullbwd_pipe3d')), ib.last_kernel()
    assert rel_err(to_np(gi), oracle.grid_pull(v64, g64, bound, o, ex)) <= 1e-5
    assert rel_err(to_np(gg), want_gg) <= 1e-5
    # ---- count: grad (B, 1, *vshape)
    gcnt = vol[:, :1].cuda()
    gg = pp.grid_count_backward(gcnt, g, bound, o, ex)
    want = oracle.grid_grad(v64[:, :1], g64, bound, o, ex)[:, 0]
    assert rel_err(to_np(gg), want) <= 1e-5
    # ---- grad: gout (B, C, *shape, 3); fused Hessian contraction (generic kernel, no (B,C,N,3,3) temporary)
    gout3 = torch.randn([B, channels, *SHAPE, 3], generator=gen)
    gi, gg = pp.grid_grad_backward(gout3.cuda(), v, g, bound, o, ex)
    assert ib.last_kernel().startswith('gather_grad_bwd_grid'), ib.last_kernel()
    o3 = gout3.double().numpy()
    assert rel_err(to_np(gi), oracle.grid_pushgrad(o3, g64, VSHAPE, bound, o, ex)) <= 1e-5
    hess = oracle.grid_hess(v64, g64, bound, o, ex)
    want = (hess * o3[..., None]).sum(axis=(1, -2))
    if np.abs(want).max() > 0:
        assert rel_err(to_np(gg), want) <= 1e-5
    else:
        assert np.abs(to_np(gg)).max() == 0


def test_fused_backward_full_size():
    """256^3 cubic, C = 1 (the headline shape): autograd through grid_pull, grid branch on the persistent
    fused kernel, checked against the unfused composition on the GPU and (grid branch) the oracle."""
    import oracle
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    bench = _bench()
    vol, grid = bench.make_workload(256, 'cuda')
    gout = torch.randn(vol.shape, generator=torch.Generator().manual_seed(5)).cuda()
    v = vol.clone().requires_grad_(); g = grid.clone().requires_grad_()
    out = ib.grid_pull(v, g, interpolation=3, bound='dct2', extrapolate=True)
    out.backward(gout)            # (runs on the autograd thread: the per-thread launch log is checked below)
    _, gg2 = pp.grid_pull_backward(gout, v.detach().requires_grad_(), g.detach().requires_grad_(), [3], [3], 1)
    assert ib.last_kernel().startswith('pullbwd_pipe3d'), ib.last_kernel()
    assert torch.equal(gg2, g.grad)
    ref_gg = pp.grid_grad(vol, grid, [3], [3], 1)[:, 0] * gout[0, 0, ..., None]
    assert rel_err(to_np(g.grad), to_np(ref_gg)) <= 1e-6
    ref_gi = pp.grid_push(gout, grid, [256] * 3, [3], [3], 1)
    assert rel_err(to_np(v.grad), to_np(ref_gi)) <= 1e-6
    want = oracle.grid_grad(to_np(vol), to_np(grid), [3], [3], 1)[:, 0] * to_np(gout)[0, 0, ..., None]
    assert rel_err(to_np(g.grad), want) <= 1e-5


# --------------------------------------------------------- displacement fields --

@pytest.mark.parametrize('dim,shape', [(1, (300,)), (2, (48, 40)), (3, (20, 24, 18)), (3, SHAPE)])
def test_displacement_mode_equals_identity_plus_grid(dim, shape):
    """`displacement=True` (IB200_FLAG_DISPLACEMENT: coordinate = lattice index + grid value, formed in registers)
    returns what add_identity_grid + the same call returns (interpol/api.py:482-520), for every op, the generic
    and the tiled kernels, float32 and float16, and through autograd."""
    import interpol_b200 as ib
    gen = torch.Generator().manual_seed(60 + dim + len(shape) * shape[0])
    B, C = 2, 2
    vol = torch.randn([B, C, *shape], generator=gen).cuda()
    disp = (smooth_grid(shape, gen, amp=3.0, batch=B) - ib.identity_grid(shape)).contiguous().cuda()
    grid = ib.add_identity_grid(disp)
    for order, bound, ex in ((3, 'dct2', True), (1, 'zero', False), (2, 'dft', 2)):
        kw = dict(interpolation=order, bound=bound, extrapolate=ex)
        for fn in (ib.grid_pull, ib.grid_grad, ib.grid_push):
            a = fn(vol, disp, displacement=True, **kw)
            b = fn(vol, grid, **kw)
            assert rel_err(to_np(a), to_np(b)) <= 2e-6, (fn.__name__, order, bound)
        a = ib.grid_count(disp, displacement=True, **kw)
        b = ib.grid_count(grid, **kw)
        assert rel_err(to_np(a), to_np(b)) <= 2e-6
    if dim == 3:
        # float16 storage: the lattice index is added in float32 (more accurate than a materialised float16 grid)
        h = ib.grid_pull(vol.half(), disp.half(), interpolation=3, bound='dct2', extrapolate=True, displacement=True)
        assert h.dtype == torch.float16
        want = ib.grid_pull(vol.half().float(), ib.add_identity_grid(disp.half().float()), interpolation=3, bound='dct2',
                            extrapolate=True)
        assert rel_err(to_np(h), to_np(want)) <= 2e-3
        lab = torch.randint(0, 5, [B, 1, *shape], generator=gen).cuda()
        assert torch.equal(ib.grid_pull(lab, disp, interpolation=1, bound='dct2', extrapolate=True, displacement=True),
                           ib.grid_pull(lab, grid, interpolation=1, bound='dct2', extrapolate=True))
    # autograd: d/d(displacement) == d/d(coordinate)
    v1 = vol.clone().requires_grad_(); d1 = disp.clone().requires_grad_()
    v2 = vol.clone().requires_grad_(); g2 = grid.clone().requires_grad_()
    gout = torch.randn(vol.shape[:2] + tuple(shape), generator=gen).cuda()
    ib.grid_pull(v1, d1, interpolation=3, bound='dct2', extrapolate=True, displacement=True).backward(gout)
    ib.grid_pull(v2, g2, interpolation=3, bound='dct2', extrapolate=True).backward(gout)
    assert rel_err(to_np(v1.grad), to_np(v2.grad)) <= 2e-6
    assert rel_err(to_np(d1.grad), to_np(g2.grad)) <= 2e-6


# ------------------------------------------------- channel-interleaved boxes --

@pytest.mark.parametrize('channels', [4, 8])
@pytest.mark.parametrize('extrapolate', [1, 0, 2])
@pytest.mark.parametrize('order', [1, 2, 3])
def test_interleaved_channel_boxes_vs_oracle(order, extrapolate, channels):
    """float32 volumes with a multiple of 4 channels: pull / grad through the channel-interleaved boxes (one LDS.128
    per node for four channels, `*_c4` kernels), every bound, partial tiles, points outside the field of view; a steep
    deformation sends some groups through the incoherent fallback of the same kernel.  Strided (padded) volumes too."""
    import oracle
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    for amp, bound in ((3.0, [(order + extrapolate) % 7, (order + 2) % 7, (order + 4 + extrapolate) % 7]), (40.0, [3]), (3.0, [0])):
        vol, _, grid = _case(order, 900 + 10 * order + extrapolate + channels, B=2, C=channels, amp=amp)
        grid = _off_threshold(grid, VSHAPE)
        v64, g64 = vol.double().numpy(), grid.double().numpy()
        big = torch.zeros([2, channels, VSHAPE[0], VSHAPE[1], VSHAPE[2] + 3])
        big[..., :VSHAPE[2]] = vol
        for v_dev in (vol.cuda(), big.cuda()[..., :VSHAPE[2]]):
            got = pp.grid_pull(v_dev, grid.cuda(), bound, [order], extrapolate)
            assert ib.last_kernel().startswith('pull_tile3d') and ib.last_kernel().endswith('_c4'), ib.last_kernel()
            assert rel_err(to_np(got), oracle.grid_pull(v64, g64, bound, [order], extrapolate)) <= 1e-5
            got = pp.grid_grad(v_dev, grid.cuda(), bound, [order], extrapolate)
            assert ib.last_kernel().startswith('grad_tile3d') and ib.last_kernel().endswith('_c4'), ib.last_kernel()
            assert rel_err(to_np(got), oracle.grid_grad(v64, g64, bound, [order], extrapolate)) <= 1e-5
