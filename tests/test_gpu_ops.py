"""GPU tests of the CUDA path against the CPU oracle on seeded inputs, at sizes
the oracle finishes in seconds; size-independent properties at full BASELINE
sizes; edge cases; autograd consistency (the reference's own gradcheck test,
interpol/tests/test_gradcheck_pushpull.py, on the CUDA device)."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

BOUNDS = ['zero', 'replicate', 'dct1', 'dct2', 'dst1', 'dst2', 'dft']


def smooth_grid(shape, gen, amp=3.0, batch=1, dtype=torch.float32):
    dim = len(shape)
    coarse = torch.randn([batch, dim] + [6] * dim, generator=gen) * amp
    mode = {1: 'linear', 2: 'bilinear', 3: 'trilinear'}[dim]
    disp = torch.nn.functional.interpolate(coarse, size=list(shape), mode=mode, align_corners=True)
    ident = torch.stack(torch.meshgrid(*[torch.arange(float(s)) for s in shape], indexing='ij'), dim=-1)
    return (disp.movedim(1, -1) + ident).to(dtype)


def to_np(x):
    return x.detach().double().cpu().numpy()


# ------------------------------------------------------------------ oracle --

@pytest.mark.parametrize('order', [0, 1, 2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize('dim', [1, 2, 3])
def test_ops_vs_oracle_f32(dim, order):
    """all six building blocks, smooth deformation leaving the field of view,
    float32 CUDA vs float64 oracle: 1e-5 (north-star tolerance)."""
    import oracle
    from interpol_b200 import pushpull as pp
    gen = torch.Generator().manual_seed(100 * dim + order)
    shape = {1: (300,), 2: (48, 40), 3: (20, 24, 18)}[dim]
    B, C = 2, 3
    vol = torch.randn([B, C, *shape], generator=gen)
    grid = smooth_grid(shape, gen, amp=4.0, batch=B)
    srcg = torch.randn([B, C, *shape, dim], generator=gen)
    bound = [(order + d) % 7 for d in range(dim)]
    if order == 0:
        grid = grid + 0.013          # keep away from exact .5 (rounding rule is tested elsewhere)
    o = [order]
    for ex in (1, 0):
        d = lambda t: t.cuda()
        got = {
            'pull': pp.grid_pull(d(vol), d(grid), bound, o, ex),
            'grad': pp.grid_grad(d(vol), d(grid), bound, o, ex),
            'hess': pp.grid_hess(d(vol), d(grid), bound, o, ex),
            'push': pp.grid_push(d(vol), d(grid), list(shape), bound, o, ex),
            'count': pp.grid_count(d(grid), list(shape), bound, o, ex),
            'pushgrad': pp.grid_pushgrad(d(srcg), d(grid), list(shape), bound, o, ex),
        }
        v64, g64, s64 = vol.double().numpy(), grid.double().numpy(), srcg.double().numpy()
        want = {
            'pull': oracle.grid_pull(v64, g64, bound, o, ex),
            'grad': oracle.grid_grad(v64, g64, bound, o, ex),
            'hess': oracle.grid_hess(v64, g64, bound, o, ex),
            'push': oracle.grid_push(v64, g64, shape, bound, o, ex),
            'count': oracle.grid_count(g64, shape, bound, o, ex),
            'pushgrad': oracle.grid_pushgrad(s64, g64, shape, bound, o, ex),
        }
        for k in got:
            if np.abs(want[k]).max() == 0:
                assert np.abs(to_np(got[k])).max() == 0
                continue
            err = rel_err(to_np(got[k]), want[k])
            assert err <= 1e-5, (k, dim, order, ex, err)


@pytest.mark.parametrize('dtype,tol', [(torch.float16, 1e-2), (torch.bfloat16, 4e-2)], ids=['f16', 'bf16'])
@pytest.mark.parametrize('order', [1, 3, 5])
def test_16bit_vs_oracle(order, dtype, tol):
    """16-bit storage, float32 arithmetic, against the float64 oracle on the
    16-bit-rounded inputs (north-star: 1e-2 rel for fp16)."""
    import oracle
    from interpol_b200 import pushpull as pp
    gen = torch.Generator().manual_seed(7 + order)
    shape = (24, 20, 28)
    vol = torch.randn([1, 2, *shape], generator=gen).to(dtype)
    grid = smooth_grid(shape, gen, amp=2.0).to(dtype)
    bound = [6]
    v64, g64 = vol.double().numpy(), grid.double().numpy()
    pull = pp.grid_pull(vol.cuda(), grid.cuda(), bound, [order], 1)
    assert pull.dtype == dtype
    assert rel_err(to_np(pull), oracle.grid_pull(v64, g64, bound, [order], 1)) <= tol
    push = pp.grid_push(vol.cuda(), grid.cuda(), list(shape), bound, [order], 1)
    assert rel_err(to_np(push), oracle.grid_push(v64, g64, shape, bound, [order], 1)) <= tol
    count = pp.grid_count(grid.cuda(), list(shape), bound, [order], 1)
    assert rel_err(to_np(count), oracle.grid_count(g64, shape, bound, [order], 1)) <= tol
    grad = pp.grid_grad(vol.cuda(), grid.cuda(), bound, [order], 1)
    assert rel_err(to_np(grad), oracle.grid_grad(v64, g64, bound, [order], 1)) <= tol


def test_cfg1_2d_linear_identity():
    """BASELINE config 0: 2-D 256x256 fp32, identity grid, order 1, bound zero,
    extrapolate=False (API default): pull at the nodes is the identity."""
    import interpol_b200 as ib
    x = torch.randn(256, 256, device='cuda')
    g = ib.identity_grid([256, 256], device='cuda')
    y = ib.grid_pull(x, g)
    assert torch.equal(x, y)


def test_cfg2_128_cubic_dct2_vs_oracle():
    """BASELINE config 1 at full size: 128^3 fp32, smooth deformation, cubic, dct2."""
    import oracle
    from interpol_b200 import pushpull as pp
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import make_workload
    vol, grid = make_workload(128, 'cuda')
    out = pp.grid_pull(vol, grid, [3], [3], 1)
    push = pp.grid_push(out, grid, [128] * 3, [3], [3], 1)
    v64, g64 = to_np(vol), to_np(grid)
    ref = oracle.grid_pull(v64, g64, [3], [3], 1)
    assert rel_err(to_np(out), ref) <= 1e-5
    refp = oracle.grid_push(to_np(out), g64, [128] * 3, [3], [3], 1, nthreads=8)
    assert rel_err(to_np(push), refp) <= 1e-5


# -------------------------------------------------------------- properties --

@pytest.mark.parametrize('size', [256])
def test_full_size_properties(size):
    """North-star size (256^3, cubic): properties that need no oracle.
    adjointness <pull(x), y> == <x, push(y)>; partition of unity under dft;
    constants are reproduced under dct2."""
    from interpol_b200 import pushpull as pp
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import make_workload
    vol, grid = make_workload(size, 'cuda')
    y = torch.randn(vol.shape, generator=torch.Generator().manual_seed(99)).cuda()
    for bound in ([3], [6], [0]):
        px = pp.grid_pull(vol, grid, bound, [3], 1)
        py = pp.grid_push(y, grid, [size] * 3, bound, [3], 1)
        lhs = (px.double() * y.double()).sum().item()
        rhs = (vol.double() * py.double()).sum().item()
        # the inner products cancel to O(sqrt(N)); the error scale is eps * sum |px| |y|
        scale = (px.double().abs() * y.double().abs()).sum().item()
        assert abs(lhs - rhs) <= 5e-6 * scale, (bound, lhs, rhs, scale)
    cnt = pp.grid_count(grid, [size] * 3, [6], [3], 1)
    assert abs(cnt.double().sum().item() - size ** 3) <= 1e-6 * size ** 3
    const = torch.full_like(vol, 2.5)
    out = pp.grid_pull(const, grid, [3], [3], 1)
    assert (out - 2.5).abs().max().item() <= 1e-5
    # tiled and direct kernels agree
    pp.flags = 1
    try:
        direct = pp.grid_pull(vol, grid, [3], [3], 1)
    finally:
        pp.flags = 0
    tiled = pp.grid_pull(vol, grid, [3], [3], 1)
    assert rel_err(to_np(tiled), to_np(direct)) <= 2e-6


def test_prefilter_pull_identity():
    """interpol/tests/test_coeff.py: resize to the same shape is the identity.
    (The reference test draws unseeded inputs and uses allclose's default atol = 1e-8; its own algorithm --
    the oracle reproduces it -- misses that by up to 3e-7 on ~0.05 % of the draws, e.g. n = 7, dft, order 7,
    because of the truncated boundary sums of coeff.py:82-105.  Seeded inputs and atol = 1e-6 here.)"""
    import interpol_b200 as ib
    gen = torch.Generator().manual_seed(1234)
    for length in [1, 2, 3, 7, 9, 11]:
        for bound in ['dct1', 'dct2', 'dft']:
            for order in range(8):
                x = torch.randn([length], dtype=torch.double, generator=gen).cuda()
                y = ib.resize(x, shape=[length], bound=bound, interpolation=order)
                assert torch.allclose(x, y, atol=1e-6), (length, bound, order)


def test_prefilter_large_axes():
    """prefilter over all three axes of a 96x80x72 volume, orders 2..7, vs oracle."""
    import oracle
    import interpol_b200 as ib
    x = torch.randn(2, 96, 80, 72)
    for order in range(2, 8):
        for bound, code in (('dct2', 3), ('dct1', 2), ('dft', 6)):
            out = ib.spline_coeff_nd(x.cuda(), interpolation=order, bound=bound, dim=3)
            ref = oracle.spline_coeff_nd(x.double().numpy(), [code], [order], 3)
            assert rel_err(to_np(out), ref) <= 1e-5, (order, bound)
    # a line too long for shared memory takes the global fallback
    z = torch.randn(3, 70000)
    out = ib.spline_coeff(z.cuda(), interpolation=3, bound='dct2', dim=-1)
    ref = oracle.spline_coeff(z.double().numpy(), 3, 3, dim=-1)
    assert rel_err(to_np(out), ref) <= 1e-5
    out = ib.spline_coeff(z.t().contiguous().cuda(), interpolation=5, bound='dft', dim=0)
    ref = oracle.spline_coeff(z.t().double().numpy(), 6, 5, dim=0)
    assert rel_err(to_np(out), ref) <= 1e-5


# --------------------------------------------------------------- edge cases --

def test_edge_cases():
    import oracle
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    dev = 'cuda'
    # empty batch / empty lattice
    out = pp.grid_pull(torch.zeros(0, 2, 4, 4, device=dev), torch.zeros(0, 3, 3, 2, device=dev), [0], [1], 1)
    assert out.shape == (0, 2, 3, 3)
    out = pp.grid_pull(torch.zeros(1, 2, 4, 4, device=dev), torch.zeros(1, 0, 3, 2, device=dev), [0], [1], 1)
    assert out.shape == (1, 2, 0, 3)
    out = pp.grid_push(torch.zeros(1, 2, 0, 3, device=dev), torch.zeros(1, 0, 3, 2, device=dev), [4, 4], [0], [1], 1)
    assert out.shape == (1, 2, 4, 4) and float(out.abs().sum()) == 0
    # singleton spatial axes, every bound
    vol = torch.randn(1, 1, 1, 5, 1)
    grid = torch.rand(1, 2, 3, 2, 3) * 6 - 2
    for b in range(7):
        for o in (0, 1, 3):
            got = pp.grid_pull(vol.to(dev), grid.to(dev), [b], [o], 1)
            want = oracle.grid_pull(vol.double().numpy(), grid.double().numpy(), [b], [o], 1)
            assert rel_err(to_np(got), want) <= 1e-5 or np.abs(want).max() == 0
    # non-contiguous inputs and broadcast batch (zero strides): no copies needed
    gen = torch.Generator().manual_seed(3)
    vol = torch.randn([1, 3, 12, 14, 10], generator=gen)
    grid = smooth_grid((12, 14, 10), gen, batch=2)
    volT = vol.cuda().permute(0, 1, 4, 3, 2).contiguous().permute(0, 1, 4, 3, 2)     # same values, odd strides
    big = torch.zeros(2, 12, 14, 10, 4, device=dev)
    big[..., :3] = grid.cuda()
    gridT = big[..., :3]                                                           # padded rows
    got = pp.grid_pull(volT, gridT, [3], [3], 1)
    want = oracle.grid_pull(vol.double().numpy(), grid.double().numpy(), [3], [3], 1)
    assert rel_err(to_np(got), want) <= 1e-5
    gridS = grid.cuda()[:, ::2, 1:, :]                                             # strided lattice
    got = pp.grid_push(volT[:, :, ::2, 1:, :].expand(2, 3, 6, 13, 10), gridS, [12, 14, 10], [1], [2], 0)
    want = oracle.grid_push(vol.double().numpy()[:, :, ::2, 1:, :], grid.double().numpy()[:, ::2, 1:, :],
                            [12, 14, 10], [1], [2], 0)
    assert rel_err(to_np(got), want) <= 1e-5
    # API-level broadcasting: (C, *spatial) volume with a (B, *spatial, D) grid
    got = ib.grid_pull(vol[0].cuda(), grid.cuda(), interpolation=2, bound='dft', extrapolate=True)
    assert got.shape == (2, 3, 12, 14, 10)
    want = oracle.grid_pull(vol.double().numpy(), grid.double().numpy(), [6], [2], 1)
    assert rel_err(to_np(got), want) <= 1e-5
    # no channel axis at all
    got = ib.grid_pull(vol[0, 0].cuda(), grid[0].cuda(), interpolation=1, bound='dct2')
    assert got.shape == (12, 14, 10)
    # NaN / inf / huge coordinates contribute nothing and do not fault
    bad = grid.clone()[:1]
    bad[0, 0, 0, 0, 0] = float('nan'); bad[0, 0, 0, 1, 1] = float('inf'); bad[0, 0, 0, 2, 2] = 3e9
    for b in (0, 3, 6):
        got = pp.grid_pull(vol.cuda(), bad.cuda(), [b], [3], 1)
        assert got[0, :, 0, 0, 0].abs().max() == 0 and got[0, :, 0, 0, 1].abs().max() == 0
        assert torch.isfinite(got).all()
        got = pp.grid_push(vol.cuda(), bad.cuda(), [12, 14, 10], [b], [3], 1)
        assert torch.isfinite(got).all()
    # mixed dtypes promote (float32 volume + float64 grid -> float64), like the reference
    got = pp.grid_pull(vol.cuda(), grid[:1].double().cuda(), [3], [1], 1)
    assert got.dtype == torch.float64
    # CPU tensors are staged through the GPU and come back on the CPU
    got = ib.grid_pull(vol, grid, interpolation=3, bound='dct2', extrapolate=True)
    assert got.device.type == 'cpu'
    want = oracle.grid_pull(vol.double().numpy(), grid.double().numpy(), [3], [3], 1)
    assert rel_err(to_np(got), want) <= 1e-5
    # errors
    with pytest.raises(ValueError):
        ib.grid_pull(vol.cuda(), grid.cuda(), interpolation=8)
    with pytest.raises(ValueError):
        ib.grid_pull(vol.cuda(), grid.cuda(), bound='nope')
    with pytest.raises(ValueError):
        pp.grid_push(vol.cuda(), grid.cuda()[:, :5], None, [0], [1], 1)
    with pytest.raises(NotImplementedError):
        pp.grid_pull(torch.zeros(1, 1, 2, 2, 2, 2, device=dev), torch.zeros(1, 2, 2, 2, 2, 4, device=dev), [0], [1], 1)


def test_nearest_rounding_rules():
    """Q5: all-nearest uses round-half-to-even (iso0.py:12); a nearest axis in a
    mixed call uses floor(g + 0.5) (nd.py:45)."""
    from interpol_b200 import pushpull as pp
    x = torch.tensor([1., 2., 3., 4.], device='cuda').reshape(1, 1, 4)
    g = torch.tensor([0.5, 1.5, 2.5], device='cuda').reshape(1, 3, 1)
    out = pp.grid_pull(x, g, [1], [0], 1).flatten().tolist()
    assert out == [1., 3., 3.]
    x2 = x.reshape(1, 1, 4, 1).expand(1, 1, 4, 3).contiguous()
    g2 = torch.stack([g.flatten(), torch.ones(3, device='cuda')], -1).reshape(1, 3, 1, 2)
    out = pp.grid_pull(x2, g2, [1], [0, 1], 1).flatten().tolist()
    assert out == [2., 3., 4.]


def test_survey_tables():
    """1-D responses at integer coordinates for x = [1,2,3,4]."""
    import interpol_b200 as ib
    x = torch.tensor([1., 2., 3., 4.], device='cuda')
    g = torch.arange(-6., 11., device='cuda').unsqueeze(-1)
    rows = {
        'dst1': [0, -4, -3, -2, -1, 0, 0, 2, 3, 4, 0, -4, -3, -2, -1, 0, 0],
        'dst2': [3, 4, -4, -3, -2, -1, 1, 2, 3, 4, -4, -3, -2, -1, 1, 2, 3],
        'dct1': [1, 2, 3, 4, 3, 2, 1, 2, 3, 4, 3, 2, 1, 2, 3, 4, 3],
        'dct2': [3, 4, 4, 3, 2, 1, 1, 2, 3, 4, 4, 3, 2, 1, 1, 2, 3],
        'dft': [3, 4, 1, 2, 3, 4, 1, 2, 3, 4, 1, 2, 3, 4, 1, 2, 3],
        'zero': [0, 0, 0, 0, 0, 0, 1, 2, 3, 4, 0, 0, 0, 0, 0, 0, 0],
        'replicate': [1, 1, 1, 1, 1, 1, 1, 2, 3, 4, 4, 4, 4, 4, 4, 4, 4],
    }
    for bound, row in rows.items():
        for order in (0, 1):
            out = ib.grid_pull(x, g, interpolation=order, bound=bound, extrapolate=True)
            assert out.tolist() == [float(v) for v in row], (bound, order)


# ------------------------------------------------------------------ autograd --

@pytest.mark.parametrize('dim', [1, 2, 3])
@pytest.mark.parametrize('order,bound', [(o, b) for o in range(3) for b in range(7)] + [(o, 3) for o in range(3, 8)])
def test_gradcheck(dim, order, bound):
    """The reference's own test (test_gradcheck_pushpull.py:65-125) on the CUDA
    device: float64, volume 3^dim, batch 2, grid = identity + randn."""
    import interpol_b200 as ib
    from torch.autograd import gradcheck
    torch.manual_seed(1000 * dim + 10 * order + bound)
    shape = (3,) * dim
    kwargs = dict(rtol=1., raise_exception=True, check_undefined_grad=False, nondet_tol=1e-3)
    grid = ib.add_identity_grid_(torch.randn([2, *shape, dim], device='cuda', dtype=torch.double))
    vol = torch.randn((2, 1) + shape, device='cuda', dtype=torch.double)
    vol.requires_grad = True
    grid.requires_grad = True
    assert gradcheck(ib.grid_pull, (vol, grid, order, bound, True), **kwargs)
    assert gradcheck(ib.grid_push, (vol, grid, shape, order, bound, True), **kwargs)
    assert gradcheck(ib.grid_count, (grid, shape, order, bound, True), **kwargs)
    assert gradcheck(ib.grid_grad, (vol, grid, order, bound, True), **kwargs)


def test_gradcheck_mixed_orders_and_prefilter():
    """mixed orders containing 1 (where the reference's gradient has the wrong
    sign, splines.py:96-97) and the prefilter's backward."""
    import interpol_b200 as ib
    from torch.autograd import gradcheck
    torch.manual_seed(5)
    kwargs = dict(rtol=1., raise_exception=True, check_undefined_grad=False, nondet_tol=1e-3)
    grid = ib.add_identity_grid_(torch.randn([2, 3, 4, 2], device='cuda', dtype=torch.double))
    vol = torch.randn((2, 2, 3, 4), device='cuda', dtype=torch.double)
    vol.requires_grad = True
    grid.requires_grad = True
    for orders in ([1, 3], [3, 1], [1, 2], [0, 1]):
        assert gradcheck(ib.grid_pull, (vol, grid, orders, 'dct2', True), **kwargs)
        assert gradcheck(ib.grid_grad, (vol, grid, orders, 'dct2', True), **kwargs)
    assert gradcheck(lambda v: ib.spline_coeff_nd(v, interpolation=3, bound='dct2', dim=2), (vol,), **kwargs)
    assert gradcheck(lambda v, g: ib.grid_pull(v, g, 3, 'dct2', True, True), (vol, grid), **kwargs)


def test_backward_matches_oracle_composition():
    """GridPull.backward == (push(grad), sum_c grad(vol) * grad): fused kernel vs
    the reference's algebra (pushpull.py:237-258) evaluated with the oracle."""
    import oracle
    import interpol_b200 as ib
    gen = torch.Generator().manual_seed(11)
    shape = (14, 12, 16)
    vol = torch.randn([2, 3, *shape], generator=gen)
    grid = smooth_grid(shape, gen, batch=2)
    gout = torch.randn([2, 3, *shape], generator=gen)
    v = vol.cuda().requires_grad_(); g = grid.cuda().requires_grad_()
    out = ib.grid_pull(v, g, interpolation=3, bound='dct2', extrapolate=True)
    out.backward(gout.cuda())
    v64, g64, o64 = vol.double().numpy(), grid.double().numpy(), gout.double().numpy()
    gi = oracle.grid_push(o64, g64, shape, [3], [3], 1)
    gg = (oracle.grid_grad(v64, g64, [3], [3], 1) * o64[..., None]).sum(1)
    assert rel_err(to_np(v.grad), gi) <= 1e-5
    assert rel_err(to_np(g.grad), gg) <= 1e-5


def test_reference_side_by_side():
    """When the driver-installed reference is present (baseline/_ref), run it on
    the CPU beside the CUDA path on the same inputs."""
    import os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref_dir = os.path.join(root, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref_dir, 'interpol')):
        pytest.skip('baseline/_ref not present')
    sys.path.insert(0, ref_dir)
    import warnings
    warnings.filterwarnings('ignore')
    try:
        import interpol as ref
    finally:
        sys.path.remove(ref_dir)
    import interpol_b200 as ib
    gen = torch.Generator().manual_seed(21)
    shape = (32, 28, 36)
    vol = torch.randn([1, 2, *shape], generator=gen)
    grid = smooth_grid(shape, gen)
    for order, bound in ((1, 'zero'), (3, 'dct2'), (2, 'dft'), (5, 'dct1'), (3, 'dst2')):
        for fn in ('grid_pull', 'grid_grad', 'grid_push'):
            want = getattr(ref, fn)(vol.double(), grid.double(), interpolation=order, bound=bound, extrapolate=True)
            got = getattr(ib, fn)(vol.cuda(), grid.cuda(), interpolation=order, bound=bound, extrapolate=True)
            assert rel_err(to_np(got), want.numpy()) <= 1e-5, (fn, order, bound)
        want = ref.grid_count(grid.double(), interpolation=order, bound=bound, extrapolate=True)
        got = ib.grid_count(grid.cuda(), interpolation=order, bound=bound, extrapolate=True)
        assert rel_err(to_np(got), want.numpy()) <= 1e-5
    want = ref.spline_coeff_nd(vol.double(), interpolation=3, bound='dct2', dim=3)
    got = ib.spline_coeff_nd(vol.cuda(), interpolation=3, bound='dct2', dim=3)
    assert rel_err(to_np(got), want.numpy()) <= 1e-5


@pytest.mark.parametrize('dim', [1, 2, 3])
def test_label_pull_fused_matches_label_loop(dim):
    """integer label maps: the one-pass kernel (ib200_pull_labels) returns what the reference's loop over
    `input.unique()` (api.py:194-205, run here through grid_pull of the soft masks) returns."""
    import interpol_b200 as ib
    import interpol_b200.api as api
    gen = torch.Generator().manual_seed(40 + dim)
    shape = {1: (200,), 2: (40, 36), 3: (20, 18, 22)}[dim]
    for dtype in (torch.int64, torch.uint8, torch.int32):
        lab = torch.randint(0, 7, [2, 2, *shape], generator=gen).to(dtype)
        if dtype != torch.uint8:
            lab = lab * 37 - 60            # negative and sparse label values
        grid = smooth_grid(shape, gen, amp=3.0, batch=2).contiguous()
        for order, bound, ex in ((1, 'dct2', True), (0, 'zero', False), ([1, 0, 1][:dim], 'dft', True), (1, 'dst2', 2), (1, 'replicate', False)):
            a = ib.grid_pull(lab.cuda(), grid.cuda(), interpolation=order, bound=bound, extrapolate=ex)
            assert ib.last_kernel() == 'pull_labels', ib.last_kernel()
            api.LABELS_FUSED = False
            try:
                b = ib.grid_pull(lab.cuda(), grid.cuda(), interpolation=order, bound=bound, extrapolate=ex)
            finally:
                api.LABELS_FUSED = True
            assert ib.last_kernel() != 'pull_labels'
            assert a.dtype == b.dtype == dtype and a.shape == b.shape
            if order == 0:
                # nearest-neighbour label maps are bit-exact (north star): a mask value is 0 or 1, nothing to round
                assert torch.equal(a, b), (dim, dtype, order, bound, ex)
            else:
                # identical except where two masks tie to rounding (both kernels sum the same terms in the same order)
                assert (a != b).float().mean().item() <= 1e-4, (dim, dtype, order, bound, ex)
    # int64 labels beyond the int32 range and int8 / int16 storage: read and written natively
    big = (torch.randint(0, 4, [1, 1, *shape], generator=gen) * (2 ** 40) - 2 ** 41).cuda()
    grid1 = smooth_grid(shape, gen, amp=2.0).contiguous().cuda()
    a = ib.grid_pull(big, grid1, interpolation=1, bound='dct2', extrapolate=True)
    assert ib.last_kernel() == 'pull_labels' and a.dtype == torch.int64
    small = torch.div(big + 2 ** 41, 2 ** 40, rounding_mode='floor')
    b = ib.grid_pull(small.to(torch.int16), grid1, interpolation=1, bound='dct2', extrapolate=True)
    c = ib.grid_pull(small.to(torch.int8), grid1, interpolation=1, bound='dct2', extrapolate=True)
    assert b.dtype == torch.int16 and c.dtype == torch.int8
    # (dct2 + extrapolate: every point is reached by some mask, and the relabelling is monotone, so ties break alike)
    assert torch.equal(torch.div(a + 2 ** 41, 2 ** 40, rounding_mode='floor'), b.long())
    assert torch.equal(b.long(), c.long())
    # exact ties (up-sampling by 2 with linear weights: 0.5 / 0.5): the smallest label wins in both
    lab = torch.randint(0, 5, [1, 1, *shape], generator=gen).cuda()
    ident = ib.identity_grid(shape, device='cuda')[None] * 0.5
    a = ib.grid_pull(lab, ident, interpolation=1, bound='replicate', extrapolate=True)
    api.LABELS_FUSED = False
    try:
        b = ib.grid_pull(lab, ident, interpolation=1, bound='replicate', extrapolate=True)
    finally:
        api.LABELS_FUSED = True
    assert torch.equal(a, b)


def test_real_reference_forwards_through_the_seam():
    """The UNMODIFIED reference (baseline/_ref) with this engine installed through its own plugin seam
    (interpol/backend.py:1 + interpol/api.py:186-441): `interpol.grid_pull(...)` on CUDA tensors must run
    this repo's kernels (launch counter advances, kernel names are ours) and return what the reference's own
    TorchScript path returns on the CPU for the same inputs."""
    import os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref_dir = os.path.join(root, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref_dir, 'interpol')):
        pytest.skip('baseline/_ref not present')
    sys.path.insert(0, ref_dir)
    import warnings
    warnings.filterwarnings('ignore')
    try:
        import interpol as ref
    finally:
        sys.path.remove(ref_dir)
    import interpol_b200 as ib
    gen = torch.Generator().manual_seed(31)
    shape = (40, 36, 44)
    vol = torch.randn([1, 2, *shape], generator=gen)
    grid = smooth_grid(shape, gen)           # (component-major strides: densified on the way in)
    kw = dict(interpolation=3, bound='dct2', extrapolate=True)
    want = {fn: getattr(ref, fn)(vol.double(), grid.double(), **kw) for fn in ('grid_pull', 'grid_grad', 'grid_push')}
    want['grid_count'] = ref.grid_count(grid.double(), **kw)
    want['coeff'] = ref.spline_coeff_nd(vol.double(), interpolation=3, bound='dct2', dim=3)
    assert ref.backend.jitfields is False
    ib.install_as_backend(ref)
    try:
        assert ref.backend.jitfields is True and ref.api.jitfields.available
        n0 = ib.launch_count()
        got = {fn: getattr(ref, fn)(vol.cuda(), grid.cuda(), **kw) for fn in ('grid_pull', 'grid_grad', 'grid_push')}
        assert ib.last_kernel().startswith(('push_box3d', 'push_tile3d')), ib.last_kernel()
        got['grid_count'] = ref.grid_count(grid.cuda(), **kw)
        got['coeff'] = ref.spline_coeff_nd(vol.cuda(), interpolation=3, bound='dct2', dim=3)
        assert ib.last_kernel().startswith('coeff_'), ib.last_kernel()
        assert ib.launch_count() - n0 >= 7                       # pull, grad, push, count, 3 prefilter passes
        for k in want:
            assert got[k].is_cuda
            assert rel_err(to_np(got[k]), want[k].numpy()) <= 1e-5, k
    finally:
        ref.backend.jitfields = False


# ------------------------------------------------------------- host tensors --

@pytest.mark.parametrize('pinned', [True, False])
def test_host_tensors_streamed_equals_plain(pinned, monkeypatch):
    """CPU tensors: the slab-streamed pull / grad (upload, gather and download overlapped, api._sample_streamed)
    returns what one upload + one launch returns (to rounding: a slab may take another kernel than the whole
    lattice), inside and outside a stage_scope, and leaves the device twins the next call of the scope needs."""
    import interpol_b200 as ib
    from interpol_b200 import api
    gen = torch.Generator().manual_seed(77)
    shape, vshape = (44, 20, 24), (30, 22, 26)
    vol = torch.randn([2, 2, *vshape], generator=gen)
    grid = smooth_grid(shape, gen, amp=3.0, batch=2).contiguous()
    if pinned:
        vol, grid = vol.pin_memory(), grid.pin_memory()
    kw = dict(interpolation=3, bound=['dct2', 'dft', 'zero'], extrapolate=True)
    plain_pull = ib.grid_pull(vol, grid, prefilter=True, **kw)
    plain_grad = ib.grid_grad(vol, grid, **kw)
    assert plain_pull.device.type == 'cpu'
    monkeypatch.setattr(api, 'STREAM_MIN_BYTES', 1)
    monkeypatch.setattr(api, 'STREAM_SLAB_BYTES', 16 << 10)
    assert api._streamable(vol, grid, False)
    got = ib.grid_pull(vol, grid, prefilter=True, **kw)
    assert got.device.type == 'cpu' and got.shape == plain_pull.shape

    def close(a, b):
        return rel_err(a.double().numpy(), b.double().numpy()) <= 3e-6
    assert close(got, plain_pull)
    assert close(ib.grid_grad(vol, grid, **kw), plain_grad)
    # no channel / batch axes, ragged last slab
    v1, g1 = vol[0, 0].contiguous(), grid[0, :37].contiguous()
    monkeypatch.setattr(api, 'STREAM_MIN_BYTES', 1 << 40)
    want = ib.grid_pull(v1, g1, **kw)
    monkeypatch.setattr(api, 'STREAM_MIN_BYTES', 1)
    got = ib.grid_pull(v1, g1, **kw)
    assert got.shape == (37, 20, 24) and close(got, want)
    # inside a scope: the push that follows finds the grid and the pulled image on the device
    with ib.stage_scope() as scope:
        pulled = ib.grid_pull(vol, grid, **kw)
        dev = torch.device('cuda', torch.cuda.current_device())
        assert scope.lookup(grid, dev) is not None and scope.lookup(pulled, dev) is not None
        pushed = ib.grid_push(pulled, grid, shape=list(vshape), **kw)
    monkeypatch.setattr(api, 'STREAM_MIN_BYTES', 1 << 40)
    with ib.stage_scope():
        pulled0 = ib.grid_pull(vol, grid, **kw)
        pushed0 = ib.grid_push(pulled0, grid, shape=list(vshape), **kw)
    assert close(pulled, pulled0) and close(pushed, pushed0)
    # tensors that need gradients, displacement fields and device tensors keep the plain path
    assert not api._streamable(vol.clone().requires_grad_(), grid, False)
    assert not api._streamable(vol, grid, True)
    assert not api._streamable(vol.cuda(), grid.cuda(), False)


def test_cuda_graph_capture_and_replay():
    """Every op is capturable in a CUDA graph (no host synchronisation, no allocation outside torch's pool, tensor maps
    passed by value): a launch-bound registration step (pull, grad, push, prefilter at 48^3) replays on new contents of
    the same buffers and returns what the eager call returns."""
    import interpol_b200 as ib
    gen = torch.Generator().manual_seed(123)
    shape = (48, 40, 32)
    vol = torch.randn([1, 2, *shape], generator=gen).cuda()
    grid = smooth_grid(shape, gen, amp=2.0).contiguous().cuda()
    kw = dict(interpolation=3, bound='dct2', extrapolate=True)

    def step():
        p = ib.grid_pull(vol, grid, **kw)
        g = ib.grid_grad(vol, grid, **kw)
        s = ib.grid_push(p, grid, **kw)
        c = ib.spline_coeff_nd(s, interpolation=3, bound='dct2', dim=3)
        k = ib.grid_count(grid, **kw)
        return p, g, s, c, k

    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        outs = step()
    vol.copy_(torch.randn(vol.shape, generator=gen))
    grid.add_(0.37)
    graph.replay()
    torch.cuda.synchronize()
    want = step()
    for a, b in zip(outs, want):
        scale = b.abs().max().item()
        assert (a - b).abs().max().item() <= 2e-6 * scale        # (scatters: float REDs in another order)
