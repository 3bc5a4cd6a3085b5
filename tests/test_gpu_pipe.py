"""GPU tests of the persistent warp-specialised ("pipe") kernels: against the
float64 oracle over every bound / extrapolate mode, and against the generic
one-thread-per-point kernels (flags=NO_TILES) on the shapes that exercise each
mode of the pipeline (plain boxes, folded boxes, empty tiles, boxes that do not
fit and fall back to global gathers, partial tiles, several channels/batches)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from test_gpu_ops import smooth_grid, to_np

pytestmark = pytest.mark.gpu

NO_TILES, NO_PIPE, FORCE_PIPE = 1, 4, 8


def _with_flags(pp, flags, fn):
    old = pp.flags
    pp.flags = flags
    try:
        return fn()
    finally:
        pp.flags = old


@pytest.mark.parametrize('extrapolate', [1, 0, 2])
@pytest.mark.parametrize('bound', range(7))
@pytest.mark.parametrize('order', [1, 2, 3])
def test_pipe_pull_grad_vs_oracle(order, bound, extrapolate):
    import oracle
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    gen = torch.Generator().manual_seed(1000 + 100 * order + 10 * bound + extrapolate)
    vshape, shape = (30, 26, 44), (36, 28, 40)        # partial tiles in x and y and z
    B, C = 2, 2
    vol = torch.randn([B, C, *vshape], generator=gen)
    grid = smooth_grid(shape, gen, amp=4.0, batch=B)
    grid = (grid * torch.tensor([vshape[d] / shape[d] for d in range(3)]) - 1.5).contiguous()    # leaves the field of view
    b, o = [bound, (bound + 1) % 7, (bound + 3) % 7], [order]
    got = _with_flags(pp, FORCE_PIPE, lambda: pp.grid_pull(vol.cuda(), grid.cuda(), b, o, extrapolate))
    assert ib.last_kernel().startswith('pull_pipe3d'), ib.last_kernel()
    want = oracle.grid_pull(vol.double().numpy(), grid.double().numpy(), b, o, extrapolate)
    assert rel_err(to_np(got), want) <= 1e-5
    got = _with_flags(pp, FORCE_PIPE, lambda: pp.grid_grad(vol.cuda(), grid.cuda(), b, o, extrapolate))
    assert ib.last_kernel().startswith('grad_pipe3d'), ib.last_kernel()
    want = oracle.grid_grad(vol.double().numpy(), grid.double().numpy(), b, o, extrapolate)
    assert rel_err(to_np(got), want) <= 1e-5


@pytest.mark.parametrize('case', ['smooth', 'steep', 'incoherent', 'far_outside', 'zoom_in', 'zoom_out', 'nan_inf'])
@pytest.mark.parametrize('order', [1, 3])
def test_pipe_matches_generic(order, case):
    """same arithmetic as the one-thread-per-point kernel on every pipeline mode"""
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    gen = torch.Generator().manual_seed(77 + order)
    shape = (72, 40, 96)
    vol = torch.randn([1, 3, *shape], generator=gen).cuda()
    grid = smooth_grid(shape, gen, amp=3.0)
    if case == 'steep':
        grid = smooth_grid(shape, gen, amp=30.0)
    elif case == 'incoherent':
        grid = grid + torch.randn(grid.shape, generator=gen) * 20
    elif case == 'far_outside':
        grid = grid + torch.tensor([0., 500., -300.])
    elif case == 'zoom_in':
        grid = grid * 0.25 + 10
    elif case == 'zoom_out':
        grid = grid * 3.0 - 50
    elif case == 'nan_inf':
        grid = grid.clone()
        grid[0, 5, 7, 9, 1] = float('nan')
        grid[0, 40, 20, 33, 0] = float('inf')
        grid[0, 41, 21, 34, 2] = -float('inf')
        grid[0, 60, 1, 2, 0] = 3e30
    grid = grid.contiguous().cuda()
    for bound, ex in (([3], 1), ([6, 0, 4], 0), ([5, 2, 1], 2)):
        for fn in (pp.grid_pull, pp.grid_grad):
            a = _with_flags(pp, FORCE_PIPE, lambda: fn(vol, grid, bound, [order], ex))
            assert 'pipe3d' in ib.last_kernel(), ib.last_kernel()
            b = _with_flags(pp, NO_TILES, lambda: fn(vol, grid, bound, [order], ex))
            assert 'box' not in ib.last_kernel() and 'tile' not in ib.last_kernel()
            scale = b.abs().max().item()
            if scale == 0:
                assert a.abs().max().item() == 0
            else:
                assert ((a - b).abs().max().item() / scale) <= 2e-6, (case, bound, ex, fn.__name__)


def test_pipe_full_size_vs_tile_kernel():
    """256^3 cubic (the bench workload): pipe == one-tile-per-CTA kernel == generic"""
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import make_workload
    vol, grid = make_workload(256, 'cuda')
    a = pp.grid_pull(vol, grid, [3], [3], 1)
    assert ib.last_kernel().startswith('pull_pipe3d')
    b = _with_flags(pp, NO_PIPE, lambda: pp.grid_pull(vol, grid, [3], [3], 1))
    assert ib.last_kernel().startswith('pull_tile3d')
    c = _with_flags(pp, NO_TILES, lambda: pp.grid_pull(vol, grid, [3], [3], 1))
    assert rel_err(to_np(a), to_np(c)) <= 2e-6
    assert rel_err(to_np(b), to_np(c)) <= 2e-6


def test_pipe_many_small_batches():
    """more batch elements than tiles per element, broadcast (stride-0) volume"""
    import oracle
    from interpol_b200 import pushpull as pp
    gen = torch.Generator().manual_seed(5)
    shape = (32, 32, 32)
    vol = torch.randn([1, 1, *shape], generator=gen)
    grid = smooth_grid(shape, gen, amp=2.0, batch=5).contiguous()
    got = _with_flags(pp, FORCE_PIPE, lambda: pp.grid_pull(vol.cuda().expand(5, 1, *shape), grid.cuda(), [1], [3], 1))
    want = oracle.grid_pull(vol.double().numpy().repeat(5, 0), grid.double().numpy(), [1], [3], 1)
    assert rel_err(to_np(got), want) <= 1e-5


# ------------------------------------------------------------------ push / count --

@pytest.mark.parametrize('extrapolate', [1, 0, 2])
@pytest.mark.parametrize('bound', range(7))
@pytest.mark.parametrize('order', [1, 2, 3])
def test_pipe_push_count_vs_oracle(order, bound, extrapolate):
    import oracle
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    gen = torch.Generator().manual_seed(2000 + 100 * order + 10 * bound + extrapolate)
    vshape, shape = (30, 26, 44), (36, 28, 40)        # partial tiles in x and y and z
    B, C = 2, 2
    img = torch.randn([B, C, *shape], generator=gen)
    grid = smooth_grid(shape, gen, amp=4.0, batch=B)
    grid = (grid * torch.tensor([vshape[d] / shape[d] for d in range(3)]) - 1.5).contiguous()    # leaves the field of view
    b, o = [bound, (bound + 1) % 7, (bound + 3) % 7], [order]
    got = pp.grid_push(img.cuda(), grid.cuda(), list(vshape), b, o, extrapolate)
    assert ib.last_kernel().startswith('push_box3d'), ib.last_kernel()
    want = oracle.grid_push(img.double().numpy(), grid.double().numpy(), vshape, b, o, extrapolate)
    assert rel_err(to_np(got), want) <= 1e-5
    got = pp.grid_count(grid.cuda(), list(vshape), b, o, extrapolate)
    assert ib.last_kernel().startswith('count_box3d'), ib.last_kernel()
    want = oracle.grid_count(grid.double().numpy(), vshape, b, o, extrapolate)
    assert rel_err(to_np(got), want) <= 1e-5


@pytest.mark.parametrize('case', ['smooth', 'steep', 'incoherent', 'far_outside', 'zoom_in', 'zoom_out', 'nan_inf', 'spiky'])
@pytest.mark.parametrize('order', [1, 3])
def test_pipe_push_matches_generic(order, case):
    """same result as the one-thread-per-point scatter on every pipeline mode (float atomics there:
    agreement to a few float32 ulps of the largest accumulated value)"""
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    gen = torch.Generator().manual_seed(177 + order)
    shape = (72, 40, 96)
    img = torch.randn([1, 2, *shape], generator=gen)
    grid = smooth_grid(shape, gen, amp=3.0)
    if case == 'steep':
        grid = smooth_grid(shape, gen, amp=30.0)
    elif case == 'incoherent':
        grid = grid + torch.randn(grid.shape, generator=gen) * 20
    elif case == 'far_outside':
        grid = grid + torch.tensor([0., 500., -300.])
    elif case == 'zoom_in':
        grid = grid * 0.25 + 10       # 64 sources per target voxel: the overflow bound matters
    elif case == 'zoom_out':
        grid = grid * 3.0 - 50
    elif case == 'nan_inf':
        grid = grid.clone()
        grid[0, 5, 7, 9, 1] = float('nan')
        grid[0, 40, 20, 33, 0] = float('inf')
        grid[0, 41, 21, 34, 2] = -float('inf')
        grid[0, 60, 1, 2, 0] = 3e30
    elif case == 'spiky':
        img = img * (torch.rand(img.shape, generator=gen) < 0.01) * 1e4 + img * 1e-3
    img, grid = img.cuda(), grid.contiguous().cuda()
    for bound, ex in (([3], 1), ([6, 0, 4], 0), ([5, 2, 1], 2)):
        for count in (False, True):
            fn = (lambda: pp.grid_count(grid, list(shape), bound, [order], ex)) if count else \
                 (lambda: pp.grid_push(img, grid, list(shape), bound, [order], ex))
            a = fn()
            assert 'box3d' in ib.last_kernel(), ib.last_kernel()
            b = _with_flags(pp, NO_TILES, fn)
            assert 'box' not in ib.last_kernel() and 'tile' not in ib.last_kernel()
            scale = b.abs().max().item()
            if scale == 0:
                assert a.abs().max().item() == 0
            else:
                assert ((a - b).abs().max().item() / scale) <= 8e-6, (case, bound, ex, count)


def test_pipe_push_full_size_adjoint():
    """256^3 cubic (the bench workload): <pull(x), y> == <x, push(y)> with the persistent pull and the boxed push"""
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import make_workload
    vol, grid = make_workload(256, 'cuda')
    y = torch.randn(vol.shape, generator=torch.Generator().manual_seed(99)).cuda()
    for bound in ([3], [6], [0], [1, 2, 4]):
        px = pp.grid_pull(vol, grid, bound, [3], 1)
        assert ib.last_kernel().startswith('pull_pipe3d')
        py = pp.grid_push(y, grid, [256] * 3, bound, [3], 1)
        assert ib.last_kernel().startswith('push_box3d')
        lhs = (px.double() * y.double()).sum().item()
        rhs = (vol.double() * py.double()).sum().item()
        scale = (px.double().abs() * y.double().abs()).sum().item()
        assert abs(lhs - rhs) <= 5e-6 * scale, (bound, lhs, rhs, scale)
    cnt = pp.grid_count(grid, [256] * 3, [6], [3], 1)
    assert ib.last_kernel().startswith('count_box3d')
    assert abs(cnt.double().sum().item() - 256 ** 3) <= 1e-6 * 256 ** 3
    ref = _with_flags(pp, NO_PIPE, lambda: pp.grid_count(grid, [256] * 3, [6], [3], 1))
    assert ib.last_kernel().startswith('count_tile3d')
    assert rel_err(to_np(cnt), to_np(ref)) <= 4e-6


def _pipe_claimed():
    import ctypes
    from interpol_b200 import _lib
    buf = (ctypes.c_int * 1)()
    rc = _lib.lib().ib200_debug_pipe_control(buf)
    return rc, buf[0]


@pytest.mark.parametrize('op', ['pull', 'grad', 'bwd_grid'])
@pytest.mark.parametrize('static', [False, True])
def test_pipe_dynamic_tile_claims(op, static, monkeypatch):
    """Tiles beyond the first of each CTA are claimed from a global counter (IB200_STATIC_TILES=1: the static
    round-robin that streams under capture use): every tile must be processed exactly once either way."""
    import oracle
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    if static:
        monkeypatch.setenv('IB200_STATIC_TILES', '1')
    gen = torch.Generator().manual_seed(5 + static)
    shape, vshape = (48, 40, 200), (48, 40, 232)
    vol = torch.randn([2, 1, *vshape], generator=gen)
    grid = smooth_grid(shape, gen, amp=2.0, batch=2)
    grid[..., 2] = grid[..., 2] * 1.12 + 0.3
    grid = grid.contiguous()
    b, o = [3, 1, 3], [3]
    if op == 'pull':
        got = _with_flags(pp, FORCE_PIPE, lambda: pp.grid_pull(vol.cuda(), grid.cuda(), b, o, 1))
        want = oracle.grid_pull(vol.double().numpy(), grid.double().numpy(), b, o, 1)
    elif op == 'grad':
        got = _with_flags(pp, FORCE_PIPE, lambda: pp.grid_grad(vol.cuda(), grid.cuda(), b, o, 1))
        want = oracle.grid_grad(vol.double().numpy(), grid.double().numpy(), b, o, 1)
    else:
        gout = torch.randn([2, 1, *shape], generator=gen)
        got = _with_flags(pp, FORCE_PIPE, lambda: pp.grid_pull_grad_grid(gout.cuda(), vol.cuda(), grid.cuda(), b, o, 1))
        want = oracle.grid_grad(vol.double().numpy(), grid.double().numpy(), b, o, 1)[:, 0] * gout.double().numpy()[:, 0, ..., None]
    assert '_pipe3d' in ib.last_kernel(), ib.last_kernel()
    rc, claimed = _pipe_claimed()
    ntiles = 2 * (shape[0] // 8) * (shape[1] // 8) * ((shape[2] + 31) // 32)
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    if static:
        assert rc == -1
    else:
        assert rc == 0 and ntiles - sms <= claimed <= ntiles + sms, (claimed, ntiles)
    assert rel_err(to_np(got), want) <= 1e-5


def test_pipe_under_stream_capture_matches_eager():
    """a captured launch takes the static schedule (no counter word baked into the graph)"""
    from interpol_b200 import pushpull as pp
    gen = torch.Generator().manual_seed(9)
    shape = (48, 40, 96)
    vol = torch.randn([1, 1, *shape], generator=gen).cuda()
    grid = smooth_grid(shape, gen, amp=2.0).cuda()
    eager = _with_flags(pp, FORCE_PIPE, lambda: pp.grid_pull(vol, grid, [3], [3], 1))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        _with_flags(pp, FORCE_PIPE, lambda: pp.grid_pull(vol, grid, [3], [3], 1))      # warm-up outside capture
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = _with_flags(pp, FORCE_PIPE, lambda: pp.grid_pull(vol, grid, [3], [3], 1))
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager)
