"""GPU parity proper: the CUDA path, called through the C ABI
(interpol_b200.pushpull / interpol_b200.coeff), against outputs of the
unmodified reference stored in tests/golden/*.npz.

Tolerances (max|a-ref|/max|ref|):
  float64 kernels  : 1e-10   (same algorithm, different summation order)
  float32 kernels  : 1e-5    vs the reference run in float64 on the same
                             (float32-representable) inputs
"""
import os
import numpy as np
import pytest
import torch

import cases
from conftest import rel_err

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
PUSHPULL = np.load(os.path.join(G, 'pushpull.npz'))
COEFF = np.load(os.path.join(G, 'coeff.npz'))
API = np.load(os.path.join(G, 'api.npz'))

TOL = {torch.float64: 1e-10, torch.float32: 1e-5}


def run_cuda(case, dtype):
    from interpol_b200 import pushpull as pp
    dim = case['dim']
    vol, grid, src, srcg = cases.make_inputs(case['name'], dim, case['B'], case['C'], np.float64)
    dev = torch.device('cuda')
    vol, grid, src, srcg = [torch.from_numpy(a).to(dev, dtype) for a in (vol, grid, src, srcg)]
    b, o, e = case['bound'], case['order'], case['extrapolate']
    ishape = list(vol.shape[2:])
    op = case['op']
    if op == 'pull':
        return pp.grid_pull(vol, grid, b, o, e)
    if op == 'grad':
        return pp.grid_grad(vol, grid, b, o, e)
    if op == 'hess':
        return pp.grid_hess(vol, grid, b, o, e)
    if op == 'push':
        return pp.grid_push(src, grid, ishape, b, o, e)
    if op == 'count':
        return pp.grid_count(grid, ishape, b, o, e)
    if op == 'pushgrad':
        return pp.grid_pushgrad(srcg, grid, ishape, b, o, e)
    raise ValueError(op)


PP_CASES = cases.pushpull_cases()


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32], ids=['f64', 'f32'])
@pytest.mark.parametrize('case', PP_CASES, ids=[c['name'] for c in PP_CASES])
def test_pushpull_golden(case, dtype):
    ref = PUSHPULL[case['name']]
    out = run_cuda(case, dtype).double().cpu().numpy()
    assert out.shape == ref.shape
    tol = TOL[dtype]
    if case['dtype'] == 'f32':      # golden itself is a float32 run of the reference
        tol = max(tol, 2e-6)
    assert rel_err(out, ref) <= tol, (case['name'], rel_err(out, ref))


C_CASES = [c for c in cases.coeff_cases() if c['dtype'] == 'f64']


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32], ids=['f64', 'f32'])
@pytest.mark.parametrize('case', C_CASES, ids=[c['name'] for c in C_CASES])
def test_coeff_golden(case, dtype):
    from interpol_b200 import coeff
    x = torch.from_numpy(cases.coeff_input(case['name'], case['n'], np.float64)).to('cuda', dtype)
    ref = COEFF[case['name']]
    out = coeff.spline_coeff(x, case['bound'], case['order'], dim=1).double().cpu().numpy()
    tol = 1e-10 if dtype == torch.float64 else 1e-5
    assert rel_err(out, ref) <= tol, rel_err(out, ref)
    # in place on a dense tensor returns the same storage
    y = x.clone()
    z = coeff.spline_coeff(y, case['bound'], case['order'], dim=1, inplace=True)
    assert z.data_ptr() == y.data_ptr()
    assert rel_err(z.double().cpu().numpy(), ref) <= tol


@pytest.mark.parametrize('case', cases.coeff_nd_cases(), ids=[c['name'] for c in cases.coeff_nd_cases()])
def test_coeff_nd_golden(case):
    from interpol_b200 import coeff
    x = torch.from_numpy(cases.coeff_nd_input(case['name'], case['shape'])).cuda()
    ref = COEFF[case['name']]
    out = coeff.spline_coeff_nd(x, case['bound'], case['order'], len(case['shape']))
    assert rel_err(out.cpu().numpy(), ref) <= 1e-10
    out32 = coeff.spline_coeff_nd(x.float(), case['bound'], case['order'], len(case['shape']))
    assert rel_err(out32.cpu().numpy(), ref) <= 1e-5


def test_coeff_unsupported_bound():
    import interpol_b200 as ib
    x = torch.zeros(4, 5, device='cuda')
    for b in ('dst1', 'dst2'):
        with pytest.raises(NotImplementedError):
            ib.spline_coeff(x, interpolation=3, bound=b)
    # orders 0/1 are no-ops whatever the bound (coeff.py:306-307)
    assert torch.equal(ib.spline_coeff(x + 1, interpolation=1, bound='dst1'), x + 1)


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32], ids=['f64', 'f32'])
def test_public_api_golden(dtype):
    """interpol.grid_pull / grid_grad / grid_push / grid_count / spline_coeff[_nd]
    / resize / restrict through the public functions (api.py)."""
    import interpol_b200 as ib
    tol = 1e-10 if dtype == torch.float64 else 2e-5
    dev = 'cuda'
    vol = torch.from_numpy(API['api_vol']).to(dev, dtype)
    grid = torch.from_numpy(API['api_grid']).to(dev, dtype)
    src = torch.from_numpy(API['api_src']).to(dev, dtype)

    def chk(out, key):
        ref = API[key]
        assert tuple(out.shape) == ref.shape, key
        assert rel_err(out.double().cpu().numpy(), ref) <= tol, (key, rel_err(out.double().cpu().numpy(), ref))

    for pf in (False, True):
        chk(ib.grid_pull(vol, grid, interpolation=3, bound='dct2', extrapolate=True, prefilter=pf), 'api_pull_pf%d' % pf)
        chk(ib.grid_grad(vol, grid, interpolation=3, bound='dct2', extrapolate=True, prefilter=pf), 'api_grad_pf%d' % pf)
    chk(ib.grid_push(src, grid, shape=(6, 7, 8), interpolation=2, bound='dft', extrapolate=False), 'api_push')
    chk(ib.grid_push(src, grid, shape=(6, 7, 8), interpolation=3, bound='dct1', extrapolate=True, prefilter=True), 'api_push_pf')
    chk(ib.grid_count(grid, shape=(6, 7, 8), interpolation=1, bound='dct1', extrapolate=2), 'api_count')
    chk(ib.spline_coeff(vol, interpolation=3, bound='dct2', dim=-2), 'api_coeff')
    chk(ib.spline_coeff_nd(vol, interpolation=[3, 5, 2], bound=['dct2', 'dft', 'dct1'], dim=3), 'api_coeff_nd')
    chk(ib.resize(vol, factor=[1.5, 0.75, 2.0], interpolation=3, anchor='e'), 'api_resize')
    chk(ib.restrict(vol, factor=[2, 2, 2], interpolation=1, anchor='e'), 'api_restrict')


def test_label_maps_bit_exact():
    """Integer inputs: label-wise soft pull + arg-max (api.py:194-205); nearest
    neighbour labels must be bit exact."""
    import interpol_b200 as ib
    lab = torch.from_numpy(API['api_label']).cuda()
    grid = torch.from_numpy(API['api_grid']).cuda()[:1]
    out = ib.grid_pull(lab, grid, interpolation=1, bound='replicate', extrapolate=True)
    assert out.dtype == lab.dtype
    assert torch.equal(out.cpu(), torch.from_numpy(API['api_pull_label']))
    out = ib.grid_pull(lab, grid, interpolation=0, bound='dct2', extrapolate=True)
    assert torch.equal(out.cpu(), torch.from_numpy(API['api_pull_label_nn']))
    out = ib.grid_pull(lab, grid.float(), interpolation=0, bound='dct2', extrapolate=True)
    assert torch.equal(out.cpu(), torch.from_numpy(API['api_pull_label_nn']))
