"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(balbasty/torch-interpol mounted read-only at /root/reference).

Run from the repo root, in the build container only:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; tests there only read the .npz files.
"""
import os
import sys
import warnings
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = os.environ.get('INTERPOL_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
warnings.filterwarnings('ignore')

import cases  # noqa: E402
import interpol  # noqa: E402  (the reference)
from interpol import pushpull as ref_pp  # noqa: E402
from interpol import coeff as ref_coeff  # noqa: E402
from interpol.bounds import Bound  # noqa: E402
from interpol.splines import Spline  # noqa: E402

TDT = {'f64': torch.float64, 'f32': torch.float32}


def run_pushpull(case):
    dim = case['dim']
    dt = cases.NP_DTYPE[case['dtype']]
    vol, grid, src, srcg = cases.make_inputs(case['name'], dim, case['B'], case['C'], dt)
    vol, grid, src, srcg = map(torch.from_numpy, (vol, grid, src, srcg))
    b, o, e = case['bound'], case['order'], case['extrapolate']
    ishape = list(vol.shape[2:])
    op = case['op']
    if op == 'pull':
        out = ref_pp.grid_pull(vol, grid, b, o, e)
    elif op == 'grad':
        out = ref_pp.grid_grad(vol, grid, b, o, e)
    elif op == 'hess':
        out = ref_pp.grid_hess(vol, grid, b, o, e)
    elif op == 'push':
        out = ref_pp.grid_push(src, grid, ishape, b, o, e)
    elif op == 'count':
        out = ref_pp.grid_count(grid, ishape, b, o, e)
    elif op == 'pushgrad':
        out = ref_pp.grid_pushgrad(srcg, grid, ishape, b, o, e)
    else:
        raise ValueError(op)
    return out.numpy()


def main():
    torch.manual_seed(0)
    out = {}
    # ---- truth tables: bounds.py:30-89 ---------------------------------
    ii = torch.arange(-23, 24)
    for n in (1, 2, 4, 5):
        for b in range(7):
            bound = Bound(b)
            out['bound_index_b%d_n%d' % (b, n)] = bound.index(ii, n).numpy().astype(np.int64)
            sgn = bound.transform(ii, n)
            sgn = torch.ones_like(ii) if sgn is None else sgn.expand(ii.shape)
            out['bound_sign_b%d_n%d' % (b, n)] = sgn.numpy().astype(np.int64)
    out['bound_i'] = ii.numpy().astype(np.int64)
    # ---- spline polynomials: splines.py:30-195 -------------------------
    x = torch.linspace(-4.25, 4.25, 1361, dtype=torch.float64)   # step 1/160: hits every knot
    out['spline_x'] = x.numpy()
    for o in range(8):
        s = Spline(o)
        out['spline_w_o%d' % o] = s.fastweight(x).numpy()
        out['spline_g_o%d' % o] = s.fastgrad(x).numpy()
        out['spline_h_o%d' % o] = s.fasthess(x).numpy()
    np.savez_compressed(os.path.join(HERE, 'tables.npz'), **out)
    print('tables.npz', len(out))

    # ---- push / pull family --------------------------------------------
    out = {}
    for case in cases.pushpull_cases():
        out[case['name']] = run_pushpull(case)
    np.savez_compressed(os.path.join(HERE, 'pushpull.npz'), **out)
    print('pushpull.npz', len(out))

    # ---- prefilter -----------------------------------------------------
    out = {}
    for case in cases.coeff_cases():
        x = torch.from_numpy(cases.coeff_input(case['name'], case['n'], cases.NP_DTYPE[case['dtype']]))
        y = ref_coeff.spline_coeff(x, case['bound'], case['order'], dim=1)
        out[case['name']] = y.numpy()
    for case in cases.coeff_nd_cases():
        x = torch.from_numpy(cases.coeff_nd_input(case['name'], case['shape']))
        y = ref_coeff.spline_coeff_nd(x, case['bound'], case['order'], len(case['shape']))
        out[case['name']] = y.numpy()
    np.savez_compressed(os.path.join(HERE, 'coeff.npz'), **out)
    print('coeff.npz', len(out))

    # ---- public API (api.py) on canonical examples ---------------------
    out = {}
    rng = np.random.default_rng(7)
    vol = torch.from_numpy(rng.standard_normal((2, 3, 6, 7, 8)).astype(np.float32).astype(np.float64))
    disp = torch.from_numpy(np.round(rng.standard_normal((2, 5, 6, 7, 3)) * 2 * 256) / 256)
    grid = interpol.add_identity_grid(disp)
    out['api_vol'] = vol.numpy()
    out['api_grid'] = grid.numpy()
    for prefilter in (False, True):
        tag = 'pf%d' % prefilter
        out['api_pull_' + tag] = interpol.grid_pull(vol, grid, interpolation=3, bound='dct2',
                                                    extrapolate=True, prefilter=prefilter).numpy()
        out['api_grad_' + tag] = interpol.grid_grad(vol, grid, interpolation=3, bound='dct2',
                                                    extrapolate=True, prefilter=prefilter).numpy()
    src = torch.from_numpy(rng.standard_normal((2, 3, 5, 6, 7)).astype(np.float32).astype(np.float64))
    out['api_src'] = src.numpy()
    out['api_push'] = interpol.grid_push(src, grid, shape=(6, 7, 8), interpolation=2, bound='dft',
                                         extrapolate=False).numpy()
    out['api_push_pf'] = interpol.grid_push(src, grid, shape=(6, 7, 8), interpolation=3, bound='dct1',
                                            extrapolate=True, prefilter=True).numpy()
    out['api_count'] = interpol.grid_count(grid, shape=(6, 7, 8), interpolation=1, bound='dct1',
                                           extrapolate=2).numpy()
    # label map pull (api.py:194-205)
    lab = torch.from_numpy(rng.integers(0, 4, size=(1, 1, 6, 7, 8)).astype(np.int64))
    out['api_label'] = lab.numpy()
    out['api_pull_label'] = interpol.grid_pull(lab, grid[:1], interpolation=1, bound='replicate',
                                               extrapolate=True).numpy()
    out['api_pull_label_nn'] = interpol.grid_pull(lab, grid[:1], interpolation=0, bound='dct2',
                                                  extrapolate=True).numpy()
    # spline_coeff / spline_coeff_nd public entry points
    out['api_coeff'] = interpol.spline_coeff(vol, interpolation=3, bound='dct2', dim=-2).numpy()
    out['api_coeff_nd'] = interpol.spline_coeff_nd(vol, interpolation=[3, 5, 2], bound=['dct2', 'dft', 'dct1'],
                                                   dim=3).numpy()
    # resize identity (tests/test_coeff.py) and a genuine resize
    out['api_resize'] = interpol.resize(vol, factor=[1.5, 0.75, 2.0], interpolation=3, anchor='e').numpy()
    out['api_restrict'] = interpol.restrict(vol, factor=[2, 2, 2], interpolation=1, anchor='e').numpy()
    np.savez_compressed(os.path.join(HERE, 'api.npz'), **out)
    print('api.npz', len(out))


if __name__ == '__main__':
    main()
