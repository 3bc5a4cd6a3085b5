"""Pin the CPU oracle (oracle/oracle.c) against outputs of the unmodified
reference stored in tests/golden/*.npz (made by tests/golden/make_golden.py).
CPU only."""
import os
import numpy as np
import pytest

import cases
from conftest import rel_err
import oracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
TABLES = np.load(os.path.join(G, 'tables.npz'))
PUSHPULL = np.load(os.path.join(G, 'pushpull.npz'))
COEFF = np.load(os.path.join(G, 'coeff.npz'))


@pytest.mark.parametrize('n', [1, 2, 4, 5])
@pytest.mark.parametrize('bound', range(7))
def test_bound_tables(bound, n):
    """bounds.py:30-89 (incl. the dst1 zero at i == 0 mod 2(n+1))."""
    ii = TABLES['bound_i']
    idx = [oracle.bound_index(bound, int(i), n) for i in ii]
    sgn = [oracle.bound_sign(bound, int(i), n) for i in ii]
    assert idx == TABLES['bound_index_b%d_n%d' % (bound, n)].tolist()
    assert sgn == TABLES['bound_sign_b%d_n%d' % (bound, n)].tolist()


def test_bound_table_survey():
    """The n=4 rows printed in SURVEY.md 8.3."""
    want = {
        2: [1, 2, 3, 2, 1, 0, 1, 2, 3, 2, 1, 0, 1, 2, 3, 2, 1, 0, 1, 2, 3, 2, 1],
        3: [2, 1, 0, 0, 1, 2, 3, 3, 2, 1, 0, 0, 1, 2, 3, 3, 2, 1, 0, 0, 1, 2, 3],
        4: [0, 0, 1, 2, 3, 3, 3, 2, 1, 0, 0, 0, 1, 2, 3, 3, 3, 2, 1, 0, 0, 0, 1],
        6: [1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3],
    }
    for b, row in want.items():
        assert [oracle.bound_index(b, i, 4) for i in range(-11, 12)] == row
    dst1 = [0, 1, 1, 1, 0, 0, -1, -1, -1, -1, 0, 0, 1, 1, 1, 0, -1, -1, -1, -1, 0, 0, 1]
    assert [oracle.bound_sign(4, i, 4) for i in range(-11, 12)] == dst1


@pytest.mark.parametrize('order', range(8))
def test_spline_polynomials(order):
    """splines.py:30-195 on a sweep that hits every knot."""
    x = TABLES['spline_x']
    np.testing.assert_allclose(oracle.weight(order, x), TABLES['spline_w_o%d' % order], rtol=0, atol=1e-14)
    # the reference's order-1 derivative in the ND path has the wrong sign
    # (splines.py:96-97): reproduce it with the quirk flag, and check that the
    # default is its negation.
    np.testing.assert_allclose(oracle.grad_weight(order, x, quirk_linear_grad=True),
                               TABLES['spline_g_o%d' % order], rtol=0, atol=1e-14)
    if order == 1:
        np.testing.assert_allclose(oracle.grad_weight(order, x), -TABLES['spline_g_o1'], rtol=0, atol=0)
    np.testing.assert_allclose(oracle.hess_weight(order, x), TABLES['spline_h_o%d' % order], rtol=0, atol=1e-13)


def run_oracle(case):
    dim = case['dim']
    dt = cases.NP_DTYPE[case['dtype']]
    vol, grid, src, srcg = cases.make_inputs(case['name'], dim, case['B'], case['C'], dt)
    b, o, e = case['bound'], case['order'], case['extrapolate']
    ishape = vol.shape[2:]
    op = case['op']
    if op == 'pull':
        return oracle.grid_pull(vol, grid, b, o, e)
    if op == 'grad':
        return oracle.grid_grad(vol, grid, b, o, e)
    if op == 'hess':
        return oracle.grid_hess(vol, grid, b, o, e)
    if op == 'push':
        return oracle.grid_push(src, grid, ishape, b, o, e)
    if op == 'count':
        return oracle.grid_count(grid, ishape, b, o, e)
    if op == 'pushgrad':
        return oracle.grid_pushgrad(srcg, grid, ishape, b, o, e)
    raise ValueError(op)


PP_CASES = cases.pushpull_cases()


@pytest.mark.parametrize('case', PP_CASES, ids=[c['name'] for c in PP_CASES])
def test_pushpull_vs_reference(case):
    """nd.py / iso0.py / iso1.py through pushpull.py:35-233."""
    ref = PUSHPULL[case['name']]
    out = run_oracle(case)
    assert out.shape == ref.shape
    tol = 1e-12 if case['dtype'] == 'f64' else 2e-6
    assert rel_err(out, ref) <= tol


C_CASES = cases.coeff_cases()


@pytest.mark.parametrize('case', C_CASES, ids=[c['name'] for c in C_CASES])
def test_coeff_vs_reference(case):
    """coeff.py:258-313, all five accepted bounds, orders 2-7, n from 1 to 40."""
    x = cases.coeff_input(case['name'], case['n'], cases.NP_DTYPE[case['dtype']])
    ref = COEFF[case['name']]
    out = oracle.spline_coeff(x, case['bound'], case['order'], dim=1)
    tol = 1e-12 if case['dtype'] == 'f64' else 5e-6
    assert rel_err(out, ref) <= tol


@pytest.mark.parametrize('case', cases.coeff_nd_cases(), ids=[c['name'] for c in cases.coeff_nd_cases()])
def test_coeff_nd_vs_reference(case):
    """coeff.py:317-347"""
    x = cases.coeff_nd_input(case['name'], case['shape'])
    ref = COEFF[case['name']]
    out = oracle.spline_coeff_nd(x, case['bound'], case['order'], len(case['shape']))
    assert rel_err(out, ref) <= 1e-12


def test_coeff_unsupported_bound():
    """coeff.py:237-254: dst1/dst2 raise NotImplementedError."""
    x = np.zeros((1, 5, 1))
    for b in (4, 5):
        with pytest.raises(NotImplementedError):
            oracle.spline_coeff(x, b, 3, dim=1)


def test_oracle_reproduces_reference_resize():
    """`resize` = prefilter + pull on the tensor-product grid of per-axis coordinates: the oracle, fed
    that grid, reproduces the reference's own resize outputs (tests/golden/resize.npz)."""
    import os
    import sys
    import torch
    import oracle
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    sys.path.insert(0, here)
    import make_golden_resize as mg
    gold = np.load(os.path.join(here, 'resize.npz'))
    codes = {'zero': 0, 'nearest': 1, 'replicate': 1, 'dct1': 2, 'dct2': 3, 'dst1': 4, 'dst2': 5, 'dft': 6}

    def aslist(v, n):
        v = list(v) if isinstance(v, (list, tuple)) else [v]
        return v + v[-1:] * (n - len(v))

    for i, c in enumerate(mg.CASES):
        if i % 5:
            continue
        want = gold['case%d' % i]
        dim = len(c['shape'])
        x = mg.make_input(i, c['shape']).numpy()
        outshape = want.shape[2:]
        lin = []
        for a, f, n_in, n_out in zip(aslist(c['anchor'], dim), aslist(c['factor'], dim), c['shape'], outshape):
            a = a[0]
            if a == 'c':
                lin.append(torch.linspace(0, n_in - 1, n_out, dtype=torch.float64).numpy())
            elif a == 'e':
                scale = n_in / n_out
                lin.append(np.arange(n_out, dtype=np.float64) * scale + 0.5 * (scale - 1))
            elif a == 'f':
                lin.append(np.arange(n_out, dtype=np.float64) / f)
            else:
                lin.append(np.arange(n_out, dtype=np.float64) / f + ((n_in - 1) - (n_out - 1) / f))
        grid = np.stack(np.meshgrid(*lin, indexing='ij'), axis=-1)[None]
        order = aslist(c['order'], dim)
        bound = [codes[b] for b in aslist(c['bound'], dim)]
        coef = oracle.spline_coeff_nd(x, bound, order, dim) if c['prefilter'] else x
        got = oracle.grid_pull(coef, grid, bound, order, int(c['extrapolate']))
        scale = np.abs(want).max()
        assert np.abs(got - want).max() <= 1e-10 * max(scale, 1e-300), (i, c)
