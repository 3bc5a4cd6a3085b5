"""CPU-only checks: the C-ABI library loads and exports every symbol declared
in include/interpol_b200.h, argument validation returns the documented status
codes without touching a GPU, the Python host mirrors the reference's option
parsing / shape algebra, and the product never imports the oracle."""
import ctypes
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'interpol_b200.h')).read()
    return sorted(set(re.findall(r'IB200_API\s+[\w\s\*]+?\b(ib200_\w+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from interpol_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 13 and 'ib200_pull' in names and 'ib200_spline_coeff' in names
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), n
    assert _lib.lib().ib200_abi_version() == 1
    assert _lib.lib().ib200_error_string(0) == b'success'


def test_problem_struct_layout_matches_header():
    """sizeof(struct ib200_problem) as laid out by ctypes == the C declaration:
    4+3+3 int32, 2 uint32, 2 int64, 3+3+5+5+6 int64."""
    from interpol_b200._lib import Problem
    assert ctypes.sizeof(Problem) == 4 * (4 + 3 + 3 + 2) + 8 * (2 + 3 + 3 + 5 + 5 + 6)


def test_validation_status_codes_without_gpu():
    """every error below is detected before any CUDA call"""
    from interpol_b200 import _lib
    from interpol_b200._lib import Problem
    L = _lib.lib()
    p = Problem()
    p.dim, p.dtype, p.extrapolate = 3, _lib.F32, 1
    p.batch = p.channels = 1
    for d in range(3):
        p.vol_shape[d] = p.pts_shape[d] = 4
    null = ctypes.c_void_p(0)

    def status(**kw):
        q = Problem.from_buffer_copy(p)
        for k, v in kw.items():
            if isinstance(v, tuple):
                getattr(q, k)[v[0]] = v[1]
            else:
                setattr(q, k, v)
        return L.ib200_pull(ctypes.byref(q), null, null, null, null)

    assert status(dim=4) == -3 and status(dim=0) == -3
    assert status(dtype=9) == -2
    assert status(bound=(1, 7)) == -4
    assert status(order=(2, 8)) == -5
    assert status(extrapolate=3) == -9
    assert status(vol_shape=(0, 0)) == -6
    assert status() == -1                                   # null pointers
    assert L.ib200_pull(None, null, null, null, null) == -1
    # prefilter: unsupported bounds, bad codes, no-op orders
    f = L.ib200_spline_coeff
    assert f(null, _lib.F32, 1, 8, 1, 4, 3, 0, null) == -7
    assert f(null, _lib.F32, 1, 8, 1, 5, 3, 0, null) == -7
    assert f(null, _lib.F32, 1, 8, 1, 4, 1, 0, null) == 0   # orders 0/1 never look at the bound
    assert f(null, 7, 1, 8, 1, 3, 3, 0, null) == -2
    assert f(null, _lib.F32, 1, 8, 1, 9, 3, 0, null) == -4
    assert f(null, _lib.F32, 1, 8, 1, 3, 9, 0, null) == -5
    assert f(null, _lib.F32, 0, 8, 1, 3, 3, 0, null) == 0   # empty
    assert f(null, _lib.F32, 1, 8, 1, 3, 3, 0, null) == -1
    # error mapping to the reference's exception classes
    with pytest.raises(NotImplementedError):
        _lib.check(-7)
    for code in (-3, -4, -5, -6, -9, -10):
        with pytest.raises(ValueError):
            _lib.check(code)
    with pytest.raises(RuntimeError):
        _lib.check(-1000 - 2)
    # scratch size: only 16-bit storage needs it
    assert L.ib200_scratch_bytes(ctypes.byref(p)) == 0
    p.dtype = _lib.F16
    assert L.ib200_scratch_bytes(ctypes.byref(p)) == 64 * 4


def test_no_cpu_fallback():
    import interpol_b200 as ib
    from interpol_b200 import pushpull as pp
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    x = torch.zeros(1, 1, 4, 4)
    g = torch.zeros(1, 4, 4, 2)
    with pytest.raises(RuntimeError):
        pp.grid_pull(x, g, [0], [1], 1)
    with pytest.raises(RuntimeError):
        ib.grid_pull(x, g)
    with pytest.raises(RuntimeError):
        ib.spline_coeff_nd(x, interpolation=3)


def test_missing_library_fails_loudly(tmp_path):
    """IB200_LIB points the loader at another build (A/B measurements); a path without a library must raise,
    not fall back to anything (checked in a fresh interpreter: the loader caches its handle)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from interpol_b200 import _lib\n"
            "try:\n    _lib.lib()\nexcept _lib.ExtensionMissing as e:\n    print('missing:', e)\n" % os.path.join(root, 'torch-interpol_b200'))
    env = dict(os.environ, IB200_LIB=str(tmp_path / 'nowhere.so'))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, timeout=300)
    assert 'missing:' in out.stdout and 'nowhere.so' in out.stdout, out.stdout + out.stderr


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'torch-interpol_b200')
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(base, f)).read()
                assert 'import oracle' not in txt and 'from oracle' not in txt and 'liboracle' not in txt, f


def test_option_parsing_matches_reference_tables():
    """autograd.py:56-154 alias tables"""
    from interpol_b200 import bound_to_nitorch, inter_to_nitorch
    from interpol_b200.bounds import BoundType
    from interpol_b200.splines import InterpolationType
    table = {0: ['zero', 'zeros', 'constant'], 1: ['replicate', 'repeat', 'border', 'nearest'],
             2: ['dct1', 'mirror'], 3: ['dct2', 'reflect', 'reflection', 'neumann'],
             4: ['dst1', 'antimirror'], 5: ['dst2', 'antireflect', 'dirichlet'], 6: ['dft', 'wrap', 'circular']}
    for code, names in table.items():
        for n in names + [n.upper() for n in names] + [code, BoundType(code)]:
            assert bound_to_nitorch(n, as_type='int') == code
        assert bound_to_nitorch([names[0]], as_type='int') == [code]
        assert bound_to_nitorch((names[0], code), as_type='str') == (BoundType(code).name,) * 2
    for bad in ('foo', 7, -1, 1.5, None):
        with pytest.raises(ValueError):
            bound_to_nitorch(bad)
    names = ['nearest', 'linear', 'quadratic', 'cubic', 'fourth', 'fifth', 'sixth', 'seventh']
    for o, n in enumerate(names):
        assert inter_to_nitorch(n, 'int') == o and inter_to_nitorch(o, 'int') == o
        assert inter_to_nitorch(InterpolationType(o), 'int') == o
        assert inter_to_nitorch([o, n.upper()], 'int') == [o, o]
        assert inter_to_nitorch(o, 'str') == n
    for bad in ('bicubic', 8, -1, None):
        with pytest.raises(ValueError):
            inter_to_nitorch(bad)


def test_shape_algebra():
    """api._Layout (the shape conventions of api.py:93-146), utils.expanded_shape, pad_list_int"""
    from interpol_b200.api import _Layout
    from interpol_b200.utils import expanded_shape, make_list
    from interpol_b200.pushpull import pad_list_int
    assert pad_list_int([1], 3) == [1, 1, 1] and pad_list_int([1, 2, 3, 4], 2) == [1, 2]
    assert expanded_shape((2, 1, 3), (4, 3)) == (2, 4, 3)
    assert expanded_shape((2, 3), (1, 1, 5), side='right') == (2, 3, 5)
    assert expanded_shape((0, 3), (1, 3)) == (0, 3) and expanded_shape() == ()
    with pytest.raises(ValueError):
        expanded_shape((2, 3), (4, 3))
    with pytest.raises(ValueError):
        expanded_shape((0, 3), (4, 3))
    assert make_list(1, 3) == [1, 1, 1] and make_list([1, 2], 3) == [1, 2, 2]
    # no batch, no channel
    lay = _Layout(torch.zeros(5, 6, 2), torch.zeros(7, 8))
    assert lay.grid.shape == (1, 5, 6, 2) and lay.volume.shape == (1, 1, 7, 8) and lay.dim == 2
    assert lay.restore(torch.zeros(1, 1, 5, 6)).shape == (5, 6)
    # channel, broadcast batch (zero-stride expand, no copy)
    base = torch.zeros(3, 7, 8)
    lay = _Layout(torch.zeros(4, 5, 6, 2), base)
    assert lay.grid.shape == (4, 5, 6, 2) and lay.volume.shape == (4, 3, 7, 8) and lay.volume.stride(0) == 0
    assert lay.volume.data_ptr() == base.data_ptr()
    assert lay.restore(torch.zeros(4, 3, 5, 6)).shape == (4, 3, 5, 6)
    assert lay.restore(torch.zeros(4, 3, 5, 6, 2)).shape == (4, 3, 5, 6, 2)         # grad: trailing feature axis
    # multiple batch axes
    lay = _Layout(torch.zeros(2, 1, 5, 6, 2), torch.zeros(3, 4, 7, 8))
    assert lay.grid.shape == (6, 5, 6, 2) and lay.volume.shape == (6, 4, 7, 8)
    assert lay.restore(torch.zeros(6, 4, 5, 6)).shape == (2, 3, 4, 5, 6)
    # push: spatial shapes broadcast together (api.py:118-119)
    lay = _Layout(torch.zeros(5, 6, 2), torch.zeros(3, 1, 6), splat=True)
    assert lay.volume.shape == (1, 3, 5, 6)
    # count
    lay = _Layout(torch.zeros(2, 5, 6, 2))
    assert lay.grid.shape == (2, 5, 6, 2) and lay.restore(torch.zeros(2, 1, 7, 7)).shape == (2, 1, 7, 7)
    assert _Layout(torch.zeros(5, 6, 2)).restore(torch.zeros(1, 1, 7, 7)).shape == (7, 7)


def test_grid_helpers_cpu():
    import interpol_b200 as ib
    g = ib.identity_grid([2, 3])
    assert g.shape == (2, 3, 2) and g[1, 2].tolist() == [1., 2.]
    d = torch.zeros(2, 3, 2)
    assert torch.equal(ib.add_identity_grid(d), g) and float(d.abs().sum()) == 0
    assert torch.equal(ib.add_identity_grid_(d), g)
    a = ib.affine_grid(torch.eye(3), [2, 3])
    assert torch.equal(a, g)
    m = torch.eye(3)[None].repeat(4, 1, 1)
    m[:, 0, 2] = torch.arange(4.)
    a = ib.affine_grid(m, [2, 3])
    assert a.shape == (4, 2, 3, 2) and torch.equal(a[3, ..., 0], g[..., 0] + 3)
    with pytest.raises(ValueError):
        ib.affine_grid(torch.eye(3), [2, 3, 4])


def test_install_as_backend_seam():
    """interpol/backend.py:1 + interpol/jitfields.py:47-114: rebinding the seam
    makes the reference's own entry points forward to this engine."""
    import types
    import interpol_b200 as ib
    fake = types.ModuleType('fake_interpol')
    fake.backend = types.ModuleType('fake_interpol.backend')
    fake.backend.jitfields = False
    api = types.ModuleType('fake_interpol.api')
    api.jitfields = None
    sys.modules['fake_interpol'] = fake
    sys.modules['fake_interpol.api'] = api
    try:
        shim = ib.install_as_backend(fake)
        assert fake.backend.jitfields is True and api.jitfields is shim and shim.available
        assert shim.grid_pull is ib.grid_pull and shim.restrict is ib.restrict
    finally:
        del sys.modules['fake_interpol'], sys.modules['fake_interpol.api']


def test_pinned_result_pool_reuses_released_buffers(monkeypatch):
    """api._pinned_empty hands a page-locked buffer out again only after the caller dropped every
    view of it (host logic; the page-locking itself is stubbed out: no CUDA here)."""
    import torch
    import interpol_b200.api as api
    real_empty = torch.empty

    def fake_empty(*a, **k):
        k.pop('pin_memory', None)
        return real_empty(*a, **k)
    monkeypatch.setattr(torch, 'empty', fake_empty)
    monkeypatch.setattr(api, '_POOL', [])
    a = api._pinned_empty([4, 5], torch.float32)
    pa = a.data_ptr()
    b = api._pinned_empty([4, 5], torch.float32)
    assert b.data_ptr() != pa                       # `a` is still alive
    view = a[1:]
    del a
    c = api._pinned_empty([4, 5], torch.float32)
    assert c.data_ptr() not in (pa,)                # a view of `a` is still alive
    del view
    d = api._pinned_empty([2, 5], torch.float64)
    assert d.data_ptr() == pa and d.shape == (2, 5) and d.dtype == torch.float64
    assert len(api._POOL) == 3


def test_label_and_separable_dispatch_rules():
    """host-side routing (no CUDA needed): which inputs take the one-pass label kernel / the separable
    resize and restrict passes, and which fall back to the reference's own formulation."""
    import importlib
    import torch
    import interpol_b200.api as api
    rz = importlib.import_module('interpol_b200.resize')
    rs = importlib.import_module('interpol_b200.restrict')
    g32 = torch.zeros(1, 4, 4, 2)
    lab = torch.randint(0, 5, [1, 1, 4, 4])
    assert api._labels_fused_ok(lab, g32, 1) and api._labels_fused_ok(lab.to(torch.uint8), g32, [0, 1])
    assert not api._labels_fused_ok(lab, g32, 3)                      # higher orders prefilter the masks
    assert not api._labels_fused_ok(lab, g32.half(), 1)               # 16-bit grids: label loop
    assert api._labels_fused_ok(lab + 2 ** 40, g32, 1)                # int64 maps are read natively (any value)
    assert api._labels_fused_ok(lab.to(torch.int8), g32, 0) and not api._labels_fused_ok(lab.bool(), g32, 0)
    x = torch.zeros(1, 1, 8, 8)
    ok_cuda = torch.cuda.is_available()
    assert rz._separable_ok(x, 2, {'bound': 'dct2'}) == ok_cuda
    assert not rz._separable_ok(x, 1, {})                             # spatial dims must be the trailing ones exactly
    assert not rz._separable_ok(x.long(), 2, {})                      # label maps
    assert not rz._separable_ok(x.half().requires_grad_(), 2, {})     # 16-bit adjoint not available
    assert not rz._separable_ok(x, 2, {'unknown_option': 1})
    assert rs._separable_ok(x, 2, {}) == ok_cuda and not rs._separable_ok(x.half(), 2, {})
    rz.SEPARABLE = False
    try:
        assert not rz._separable_ok(x, 2, {})
    finally:
        rz.SEPARABLE = True
