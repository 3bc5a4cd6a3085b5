"""ctypes front-end of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never
imports this module.

The functions mirror the reference's non-differentiable building blocks
(interpol/pushpull.py:35-233 and interpol/coeff.py:288-347): canonical layouts
`(B, C, *spatial)` for volumes and `(B, *spatial, D)` for grids, integer
`bound` / `interpolation` lists, integer `extrapolate`.  Inputs are numpy
arrays or CPU torch tensors (float32 / float64; anything else is computed in
float64); outputs are numpy arrays of the computation dtype.
"""
import ctypes
import os
import subprocess
import numpy as np

__all__ = [
    'build', 'lib', 'bound_index', 'bound_sign', 'weight', 'grad_weight',
    'hess_weight', 'grid_pull', 'grid_push', 'grid_count', 'grid_grad',
    'grid_pushgrad', 'grid_hess', 'spline_coeff', 'spline_coeff_nd',
    'get_poles', 'set_num_threads',
]

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

i64 = ctypes.c_longlong
_pi64 = ctypes.POINTER(i64)
_pint = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    """Compile oracle/liboracle.so with gcc (a few seconds)."""
    so = os.path.join(_HERE, 'liboracle.so')
    src = os.path.join(_HERE, 'oracle.c')
    if force or not os.path.exists(so) or \
            os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-C', _HERE, 'liboracle.so'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, 'liboracle.so')
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
    return _LIB


def set_num_threads(n):
    """Set the OpenMP team size used by the oracle."""
    try:
        omp = ctypes.CDLL('libgomp.so.1')
        omp.omp_set_num_threads(int(n))
    except OSError:
        os.environ['OMP_NUM_THREADS'] = str(int(n))


def _np(x, dtype=None):
    if hasattr(x, 'detach'):
        x = x.detach().cpu()
        if str(x.dtype) in ('torch.float16', 'torch.bfloat16'):
            x = x.double()
        x = x.numpy()
    x = np.asarray(x)
    if dtype is None:
        dtype = x.dtype if x.dtype in (np.float32, np.float64) else np.float64
    return np.ascontiguousarray(x, dtype=dtype)


def _sfx(dtype):
    return '_f32' if dtype == np.float32 else '_f64'


def _ctype(dtype):
    return ctypes.c_float if dtype == np.float32 else ctypes.c_double


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _ilist(x, n, ctype=ctypes.c_int):
    x = list(x) if isinstance(x, (list, tuple)) else [x]
    if len(x) < n:                      # jit_utils.py:10-15 pad_list_int
        x = x + x[-1:] * (n - len(x))
    x = x[:n]
    x = x + [0] * (3 - len(x))
    return (ctype * 3)(*[int(v) for v in x])


def _common(*arrays):
    dt = np.float32
    for a in arrays:
        if a is not None and a.dtype != np.float32:
            dt = np.float64
    return dt


# --------------------------------------------------------------------------
# scalar helpers (truth tables)
# --------------------------------------------------------------------------

def bound_index(bound, i, n):
    f = lib().orc_bound_index_f64
    f.restype = i64
    f.argtypes = [ctypes.c_int, i64, i64]
    return int(f(int(bound), int(i), int(n)))


def bound_sign(bound, i, n):
    f = lib().orc_bound_sign_f64
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_int, i64, i64]
    return int(f(int(bound), int(i), int(n)))


def _scalar_fn(name, dtype, extra=()):
    f = getattr(lib(), name + _sfx(dtype))
    ct = _ctype(dtype)
    f.restype = ct
    f.argtypes = [ctypes.c_int, ct] + list(extra)
    return f


def weight(order, x, dtype=np.float64):
    f = _scalar_fn('orc_weight', dtype)
    return np.array([f(int(order), float(v)) for v in np.ravel(x)], dtype=dtype)


def grad_weight(order, x, dtype=np.float64, quirk_linear_grad=False):
    f = _scalar_fn('orc_grad', dtype, [ctypes.c_int])
    return np.array([f(int(order), float(v), int(quirk_linear_grad))
                     for v in np.ravel(x)], dtype=dtype)


def hess_weight(order, x, dtype=np.float64):
    f = _scalar_fn('orc_hess', dtype)
    return np.array([f(int(order), float(v)) for v in np.ravel(x)], dtype=dtype)


def get_poles(order):
    buf = (ctypes.c_double * 3)()
    f = lib().orc_get_poles_f64
    f.restype = ctypes.c_int
    n = f(int(order), buf)
    return [buf[i] for i in range(max(n, 0))]


# --------------------------------------------------------------------------
# gather family: pull / grad / hess  (pushpull.py:35-66, 146-172, 207-233)
# --------------------------------------------------------------------------

def _prep_gather(inp, grid):
    inp = _np(inp)
    grid = _np(grid)
    dt = _common(inp, grid)
    inp = _np(inp, dt)
    grid = _np(grid, dt)
    dim = grid.shape[-1]
    B = max(inp.shape[0], grid.shape[0])
    if inp.shape[0] != B:
        inp = np.ascontiguousarray(np.broadcast_to(inp, (B,) + inp.shape[1:]))
    if grid.shape[0] != B:
        grid = np.ascontiguousarray(np.broadcast_to(grid, (B,) + grid.shape[1:]))
    C = inp.shape[1]
    ishape = inp.shape[2:]
    oshape = grid.shape[1:-1]
    assert len(ishape) == dim and len(oshape) == dim and 1 <= dim <= 3
    return inp, grid, dt, dim, B, C, ishape, oshape


def grid_pull(inp, grid, bound, interpolation, extrapolate):
    """pushpull.grid_pull (pushpull.py:35-66) -> (B, C, *oshape)"""
    inp, grid, dt, dim, B, C, ishape, oshape = _prep_gather(inp, grid)
    out = np.empty((B, C) + tuple(oshape), dtype=dt)
    f = getattr(lib(), 'orc_pull' + _sfx(dt))
    f.restype = None
    f(_ptr(inp), _ptr(grid), _ptr(out), i64(B), i64(C), ctypes.c_int(dim),
      _ilist(ishape, dim, i64), _ilist(oshape, dim, i64),
      _ilist(bound, dim), _ilist(interpolation, dim), ctypes.c_int(int(extrapolate)))
    return out


def grid_grad(inp, grid, bound, interpolation, extrapolate, quirk_linear_grad=False):
    """pushpull.grid_grad (pushpull.py:146-172) -> (B, C, *oshape, D)"""
    inp, grid, dt, dim, B, C, ishape, oshape = _prep_gather(inp, grid)
    out = np.empty((B, C) + tuple(oshape) + (dim,), dtype=dt)
    f = getattr(lib(), 'orc_grad_pull' + _sfx(dt))
    f.restype = None
    f(_ptr(inp), _ptr(grid), _ptr(out), i64(B), i64(C), ctypes.c_int(dim),
      _ilist(ishape, dim, i64), _ilist(oshape, dim, i64),
      _ilist(bound, dim), _ilist(interpolation, dim), ctypes.c_int(int(extrapolate)),
      ctypes.c_int(int(quirk_linear_grad)))
    return out


def grid_hess(inp, grid, bound, interpolation, extrapolate, quirk_linear_grad=False):
    """pushpull.grid_hess (pushpull.py:207-233) -> (B, C, *oshape, D, D)"""
    inp, grid, dt, dim, B, C, ishape, oshape = _prep_gather(inp, grid)
    out = np.empty((B, C) + tuple(oshape) + (dim, dim), dtype=dt)
    f = getattr(lib(), 'orc_hess_pull' + _sfx(dt))
    f.restype = None
    f(_ptr(inp), _ptr(grid), _ptr(out), i64(B), i64(C), ctypes.c_int(dim),
      _ilist(ishape, dim, i64), _ilist(oshape, dim, i64),
      _ilist(bound, dim), _ilist(interpolation, dim), ctypes.c_int(int(extrapolate)),
      ctypes.c_int(int(quirk_linear_grad)))
    return out


# --------------------------------------------------------------------------
# scatter family: push / count / pushgrad (pushpull.py:70-142, 175-204)
# --------------------------------------------------------------------------

def _prep_scatter(inp, grid, shape, ncomp):
    grid = _np(grid)
    inp = None if inp is None else _np(inp)
    dt = _common(inp, grid)
    grid = _np(grid, dt)
    dim = grid.shape[-1]
    gshape = grid.shape[1:-1]
    B = grid.shape[0]
    C = 1
    if inp is not None:
        inp = _np(inp, dt)
        B = max(B, inp.shape[0])
        if inp.shape[0] != B:
            inp = np.ascontiguousarray(np.broadcast_to(inp, (B,) + inp.shape[1:]))
        C = inp.shape[1]
        ispatial = inp.shape[2:2 + dim]
        if tuple(ispatial) != tuple(gshape):
            # iso1.py:150 / iso0.py:78
            raise ValueError('Input and grid should have the same spatial shape')
    if grid.shape[0] != B:
        grid = np.ascontiguousarray(np.broadcast_to(grid, (B,) + grid.shape[1:]))
    if shape is None:
        shape = gshape
    shape = tuple(int(s) for s in shape)
    assert len(shape) == dim and 1 <= dim <= 3
    return inp, grid, dt, dim, B, C, gshape, shape


def grid_push(inp, grid, shape, bound, interpolation, extrapolate, nthreads=0):
    """pushpull.grid_push (pushpull.py:70-102) -> (B, C, *shape)"""
    inp, grid, dt, dim, B, C, gshape, shape = _prep_scatter(inp, grid, shape, 1)
    out = np.empty((B, C) + shape, dtype=dt)
    args = [None if inp is None else _ptr(inp), _ptr(grid), _ptr(out), i64(B), i64(C),
            ctypes.c_int(dim), _ilist(gshape, dim, i64), _ilist(shape, dim, i64),
            _ilist(bound, dim), _ilist(interpolation, dim), ctypes.c_int(int(extrapolate))]
    if nthreads and nthreads > 1:
        f = getattr(lib(), 'orc_push_mt' + _sfx(dt))
        args.append(ctypes.c_int(int(nthreads)))
    else:
        f = getattr(lib(), 'orc_push' + _sfx(dt))
    f.restype = None
    f(*args)
    return out


def grid_count(grid, shape, bound, interpolation, extrapolate, nthreads=0):
    """pushpull.grid_count (pushpull.py:106-142) -> (B, 1, *shape)"""
    return grid_push(None, grid, shape, bound, interpolation, extrapolate, nthreads)


def grid_pushgrad(inp, grid, shape, bound, interpolation, extrapolate, quirk_linear_grad=False):
    """pushpull.grid_pushgrad (pushpull.py:175-204): inp (B, C, *gshape, D)"""
    inp, grid, dt, dim, B, C, gshape, shape = _prep_scatter(inp, grid, shape, 0)
    assert inp.shape[-1] == dim
    out = np.empty((B, C) + shape, dtype=dt)
    f = getattr(lib(), 'orc_pushgrad' + _sfx(dt))
    f.restype = None
    f(_ptr(inp), _ptr(grid), _ptr(out), i64(B), i64(C), ctypes.c_int(dim),
      _ilist(gshape, dim, i64), _ilist(shape, dim, i64),
      _ilist(bound, dim), _ilist(interpolation, dim), ctypes.c_int(int(extrapolate)),
      ctypes.c_int(int(quirk_linear_grad)))
    return out


# --------------------------------------------------------------------------
# prefilter (coeff.py:288-347)
# --------------------------------------------------------------------------

def spline_coeff(inp, bound, order, dim=-1):
    """coeff.spline_coeff (coeff.py:288-313); returns a new array."""
    x = _np(inp).copy()
    dt = x.dtype
    ax = dim % x.ndim
    outer = int(np.prod(x.shape[:ax], dtype=np.int64))
    n = x.shape[ax]
    inner = int(np.prod(x.shape[ax + 1:], dtype=np.int64))
    f = getattr(lib(), 'orc_spline_coeff' + _sfx(dt))
    f.restype = ctypes.c_int
    r = f(_ptr(x), i64(outer), i64(n), i64(inner), ctypes.c_int(int(bound)),
          ctypes.c_int(int(order)))
    if r != 0:
        raise NotImplementedError('prefilter bound %r' % (bound,))
    return x


def spline_coeff_nd(inp, bound, order, dim=None):
    """coeff.spline_coeff_nd (coeff.py:317-347); returns a new array."""
    x = _np(inp).copy()
    if dim is None:
        dim = x.ndim
    bound = list(bound) if isinstance(bound, (list, tuple)) else [bound]
    order = list(order) if isinstance(order, (list, tuple)) else [order]
    bound = (bound + bound[-1:] * dim)[:dim] if len(bound) < dim else bound[:dim]
    order = (order + order[-1:] * dim)[:dim] if len(order) < dim else order[:dim]
    for d, b, o in zip(range(dim), bound, order):
        x = spline_coeff(x, b, o, dim=-dim + d)
    return x
