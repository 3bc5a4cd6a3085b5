"""Deterministic case list + input generators shared by make_golden.py (which
runs the *reference*, only possible where /root/reference is mounted) and by
the tests (which re-create the same inputs from the seeds and compare the
oracle / the CUDA path with the stored reference outputs).

Only numpy is used to make inputs so the bytes are identical everywhere
(numpy's PCG64 stream is stable across versions).
"""
import zlib
import numpy as np

ISHAPES = {1: (7,), 2: (5, 6), 3: (4, 5, 6)}
OSHAPES = {1: (9,), 2: (4, 5), 3: (3, 4, 5)}


def _seed(name):
    return zlib.crc32(name.encode())


def make_inputs(name, dim, B=1, C=2, dtype=np.float64, spread=2.0,
                ishape=None, oshape=None):
    """Volume (B,C,*ishape), grid (B,*oshape,dim) in voxel units reaching well
    outside the field of view, plus a source image aligned with the grid for
    push (B,C,*oshape) and pushgrad (B,C,*oshape,dim)."""
    rng = np.random.default_rng(_seed(name))
    ishape = ishape or ISHAPES[dim]
    oshape = oshape or OSHAPES[dim]
    vol = rng.standard_normal((B, C) + tuple(ishape))
    # coordinates: uniform over [-spread*n, (1+spread)*n] mixed with in-bounds
    grid = np.empty((B,) + tuple(oshape) + (dim,))
    for d in range(dim):
        n = ishape[d]
        wide = rng.uniform(-spread * n, (1 + spread) * n, size=(B,) + tuple(oshape))
        near = rng.uniform(-1.0, n, size=(B,) + tuple(oshape))
        pick = rng.uniform(size=(B,) + tuple(oshape)) < 0.5
        grid[..., d] = np.where(pick, wide, near)
    # a few exact integers / half-integers (knots of the splines)
    flat = grid.reshape(-1)
    k = max(1, flat.size // 10)
    pos = rng.choice(flat.size, size=k, replace=False)
    flat[pos[: k // 2]] = np.round(flat[pos[: k // 2]])
    flat[pos[k // 2:]] = np.floor(flat[pos[k // 2:]]) + 0.5
    src = rng.standard_normal((B, C) + tuple(oshape))
    srcg = rng.standard_normal((B, C) + tuple(oshape) + (dim,))
    # make every input exactly representable in float32 so that float32 and
    # float64 runs see the same numbers (coordinates: multiples of 1/256)
    grid = np.round(grid * 256) / 256
    vol = vol.astype(np.float32).astype(np.float64)
    src = src.astype(np.float32).astype(np.float64)
    srcg = srcg.astype(np.float32).astype(np.float64)
    return (vol.astype(dtype), grid.astype(dtype), src.astype(dtype),
            srcg.astype(dtype))


def pushpull_cases():
    """Yield dicts: name, op, dim, order(list), bound(list), extrapolate, dtype."""
    cases = []

    def add(op, dim, order, bound, extrapolate, dtype='f64', C=2, B=1):
        order = list(order) if isinstance(order, (list, tuple)) else [order]
        bound = list(bound) if isinstance(bound, (list, tuple)) else [bound]
        name = '%s_%dd_o%s_b%s_e%d_%s_B%dC%d' % (
            op, dim, ''.join(map(str, order)), ''.join(map(str, bound)),
            extrapolate, dtype, B, C)
        cases.append(dict(name=name, op=op, dim=dim, order=order, bound=bound,
                          extrapolate=extrapolate, dtype=dtype, C=C, B=B))

    # A/B: every order x bound, extrapolate=1, the four public ops
    for dim in (1, 2, 3):
        for order in range(8):
            for bound in range(7):
                for op in ('pull', 'push', 'count', 'grad'):
                    add(op, dim, order, bound, 1)
    # C: extrapolate 0 / 2
    for dim in (1, 2, 3):
        for order in (0, 1, 3):
            for bound in (0, 3, 6):
                for ex in (0, 2):
                    for op in ('pull', 'push', 'count', 'grad'):
                        if op == 'pull' and dim == 2 and order == 0:
                            continue  # reference bug iso0.py:155 (returns a bool mask)
                        add(op, dim, order, bound, ex)
    # D: backward-only ops
    for dim in (1, 2, 3):
        for order in (0, 1, 2, 3, 5):
            for bound in (0, 1, 3, 4, 5, 6):
                add('pushgrad', dim, order, bound, 1)
                add('hess', dim, order, bound, 1)
        for order in (1, 3):
            for ex in (0, 2):
                add('pushgrad', dim, order, 3, ex)
                add('hess', dim, order, 3, ex, C=1)  # C>1: reference mask bug nd.py:455
    # E: per-dimension mixtures (no order-1 axis for derivative ops: splines.py:96)
    for op in ('pull', 'push', 'count'):
        add(op, 2, [1, 3], [3, 6], 1)
        add(op, 2, [0, 2], [0, 4], 1)
        add(op, 3, [1, 3, 2], [3, 6, 0], 1)
        add(op, 3, [0, 1, 5], [5, 2, 1], 1)
        add(op, 3, [3, 3, 3], [3, 6, 0], 0)
    for op in ('grad', 'pushgrad', 'hess'):
        add(op, 2, [2, 3], [3, 6], 1)
        add(op, 3, [3, 2, 4], [3, 6, 0], 1)
        add(op, 3, [0, 3, 2], [1, 2, 5], 1)
    # F: float32 runs of the reference
    for dim in (1, 2, 3):
        for order in (1, 3):
            for op in ('pull', 'push', 'count', 'grad'):
                add(op, dim, order, 3, 1, dtype='f32')
    # G: batch of two with broadcast channel count 1
    for op in ('pull', 'push', 'grad'):
        add(op, 3, 3, 3, 1, B=2, C=1)
        add(op, 2, 1, 0, 0, B=2, C=3)
    return cases


COEFF_LENGTHS = [1, 2, 3, 7, 9, 11, 40]
COEFF_BOUNDS = [0, 1, 2, 3, 6]


def coeff_cases():
    cases = []
    for n in COEFF_LENGTHS:
        for bound in COEFF_BOUNDS:
            for order in range(2, 8):
                for dtype in ('f64', 'f32'):
                    cases.append(dict(name='coeff_n%d_b%d_o%d_%s' % (n, bound, order, dtype),
                                      n=n, bound=bound, order=order, dtype=dtype))
    return cases


def coeff_input(name, n, dtype=np.float64):
    rng = np.random.default_rng(_seed(name))
    x = rng.standard_normal((3, n, 2)).astype(np.float32)
    return x.astype(dtype)   # filter along axis 1


def coeff_nd_cases():
    cases = []
    for shape, order, bound in [((6, 7), [3, 3], [3, 3]), ((5, 6, 7), [3, 2, 5], [3, 6, 2]),
                                ((4, 9, 8), [7, 4, 6], [1, 0, 6]), ((12,), [3], [2])]:
        name = 'coeffnd_%s_o%s_b%s' % ('x'.join(map(str, shape)), ''.join(map(str, order)),
                                       ''.join(map(str, bound)))
        cases.append(dict(name=name, shape=shape, order=order, bound=bound))
    return cases


def coeff_nd_input(name, shape):
    rng = np.random.default_rng(_seed(name))
    return rng.standard_normal((2,) + tuple(shape)).astype(np.float32).astype(np.float64)


NP_DTYPE = {'f64': np.float64, 'f32': np.float32}
