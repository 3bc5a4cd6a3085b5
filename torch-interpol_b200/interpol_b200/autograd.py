"""Autograd boundary: the six `torch.autograd.Function`s of the reference
(`interpol/autograd.py:157-333`) with identical `.apply` signatures and
gradient tuples, backed by the CUDA kernels behind the C ABI."""
import torch
from torch.amp import custom_fwd, custom_bwd

from .bounds import BoundType
from .splines import InterpolationType
from .coeff import spline_coeff, spline_coeff_nd
from .pushpull import (
    grid_pull, grid_pull_backward,
    grid_push, grid_push_backward,
    grid_count, grid_count_backward,
    grid_grad, grid_grad_backward)


def make_list(x):
    if not isinstance(x, (list, tuple)):
        x = [x]
    return list(x)


_BOUND_ALIASES = {
    'replicate': 'replicate', 'repeat': 'replicate', 'border': 'replicate', 'nearest': 'replicate',
    'zero': 'zero', 'zeros': 'zero', 'constant': 'zero',
    'dct2': 'dct2', 'reflect': 'dct2', 'reflection': 'dct2', 'neumann': 'dct2',
    'dct1': 'dct1', 'mirror': 'dct1',
    'dft': 'dft', 'wrap': 'dft', 'circular': 'dft',
    'dst2': 'dst2', 'antireflect': 'dst2', 'dirichlet': 'dst2',
    'dst1': 'dst1', 'antimirror': 'dst1',
}


def bound_to_nitorch(bound, as_type='str'):
    """Canonicalise boundary names (reference: interpol/autograd.py:56-103).

    bound : [list of] str, int or BoundType;  as_type : 'str' | 'enum' | 'int'
    """
    intype = type(bound)
    if not isinstance(bound, (list, tuple)):
        bound = [bound]
    obound = []
    for b in bound:
        if isinstance(b, str):
            name = _BOUND_ALIASES.get(b.lower())
            if name is None:
                raise ValueError(f'Unknown boundary condition {b}')
            obound.append(BoundType[name])
        elif isinstance(b, BoundType):
            obound.append(b)
        elif isinstance(b, int) and not isinstance(b, bool):
            obound.append(BoundType(b))     # ValueError if not 0..6
        else:
            raise ValueError(f'Unknown boundary condition {b}')
    if as_type in ('int', int):
        obound = [b.value for b in obound]
    elif as_type in ('str', str):
        obound = [b.name for b in obound]
    if issubclass(intype, (list, tuple)):
        return intype(obound)
    return obound[0]


_ORDER_NAMES = ['nearest', 'linear', 'quadratic', 'cubic', 'fourth', 'fifth', 'sixth', 'seventh']


def inter_to_nitorch(inter, as_type='str'):
    """Canonicalise interpolation orders (reference: interpol/autograd.py:106-154)."""
    intype = type(inter)
    if not isinstance(inter, (list, tuple)):
        inter = [inter]
    ointer = []
    for o in inter:
        if isinstance(o, str):
            o = o.lower()
            if o not in _ORDER_NAMES:
                raise ValueError(f'Unknown interpolation order {o}')
            ointer.append(_ORDER_NAMES.index(o))
        elif isinstance(o, InterpolationType):
            ointer.append(o.value)
        elif isinstance(o, int) and not isinstance(o, bool) and 0 <= o <= 7:
            ointer.append(int(o))
        else:
            raise ValueError(f'Unknown interpolation order {o}')
    if as_type in ('enum', 'str', str):
        ointer = [InterpolationType(o) for o in ointer]
        if as_type in ('str', str):
            ointer = [o.name for o in ointer]
    if issubclass(intype, (list, tuple)):
        return intype(ointer)
    return ointer[0]


def _options(interpolation, bound, extrapolate):
    bound = bound_to_nitorch(make_list(bound), as_type='int')
    interpolation = inter_to_nitorch(make_list(interpolation), as_type='int')
    return bound, interpolation, int(extrapolate)


# The six Functions differ only in which kernels they call and in how many of their leading arguments are
# tensors, so they are stamped out of two templates.  `.apply` keeps the reference's positional signatures
# (interpol/autograd.py:157-333); the sampling Functions accept one optional trailing argument, `displacement`
# (an extension: the grid holds displacements), and return one gradient slot per argument they were given.

def _sampling_function(name, kernel, adjoint, n_tensors, takes_shape, signature, lines):
    def forward(ctx, *args):
        tensors, rest = args[:n_tensors], list(args[n_tensors:])
        shape = [rest.pop(0)] if takes_shape else []
        interpolation, bound, extrapolate = rest[:3]
        displacement = bool(rest[3]) if len(rest) > 3 else False
        ctx.opt = _options(interpolation, bound, extrapolate) + (displacement,)
        ctx.n_args = len(args)
        ctx.save_for_backward(*tensors)
        return kernel(*tensors, *shape, *ctx.opt)

    def backward(ctx, grad):
        grads = (None,) * n_tensors
        if any(ctx.needs_input_grad[:n_tensors]):
            grads = adjoint(grad, *ctx.saved_tensors, *ctx.opt)
            if n_tensors == 1:
                grads = (grads,)
        return tuple(grads) + (None,) * (ctx.n_args - n_tensors)

    body = {
        '__doc__': '`%s.apply(%s[, displacement])` -- reference: interpol/autograd.py:%s' % (name, signature, lines),
        'forward': staticmethod(custom_fwd(device_type='cuda', cast_inputs=torch.float32)(forward)),
        'backward': staticmethod(custom_bwd(device_type='cuda')(backward)),
    }
    return type(name, (torch.autograd.Function,), body)


GridPull = _sampling_function('GridPull', grid_pull, grid_pull_backward, 2, False,
                              'input, grid, interpolation, bound, extrapolate', '157-184')
GridPush = _sampling_function('GridPush', grid_push, grid_push_backward, 2, True,
                              'input, grid, shape, interpolation, bound, extrapolate', '187-214')
GridCount = _sampling_function('GridCount', grid_count, grid_count_backward, 1, True,
                               'grid, shape, interpolation, bound, extrapolate', '217-245')
GridGrad = _sampling_function('GridGrad', grid_grad, grid_grad_backward, 2, False,
                              'input, grid, interpolation, bound, extrapolate', '248-277')


def _prefilter_function(name, kernel, per_axis, lines):
    """`apply(input, bound, interpolation, dim, inplace)` -- note: bound BEFORE interpolation, like the reference.
    The filter is symmetric, so the backward pass is the same filter applied to the gradient (autograd.py:301-305)."""
    def forward(ctx, input, bound, interpolation, dim, inplace):
        bound, interpolation = make_list(bound), make_list(interpolation)
        if not per_axis:
            bound, interpolation = bound[0], interpolation[0]
        opt = (bound_to_nitorch(bound, as_type='int'), inter_to_nitorch(interpolation, as_type='int'), dim)
        if inplace:
            ctx.mark_dirty(input)
        if input.requires_grad:
            ctx.opt = opt
        return kernel(input, *opt, inplace)

    def backward(ctx, grad):
        return kernel(grad, *ctx.opt, inplace=False), None, None, None, None

    body = {
        '__doc__': '`%s.apply(input, bound, interpolation, dim, inplace)` -- reference: interpol/autograd.py:%s' % (name, lines),
        'forward': staticmethod(custom_fwd(device_type='cuda')(forward)),
        'backward': staticmethod(custom_bwd(device_type='cuda')(backward)),
    }
    return type(name, (torch.autograd.Function,), body)


SplineCoeff = _prefilter_function('SplineCoeff', spline_coeff, False, '280-305')
SplineCoeffND = _prefilter_function('SplineCoeffND', spline_coeff_nd, True, '308-333')
