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


class GridPull(torch.autograd.Function):
    """reference: interpol/autograd.py:157-184"""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, input, grid, interpolation, bound, extrapolate, displacement=False):
        # (`displacement`: optional sixth argument, an extension -- the grid holds displacements)
        opt = _options(interpolation, bound, extrapolate) + (bool(displacement),)
        output = grid_pull(input, grid, *opt)
        ctx.opt = opt
        ctx.save_for_backward(input, grid)
        return output

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        grad_input, grad_grid = grid_pull_backward(grad, *ctx.saved_tensors, *ctx.opt)
        return grad_input, grad_grid, None, None, None, None


class GridPush(torch.autograd.Function):
    """reference: interpol/autograd.py:187-214"""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, input, grid, shape, interpolation, bound, extrapolate, displacement=False):
        opt = _options(interpolation, bound, extrapolate) + (bool(displacement),)
        output = grid_push(input, grid, shape, *opt)
        ctx.opt = opt
        ctx.save_for_backward(input, grid)
        return output

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        grad_input, grad_grid = grid_push_backward(grad, *ctx.saved_tensors, *ctx.opt)
        return grad_input, grad_grid, None, None, None, None, None


class GridCount(torch.autograd.Function):
    """reference: interpol/autograd.py:217-245"""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, grid, shape, interpolation, bound, extrapolate, displacement=False):
        opt = _options(interpolation, bound, extrapolate) + (bool(displacement),)
        output = grid_count(grid, shape, *opt)
        ctx.opt = opt
        ctx.save_for_backward(grid)
        return output

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        grad_grid = None
        if ctx.needs_input_grad[0]:
            grad_grid = grid_count_backward(grad, *ctx.saved_tensors, *ctx.opt)
        return grad_grid, None, None, None, None, None


class GridGrad(torch.autograd.Function):
    """reference: interpol/autograd.py:248-277"""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, input, grid, interpolation, bound, extrapolate, displacement=False):
        opt = _options(interpolation, bound, extrapolate) + (bool(displacement),)
        output = grid_grad(input, grid, *opt)
        ctx.opt = opt
        ctx.save_for_backward(input, grid)
        return output

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        grad_input = grad_grid = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            grad_input, grad_grid = grid_grad_backward(grad, *ctx.saved_tensors, *ctx.opt)
        return grad_input, grad_grid, None, None, None, None


class SplineCoeff(torch.autograd.Function):
    """reference: interpol/autograd.py:280-305 (note: bound before interpolation)"""

    @staticmethod
    @custom_fwd(device_type='cuda')
    def forward(ctx, input, bound, interpolation, dim, inplace):
        bound = bound_to_nitorch(make_list(bound)[0], as_type='int')
        interpolation = inter_to_nitorch(make_list(interpolation)[0], as_type='int')
        opt = (bound, interpolation, dim, inplace)
        if inplace:
            ctx.mark_dirty(input)
        output = spline_coeff(input, *opt)
        if input.requires_grad:
            ctx.opt = opt
        return output

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        # symmetric filter -> backward == forward (autograd.py:301-305)
        grad = spline_coeff(grad, *ctx.opt[:-1], inplace=False)
        return grad, None, None, None, None


class SplineCoeffND(torch.autograd.Function):
    """reference: interpol/autograd.py:308-333"""

    @staticmethod
    @custom_fwd(device_type='cuda')
    def forward(ctx, input, bound, interpolation, dim, inplace):
        bound = bound_to_nitorch(make_list(bound), as_type='int')
        interpolation = inter_to_nitorch(make_list(interpolation), as_type='int')
        opt = (bound, interpolation, dim, inplace)
        if inplace:
            ctx.mark_dirty(input)
        output = spline_coeff_nd(input, *opt)
        if input.requires_grad:
            ctx.opt = opt
        return output

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        grad = spline_coeff_nd(grad, *ctx.opt[:-1], inplace=False)
        return grad, None, None, None, None
