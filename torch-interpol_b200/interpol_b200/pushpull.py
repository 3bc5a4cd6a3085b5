"""Non-differentiable forward/backward building blocks -- same names, argument
order and shapes as the reference's `interpol/pushpull.py`, but every function
is a single call into the C ABI (include/interpol_b200.h) instead of a loop of
ATen ops.

    inp  : (B, C, *spatial_in)      grid : (B, *spatial_out, D)
    bound: List[int]  interpolation: List[int]  extrapolate: int

Reference: interpol/pushpull.py:35-233 (forward), :237-325 (backward algebra).
"""
import ctypes
from typing import List, Optional

import torch

from . import _lib
from ._lib import Problem

Tensor = torch.Tensor

# module-level switches (A/B testing, parity experiments)
flags = 0


def pad_list_int(x: List[int], dim: int) -> List[int]:
    """jit_utils.py:10-15"""
    x = list(x)
    if len(x) < dim:
        x = x + x[-1:] * (dim - len(x))
    if len(x) > dim:
        x = x[:dim]
    return x


def _common_dtype(*tensors):
    dt = None
    for t in tensors:
        if t is None:
            continue
        if not t.dtype.is_floating_point:
            raise TypeError('interpol_b200 kernels need floating point tensors (got %s)' % t.dtype)
        dt = t.dtype if dt is None else torch.promote_types(dt, t.dtype)
    if dt not in _lib.DTYPE_CODE:
        raise TypeError('unsupported dtype %s' % dt)
    return dt


DISPLACEMENT = 16     # IB200_FLAG_DISPLACEMENT: the grid holds displacements (coordinate = lattice index + value)


def _problem(dim, dtype, device, bound, interpolation, extrapolate, batch, channels,
             vol_shape, pts_shape, displacement=False):
    if dim < 1 or dim > 3:
        # nd.py handles any dimension through ATen; there is no CPU fallback here
        raise NotImplementedError('interpol_b200 supports 1, 2 or 3 spatial dimensions (got %d)' % dim)
    p = Problem()
    p.dim = dim
    p.dtype = _lib.DTYPE_CODE[dtype]
    p.extrapolate = int(extrapolate)
    p.device = device.index if device.index is not None else torch.cuda.current_device()
    b = pad_list_int(bound, dim)
    o = pad_list_int(interpolation, dim)
    for d in range(dim):
        p.bound[d] = int(b[d])
        p.order[d] = int(o[d])
        p.vol_shape[d] = int(vol_shape[d])
        p.pts_shape[d] = int(pts_shape[d])
    p.flags = flags | (DISPLACEMENT if displacement else 0)
    p.batch = batch
    p.channels = channels
    return p


def _dense_grid(grid, dim):
    """The tiled kernels stage (B, *spatial, D) grids as dense array-of-structs tiles; a strided grid (e.g.
    component-major, the layout `disp.movedim(1, -1)` leaves behind) runs on the one-thread-per-point kernels,
    3-4x slower on large 3-D problems than paying one copy of the grid."""
    if dim != 3 or grid.is_contiguous() or grid.shape[0] == 0:
        return grid
    if grid[0].numel() >= 3 * 32768 and not grid[0].is_contiguous():
        return grid.contiguous()
    return grid


def _bstride(t, batch):
    """batch stride with broadcasting of a singleton batch (reference: expand)"""
    return 0 if (t.shape[0] == 1 and batch != 1) else t.stride(0)


def _set_vol(p, vol, batch, dim):
    p.vol_stride[0] = _bstride(vol, batch)
    p.vol_stride[1] = vol.stride(1)
    for d in range(dim):
        p.vol_stride[2 + d] = vol.stride(2 + d)


def _set_grid(p, grid, batch, dim):
    p.grid_stride[0] = _bstride(grid, batch)
    for d in range(dim):
        p.grid_stride[1 + d] = grid.stride(1 + d)
    p.grid_stride[1 + dim] = grid.stride(1 + dim)


def _set_img(p, img, batch, dim, comp=False):
    p.img_stride[0] = _bstride(img, batch)
    p.img_stride[1] = img.stride(1)
    for d in range(dim):
        p.img_stride[2 + d] = img.stride(2 + d)
    if comp:
        p.img_stride[2 + dim] = img.stride(2 + dim)


def _batch(*tensors):
    b = 1
    for t in tensors:
        if t.shape[0] != 1:
            if b != 1 and t.shape[0] != b:
                raise ValueError('Incompatible batch sizes: %d and %d' % (b, t.shape[0]))
            b = t.shape[0]
    if any(t.shape[0] == 0 for t in tensors):
        b = 0
    return b


def _gather(fn_name, inp, grid, bound, interpolation, extrapolate, trailing, gout=None, gout_comp=False,
            displacement=False):
    _lib.require_cuda(inp, grid, gout)
    dim = grid.shape[-1]
    if grid.dim() != dim + 2 or inp.dim() != dim + 2:
        raise ValueError('expected inp (B, C, *spatial) and grid (B, *spatial, D)')
    dtype = _common_dtype(inp, grid, gout)
    inp = inp.to(dtype)
    grid = _dense_grid(grid.to(dtype), dim)
    batch = _batch(inp, grid) if gout is None else _batch(inp, grid, gout)
    channels = inp.shape[1]
    ishape = inp.shape[2:]
    oshape = grid.shape[1:-1]
    p = _problem(dim, dtype, grid.device, bound, interpolation, extrapolate, batch, channels,
                 ishape, oshape, displacement)
    _set_vol(p, inp, batch, dim)
    _set_grid(p, grid, batch, dim)
    L = _lib.lib()
    with _lib.on_device(grid.device):
        s = _lib.stream_ptr(grid.device)
        if gout is None:
            out = torch.empty([batch, channels, *oshape, *trailing(dim)], dtype=dtype, device=grid.device)
            st = getattr(L, fn_name)(ctypes.byref(p), _lib.ptr(inp), _lib.ptr(grid), _lib.ptr(out), s)
        else:
            gout = gout.to(dtype)
            _set_img(p, gout, batch, dim, comp=gout_comp)
            out = torch.empty([batch, *oshape, dim], dtype=dtype, device=grid.device)
            fn = L.ib200_grad_backward_grid if gout_comp else L.ib200_pull_backward_grid
            st = fn(ctypes.byref(p), _lib.ptr(inp), _lib.ptr(grid), _lib.ptr(gout), _lib.ptr(out), s)
    _lib.check(st)
    return out


LABEL_TYPES = {torch.int32: 0, torch.int64: 1, torch.uint8: 2, torch.int16: 3}      # IB200_LABEL_*


def grid_pull_labels(inp, grid, bound: List[int], interpolation: List[int], extrapolate: int, displacement=False):
    """(B, C, *spatial_in) integer label map (int32 / int64 / uint8 / int16, read as stored), (B, *spatial_out, D)
    float32/64 grid -> (B, C, *spatial_out) of the same integer type: per point, the label whose soft mask
    interpolates to the largest value (orders 0 / 1).  One pass instead of the loop over `input.unique()` of
    interpol/api.py:194-205."""
    _lib.require_cuda(inp, grid)
    dim = grid.shape[-1]
    if grid.dim() != dim + 2 or inp.dim() != dim + 2:
        raise ValueError('expected inp (B, C, *spatial) and grid (B, *spatial, D)')
    if inp.dtype not in LABEL_TYPES or grid.dtype not in (torch.float32, torch.float64):
        raise TypeError('grid_pull_labels: int32 / int64 / uint8 / int16 labels and a float32 / float64 grid expected')
    batch = _batch(inp, grid)
    channels = inp.shape[1]
    oshape = grid.shape[1:-1]
    p = _problem(dim, grid.dtype, grid.device, bound, interpolation, extrapolate, batch, channels, inp.shape[2:], oshape,
                 displacement)
    p.reserved = LABEL_TYPES[inp.dtype]
    _set_vol(p, inp, batch, dim)
    _set_grid(p, grid, batch, dim)
    L = _lib.lib()
    with _lib.on_device(grid.device):
        out = torch.empty([batch, channels, *oshape], dtype=inp.dtype, device=grid.device)
        st = L.ib200_pull_labels(ctypes.byref(p), _lib.ptr(inp), _lib.ptr(grid), _lib.ptr(out), _lib.stream_ptr(grid.device))
    _lib.check(st)
    return out


def grid_pull(inp, grid, bound: List[int], interpolation: List[int], extrapolate: int, displacement=False):
    """(B, C, *spatial_in), (B, *spatial_out, D) -> (B, C, *spatial_out)
    Reference: interpol/pushpull.py:35-66.  `displacement`: the grid holds displacements (every function below)."""
    return _gather('ib200_pull', inp, grid, bound, interpolation, extrapolate, lambda d: [], displacement=displacement)


def grid_grad(inp, grid, bound: List[int], interpolation: List[int], extrapolate: int, displacement=False):
    """-> (B, C, *spatial_out, D).  Reference: interpol/pushpull.py:146-172."""
    return _gather('ib200_grad', inp, grid, bound, interpolation, extrapolate, lambda d: [d], displacement=displacement)


def grid_hess(inp, grid, bound: List[int], interpolation: List[int], extrapolate: int, displacement=False):
    """-> (B, C, *spatial_out, D, D).  Reference: interpol/pushpull.py:207-233."""
    return _gather('ib200_hess', inp, grid, bound, interpolation, extrapolate, lambda d: [d, d], displacement=displacement)


def grid_pull_grad_grid(gout, inp, grid, bound, interpolation, extrapolate, displacement=False):
    """Fused `(grid_grad(inp, grid) * gout.unsqueeze(-1)).sum(1)` -> (B, *spatial_out, D)
    (the grid branch of interpol/pushpull.py:254-257 without the (B,C,*,D) temporary)."""
    return _gather('', inp, grid, bound, interpolation, extrapolate, None, gout=gout, displacement=displacement)


def grid_grad_grad_grid(gout, inp, grid, bound, interpolation, extrapolate, displacement=False):
    """Fused `(grid_hess(inp, grid) * gout.unsqueeze(-1)).sum(dim=[1, -2])` -> (B, *spatial_out, D), gout
    (B, C, *spatial_out, D): the grid branch of interpol/pushpull.py:318-324 without the (B,C,*,D,D) Hessian."""
    return _gather('', inp, grid, bound, interpolation, extrapolate, None, gout=gout, gout_comp=True,
                   displacement=displacement)


def _scatter(fn_name, inp, grid, shape, bound, interpolation, extrapolate, comp=False, displacement=False):
    _lib.require_cuda(inp, grid)
    dim = grid.shape[-1]
    if grid.dim() != dim + 2:
        raise ValueError('expected grid (B, *spatial, D)')
    gshape = grid.shape[1:-1]
    dtype = _common_dtype(inp, grid)
    grid = _dense_grid(grid.to(dtype), dim)
    if inp is not None:
        if inp.dim() != dim + 2 + int(comp):
            raise ValueError('expected inp (B, C, *spatial%s)' % (', D' if comp else ''))
        inp = inp.to(dtype)
        if tuple(inp.shape[2:2 + dim]) != tuple(gshape):
            # iso1.py:150, iso0.py:78
            raise ValueError('Input and grid should have the same spatial shape')
        batch = _batch(inp, grid)
        channels = inp.shape[1]
    else:
        batch = grid.shape[0]
        channels = 1
    if shape is None:
        shape = gshape
    shape = [int(s) for s in shape]
    if len(shape) != dim:
        raise ValueError('`shape` should have %d elements' % dim)
    p = _problem(dim, dtype, grid.device, bound, interpolation, extrapolate, batch, channels,
                 shape, gshape, displacement)
    _set_grid(p, grid, batch, dim)
    if inp is not None:
        _set_img(p, inp, batch, dim, comp)
    L = _lib.lib()
    with _lib.on_device(grid.device):
        out = torch.empty([batch, channels, *shape], dtype=dtype, device=grid.device)
        nscratch = L.ib200_scratch_bytes(ctypes.byref(p))
        scratch = torch.empty([nscratch // 4], dtype=torch.float32, device=grid.device) if nscratch else None
        s = _lib.stream_ptr(grid.device)
        if inp is None:
            st = L.ib200_count(ctypes.byref(p), _lib.ptr(grid), _lib.ptr(out), _lib.ptr(scratch), s)
        else:
            st = getattr(L, fn_name)(ctypes.byref(p), _lib.ptr(inp), _lib.ptr(grid), _lib.ptr(out),
                                     _lib.ptr(scratch), s)
    _lib.check(st)
    return out


def grid_push(inp, grid, shape: Optional[List[int]], bound: List[int], interpolation: List[int],
              extrapolate: int, displacement=False):
    """(B, C, *spatial_in), (B, *spatial_in, D) -> (B, C, *shape)
    Reference: interpol/pushpull.py:70-102."""
    return _scatter('ib200_push', inp, grid, shape, bound, interpolation, extrapolate, displacement=displacement)


def grid_count(grid, shape: Optional[List[int]], bound: List[int], interpolation: List[int],
               extrapolate: int, displacement=False):
    """(B, *spatial_in, D) -> (B, 1, *shape).  Reference: interpol/pushpull.py:106-142."""
    return _scatter('ib200_count', None, grid, shape, bound, interpolation, extrapolate, displacement=displacement)


def grid_pushgrad(inp, grid, shape: List[int], bound: List[int], interpolation: List[int],
                  extrapolate: int, displacement=False):
    """(B, C, *spatial_in, D) -> (B, C, *shape).  Reference: interpol/pushpull.py:175-204."""
    return _scatter('ib200_pushgrad', inp, grid, shape, bound, interpolation, extrapolate, comp=True,
                    displacement=displacement)


# ---------------------------------------------------------------------------
# backward compositions (interpol/pushpull.py:237-325)
# ---------------------------------------------------------------------------

def grid_pull_backward(grad, inp, grid, bound, interpolation, extrapolate, displacement=False):
    """-> (B, C, *spatial_in), (B, *spatial_out, D).  Reference: pushpull.py:237-258.
    (the derivative w.r.t. a displacement equals the derivative w.r.t. the coordinate)"""
    dim = grid.shape[-1]
    grad_inp = grad_grid = None
    if inp.requires_grad:
        grad_inp = grid_push(grad, grid, inp.shape[-dim:], bound, interpolation, extrapolate, displacement)
    if grid.requires_grad:
        grad_grid = grid_pull_grad_grid(grad, inp, grid, bound, interpolation, extrapolate, displacement)
    return grad_inp, grad_grid


def grid_push_backward(grad, inp, grid, bound, interpolation, extrapolate, displacement=False):
    """-> (B, C, *spatial_in), (B, *spatial_in, D).  Reference: pushpull.py:262-282."""
    grad_inp = grad_grid = None
    if inp.requires_grad:
        grad_inp = grid_pull(grad, grid, bound, interpolation, extrapolate, displacement)
    if grid.requires_grad:
        # sum_c grad(grad_vol, grid)[b,c] * inp[b,c]: same fused kernel, roles swapped
        grad_grid = grid_pull_grad_grid(inp, grad, grid, bound, interpolation, extrapolate, displacement)
    return grad_inp, grad_grid


def grid_count_backward(grad, grid, bound, interpolation, extrapolate, displacement=False):
    """-> (B, *spatial_in, D).  Reference: pushpull.py:286-299."""
    if grid.requires_grad and grad.shape[1] == 1:
        # one channel (what GridCount produces): the gradient kernels' own fast paths
        return grid_grad(grad, grid, bound, interpolation, extrapolate, displacement)[:, 0]
    if grid.requires_grad:
        ones = torch.ones([1, 1, *([1] * (grid.dim() - 2))], dtype=grad.dtype, device=grad.device)
        ones = ones.expand([grid.shape[0], grad.shape[1], *grid.shape[1:-1]])
        return grid_pull_grad_grid(ones, grad, grid, bound, interpolation, extrapolate, displacement)
    return None


def grid_grad_backward(grad, inp, grid, bound, interpolation, extrapolate, displacement=False):
    """grad (B, C, *spatial_out, D) -> (B, C, *spatial_in), (B, *spatial_out, D).
    Reference: pushpull.py:303-325."""
    dim = grid.shape[-1]
    shape = inp.shape[-dim:]
    grad_inp = grad_grid = None
    if inp.requires_grad:
        grad_inp = grid_pushgrad(grad, grid, shape, bound, interpolation, extrapolate, displacement)
    if grid.requires_grad:
        grad_grid = grid_grad_grad_grid(grad, inp, grid, bound, interpolation, extrapolate, displacement)
    return grad_inp, grad_grid
