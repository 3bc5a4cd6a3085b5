"""Separable resampling along one axis (`ib200_resample_axis`) and its adjoint
(`ib200_resample_axis_adjoint`): the building blocks of `resize` / `restrict`, whose sampling grids
are tensor products of one coordinate vector per axis (reference: interpol/resize.py:91-117,
interpol/restrict.py:86-120 build the dense grid and call grid_pull / grid_push)."""
import torch

from . import _lib


def _geom(x, axis):
    outer = 1
    for s in x.shape[:axis]:
        outer *= s
    inner = 1
    for s in x.shape[axis + 1:]:
        inner *= s
    return outer, inner


def _call(fn_name, x, coords, axis, n_out, bound, order, extrapolate, all_nearest, all_linear):
    _lib.require_cuda(x)
    if x.dtype not in _lib.DTYPE_CODE:
        raise TypeError('unsupported dtype %s' % x.dtype)
    x = x.contiguous()
    coords = coords.detach().to(device=x.device, dtype=x.dtype).contiguous()
    outer, inner = _geom(x, axis)
    out = torch.empty([*x.shape[:axis], n_out, *x.shape[axis + 1:]], dtype=x.dtype, device=x.device)
    L = _lib.lib()
    with _lib.on_device(x.device):
        st = getattr(L, fn_name)(_lib.ptr(x), _lib.ptr(out), _lib.ptr(coords), _lib.DTYPE_CODE[x.dtype],
                                 outer, x.shape[axis], n_out, inner, int(order), int(bound), int(extrapolate),
                                 int(bool(all_nearest)), int(bool(all_linear)), x.device.index,
                                 _lib.stream_ptr(x.device))
    _lib.check(st)
    return out


def resample_axis(x, coords, axis, bound, order, extrapolate, all_nearest, all_linear):
    """out[..., i, ...] = sum_k w_k(coords[i]) x[..., fold(start + k), ...] along `axis`."""
    return _call('ib200_resample_axis', x, coords, axis, coords.numel(), bound, order, extrapolate, all_nearest, all_linear)


def resample_axis_adjoint(x, coords, axis, n_out, bound, order, extrapolate, all_nearest, all_linear):
    """out[..., fold(start(coords[i]) + k), ...] += w_k(coords[i]) x[..., i, ...] along `axis` (float32 / float64)."""
    if coords.numel() != x.shape[axis]:
        raise ValueError('one coordinate per input sample expected along the axis')
    return _call('ib200_resample_axis_adjoint', x, coords, axis, int(n_out), bound, order, extrapolate, all_nearest, all_linear)


class ResampleAxis(torch.autograd.Function):
    """resample_axis with its adjoint as backward (coordinates are constants of a resize: no gradient)."""

    @staticmethod
    def forward(ctx, x, coords, axis, bound, order, extrapolate, all_nearest, all_linear):
        ctx.opt = (coords, axis, x.shape[axis], bound, order, extrapolate, all_nearest, all_linear)
        return resample_axis(x, coords, axis, bound, order, extrapolate, all_nearest, all_linear)

    @staticmethod
    def backward(ctx, grad):
        coords, axis, n_in, bound, order, extrapolate, all_nearest, all_linear = ctx.opt
        g = ResampleAxisAdjoint.apply(grad, coords, axis, n_in, bound, order, extrapolate, all_nearest, all_linear)
        return (g,) + (None,) * 7


class ResampleAxisAdjoint(torch.autograd.Function):
    """resample_axis_adjoint with the forward pass as backward."""

    @staticmethod
    def forward(ctx, x, coords, axis, n_out, bound, order, extrapolate, all_nearest, all_linear):
        ctx.opt = (coords, axis, bound, order, extrapolate, all_nearest, all_linear)
        return resample_axis_adjoint(x, coords, axis, n_out, bound, order, extrapolate, all_nearest, all_linear)

    @staticmethod
    def backward(ctx, grad):
        coords, axis, bound, order, extrapolate, all_nearest, all_linear = ctx.opt
        g = ResampleAxis.apply(grad, coords, axis, bound, order, extrapolate, all_nearest, all_linear)
        return (g,) + (None,) * 8
