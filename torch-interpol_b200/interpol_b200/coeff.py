"""Spline prefilter host side: `spline_coeff` / `spline_coeff_nd` with the
signatures of the reference's `interpol/coeff.py:288-347`, each axis being one
launch of the line-batched IIR kernel (csrc/coeff.cu) through the C ABI."""
from typing import List, Optional

import torch

from . import _lib
from .pushpull import pad_list_int


def _filter_axis_(x, bound: int, order: int, axis: int):
    """in place on a dense tensor"""
    n = x.shape[axis]
    outer = 1
    for s in x.shape[:axis]:
        outer *= s
    inner = 1
    for s in x.shape[axis + 1:]:
        inner *= s
    L = _lib.lib()
    with _lib.on_device(x.device):
        st = L.ib200_spline_coeff(_lib.ptr(x), _lib.DTYPE_CODE[x.dtype], outer, n, inner,
                                  int(bound), int(order), x.device.index, _lib.stream_ptr(x.device))
    _lib.check(st)


def _check(inp):
    _lib.require_cuda(inp)
    if inp.dtype not in _lib.DTYPE_CODE:
        raise TypeError('unsupported dtype %s' % inp.dtype)


def spline_coeff(inp, bound: int, order: int, dim: int = -1, inplace: bool = False):
    """Interpolating spline coefficients along one dimension.
    Reference: interpol/coeff.py:288-313 (orders 0/1: copy; n == 1: copy)."""
    _check(inp)
    if order in (0, 1):
        return inp if inplace else inp.clone()
    if bound in (4, 5):
        # coeff.py:244,254 (raised inside TorchScript in the reference)
        raise NotImplementedError('spline prefilter: dst1/dst2 boundary conditions are not implemented')
    axis = dim % inp.dim() if inp.dim() else 0
    if inp.dim() == 0:
        return inp if inplace else inp.clone()
    if inplace and inp.is_contiguous():
        _filter_axis_(inp, bound, order, axis)
        return inp
    work = inp.contiguous() if inplace else inp.clone(memory_format=torch.contiguous_format)
    _filter_axis_(work, bound, order, axis)
    if inplace:
        inp.copy_(work)
        return inp
    return work


def spline_coeff_nd(inp, bound: List[int], order: List[int], dim: Optional[int] = None,
                    inplace: bool = False):
    """Interpolating spline coefficients along the last `dim` dimensions.
    Reference: interpol/coeff.py:317-347."""
    _check(inp)
    if dim is None:
        dim = inp.dim()
    bound = pad_list_int(list(bound), dim)
    order = pad_list_int(list(order), dim)
    for b, o in zip(bound, order):
        if o > 1 and b in (4, 5):
            raise NotImplementedError('spline prefilter: dst1/dst2 boundary conditions are not implemented')
    dense = inp.is_contiguous()
    work = inp if (inplace and dense) else \
        (inp.contiguous() if inplace else inp.clone(memory_format=torch.contiguous_format))
    for d, b, o in zip(range(dim), bound, order):
        if o > 1:
            _filter_axis_(work, b, o, work.dim() - dim + d)
    if inplace and work is not inp:
        inp.copy_(work)
        return inp
    return work
