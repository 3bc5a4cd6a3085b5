"""`resize`: change the sampling density of an image (reference: interpol/resize.py:13-119; same
arguments and anchor conventions).

The reference builds the dense sampling grid and calls `grid_pull`.  That grid is the tensor product
of one coordinate vector per axis and B-spline weights are separable, so the same result is obtained
with one 1-D resampling pass per axis (`ib200_resample_axis`): no (B, *out, D) grid in HBM and
3 (order+1) taps per voxel instead of (order+1)^3.  The dense-grid path is kept for what the separable
kernels do not cover (integer label maps, 16-bit inputs that require grad, > 3 spatial dims)."""
import torch

from .api import grid_pull, _stage
from .autograd import bound_to_nitorch, inter_to_nitorch
from .geometry import SamplingPlan
from .utils import meshgrid_ij

__all__ = ['resize']

SEPARABLE = True     # False: always build the dense grid and call grid_pull (A/B testing)


def resize(image, factor=None, shape=None, anchor='c',
           interpolation=1, prefilter=True, **kwargs):
    """Resize an image by a factor or to a specific shape.

    image : (batch, channel, *inshape);  factor : float | list;  shape : list[int]
    anchor : {'centers', 'edges', 'first', 'last'} | list
    returns (batch, channel, *shape)
    """
    plan = SamplingPlan(image, factor, shape, anchor, upsample=True)
    opts = dict(bound='nearest', extrapolate=True, interpolation=interpolation, prefilter=prefilter)
    opts.update(kwargs)
    if _separable_ok(image, plan.ndim, opts):
        return _resize_separable(image, plan.coords, plan.ndim, **opts)
    return grid_pull(image, torch.stack(meshgrid_ij(*plan.coords), dim=-1), **opts)


def _separable_ok(image, nb_dim, kwargs):
    if not SEPARABLE:
        return False
    if not torch.is_tensor(image) or not image.dtype.is_floating_point:
        return False
    if image.requires_grad and image.dtype not in (torch.float32, torch.float64):
        return False        # the adjoint pass accumulates with float32 / float64 global atomics
    if nb_dim < 1 or nb_dim > 3 or image.dim() != nb_dim + 2 or image.numel() == 0:
        return False
    return set(kwargs) <= {'bound', 'extrapolate', 'interpolation', 'prefilter'} and \
        (image.is_cuda or torch.cuda.is_available())


def _resize_separable(image, lin, nb_dim, interpolation=1, bound='nearest', extrapolate=True, prefilter=True):
    from .separable import ResampleAxis
    from .api import spline_coeff_nd
    from .pushpull import pad_list_int
    from .autograd import _options
    bnd, order, extrapolate = _options(interpolation, bound, extrapolate)
    order, bnd = pad_list_int(order, nb_dim), pad_list_int(bnd, nb_dim)
    (x,), back = _stage(image)
    if prefilter:
        x = spline_coeff_nd(x, interpolation=interpolation, bound=bound, dim=nb_dim)
    all_nearest = all(o == 0 for o in order)
    all_linear = all(o == 1 for o in order)
    # the axis that shrinks the most first: later passes stream less data
    axes = sorted(range(nb_dim), key=lambda d: lin[d].numel() / max(x.shape[2 + d], 1))
    for d in axes:
        x = ResampleAxis.apply(x, lin[d], 2 + d, bnd[d], order[d], extrapolate, all_nearest, all_linear)
    return back(x)
