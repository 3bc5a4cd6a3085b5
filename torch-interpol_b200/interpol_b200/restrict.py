"""`restrict`: adjoint of `resize` (reference: interpol/restrict.py:9-122; same arguments).

The reference splats through `grid_push` on the dense tensor-product grid; here float32 / float64 images
take one adjoint 1-D pass per axis (`ib200_resample_axis_adjoint`, see resize.py), everything else the
dense grid + `grid_push`."""
import torch

from .api import grid_push, _stage
from .geometry import SamplingPlan
from .utils import meshgrid_ij

SEPARABLE = True     # False: always build the dense grid and call grid_push (A/B testing)

__all__ = ['restrict']


def restrict(image, factor=None, shape=None, anchor='c',
             interpolation=1, reduce_sum=False, **kwargs):
    """Restrict (down-sample by splatting) an image by a factor or to a shape.

    image : (batch, channel, *inshape);  returns (batch, channel, *shape)
    reduce_sum : if False the result is divided by the product of the scales
    """
    plan = SamplingPlan(image, factor, shape, anchor, upsample=False)
    opts = dict(bound='nearest', extrapolate=True, interpolation=interpolation, prefilter=False)
    opts.update(kwargs)
    if _separable_ok(image, plan.ndim, opts):
        out = _restrict_separable(image, plan.coords, plan.outshape, plan.ndim, **opts)
    else:
        out = grid_push(image, torch.stack(meshgrid_ij(*plan.coords), dim=-1), plan.outshape, **opts)
    if reduce_sum:
        return out
    return out / plan.scale if out.requires_grad else out.div_(plan.scale)


def _separable_ok(image, nb_dim, kwargs):
    if not SEPARABLE or not torch.is_tensor(image) or image.dtype not in (torch.float32, torch.float64):
        return False
    if nb_dim < 1 or nb_dim > 3 or image.dim() != nb_dim + 2 or image.numel() == 0:
        return False
    return set(kwargs) <= {'bound', 'extrapolate', 'interpolation', 'prefilter'} and \
        (image.is_cuda or torch.cuda.is_available())


def _restrict_separable(image, lin, shape, nb_dim, interpolation=1, bound='nearest', extrapolate=True, prefilter=False):
    from .api import spline_coeff_nd
    from .autograd import _options
    from .pushpull import pad_list_int
    from .separable import ResampleAxisAdjoint
    bnd, order, extrapolate = _options(interpolation, bound, extrapolate)
    order, bnd = pad_list_int(order, nb_dim), pad_list_int(bnd, nb_dim)
    (x,), back = _stage(image)
    all_nearest = all(o == 0 for o in order)
    all_linear = all(o == 1 for o in order)
    # the axis that shrinks the most first: later passes stream less data
    axes = sorted(range(nb_dim), key=lambda d: shape[d] / max(x.shape[2 + d], 1))
    for d in axes:
        x = ResampleAxisAdjoint.apply(x, lin[d], 2 + d, int(shape[d]), bnd[d], order[d], extrapolate, all_nearest, all_linear)
    if prefilter:
        x = spline_coeff_nd(x, interpolation=interpolation, bound=bound, dim=nb_dim, inplace=not x.requires_grad)
    return back(x)
