"""`restrict`: adjoint of `resize` (reference: interpol/restrict.py:9-122; same arguments).

The reference splats through `grid_push` on the dense tensor-product grid; here float32 / float64 images
take one adjoint 1-D pass per axis (`ib200_resample_axis_adjoint`, see resize.py), everything else the
dense grid + `grid_push`."""
import torch

from .api import grid_push, _stage
from .utils import make_list, meshgrid_ij

SEPARABLE = True     # False: always build the dense grid and call grid_push (A/B testing)

__all__ = ['restrict']


def restrict(image, factor=None, shape=None, anchor='c',
             interpolation=1, reduce_sum=False, **kwargs):
    """Restrict (down-sample by splatting) an image by a factor or to a shape.

    image : (batch, channel, *inshape);  returns (batch, channel, *shape)
    reduce_sum : if False the result is divided by the product of the scales
    """
    factor = make_list(factor) if factor else []
    shape = make_list(shape) if shape else []
    anchor = make_list(anchor)
    nb_dim = max(len(factor), len(shape), len(anchor)) or (image.dim() - 2)
    anchor = [a[0].lower() for a in make_list(anchor, nb_dim)]
    bck = dict(dtype=image.dtype, device=image.device)

    inshape = image.shape[-nb_dim:]
    if factor:
        factor = make_list(factor, nb_dim)
    elif not shape:
        raise ValueError('One of `factor` or `shape` must be provided')
    if shape:
        shape = make_list(shape, nb_dim)
    else:
        shape = [int(i/f) for i, f in zip(inshape, factor)]
    if not factor:
        factor = [i/o for o, i in zip(shape, inshape)]

    lin = []
    fullscale = 1
    for anch, f, inshp, outshp in zip(anchor, factor, inshape, shape):
        if anch == 'c':
            lin.append(torch.linspace(0, outshp - 1, inshp, **bck))
            fullscale *= (inshp - 1) / (outshp - 1)
        elif anch == 'e':
            scale = outshp / inshp
            shift = 0.5 * (scale - 1)
            fullscale *= scale
            lin.append(torch.arange(0., inshp, **bck) * scale + shift)
        elif anch == 'f':
            fullscale *= 1/f
            lin.append(torch.arange(0., inshp, **bck) / f)
        elif anch == 'l':
            shift = (outshp - 1) - (inshp - 1) / f
            fullscale *= 1/f
            lin.append(torch.arange(0., inshp, **bck) / f + shift)
        else:
            raise ValueError('Unknown anchor {}'.format(anch))

    kwargs.setdefault('bound', 'nearest')
    kwargs.setdefault('extrapolate', True)
    kwargs.setdefault('interpolation', interpolation)
    kwargs.setdefault('prefilter', False)
    if _separable_ok(image, nb_dim, kwargs):
        resized = _restrict_separable(image, lin, shape, nb_dim, **kwargs)
    else:
        grid = torch.stack(meshgrid_ij(*lin), dim=-1)
        resized = grid_push(image, grid, shape, **kwargs)
    if not reduce_sum:
        resized = resized / fullscale if resized.requires_grad else resized.div_(fullscale)
    return resized


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
