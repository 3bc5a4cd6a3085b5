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
from .utils import make_list, meshgrid_ij

__all__ = ['resize']

SEPARABLE = True     # False: always build the dense grid and call grid_pull (A/B testing)


def _lattice(anchor, factor, inshape, outshape, bck, restrict=False):
    """1-D sampling coordinates per axis for the four anchor modes."""
    lin = []
    scales = []
    for anch, f, inshp, outshp in zip(anchor, factor, inshape, outshape):
        if anch == 'c':      # centres of the corner voxels are aligned
            lin.append(torch.linspace(0, inshp - 1, outshp, **bck))
            scales.append((outshp - 1) / (inshp - 1) if inshp > 1 else 1.)
        elif anch == 'e':    # edges of the field of view are aligned
            scale = inshp / outshp
            shift = 0.5 * (scale - 1)
            lin.append(torch.arange(0., outshp, **bck) * scale + shift)
            scales.append(1 / scale)
        elif anch == 'f':    # first voxel aligned, exact factor
            lin.append(torch.arange(0., outshp, **bck) / f)
            scales.append(f)
        elif anch == 'l':    # last voxel aligned, exact factor
            shift = (inshp - 1) - (outshp - 1) / f
            lin.append(torch.arange(0., outshp, **bck) / f + shift)
            scales.append(f)
        else:
            raise ValueError('Unknown anchor {}'.format(anch))
    return lin, scales


def resize(image, factor=None, shape=None, anchor='c',
           interpolation=1, prefilter=True, **kwargs):
    """Resize an image by a factor or to a specific shape.

    image : (batch, channel, *inshape);  factor : float | list;  shape : list[int]
    anchor : {'centers', 'edges', 'first', 'last'} | list
    returns (batch, channel, *shape)
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
        shape = [int(i*f) for i, f in zip(inshape, factor)]
    if not factor:
        factor = [o/i for o, i in zip(shape, inshape)]

    lin, _ = _lattice(anchor, factor, inshape, shape, bck)

    kwargs.setdefault('bound', 'nearest')
    kwargs.setdefault('extrapolate', True)
    kwargs.setdefault('interpolation', interpolation)
    kwargs.setdefault('prefilter', prefilter)
    if _separable_ok(image, nb_dim, kwargs):
        return _resize_separable(image, lin, nb_dim, **kwargs)
    grid = torch.stack(meshgrid_ij(*lin), dim=-1)
    return grid_pull(image, grid, **kwargs)


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
