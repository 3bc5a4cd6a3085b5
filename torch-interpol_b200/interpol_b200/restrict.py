"""`restrict`: adjoint of `resize`, through `grid_push`
(reference: interpol/restrict.py:9-122)."""
import torch

from .api import grid_push
from .utils import make_list, meshgrid_ij

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
    grid = torch.stack(meshgrid_ij(*lin), dim=-1)
    resized = grid_push(image, grid, shape, **kwargs)
    if not reduce_sum:
        resized /= fullscale
    return resized
