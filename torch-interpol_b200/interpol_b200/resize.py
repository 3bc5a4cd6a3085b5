"""`resize`: change the sampling density of an image through `grid_pull`
(reference: interpol/resize.py:13-119; same arguments and anchor conventions)."""
import torch

from .api import grid_pull
from .utils import make_list, meshgrid_ij

__all__ = ['resize']


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
    grid = torch.stack(meshgrid_ij(*lin), dim=-1)
    return grid_pull(image, grid, **kwargs)
