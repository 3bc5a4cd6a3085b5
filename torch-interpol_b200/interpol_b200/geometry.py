"""Host-side geometry of `resize` / `restrict`: which lattice is sampled where.

Both functions relate a fine and a coarse lattice along every axis through one of four anchors
(reference: interpol/resize.py:53-89, interpol/restrict.py:52-84):

    'c' centres of the corner voxels coincide      'e' edges of the field of view coincide
    'f' first voxels coincide, exact factor        'l' last voxels coincide, exact factor

`resize` walks the OUTPUT lattice and reads the input; `restrict` walks the INPUT lattice and splats into the
output.  In both cases one lattice of `n_walk` points is expressed in the index space of the other (`n_ref`
points), so one table of closed forms serves the two."""
import torch

from .utils import make_list


def _axis_coords(anchor, n_ref, n_walk, factor, backend):
    """Coordinates of the `n_walk` lattice points in the index space of the `n_ref`-point lattice."""
    steps = torch.arange(0., n_walk, **backend)
    if anchor == 'c':
        return torch.linspace(0, n_ref - 1, n_walk, **backend)
    if anchor == 'e':
        ratio = n_ref / n_walk
        return steps * ratio + 0.5 * (ratio - 1)
    if anchor == 'f':
        return steps / factor
    if anchor == 'l':
        return steps / factor + ((n_ref - 1) - (n_walk - 1) / factor)
    raise ValueError('Unknown anchor {}'.format(anchor))


def _axis_scale(anchor, n_ref, n_walk, factor):
    """Spacing of the walked lattice in units of the reference lattice (restrict divides its sums by it)."""
    if anchor == 'c':
        return (n_walk - 1) / (n_ref - 1)
    if anchor == 'e':
        return n_ref / n_walk
    return 1 / factor


class SamplingPlan:
    """Resolved arguments of a resize / restrict call.

    ndim, inshape, outshape, factor (per axis), anchor (per axis, one letter), coords (one 1-D tensor per
    axis, in the index space of the lattice that is NOT walked), scale (product over axes, restrict only)."""

    def __init__(self, image, factor, shape, anchor, upsample):
        factor = make_list(factor) if factor else []
        shape = make_list(shape) if shape else []
        anchor = make_list(anchor)
        if not factor and not shape:
            raise ValueError('One of `factor` or `shape` must be provided')
        self.ndim = ndim = max(len(factor), len(shape), len(anchor)) or (image.dim() - 2)
        self.anchor = [a[0].lower() for a in make_list(anchor, ndim)]
        self.inshape = inshape = tuple(image.shape[-ndim:])
        factor = make_list(factor, ndim) if factor else None
        if shape:
            outshape = [int(s) for s in make_list(shape, ndim)]
        elif upsample:
            outshape = [int(n * f) for n, f in zip(inshape, factor)]
        else:
            outshape = [int(n / f) for n, f in zip(inshape, factor)]
        if factor is None:
            factor = [(o / i) if upsample else (i / o) for i, o in zip(inshape, outshape)]
        self.factor, self.outshape = factor, outshape
        backend = dict(dtype=image.dtype, device=image.device)
        walk, ref = (outshape, inshape) if upsample else (inshape, outshape)
        self.coords = [_axis_coords(a, r, w, f, backend) for a, r, w, f in zip(self.anchor, ref, walk, factor)]
        self.scale = 1
        if not upsample:
            for a, r, w, f in zip(self.anchor, ref, walk, factor):
                self.scale *= _axis_scale(a, r, w, f)
