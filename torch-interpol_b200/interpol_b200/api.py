"""High-level API: same functions, keyword arguments and defaults as the
reference's `interpol/api.py` (grid_pull :149, grid_push :215, grid_count :265,
grid_grad :302, spline_coeff :347, spline_coeff_nd :398, grid helpers :467-572).

Shapes follow the reference: volumes `(..., [channel], *spatial)`, grids
`(..., *spatial, dim)` in voxel units.  Tensors on a CUDA device are processed
in place on that device; CPU tensors are staged to the current CUDA device
(asynchronously when pinned), processed by the same kernels and returned on
the CPU -- there is no CPU implementation.
"""
import threading

import torch

from .utils import expanded_shape, matvec, meshgrid_ij
from . import pushpull as _pp
from .autograd import (GridPull, GridPush, GridCount, GridGrad,
                       SplineCoeff, SplineCoeffND)

__all__ = [
    'pull', 'push', 'count', 'stage_scope',
    'grid_pull', 'grid_push', 'grid_count', 'grid_grad',
    'spline_coeff', 'spline_coeff_nd',
    'identity_grid', 'add_identity_grid', 'add_identity_grid_', 'affine_grid',
]


# --------------------------------------------------------------------------
# host <-> device staging
# --------------------------------------------------------------------------

# Page-locking a fresh 64 MB buffer costs ~11 ms (cudaHostAlloc; measured, profiles/xfer_rates.py) -- ten
# times the copy it serves -- so result buffers come from a small pool and are handed out again once the
# caller has dropped every tensor that views them (storage use count back to the pool's own references).
# Buffers are reused for results of exactly the same byte size only, so a result never views a storage
# larger than itself (torch.save / pickling serialise whole storages).  The pool is capped
# (IB200_PINNED_POOL_MB, default 1024; 0 disables it) and switches itself off if torch's storage use
# count does not behave as the reuse test assumes.
import os as _os

_POOL, _POOL_LOCK = [], threading.Lock()
_POOL_MAX_BYTES = int(_os.environ.get('IB200_PINNED_POOL_MB', '1024')) << 20
_POOL_STATE = {'checked': False, 'ok': False, 'idle': 0}


def _use_count_fn():
    """torch._C._storage_Use_Count, after a self-test of the semantics the pool relies on: a view adds one
    reference, dropping it gives the reference back."""
    st = _POOL_STATE
    if not st['checked']:
        st['checked'] = True
        fn = getattr(torch._C, '_storage_Use_Count', None)
        try:
            probe = torch.empty(16, dtype=torch.uint8)
            idle = fn(probe.untyped_storage()._cdata)
            view = probe[:8]
            held = fn(probe.untyped_storage()._cdata)
            del view
            st['ok'] = held == idle + 1 and fn(probe.untyped_storage()._cdata) == idle
            st['idle'] = idle
        except Exception:
            st['ok'] = False
    return getattr(torch._C, '_storage_Use_Count') if st['ok'] else None


def _pinned_empty(shape, dtype):
    nbytes = 1
    for n in shape:
        nbytes *= int(n)
    nbytes *= torch.empty(0, dtype=dtype).element_size()
    use_count = _use_count_fn() if _POOL_MAX_BYTES > 0 else None
    if use_count is None or nbytes == 0 or nbytes > _POOL_MAX_BYTES:
        return torch.empty(shape, dtype=dtype, pin_memory=True)
    with _POOL_LOCK:
        best = None
        for buf in _POOL:
            if buf.numel() == nbytes and use_count(buf.untyped_storage()._cdata) <= _POOL_STATE['idle']:
                best = buf
                break
        if best is None:
            total = sum(b.numel() for b in _POOL)
            while _POOL and total + nbytes > _POOL_MAX_BYTES:
                total -= _POOL.pop(0).numel()
            best = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            _POOL.append(best)
        return best.view(dtype).view(list(shape))


_SCOPES = threading.local()


def _host_key(t):
    return (t.untyped_storage().data_ptr(), t.storage_offset(), tuple(t.shape), tuple(t.stride()), t.dtype, t._version)


class stage_scope:
    """Context manager for chains of calls on CPU tensors (`with interpol_b200.stage_scope(): ...`).

    Inside the scope every distinct CPU tensor is uploaded ONCE (keyed on storage address, view
    geometry and torch's version counter, so an in-place torch update re-uploads), and a CPU result
    keeps its device twin: feeding it to the next call costs no transfer.  A registration step
    (pull, then push back through the same grid) therefore moves volume + grid up and the two
    results down, instead of uploading the grid twice and the pulled image once more.  The scope
    keeps the host tensors it has seen alive until it exits; writes through aliases torch does not
    track (numpy views) are not seen -- leave the scope, or call `.clear()`, after such writes.
    """

    def __init__(self):
        self.twins = {}

    def clear(self):
        self.twins.clear()

    def __enter__(self):
        stack = getattr(_SCOPES, 'stack', None)
        if stack is None:
            stack = _SCOPES.stack = []
        stack.append(self)
        return self

    def __exit__(self, *exc):
        _SCOPES.stack.pop()
        self.clear()
        return False

    def lookup(self, host, dev):
        hit = self.twins.get(_host_key(host))
        if hit is not None and hit[1].device == dev:
            return hit[1]
        return None

    def remember(self, host, device_tensor):
        self.twins[_host_key(host)] = (host, device_tensor)


def _scope():
    stack = getattr(_SCOPES, 'stack', None)
    return stack[-1] if stack else None


def _stage(*tensors):
    """Move CPU tensors to the current CUDA device.  Returns (tensors, back)
    where `back(out)` returns the result to where the inputs lived."""
    if all(t is None or t.is_cuda for t in tensors):
        return tensors, (lambda out: out)
    if not torch.cuda.is_available():
        raise RuntimeError('interpol_b200 needs a CUDA device (no CPU fallback)')
    dev = None
    for t in tensors:
        if t is not None and t.is_cuda:
            dev = t.device
    if dev is None:
        dev = torch.device('cuda', torch.cuda.current_device())
    scope = _scope()

    def up(t):
        if t is None or t.is_cuda:
            return t
        if scope is not None and not t.requires_grad:
            twin = scope.lookup(t, dev)
            if twin is None:
                twin = t.to(dev, non_blocking=True)
                scope.remember(t, twin)
            return twin
        return t.to(dev, non_blocking=True)

    staged = tuple(up(t) for t in tensors)

    def back(out):
        if out.requires_grad or out.numel() == 0:
            return out.cpu()
        # page-locked result: the device->host copy runs at PCIe speed, and a result fed back into the
        # next call uploads at PCIe speed as well
        host = _pinned_empty(out.shape, out.dtype)
        host.copy_(out, non_blocking=True)
        torch.cuda.current_stream(out.device).synchronize()
        if scope is not None:
            scope.remember(host, out)
        return host

    return staged, back


# --------------------------------------------------------------------------
# streamed sampling of host tensors: upload / gather / download overlap slab by slab
# --------------------------------------------------------------------------

# A pull (or grad) of HOST tensors is transfer-bound: at 256^3 the grid alone is 201 MB (3.7 ms over PCIe)
# against 0.34 ms of kernel time, and the result (67 MB, 1.3 ms) used to wait for the kernel, which waited
# for the whole upload.  Output voxels are independent, so the lattice is cut into slabs along its first
# axis: slab k is gathered while slab k+1 uploads and slab k-1 downloads (PCIe is full duplex), on two side
# streams.  The volume goes up whole first (every slab may read any of it).
STREAM_MIN_BYTES = 16 << 20      # smaller grids: one upload + one launch is faster
STREAM_SLAB_BYTES = 48 << 20     # target slab size (at least 2, at most 8 slabs per batch element)
_SIDE_STREAMS = {}


def _side_streams(dev):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
    return _SIDE_STREAMS[key]


def _streamable(input, grid, displacement):
    if input.is_cuda or grid.is_cuda or displacement or not torch.cuda.is_available():
        return False
    if not (input.dtype.is_floating_point and grid.dtype == input.dtype):
        return False
    if input.requires_grad or grid.requires_grad or not grid.is_contiguous():
        return False
    dim = grid.shape[-1]
    if grid.dim() < dim + 1 or input.dim() < dim or grid.shape[-dim - 1] < 16:
        return False
    if grid.numel() * grid.element_size() < STREAM_MIN_BYTES:
        return False
    scope = _scope()
    if scope is not None and scope.lookup(grid, torch.device('cuda', torch.cuda.current_device())) is not None:
        return False                      # the grid is on the device already: nothing to overlap
    return True


def _sample_streamed(fn, mode, input, grid, interpolation, bound, extrapolate, prefilter):
    """grid_pull / grid_grad of host tensors, slab by slab (see above).  Returns None when the shapes do not
    allow it (the caller then takes the plain path)."""
    dev = torch.device('cuda', torch.cuda.current_device())
    dim = grid.shape[-1]
    host = _Layout(grid, input)
    grid_c = host.grid
    if grid_c.numel() != grid.numel() or not grid_c.is_contiguous():
        return None                       # broadcast batch: the plain path uploads the grid once
    (input_d,), _ = _stage(input)         # whole volume, through the scope's twin cache
    main = torch.cuda.current_stream(dev)
    up, down = _side_streams(dev)
    grid_full = torch.empty(grid.shape, dtype=grid.dtype, device=dev)
    dev_layout = _Layout(grid_full, input_d)
    grid_d, input_d = dev_layout.grid, dev_layout.volume
    if prefilter:
        input_d = spline_coeff_nd(input_d, interpolation=interpolation, bound=bound, dim=dim)
    B, C = input_d.shape[:2]
    out_shape = [B, C, *grid_d.shape[1:-1]] + ([dim] if mode == 'grad' else [])
    out_d = torch.empty(out_shape, dtype=input_d.dtype, device=dev)
    out_h = _pinned_empty(out_shape, input_d.dtype)
    nx = grid_d.shape[1]
    per_batch = grid_c[0].numel() * grid_c.element_size()
    nslab = max(2, min(8, -(-per_batch // STREAM_SLAB_BYTES)))
    step = max(8, -(-(-(-nx // nslab)) // 8) * 8)
    up.wait_stream(main)                  # the buffers above may reuse memory that `main` still reads
    grid_full.record_stream(up)
    out_d.record_stream(down)
    for b in range(B):
        for x0 in range(0, nx, step):
            x1 = min(nx, x0 + step)
            with torch.cuda.stream(up):
                grid_d[b, x0:x1].copy_(grid_c[b, x0:x1], non_blocking=True)
                arrived = up.record_event()
            main.wait_event(arrived)
            out_d[b:b + 1, :, x0:x1] = fn.apply(input_d[b:b + 1], grid_d[b:b + 1, x0:x1], interpolation, bound, extrapolate, False)
            computed = main.record_event()
            with torch.cuda.stream(down):
                down.wait_event(computed)
                for c in range(C):
                    out_h[b, c, x0:x1].copy_(out_d[b, c, x0:x1], non_blocking=True)
    down.synchronize()
    result = host.restore(out_h)
    scope = _scope()
    if scope is not None:
        scope.remember(grid, grid_full)
        scope.remember(result, host.restore(out_d))
    return result


# --------------------------------------------------------------------------
# shape canonicalisation
# --------------------------------------------------------------------------

class _Layout:
    """Canonical views of the tensors of one call and the way back.

    The kernels want volumes as (B, C, *spatial) and grids as (B, *spatial, D); the public functions accept
    any leading batch axes (broadcast between volume and grid, as in interpol/api.py:93-146), an optional
    channel axis, or no leading axes at all.  Broadcasting is done with `expand` (zero strides, which the
    kernels honour): nothing is copied unless `reshape` has to merge axes that cannot be merged.
    """

    def __init__(self, grid, volume=None, splat=False):
        self.dim = dim = grid.shape[-1]
        self.canonical = False
        if volume is not None and grid.dim() == dim + 2 and volume.dim() == dim + 2 and grid.shape[0] == volume.shape[0] \
                and (not splat or grid.shape[1:-1] == volume.shape[2:]):
            # already (B, C, *spatial) / (B, *spatial, D) with one batch: nothing to expand, reshape or restore
            # (15 us of view construction per call otherwise -- it shows at 64^3)
            self.canonical, self.grid, self.volume = True, grid, volume
            return
        lattice, lead_g = tuple(grid.shape[-dim - 1:-1]), tuple(grid.shape[:-dim - 1])
        if volume is None:                           # count: the output has one channel iff there is a batch
            self.lead, self.channels = lead_g, ([1] if lead_g else [])
            self.grid, self.volume = grid.reshape([-1, *lattice, dim]), None
            return
        vshape = tuple(volume.shape[-dim:])
        nchan = volume.shape[-dim - 1] if volume.dim() > dim else 0      # 0: no channel axis
        if splat:                                    # push: the image lives on the grid's lattice
            lattice = vshape = expanded_shape(lattice, vshape)
        self.lead = lead = expanded_shape(lead_g, tuple(volume.shape[:-dim - 1]))
        self.channels = [nchan] if nchan else ([1] if lead else [])
        self.grid = grid.expand([*lead, *lattice, dim]).reshape([-1, *lattice, dim])
        self.volume = volume.expand([*lead, nchan or 1, *vshape]).reshape([-1, nchan or 1, *vshape])

    def restore(self, out, features=0):
        """(B, C, *spatial[, features]) back to the caller's leading axes."""
        if self.canonical:
            return out
        tail = out.shape[2:]
        return out.reshape([*self.lead, *self.channels, *tail])


def _apply(function, kernel, tensors, shape, interpolation, bound, extrapolate, displacement):
    """`function.apply(...)`, or the kernel behind it called directly when neither autograd (no tensor requires a
    gradient, or grad mode is off) nor autocast has anything to do: the Function machinery costs ~12 us per call."""
    needs_graph = torch.is_grad_enabled() and any(t.requires_grad for t in tensors)
    if needs_graph or torch.is_autocast_enabled():
        return function.apply(*tensors, *shape, interpolation, bound, extrapolate, displacement)
    from .autograd import _options
    return kernel(*tensors, *shape, *_options(interpolation, bound, extrapolate), bool(displacement))


LABELS_FUSED = True      # False: always loop over the labels like the reference (A/B testing)


def _labels_fused_ok(input, grid, interpolation):
    """The fused label kernel covers orders 0 / 1 (no prefilter involved), float32 / float64 grids and the integer
    storage types it reads natively (int8 maps are widened to int16)."""
    if not LABELS_FUSED or grid.dtype not in (torch.float32, torch.float64) or input.numel() == 0:
        return False
    from .autograd import inter_to_nitorch, make_list
    if any(o > 1 for o in inter_to_nitorch(make_list(interpolation), as_type='int')):
        return False
    return input.dtype in (torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64)


def _pull_labels_by_mask(labels, grid, interpolation, bound, extrapolate, prefilter, displacement):
    """Label maps beyond what the fused kernel covers (orders >= 2, prefiltered masks): every label present is
    resampled as a soft mask and each point keeps the label whose mask is largest -- ties and all-zero points go
    to the label met first in ascending order, and a point no mask reaches stays 0 (interpol/api.py:194-205)."""
    lattice = grid.shape[1:-1]
    best = grid.new_zeros([*labels.shape[:2], *lattice])
    winner = labels.new_zeros(best.shape)
    for value in labels.unique():
        mask = (labels == value).to(grid.dtype)
        if prefilter:
            mask = spline_coeff_nd(mask, interpolation=interpolation, bound=bound, dim=grid.shape[-1], inplace=True)
        mask = GridPull.apply(mask, grid, interpolation, bound, extrapolate, displacement)
        winner[mask > best] = value
        best = torch.maximum(best, mask)
    return winner


# --------------------------------------------------------------------------
# public functions
# --------------------------------------------------------------------------

def grid_pull(input, grid, interpolation='linear', bound='zero',
              extrapolate=False, prefilter=False, *, displacement=False):
    """Sample an image with respect to a deformation field.

    input : (..., [channel], *inshape) tensor;  grid : (..., *outshape, dim) tensor
    interpolation : int | str | list, default 'linear' (orders 0..7)
    bound : str | int | list, default 'zero' (zero, replicate, dct1, dct2, dst1, dst2, dft)
    extrapolate : bool | int, default False;  prefilter : bool, default False
    displacement : bool, default False (extension, keyword only; also on grid_push / grid_count / grid_grad):
        `grid` holds displacements in voxels -- the coordinate of lattice point x is x + grid[x], formed in
        registers: same result as `grid_pull(input, add_identity_grid(grid), ...)` (api.py:482-520) without
        reading or writing the identity grid
    returns (..., [channel], *outshape)

    Integer inputs are treated as label maps: every label is resampled as a
    soft mask and the arg-max label is returned (reference: api.py:194-205).
    """
    if _streamable(input, grid, displacement):
        out = _sample_streamed(GridPull, 'pull', input, grid, interpolation, bound, extrapolate, prefilter)
        if out is not None:
            return out
    (input, grid), back = _stage(input, grid)
    layout = _Layout(grid, input)
    grid, input = layout.grid, layout.volume
    dim = layout.dim

    if input.dtype.is_floating_point:
        if prefilter:
            input = spline_coeff_nd(input, interpolation=interpolation, bound=bound, dim=dim)
        out = _apply(GridPull, _pp.grid_pull, (input, grid), (), interpolation, bound, extrapolate, displacement)
    elif _labels_fused_ok(input, grid, interpolation):
        # one pass: every point looks for the arg-max among the labels of its own (order+1)^dim nodes
        from .autograd import _options
        bnd, order, extr = _options(interpolation, bound, extrapolate)
        stored = input.to(torch.int16) if input.dtype == torch.int8 else input
        out = _pp.grid_pull_labels(stored, grid, bnd, order, extr, displacement).to(input.dtype)
    else:
        out = _pull_labels_by_mask(input, grid, interpolation, bound, extrapolate, prefilter, displacement)
    return back(layout.restore(out))


def grid_push(input, grid, shape=None, interpolation='linear', bound='zero',
              extrapolate=False, prefilter=False, *, displacement=False):
    """Splat an image with respect to a deformation field (adjoint of pull).

    input : (..., [channel], *inshape);  grid : (..., *inshape, dim)
    shape : output spatial shape, default inshape
    returns (..., [channel], *shape)        (reference: api.py:215-262)
    """
    (input, grid), back = _stage(input, grid)
    layout = _Layout(grid, input, splat=True)
    if shape is None:
        shape = tuple(layout.volume.shape[2:])
    out = _apply(GridPush, _pp.grid_push, (layout.volume, layout.grid), (shape,), interpolation, bound, extrapolate, displacement)
    if prefilter:
        out = spline_coeff_nd(out, interpolation=interpolation, bound=bound, dim=layout.dim, inplace=True)
    return back(layout.restore(out))


def grid_count(grid, shape=None, interpolation='linear', bound='zero',
               extrapolate=False, *, displacement=False):
    """Splatting weights of a deformation field (push of an image of ones).

    grid : (..., *inshape, dim);  returns (..., [1], *shape)   (reference: api.py:265-299)
    """
    (grid,), back = _stage(grid)
    layout = _Layout(grid)
    out = _apply(GridCount, _pp.grid_count, (layout.grid,), (shape,), interpolation, bound, extrapolate, displacement)
    return back(layout.restore(out))


def grid_grad(input, grid, interpolation='linear', bound='zero',
              extrapolate=False, prefilter=False, *, displacement=False):
    """Sample the spatial gradients of an image with respect to a deformation field.

    returns (..., [channel], *outshape, dim)                  (reference: api.py:302-344)
    """
    if _streamable(input, grid, displacement):
        out = _sample_streamed(GridGrad, 'grad', input, grid, interpolation, bound, extrapolate, prefilter)
        if out is not None:
            return out
    (input, grid), back = _stage(input, grid)
    layout = _Layout(grid, input)
    input = layout.volume
    if prefilter:
        input = spline_coeff_nd(input, interpolation, bound, layout.dim)
    out = _apply(GridGrad, _pp.grid_grad, (input, layout.grid), (), interpolation, bound, extrapolate, displacement)
    return back(layout.restore(out))


def spline_coeff(input, interpolation='linear', bound='dct2', dim=-1,
                 inplace=False):
    """Interpolating spline coefficients along one dimension
    (reference: api.py:347-395; only dct1/dct2/dft and their aliases zero/replicate)."""
    if not input.is_cuda:
        (x,), back = _stage(input)
        out = back(SplineCoeff.apply(x, bound, interpolation, dim, False))
        if inplace:
            input.copy_(out)
            return input
        return out
    return SplineCoeff.apply(input, bound, interpolation, dim, inplace)


def spline_coeff_nd(input, interpolation='linear', bound='dct2', dim=None,
                    inplace=False):
    """Interpolating spline coefficients along the last `dim` dimensions
    (reference: api.py:398-445)."""
    if not input.is_cuda:
        (x,), back = _stage(input)
        out = back(SplineCoeffND.apply(x, bound, interpolation, dim, False))
        if inplace:
            input.copy_(out)
            return input
        return out
    return SplineCoeffND.apply(input, bound, interpolation, dim, inplace)


# aliases
pull = grid_pull
push = grid_push
count = grid_count


# --------------------------------------------------------------------------
# grid helpers (reference: api.py:467-572) -- plain torch ops, any device
# --------------------------------------------------------------------------

def identity_grid(shape, dtype=None, device=None):
    """Identity deformation field in voxel units: (*shape, dim)."""
    mesh1d = [torch.arange(float(s), dtype=dtype, device=device) for s in shape]
    return torch.stack(meshgrid_ij(*mesh1d), dim=-1)


def add_identity_grid_(disp):
    """Add the identity grid to a displacement field (..., *spatial, dim), in place."""
    dim = disp.shape[-1]
    spatial = disp.shape[-dim-1:-1]
    mesh1d = [torch.arange(s, dtype=disp.dtype, device=disp.device) for s in spatial]
    for i, g in enumerate(meshgrid_ij(*mesh1d)):
        disp[..., i].add_(g)
    return disp


def add_identity_grid(disp):
    """Add the identity grid to a displacement field (out of place)."""
    return add_identity_grid_(disp.clone())


def affine_grid(mat, shape):
    """Dense transformation grid (..., *shape, D) from affine matrices (..., D[+1], D+1)."""
    mat = torch.as_tensor(mat)
    shape = list(shape)
    nb_dim = mat.shape[-1] - 1
    if nb_dim != len(shape):
        raise ValueError('Dimension of the affine matrix ({}) and shape ({}) '
                         'are not the same.'.format(nb_dim, len(shape)))
    if mat.shape[-2] not in (nb_dim, nb_dim+1):
        raise ValueError('First argument should be matrices of shape '
                         '(..., {0}, {1}) or (..., {1}, {1}) but got {2}.'
                         .format(nb_dim, nb_dim+1, mat.shape))
    grid = identity_grid(shape, mat.dtype, mat.device)
    lin = mat[..., :nb_dim, :nb_dim]
    off = mat[..., :nb_dim, -1]
    # (the reference's batched branch, api.py:560-565, indexes the wrong axes and
    #  raises; the intended broadcast is implemented here)
    for _ in range(nb_dim):
        lin = lin.unsqueeze(-3)
        off = off.unsqueeze(-2)
    return matvec(lin, grid) + off
