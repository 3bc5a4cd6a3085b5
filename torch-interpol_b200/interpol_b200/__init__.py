"""interpol_b200 -- B200-native high-order spline sampling.

Drop-in for the hot path of balbasty/torch-interpol: `grid_pull`, `grid_push`,
`grid_count`, `grid_grad`, `spline_coeff`, `spline_coeff_nd` and their
`torch.autograd.Function` surface, running hand-written sm_100a CUDA kernels
behind a C ABI (include/interpol_b200.h).  No CPU fallback.
"""
from .api import *          # noqa: F401,F403
from .resize import *       # noqa: F401,F403
from .restrict import *     # noqa: F401,F403
from .autograd import (GridPull, GridPush, GridCount, GridGrad,   # noqa: F401
                       SplineCoeff, SplineCoeffND, bound_to_nitorch, inter_to_nitorch)
from . import backend, pushpull, coeff, bounds, splines, distributed  # noqa: F401
from ._lib import launch_count, last_kernel, LIB_PATH  # noqa: F401

__version__ = '0.1.0'


def install_as_backend(interpol_module=None):
    """Plug this engine into an *installed* reference package through its own
    plugin seam (interpol/backend.py:1 + interpol/jitfields.py:47-114): the six
    API entry points, `resize` and `restrict` of the reference then forward here.
    See INTEGRATION.md."""
    import sys
    import types
    if interpol_module is None:
        import interpol as interpol_module    # the reference
    me = sys.modules[__name__]
    shim = types.ModuleType('interpol_b200_seam')
    shim.available = True
    for name in ('grid_pull', 'grid_push', 'grid_count', 'grid_grad',
                 'spline_coeff', 'spline_coeff_nd', 'resize', 'restrict'):
        setattr(shim, name, getattr(me, name))
    for sub in ('api', 'resize', 'restrict'):
        mod = sys.modules.get(interpol_module.__name__ + '.' + sub)
        if mod is not None and hasattr(mod, 'jitfields'):
            mod.jitfields = shim
    interpol_module.backend.jitfields = True
    return shim
