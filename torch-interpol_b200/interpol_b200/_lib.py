"""ctypes binding of libinterpol_b200.so (the C ABI of include/interpol_b200.h).

PyTorch is only used for device memory and streams: tensors are passed as raw
device pointers + element strides, work is enqueued on torch's current stream.
There is NO fallback: if the shared library is missing or the tensors do not
live on a CUDA device, the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# (IB200_LIB: another build of the same library, e.g. a variant compiled for an A/B measurement under profiles/)
LIB_PATH = os.environ.get('IB200_LIB') or os.path.join(_HERE, '_C', 'libinterpol_b200.so')

F16, F32, F64, BF16 = 0, 1, 2, 3
DTYPE_CODE = {torch.float16: F16, torch.float32: F32, torch.float64: F64, torch.bfloat16: BF16}

ERR_BOUND_UNSUPPORTED = -7
_VALUE_ERRORS = (-3, -4, -5, -6, -9, -10)

FLAG_NO_TILES = 1
FLAG_REF_LINEAR_GRAD_SIGN = 2

i32, i64, u32 = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32


class Problem(ctypes.Structure):
    """struct ib200_problem (include/interpol_b200.h)"""
    _fields_ = [
        ('dim', i32), ('dtype', i32), ('extrapolate', i32), ('device', i32),
        ('bound', i32 * 3), ('order', i32 * 3),
        ('flags', u32), ('reserved', u32),
        ('batch', i64), ('channels', i64),
        ('vol_shape', i64 * 3), ('pts_shape', i64 * 3),
        ('vol_stride', i64 * 5), ('grid_stride', i64 * 5), ('img_stride', i64 * 6),
    ]


_lib = None


class ExtensionMissing(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ExtensionMissing(
                'interpol_b200: %s not found -- build it with '
                '`python torch-interpol_b200/build.py` (there is no CPU / PyTorch fallback)' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        vp = ctypes.c_void_p
        pp = ctypes.POINTER(Problem)
        for name in ('ib200_pull', 'ib200_grad', 'ib200_hess', 'ib200_pull_labels'):
            getattr(L, name).argtypes = [pp, vp, vp, vp, vp]
        L.ib200_pull_backward_grid.argtypes = [pp, vp, vp, vp, vp, vp]
        L.ib200_grad_backward_grid.argtypes = [pp, vp, vp, vp, vp, vp]
        L.ib200_push.argtypes = [pp, vp, vp, vp, vp, vp]
        L.ib200_pushgrad.argtypes = [pp, vp, vp, vp, vp, vp]
        L.ib200_count.argtypes = [pp, vp, vp, vp, vp]
        L.ib200_scratch_bytes.argtypes = [pp]
        L.ib200_scratch_bytes.restype = ctypes.c_size_t
        L.ib200_spline_coeff.argtypes = [vp, i32, i64, i64, i64, i32, i32, i32, vp]
        L.ib200_resample_axis.argtypes = [vp, vp, vp, i32, i64, i64, i64, i64, i32, i32, i32, i32, i32, i32, vp]
        L.ib200_resample_axis.restype = ctypes.c_int
        L.ib200_resample_axis_adjoint.argtypes = [vp, vp, vp, i32, i64, i64, i64, i64, i32, i32, i32, i32, i32, i32, vp]
        L.ib200_resample_axis_adjoint.restype = ctypes.c_int
        L.ib200_error_string.argtypes = [ctypes.c_int]
        L.ib200_error_string.restype = ctypes.c_char_p
        L.ib200_last_kernel.restype = ctypes.c_char_p
        L.ib200_launch_count.restype = ctypes.c_uint64
        for name in ('ib200_pull', 'ib200_grad', 'ib200_hess', 'ib200_pull_backward_grid', 'ib200_grad_backward_grid', 'ib200_push',
                     'ib200_pushgrad', 'ib200_count', 'ib200_spline_coeff', 'ib200_abi_version', 'ib200_pull_labels'):
            getattr(L, name).restype = ctypes.c_int
        if L.ib200_abi_version() != 1:
            raise ExtensionMissing('interpol_b200: ABI version mismatch, rebuild the extension')
        _lib = L
    return _lib


def check(status):
    """Map a C status code to the exception class the reference raises."""
    if status == 0:
        return
    msg = lib().ib200_error_string(status).decode()
    if status == ERR_BOUND_UNSUPPORTED:
        raise NotImplementedError(msg)          # coeff.py:244,254
    if status in _VALUE_ERRORS:
        raise ValueError(msg)                   # autograd.py:95,145
    raise RuntimeError('interpol_b200: %s (status %d)' % (msg, status))


def launch_count():
    return int(lib().ib200_launch_count())


def last_kernel():
    return lib().ib200_last_kernel().decode()


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('interpol_b200 kernels run on CUDA tensors only (got %s); '
                               'there is no CPU fallback' % t.device)


def stream_ptr(device):
    # (plain ints: the entry points declare `c_void_p` argtypes, ctypes converts)
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t):
    return t.data_ptr() if t is not None else None


class on_device:
    """`with torch.cuda.device(d)` only when `d` is not the current device already (the context manager costs ~4 us)."""

    def __init__(self, device):
        idx = device.index
        self.ctx = None if idx is None or idx == torch.cuda.current_device() else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        return False
