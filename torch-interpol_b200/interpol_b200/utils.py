"""Small host-side helpers (shape algebra).  Reference: interpol/utils.py."""
import torch


def make_list(x, n=None, **kwargs):
    """Ensure `x` is a list, right-padded to length `n` with its last value
    (or `default`).  Reference: interpol/utils.py:11-35."""
    if not isinstance(x, (list, tuple)):
        x = [x]
    x = list(x)
    if n and len(x) < n:
        default = kwargs.get('default', x[-1])
        x = x + [default] * max(0, n - len(x))
    return x


def expanded_shape(*shapes, side='left'):
    """Broadcast shapes; raises ValueError when incompatible.
    Reference: interpol/utils.py:38-78."""
    nb_dim = max([len(s) for s in shapes], default=0)
    shape = [1] * nb_dim
    for shape1 in shapes:
        pad = [1] * (nb_dim - len(shape1))
        shape1 = [*pad, *shape1] if side == 'left' else [*shape1, *pad]
        new = []
        for s0, s1 in zip(shape, shape1):
            if s0 != 1 and s1 != 1 and s0 != s1:
                raise ValueError('Incompatible shapes for broadcasting: {} and {}.'
                                 .format(s0, s1))
            new.append(max(s0, s1) if (s0 != 0 and s1 != 0) else 0)
        shape = new
    return tuple(shape)


def matvec(mat, vec, out=None):
    """Matrix-vector product supporting broadcasting: (..., M, N) x (..., N) -> (..., M).
    Reference: interpol/utils.py:81-109."""
    mv = torch.matmul(mat, vec.unsqueeze(-1)).squeeze(-1)
    if out is not None:
        out.copy_(mv)
        return out
    return mv


def meshgrid_ij(*x):
    return torch.meshgrid(*x, indexing='ij')
