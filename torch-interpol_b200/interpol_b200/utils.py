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
    """Shape that all `shapes` broadcast to (singleton axes stretch, missing axes are added on `side`);
    ValueError when two sizes other than 1 disagree.  Same contract as interpol/utils.py:38-78."""
    rank = max(map(len, shapes), default=0)

    def aligned(shape):
        fill = (1,) * (rank - len(shape))
        return fill + tuple(shape) if side == 'left' else tuple(shape) + fill

    out = []
    for sizes in zip(*map(aligned, shapes)):
        wide = sorted({int(n) for n in sizes if n != 1})
        if len(wide) > 1:
            raise ValueError('Incompatible shapes for broadcasting: {} and {}.'.format(wide[0], wide[1]))
        out.append(wide[0] if wide else 1)
    return tuple(out)


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
