"""``ptu.soft_argmax`` on the device (/root/reference/metrabs_pytorch/ptu.py:47-75).

Supported calls are the two the path makes (models/metrabs.py:80,82): ``soft_argmax(x[B,D,J,H,W], dim=(4,3,1))``
and ``soft_argmax(x[B,J,H,W], dim=(3,2))``; both run the single-pass kernel in csrc/decode.cuh."""
from metrabs_b200 import _lib
from metrabs_b200.engine import soft_argmax_device


def soft_argmax(inp, dim):
    dim = tuple(d if d >= 0 else inp.ndim + d for d in dim)
    if inp.ndim == 5 and dim == (4, 3, 1):
        b, d, j, h, w = inp.shape
        return soft_argmax_device(inp, _lib.LAYOUT_BDJHW, j, d, h, w)[1]
    if inp.ndim == 4 and dim == (3, 2):
        b, j, h, w = inp.shape
        return soft_argmax_device(inp, _lib.LAYOUT_BDJHW, j, 0, h, w)[0]
    raise NotImplementedError(f'soft_argmax over dims {dim} of a {inp.ndim}-D tensor is not on the MeTRAbs path')
