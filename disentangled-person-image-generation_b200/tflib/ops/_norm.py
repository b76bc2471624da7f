import torch

from ... import _lib
from ...tensor import SplitTensor, ptr
from .. import context, stream


def norm_nchw(inputs, mode, scale, offset, eps=1e-5):
    """Shared body of Layernorm / Batchnorm on an NCHW tensor: stats -> normalise (no activation) -> NCHW fp32."""
    ctx = context()
    x = inputs.permute(0, 2, 3, 1).contiguous().float()
    n, h, w, c = x.shape
    groups = n if mode == _lib.NORM_LAYER else c
    count = float(h * w * c) if mode == _lib.NORM_LAYER else float(n * h * w)
    sums = torch.zeros((2, groups), dtype=torch.float64, device=x.device)
    stats = torch.zeros((2, groups), device=x.device)
    out = SplitTensor(n, h, w, c, x.device)
    ctx.norm_stats(ptr(x), n, h, w, c, mode, ptr(sums), stream())
    ctx.norm_act_fwd(ptr(x), n, h, w, c, mode, eps, ptr(sums), count, ptr(scale), ptr(offset), _lib.ACT_NONE, 0.0,
                     ptr(stats), out.ref(), None, stream())
    return out.float().permute(0, 3, 1, 2).contiguous()
