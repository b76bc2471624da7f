"""Layout helpers of the eager op mirrors: the reference's tflib ops take NCHW tensors
(tflib/ops/conv2d.py:106-112), the kernels take NHWC split-bf16."""
import torch

from ...tensor import SplitTensor


def nchw_to_split(x, cpad=None):
    x = x.permute(0, 2, 3, 1).contiguous().float()
    c = x.shape[-1]
    cpad = cpad or (c + 7) // 8 * 8
    if cpad != c:
        x = torch.cat([x, torch.zeros(x.shape[:-1] + (cpad - c,), device=x.device)], dim=-1)
    return SplitTensor.from_float(x)


def nhwc_to_nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()
