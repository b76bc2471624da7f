"""lib.ops.conv2d.Conv2D (reference tflib/ops/conv2d.py:20-123): NCHW in / NCHW out, SAME padding, HWIO
filter `name.Filters`, bias `name.Biases`, uniform(+-stdev*sqrt(3)) init with the He / Glorot stdev or the
global override set by set_weights_stdev (wgan_gp.py:411-413)."""
import ctypes as C

import numpy as np
import torch

from ... import _lib
from ...tensor import ptr
from .. import context, param, stream
from ._common import nchw_to_split, nhwc_to_nchw

_weights_stdev = None


def set_weights_stdev(weights_stdev):
    global _weights_stdev
    _weights_stdev = weights_stdev


def unset_weights_stdev():
    global _weights_stdev
    _weights_stdev = None


def Conv2D(name, input_dim, output_dim, filter_size, inputs, he_init=True, mask_type=None, stride=1,
           weightnorm=None, biases=True, gain=1.):
    if mask_type is not None or weightnorm:
        raise Exception("mask_type / weightnorm are not on the reference's live path and are not implemented")
    fan_in = input_dim * filter_size ** 2
    fan_out = output_dim * filter_size ** 2 / (stride ** 2)
    stdev = _weights_stdev if _weights_stdev is not None else (np.sqrt(4. / (fan_in + fan_out)) if he_init
                                                               else np.sqrt(2. / (fan_in + fan_out)))
    shape = (filter_size, filter_size, input_dim, output_dim)
    w = param(name + ".Filters", lambda_init(lambda: np.random.uniform(-stdev * np.sqrt(3), stdev * np.sqrt(3), shape) * gain, name + ".Filters"))
    b = param(name + ".Biases", np.zeros(output_dim, np.float32)) if biases else None
    ctx = context()
    x = nchw_to_split(inputs)
    taps = filter_size * filter_size
    cout_pad = (output_dim + 7) // 8 * 8
    wf = torch.zeros((2, taps, output_dim, x.c), dtype=torch.bfloat16, device=inputs.device)
    ctx.weight_pack(ptr(w), taps, input_dim, output_dim, x.c, cout_pad, ptr(wf[0]), ptr(wf[1]), None, None, stream())
    n, _, h, wd = inputs.shape
    oh, ow = -(-h // stride), -(-wd // stride)
    out = torch.empty((n, oh, ow, output_dim), device=inputs.device)
    ep = _lib.ConvEpilogue()
    ep.bias = b.data_ptr() if b is not None else None
    ep.act = _lib.ACT_NONE
    ep.out_f32 = out.data_ptr()
    ep.out_f32_pix_stride = output_dim
    ep.upsample = 1
    ctx.conv2d_fwd(x.ref(), ptr(wf[0]), ptr(wf[1]), filter_size, filter_size, stride, output_dim, C.byref(ep), stream())
    return nhwc_to_nchw(out)


def lambda_init(fn, name):
    """Evaluates the initialiser only when the parameter does not exist yet."""
    from .. import _params
    return None if name in _params else fn()
