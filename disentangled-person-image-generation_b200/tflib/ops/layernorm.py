"""lib.ops.layernorm.Layernorm (reference tflib/ops/layernorm.py:6-20): per-sample moments over norm_axes
(= [1,2,3] on BCHW), params `name.offset` / `name.scale` indexed by channel, eps 1e-5."""
import numpy as np

from ... import _lib
from .. import param
from ._norm import norm_nchw


def Layernorm(name, norm_axes, inputs):
    if list(norm_axes) != [1, 2, 3]:
        raise Exception("Layernorm over non-standard axes is unsupported")
    n_neurons = inputs.shape[norm_axes[0]]
    offset = param(name + ".offset", np.zeros(n_neurons, np.float32))
    scale = param(name + ".scale", np.ones(n_neurons, np.float32))
    return norm_nchw(inputs, _lib.NORM_LAYER, scale, offset)
