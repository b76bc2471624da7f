"""lib.ops.linear.Linear (reference tflib/ops/linear.py:24-147): y = x W + b with W `name.W` [in,out],
b `name.b`; initialisations None/'glorot', 'he', 'lecun', 'glorot_he', ('uniform', r)."""
import numpy as np
import torch

from ... import _lib
from ...tensor import ptr
from .. import _params, context, param, stream

_weights_stdev = None


def set_weights_stdev(weights_stdev):
    global _weights_stdev
    _weights_stdev = weights_stdev


def unset_weights_stdev():
    global _weights_stdev
    _weights_stdev = None


def Linear(name, input_dim, output_dim, inputs, biases=True, initialization=None, weightnorm=None, gain=1.):
    if weightnorm:
        raise Exception("weightnorm is not on the reference's live path and is not implemented")

    def uniform(stdev):
        if _weights_stdev is not None:
            stdev = _weights_stdev
        return np.random.uniform(-stdev * np.sqrt(3), stdev * np.sqrt(3), (input_dim, output_dim)).astype("float32") * gain

    if name + ".W" not in _params:
        if initialization in (None, "glorot"):
            val = uniform(np.sqrt(2. / (input_dim + output_dim)))
        elif initialization == "he":
            val = uniform(np.sqrt(2. / input_dim))
        elif initialization == "lecun":
            val = uniform(np.sqrt(1. / input_dim))
        elif initialization == "glorot_he":
            val = uniform(np.sqrt(4. / (input_dim + output_dim)))
        elif initialization[0] == "uniform":
            val = np.random.uniform(-initialization[1], initialization[1], (input_dim, output_dim)).astype("float32")
        else:
            raise Exception("Invalid initialization!")
        param(name + ".W", val)
    w = param(name + ".W")
    b = param(name + ".b", np.zeros((output_dim,), np.float32)) if biases else None
    x = inputs.reshape(-1, input_dim).contiguous().float()
    y = torch.empty((x.shape[0], output_dim), device=x.device)
    context().linear_fwd(ptr(x), ptr(w), ptr(b), ptr(y), x.shape[0], input_dim, output_dim, _lib.ACT_NONE, 0.0, stream())
    return y.reshape(tuple(inputs.shape[:-1]) + (output_dim,))
