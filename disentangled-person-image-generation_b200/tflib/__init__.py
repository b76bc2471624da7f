"""Mirror of the reference's `tflib` package (tflib/__init__.py:8-48): a process-global, name-keyed parameter
registry -- calling an op twice with the same name SHARES its weights (how the reference applies one critic to
x, G and x_hat) -- plus the op modules `tflib.ops.{conv2d,linear,layernorm,batchnorm}`.

Here parameters are fp32 CUDA tensors (device memory only) and the ops run EAGERLY through the C ABI
(libdpig.so); they exist so reference-style code and tests can call the kernels op by op.  The training hot
path does not go through this module (engine.py replays static launch programs instead)."""
import numpy as np
import torch

_params = {}
_ctx = None


def context():
    global _ctx
    if _ctx is None:
        from .. import _lib
        _ctx = _lib.Context(torch.cuda.current_device())
    return _ctx


def stream():
    return torch.cuda.current_stream().cuda_stream


def param(name, value=None, **kwargs):
    """Returns the parameter called `name`, creating it from `value` on first use (tflib/__init__.py:10-31)."""
    if name not in _params:
        if value is None:
            raise KeyError(name)
        t = torch.as_tensor(np.asarray(value), dtype=torch.float32).cuda().contiguous()
        t.trainable = kwargs.get("trainable", True)
        _params[name] = t
    return _params[name]


def params_with_name(name):
    return [p for n, p in _params.items() if name in n]


def param_names():
    return list(_params)


def delete_all_params():
    _params.clear()


def print_model_settings(locals_):
    for k, v in sorted(locals_.items()):
        if k.isupper():
            print("\t{}: {}".format(k, v))
