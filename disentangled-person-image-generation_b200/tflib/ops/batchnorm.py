"""lib.ops.batchnorm.Batchnorm (reference tflib/ops/batchnorm.py:6-74), fused path with axes [0,2,3]:
tf.nn.fused_batch_norm in TRAINING mode -- batch statistics, always (is_training=None at every live call site);
`name.moving_mean/.moving_variance` are created for checkpoint compatibility and never updated (quirk q4)."""
import numpy as np

from ... import _lib
from .. import param
from ._norm import norm_nchw


def Batchnorm(name, axes, inputs, is_training=None, stats_iter=None, update_moving_stats=True, fused=True):
    if list(axes) != [0, 2, 3] or not fused or is_training is not None:
        raise Exception("only the fused training-mode BatchNorm over [0,2,3] is on the reference's live path")
    c = inputs.shape[1]
    offset = param(name + ".offset", np.zeros(c, np.float32))
    scale = param(name + ".scale", np.ones(c, np.float32))
    param(name + ".moving_mean", np.zeros(c, np.float32), trainable=False)
    param(name + ".moving_variance", np.ones(c, np.float32), trainable=False)
    return norm_nchw(inputs, _lib.NORM_BATCH, scale, offset)
