"""Known-answer vectors restated from TensorFlow's OWN published unit tests (r1.4 source tree), so that the oracle
(oracle/tf_ops.py) is pinned against numbers that were produced by -- and are asserted inside -- TensorFlow itself,
not by this repository.

HOW THIS PIN IS DERIVED (read before trusting it): TensorFlow 1.4.1 cannot be installed or run here and its source is
not under /root/reference, so nothing below comes from a TensorFlow run in this environment.  Each entry restates the
inputs and the expected outputs that a named test of TensorFlow's test-suite hard-codes, with the file and test name,
as known from the public r1.4 sources.  Every expected array was ALSO re-derived here with exact rational arithmetic
(tests/test_tf_known_answers.py::test_vectors_are_self_consistent): an expected value misremembered by a digit would
not survive that check, and a wrong SEMANTIC rule (e.g. symmetric instead of bottom/right 'SAME' padding, or sampling
at pixel centres instead of corners) reproduces none of them.  The parity claim this supports: "the oracle implements
the kernel semantics TensorFlow's tests assert", not "the oracle was compared with TensorFlow output".

Conventions of the TF tests quoted: conv_ops_test fills input and filter with 1, 2, 3, ... in row-major order
(`x1 = [f * 1.0 for f in range(1, total_size_1 + 1)]`), NHWC input, HWIO filter.
"""
import numpy as np


def _seq(shape):
    return np.arange(1, int(np.prod(shape)) + 1, dtype=np.float64).reshape(shape)


# ---- tensorflow/python/kernel_tests/conv_ops_test.py (Conv2DTest._VerifyValues callers) -----------------------------
# (test name, input NHWC shape, filter HWIO shape, stride, padding, expected flat output)
CONV2D = [
    ("testConv2D1x1Filter", (1, 2, 3, 3), (1, 1, 3, 3), 1, "VALID",
     [30.0, 36.0, 42.0, 66.0, 81.0, 96.0, 102.0, 126.0, 150.0, 138.0, 171.0, 204.0, 174.0, 216.0, 258.0, 210.0, 261.0,
      312.0]),
    ("testConv2D2x2Filter", (1, 2, 3, 3), (2, 2, 3, 3), 1, "VALID", [2271.0, 2367.0, 2463.0, 2901.0, 3033.0, 3165.0]),
    ("testConv2D1x2Filter", (1, 2, 3, 3), (1, 2, 3, 3), 1, "VALID",
     [231.0, 252.0, 273.0, 384.0, 423.0, 462.0, 690.0, 765.0, 840.0, 843.0, 936.0, 1029.0]),
    ("testConv2D2x2FilterStride2", (1, 2, 3, 3), (2, 2, 3, 3), 2, "VALID", [2271.0, 2367.0, 2463.0]),
    ("testConv2D2x2FilterStride2Same", (1, 2, 3, 3), (2, 2, 3, 3), 2, "SAME",
     [2271.0, 2367.0, 2463.0, 1230.0, 1305.0, 1380.0]),
    # testConv2DKernelSmallerThanStrideSame: the three _VerifyValues calls.  The last one is the asymmetric case the
    # reference's stride-2 convs rely on: 4 -> ceil(4/3) = 2 outputs need 1 pad pixel, and it goes to the BOTTOM / RIGHT.
    ("testConv2DKernelSmallerThanStrideSame[0]", (1, 3, 3, 1), (1, 1, 1, 1), 2, "SAME", [1.0, 3.0, 7.0, 9.0]),
    ("testConv2DKernelSmallerThanStrideSame[1]", (1, 4, 4, 1), (1, 1, 1, 1), 2, "SAME", [1.0, 3.0, 9.0, 11.0]),
    ("testConv2DKernelSmallerThanStrideSame[2]", (1, 4, 4, 1), (2, 2, 1, 1), 3, "SAME", [44.0, 28.0, 41.0, 16.0]),
]


def conv2d_inputs(case):
    _, xs, ws, _, _, _ = case
    return _seq(xs), _seq(ws)


# ---- tensorflow/python/kernel_tests/crop_and_resize_op_test.py (CropAndResizeOpTest), method='bilinear' ---------------
# (test name, image [H][W] (one channel, batch 1), boxes, box_ind, crop_size, extrapolation_value, expected [nbox][ch][cw])
CROP_AND_RESIZE = [
    ("testCropAndResize2x2To1x1", [[1, 2], [3, 4]], [[0, 0, 1, 1]], [0], (1, 1), 0.0, [[[2.5]]]),
    ("testCropAndResize2x2To1x1Flipped", [[1, 2], [3, 4]], [[1, 1, 0, 0]], [0], (1, 1), 0.0, [[[2.5]]]),
    ("testCropAndResize2x2To3x3", [[1, 2], [3, 4]], [[0, 0, 1, 1]], [0], (3, 3), 0.0,
     [[[1, 1.5, 2], [2, 2.5, 3], [3, 3.5, 4]]]),
    ("testCropAndResize2x2To3x3Flipped", [[1, 2], [3, 4]], [[1, 1, 0, 0]], [0], (3, 3), 0.0,
     [[[4, 3.5, 3], [3, 2.5, 2], [2, 1.5, 1]]]),
    ("testCropAndResize3x3To2x2", [[1, 2, 3], [4, 5, 6], [7, 8, 9]], [[0, 0, 1, 1], [0, 0, 0.5, 0.5]], [0, 0], (2, 2), 0.0,
     [[[1, 3], [7, 9]], [[1, 2], [4, 5]]]),
    ("testCropAndResize3x3To2x2Flipped", [[1, 2, 3], [4, 5, 6], [7, 8, 9]], [[1, 1, 0, 0], [0.5, 0.5, 0, 0]], [0, 0], (2, 2),
     0.0, [[[9, 7], [3, 1]], [[5, 4], [2, 1]]]),
    # TF's test passes extrapolation_value = -1; the reference calls crop_and_resize with the default (0), which is the
    # only value the oracle / kernels implement, so the vector is restated with v = 0 (the test's expected array is
    # written in terms of v: [[v, v, v], [v, 1, 2], [v, 3, 4]])
    ("testCropAndResize2x2To3x3Extrapolated", [[1, 2], [3, 4]], [[-1, -1, 1, 1]], [0], (3, 3), 0.0,
     [[[0, 0, 0], [0, 1, 2], [0, 3, 4]]]),
]


# ---- tensorflow/python/training/adam_test.py: adam_update_numpy (the reference implementation the test asserts
# AdamOptimizer against for 3 steps) and testBasic's variables / gradients ---------------------------------------------
def adam_update_numpy(param, g_t, t, m, v, alpha=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
    alpha_t = alpha * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    m_t = beta1 * m + (1 - beta1) * g_t
    v_t = beta2 * v + (1 - beta2) * g_t * g_t
    param_t = param - alpha_t * m_t / (np.sqrt(v_t) + epsilon)
    return param_t, m_t, v_t


ADAM_BASIC = dict(var0=[1.0, 2.0], grads0=[0.1, 0.1], var1=[3.0, 4.0], grads1=[0.01, 0.01], steps=3)


# ---- tensorflow/python/training/rmsprop_test.py: _rmsprop_update_numpy, non-centered, momentum 0: the slot `rms` starts
# at ONES and epsilon sits INSIDE the square root -- both differ from PyTorch's RMSprop ---------------------------------
def rmsprop_update_numpy(var, g, rms, lr, decay, epsilon):
    rms_t = rms * decay + (1 - decay) * g * g
    var_t = var - lr * g / np.sqrt(rms_t + epsilon)
    return var_t, rms_t


RMSPROP_BASIC = dict(var0=[1.0, 2.0], grads0=[0.1, 0.2], var1=[3.0, 4.0], grads1=[0.01, 0.2], steps=4)


# ---- tensorflow/python/ops/nn_fused_batchnorm_test.py: _training_ref = moments over (N,H,W) + batch_normalization with
# the BIASED variance for y (the returned variance is Bessel-corrected, y is not); epsilon 0.001 there, 1e-5 in the
# reference's call (tflib/ops/batchnorm.py:30) -- an identity, restated as a function ------------------------------------
def fused_batch_norm_training_ref(x, scale, offset, epsilon):
    mean = x.mean(axis=(0, 1, 2))
    var = x.var(axis=(0, 1, 2))           # numpy default ddof = 0 == tf.nn.moments
    return (x - mean) / np.sqrt(var + epsilon) * scale + offset


# ---- tensorflow/python/ops/image_ops_test.py ResizeImagesTest.testResizeUp, NEAREST_NEIGHBOR, align_corners=False -------
RESIZE_NN_UP = dict(data=[128, 128, 64, 64, 32, 32], in_shape=(1, 3, 2, 1), out_shape=(1, 6, 4, 1),
                    expected=[128.0] * 8 + [64.0] * 8 + [32.0] * 8)


# ---- tensorflow/python/ops/nn_impl.py sigmoid_cross_entropy_with_logits docstring: the numerically stable form
#      max(x, 0) - x * z + log(1 + exp(-abs(x)))   ==   z * -log(sigmoid(x)) + (1 - z) * -log(1 - sigmoid(x)) ----------
def sigmoid_ce_doc(x, z):
    return np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))
