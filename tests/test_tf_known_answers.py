"""The oracle against TensorFlow's own published unit-test vectors (tests/golden/tf_known_answers.py: which TF test
each vector restates, and how far this pin reaches -- doc-derived, not a TensorFlow run).  CPU tests; the GPU kernels
are run against the same vectors in tests/test_ops_gpu.py::test_kernels_match_tf_known_answers."""
import os
import sys
from fractions import Fraction

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

import tf_known_answers as KA  # noqa: E402
from oracle import tf_ops as T  # noqa: E402


# ----------------------------------------------------------------- exact re-derivation of the restated vectors
def _conv_exact(x, w, stride, padding):
    """Direct nested-loop cross-correlation with TensorFlow's documented padding rule, exact integers."""
    n, h, wd, ci = x.shape
    kh, kw, _, co = w.shape
    if padding == "SAME":
        oh, ow = -(-h // stride), -(-wd // stride)
        pt = max((oh - 1) * stride + kh - h, 0) // 2
        pl = max((ow - 1) * stride + kw - wd, 0) // 2
    else:
        oh, ow, pt, pl = (h - kh) // stride + 1, (wd - kw) // stride + 1, 0, 0
    out = np.zeros((n, oh, ow, co))
    for b in range(n):
        for i in range(oh):
            for j in range(ow):
                for o in range(co):
                    s = 0
                    for a in range(kh):
                        for c in range(kw):
                            y, xx = i * stride + a - pt, j * stride + c - pl
                            if 0 <= y < h and 0 <= xx < wd:
                                s += int(sum(int(x[b, y, xx, k]) * int(w[a, c, k, o]) for k in range(ci)))
                    out[b, i, j, o] = s
    return out


def _crop_exact(img, box, crop):
    """crop_and_resize of one box with rational arithmetic (bilinear, extrapolation 0), corner-aligned sampling."""
    H, W = len(img), len(img[0])
    y1, x1, y2, x2 = [Fraction(v).limit_denominator(1000) for v in box]
    ch, cw = crop
    out = []
    for i in range(ch):
        in_y = y1 * (H - 1) + i * (y2 - y1) * (H - 1) / (ch - 1) if ch > 1 else (y1 + y2) * (H - 1) / 2
        row = []
        for j in range(cw):
            in_x = x1 * (W - 1) + j * (x2 - x1) * (W - 1) / (cw - 1) if cw > 1 else (x1 + x2) * (W - 1) / 2
            if in_y < 0 or in_y > H - 1 or in_x < 0 or in_x > W - 1:
                row.append(Fraction(0))
                continue
            t, l = int(in_y // 1), int(in_x // 1)
            b, r = min(t + (in_y != t), H - 1), min(l + (in_x != l), W - 1)
            fy, fx = in_y - t, in_x - l
            top = img[t][l] + (img[t][r] - img[t][l]) * fx
            bot = img[b][l] + (img[b][r] - img[b][l]) * fx
            row.append(top + (bot - top) * fy)
        out.append(row)
    return out


def test_vectors_are_self_consistent():
    """Every restated expected array equals an exact re-derivation from the documented rule: guards the restatement
    itself (a mistyped digit, a test attributed to the wrong rule)."""
    for case in KA.CONV2D:
        name, _, _, stride, padding, expected = case
        x, w = KA.conv2d_inputs(case)
        assert _conv_exact(x, w, stride, padding).reshape(-1).tolist() == expected, name
    for name, img, boxes, ind, crop, ev, expected in KA.CROP_AND_RESIZE:
        for b, box in enumerate(boxes):
            got = [[float(v) for v in row] for row in _crop_exact(img, box, crop)]
            assert got == [[float(v) for v in row] for row in expected[b]], name


# ----------------------------------------------------------------- the oracle against the vectors
@pytest.mark.parametrize("case", KA.CONV2D, ids=[c[0] for c in KA.CONV2D])
def test_oracle_conv2d(case):
    name, _, _, stride, padding, expected = case
    if padding != "SAME":
        # the reference path only ever uses SAME (slim.conv2d default, tflib conv2d.py:106-120); the VALID vectors pin the
        # cross-correlation orientation / HWIO layout: run them through the same oracle function on a pre-cropped view
        x, w = KA.conv2d_inputs(case)
        y = torch.nn.functional.conv2d(torch.tensor(x).permute(0, 3, 1, 2), torch.tensor(w).permute(3, 2, 0, 1),
                                       stride=stride).permute(0, 2, 3, 1)
        ys = T.conv2d_same(torch.tensor(x), torch.tensor(w), None, stride)
        assert y.reshape(-1).tolist() == expected, name
        # SAME and VALID agree wherever the window does not touch the padding (top-left aligned: pad_before = 0 here when
        # pad_total <= 1, else offset by pad_before)
        kh, kw = w.shape[0], w.shape[1]
        _, pt, _ = T.same_pads(x.shape[1], kh, stride)
        _, pl, _ = T.same_pads(x.shape[2], kw, stride)
        if stride == 1:
            assert torch.equal(ys[:, pt:pt + y.shape[1], pl:pl + y.shape[2]], y), name
        return
    x, w = KA.conv2d_inputs(case)
    y = T.conv2d_same(torch.tensor(x), torch.tensor(w), None, stride)
    assert y.reshape(-1).tolist() == expected, name


@pytest.mark.parametrize("case", KA.CROP_AND_RESIZE, ids=[c[0] for c in KA.CROP_AND_RESIZE])
def test_oracle_crop_and_resize(case):
    name, img, boxes, ind, crop, ev, expected = case
    image = torch.tensor(img, dtype=torch.float64)[None, :, :, None]
    y = T.crop_and_resize(image, torch.tensor(boxes, dtype=torch.float64), torch.tensor(ind, dtype=torch.int32), crop)
    assert torch.allclose(y[..., 0], torch.tensor(expected, dtype=torch.float64), atol=1e-12, rtol=0), (name, y[..., 0])


def test_oracle_adam_matches_tf_adam_test_reference():
    """adam_test.py testBasic: 3 steps, lr 0.001, beta1 0.9, beta2 0.999, eps 1e-8, against adam_update_numpy; then the
    reference's own hyper-parameters (trainer.py:130-140: beta1 0.5, lr 2e-5) through the same formula."""
    c = KA.ADAM_BASIC
    for lr, b1 in ((0.001, 0.9), (2e-5, 0.5)):
        for var, grad in ((c["var0"], c["grads0"]), (c["var1"], c["grads1"])):
            p_np, m_np, v_np = np.array(var), np.zeros(2), np.zeros(2)
            p, m, v = torch.tensor(var, dtype=torch.float64), torch.zeros(2, dtype=torch.float64), torch.zeros(2, dtype=torch.float64)
            g = torch.tensor(grad, dtype=torch.float64)
            for t in range(1, c["steps"] + 1):
                p_np, m_np, v_np = KA.adam_update_numpy(p_np, np.array(grad), t, m_np, v_np, alpha=lr, beta1=b1)
                T.adam_step(p, g, m, v, lr, t, beta1=b1, beta2=0.999, eps=1e-8)
                assert np.allclose(p.numpy(), p_np, rtol=1e-14, atol=0) and np.allclose(m.numpy(), m_np, rtol=1e-14)
                assert np.allclose(v.numpy(), v_np, rtol=1e-14)


def test_oracle_rmsprop_matches_tf_rmsprop_test_reference():
    """rmsprop_test.py _rmsprop_update_numpy (non-centered, momentum 0): rms slot starts at ONES, epsilon inside the
    square root; with the reference's RMSPropOptimizer(lr) defaults decay 0.9, epsilon 1e-10 (trainer.py:119-128)."""
    c = KA.RMSPROP_BASIC
    for var, grad in ((c["var0"], c["grads0"]), (c["var1"], c["grads1"])):
        p_np, rms_np = np.array(var), np.ones(2)
        p, ms = torch.tensor(var, dtype=torch.float64), torch.ones(2, dtype=torch.float64)
        g = torch.tensor(grad, dtype=torch.float64)
        for _ in range(c["steps"]):
            p_np, rms_np = KA.rmsprop_update_numpy(p_np, np.array(grad), rms_np, 2.0, 0.9, 1e-10)
            T.rmsprop_step(p, g, ms, 2.0, decay=0.9, eps=1e-10)
            assert np.allclose(p.numpy(), p_np, rtol=1e-14, atol=0) and np.allclose(ms.numpy(), rms_np, rtol=1e-14)


def test_oracle_fused_batch_norm_training_identity():
    rng = np.random.default_rng(3)
    x = rng.normal(1.0, 2.0, size=(3, 4, 5, 6))
    scale, offset = rng.normal(size=6), rng.normal(size=6)
    for eps in (0.001, 1e-5):
        want = KA.fused_batch_norm_training_ref(x, scale, offset, eps)
        got = T.batchnorm_train(torch.tensor(x), torch.tensor(scale), torch.tensor(offset), eps=eps)
        assert np.allclose(got.numpy(), want, rtol=1e-12, atol=1e-12)


def test_oracle_resize_nearest_and_sigmoid_ce():
    c = KA.RESIZE_NN_UP
    x = torch.tensor(c["data"], dtype=torch.float64).reshape(c["in_shape"])
    assert T.upscale2(x).reshape(-1).tolist() == c["expected"]
    xs = np.array([100.0, -100.0, 0.05, -0.3, 7.0, 0.0])
    zs = np.array([0.0, 0.0, 1.0, 1.0, 0.5, 1.0])
    got = T.sigmoid_ce(torch.tensor(xs), torch.tensor(zs)).numpy()
    assert np.allclose(got, KA.sigmoid_ce_doc(xs, zs), rtol=1e-14, atol=0)
    sig = 1.0 / (1.0 + np.exp(-xs[2:]))
    assert np.allclose(got[2:], -zs[2:] * np.log(sig) - (1 - zs[2:]) * np.log(1 - sig), rtol=1e-10)
