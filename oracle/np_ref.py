"""ORACLE (test infrastructure only).  Scalar-loop numpy restatements of the same TensorFlow-1.4 kernel
semantics as oracle/tf_ops.py, written independently (index arithmetic spelled out, no library conv /
gather) so that the two can be checked against each other on small cases.  Parity with real TensorFlow
remains unpinned (TF cannot be installed here); see the header of oracle/tf_ops.py.
"""
import math

import numpy as np


def conv2d_same(x, w, b=None, stride=1):
    """x NHWC, w HWIO.  out[n,oy,ox,co] = sum_{ky,kx,ci} x[n, oy*s+ky-pt, ox*s+kx-pl, ci] * w[ky,kx,ci,co]."""
    N, H, W, Ci = x.shape
    KH, KW, _, Co = w.shape
    OH, OW = -(-H // stride), -(-W // stride)
    pt = max((OH - 1) * stride + KH - H, 0) // 2
    pl = max((OW - 1) * stride + KW - W, 0) // 2
    y = np.zeros((N, OH, OW, Co), dtype=np.float64)
    for n in range(N):
        for oy in range(OH):
            for ox in range(OW):
                acc = np.zeros(Co, dtype=np.float64)
                for ky in range(KH):
                    iy = oy * stride + ky - pt
                    if iy < 0 or iy >= H:
                        continue
                    for kx in range(KW):
                        ix = ox * stride + kx - pl
                        if ix < 0 or ix >= W:
                            continue
                        acc += x[n, iy, ix, :].astype(np.float64) @ w[ky, kx].astype(np.float64)
                y[n, oy, ox] = acc
    if b is not None:
        y += b
    return y


def crop_and_resize(image, boxes, box_ind, crop_size, extrapolation=0.0):
    N, H, W, C = image.shape
    ch, cw = crop_size
    out = np.zeros((len(boxes), ch, cw, C), dtype=np.float64)
    for b, (y1, x1, y2, x2) in enumerate(boxes):
        n = int(box_ind[b])
        hs = (y2 - y1) * (H - 1) / (ch - 1) if ch > 1 else 0.0
        ws = (x2 - x1) * (W - 1) / (cw - 1) if cw > 1 else 0.0
        for y in range(ch):
            in_y = y1 * (H - 1) + y * hs if ch > 1 else 0.5 * (y1 + y2) * (H - 1)
            if in_y < 0 or in_y > H - 1:
                out[b, y] = extrapolation
                continue
            t, bo = int(math.floor(in_y)), int(math.ceil(in_y))
            yl = in_y - t
            for x in range(cw):
                in_x = x1 * (W - 1) + x * ws if cw > 1 else 0.5 * (x1 + x2) * (W - 1)
                if in_x < 0 or in_x > W - 1:
                    out[b, y, x] = extrapolation
                    continue
                l, r = int(math.floor(in_x)), int(math.ceil(in_x))
                xl = in_x - l
                top = image[n, t, l] + (image[n, t, r] - image[n, t, l]) * xl
                bot = image[n, bo, l] + (image[n, bo, r] - image[n, bo, l]) * xl
                out[b, y, x] = top + (bot - top) * yl
    return out


def resize_nn2(x):
    N, H, W, C = x.shape
    y = np.zeros((N, 2 * H, 2 * W, C), dtype=x.dtype)
    for i in range(2 * H):
        for j in range(2 * W):
            y[:, i, j] = x[:, i // 2, j // 2]
    return y


def batchnorm_train(x, scale, offset, eps=1e-5):
    y = np.zeros_like(x, dtype=np.float64)
    for c in range(x.shape[3]):
        v = x[..., c].astype(np.float64)
        m = v.mean()
        var = ((v - m) ** 2).mean()
        y[..., c] = (v - m) / math.sqrt(var + eps) * scale[c] + offset[c]
    return y


def layernorm(x, scale, offset, eps=1e-5):
    y = np.zeros_like(x, dtype=np.float64)
    for n in range(x.shape[0]):
        v = x[n].astype(np.float64)
        m = v.mean()
        var = ((v - m) ** 2).mean()
        y[n] = (v - m) / math.sqrt(var + eps) * scale + offset
    return y


def sigmoid_ce(z, l):
    return np.array([max(a, 0.0) - a * b + math.log1p(math.exp(-abs(a))) for a, b in zip(z, l)])


def adam_steps(p, grads, lr, beta1=0.5, beta2=0.999, eps=1e-8):
    """Runs len(grads) TF-Adam steps on a copy of p."""
    p = p.astype(np.float64).copy()
    m = np.zeros_like(p)
    v = np.zeros_like(p)
    for t, g in enumerate(grads, start=1):
        lr_t = lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
        m = beta1 * m + (1 - beta1) * g
        v = beta2 * v + (1 - beta2) * g * g
        p = p - lr_t * m / (np.sqrt(v) + eps)
    return p


def pose_inflate(rcv, H=128, W=64):
    """coord2channel_simple_rcv + tf_poseInflate exactly as the reference enumerates them
    (utils.py:259-318): scatter 2*V at the truncated (r,c), subtract 1, then OR over the shifted copies
    listed at utils.py:300-314 (x_offset shifts rows, y_offset shifts columns)."""
    n, k, _ = rcv.shape
    land = np.zeros((n, H, W, k))
    for b in range(n):
        for j in range(k):
            r, c, v = int(rcv[b, j, 0]), int(rcv[b, j, 1]), rcv[b, j, 2]
            land[b, r, c, j] = 2.0 * v
    land = land - 1.0
    g = (land + 1.0) / 2.0
    offsets = [(xo, 0) for xo in (-4, 4)]
    offsets += [(xo, yo) for xo in (-3, 3) for yo in range(-2, 3)]
    offsets += [(xo, yo) for xo in (-2, 2) for yo in range(-3, 4)]
    offsets += [(xo, yo) for xo in (-1, 1) for yo in range(-3, 4)]
    offsets += [(0, yo) for yo in range(-4, 5)]
    out = g.copy()
    for xo, yo in offsets:
        # pad by 4, crop at (xo+4, yo+4): shifted[r, c] = g[r + xo, c + yo], zero outside
        sh = np.zeros_like(g)
        for r in range(H):
            rr = r + xo
            if rr < 0 or rr >= H:
                continue
            c_lo, c_hi = max(0, -yo), min(W, W - yo)
            sh[:, r, c_lo:c_hi] = g[:, rr, c_lo + yo:c_hi + yo]
        out += sh
    out = np.minimum(out, 1.0)
    return out * 2.0 - 1.0
