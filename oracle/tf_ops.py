"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement, in PyTorch tensor ops, of the TensorFlow-1.4 kernels the reference's hot path calls.
TensorFlow is an UN-VENDORED third-party dependency of the reference (pinned only in prose:
README.md:17 `tensorflow-gpu 1.4.1`) and is not installable here, so **parity is unpinned**: each
function restates the published TF kernel semantics and cites the reference call site it serves.
`oracle/np_ref.py` holds independent scalar-loop restatements used to cross-check these on small
cases (tests/test_oracle.py).

All functions are dtype-agnostic (run them in float64 for truth, float32 for the CPU baseline) and
differentiable through torch.autograd, which plays the role of `tf.gradients` (trainer.py:233,
`Optimizer.minimize` trainer.py:119-146).
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- convolution
def same_pads(size, k, s):
    """TF 'SAME' (core/framework/common_shape_fns.cc): out=ceil(in/s), pad_total=max((out-1)s+k-in,0),
    pad_before = pad_total//2 (the extra pixel goes to the bottom / right)."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return out, total // 2, total - total // 2


def conv2d_same(x, w, b=None, stride=1):
    """slim.conv2d / tf.nn.conv2d(padding='SAME') + bias_add on NHWC x, HWIO w.
    Reference call sites: models.py:396-399, 425-429, 458-462, 528-539, 564-573;
    tflib/ops/conv2d.py:106-120 (there NCHW; the arithmetic is layout independent)."""
    kh, kw = w.shape[0], w.shape[1]
    _, pt, pb = same_pads(x.shape[1], kh, stride)
    _, pl, pr = same_pads(x.shape[2], kw, stride)
    xn = x.permute(0, 3, 1, 2)
    xn = F.pad(xn, (pl, pr, pt, pb))
    y = F.conv2d(xn, w.permute(3, 2, 0, 1), b, stride=stride)  # cross-correlation, like TF
    return y.permute(0, 2, 3, 1)


def upscale2(x):
    """utils.upscale -> tf.image.resize_nearest_neighbor, align_corners=False, scale 2 (utils.py:61-72):
    out[i, j] = in[i // 2, j // 2]."""
    return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)


def crop_and_resize(image, boxes, box_ind, crop_size):
    """tf.image.crop_and_resize(image, boxes, box_ind, crop_size), bilinear, extrapolation_value=0
    (models.py:350, 415).  Restates core/kernels/crop_and_resize_op.cc:
      in_y = y1*(H-1) + y*(y2-y1)*(H-1)/(ch-1)  if ch>1 else 0.5*(y1+y2)*(H-1); same for x;
      samples with in_y not in [0,H-1] or in_x not in [0,W-1] give 0;
      top=floor, bottom=ceil, lerp = in - top; value = top + (bottom-top)*y_lerp with
      top = tl + (tr-tl)*x_lerp, bottom = bl + (br-bl)*x_lerp.  Gradient flows to `image` only."""
    N, H, W, C = image.shape
    ch, cw = crop_size
    dt = image.dtype
    boxes = boxes.to(dt)
    y1, x1, y2, x2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    ys = torch.arange(ch, dtype=dt)
    xs = torch.arange(cw, dtype=dt)
    if ch > 1:
        in_y = y1[:, None] * (H - 1) + ys[None, :] * ((y2 - y1) * (H - 1) / (ch - 1))[:, None]
    else:
        in_y = (0.5 * (y1 + y2) * (H - 1))[:, None].expand(-1, ch)
    if cw > 1:
        in_x = x1[:, None] * (W - 1) + xs[None, :] * ((x2 - x1) * (W - 1) / (cw - 1))[:, None]
    else:
        in_x = (0.5 * (x1 + x2) * (W - 1))[:, None].expand(-1, cw)
    vy = (in_y >= 0) & (in_y <= H - 1)
    vx = (in_x >= 0) & (in_x <= W - 1)
    top = in_y.floor().clamp(0, H - 1).long()
    bot = in_y.ceil().clamp(0, H - 1).long()
    left = in_x.floor().clamp(0, W - 1).long()
    right = in_x.ceil().clamp(0, W - 1).long()
    yl = (in_y - in_y.floor())[:, :, None, None]
    xl = (in_x - in_x.floor())[:, None, :, None]
    b = box_ind.long()[:, None, None]

    def g(yy, xx):
        return image[b, yy[:, :, None], xx[:, None, :]]  # [nbox, ch, cw, C]

    tl, tr, bl, br = g(top, left), g(top, right), g(bot, left), g(bot, right)
    t = tl + (tr - tl) * xl
    bo = bl + (br - bl) * xl
    out = t + (bo - t) * yl
    valid = (vy[:, :, None] & vx[:, None, :])[..., None]
    return torch.where(valid, out, torch.zeros((), dtype=dt))


# ----------------------------------------------------------------------------- normalisation
def layernorm(x_nhwc, scale, offset, eps=1e-5):
    """tflib/ops/layernorm.py:6-20 with norm_axes=[1,2,3] on BCHW == all of (C,H,W) per sample;
    tf.nn.moments -> biased variance; tf.nn.batch_normalization: (x-m)*rsqrt(v+eps)*scale+offset."""
    m = x_nhwc.mean(dim=(1, 2, 3), keepdim=True)
    v = ((x_nhwc - m) ** 2).mean(dim=(1, 2, 3), keepdim=True)
    return (x_nhwc - m) * torch.rsqrt(v + eps) * scale + offset


def batchnorm_train(x_nhwc, scale, offset, eps=1e-5):
    """tflib/ops/batchnorm.py:29-30, 52-53: tf.nn.fused_batch_norm in training mode (always -- the
    reference never passes is_training): per-channel batch statistics over (N,H,W), biased variance."""
    m = x_nhwc.mean(dim=(0, 1, 2), keepdim=True)
    v = ((x_nhwc - m) ** 2).mean(dim=(0, 1, 2), keepdim=True)
    return (x_nhwc - m) * torch.rsqrt(v + eps) * scale + offset


def instance_norm(x_nhwc, scale, shift, eps=1e-3):
    """models.Instance_norm (models.py:154-166; defined but never called by the shipped graphs)."""
    m = x_nhwc.mean(dim=(1, 2), keepdim=True)
    v = ((x_nhwc - m) ** 2).mean(dim=(1, 2), keepdim=True)
    return scale * (x_nhwc - m) / (v + eps) ** 0.5 + shift


def leaky_relu(x, alpha=0.2):
    """wgan_gp.LeakyReLU (wgan_gp.py:23-24): tf.maximum(alpha*x, x)."""
    return torch.maximum(alpha * x, x)


# ----------------------------------------------------------------------------- losses
def sigmoid_ce(logits, labels):
    """tf.nn.sigmoid_cross_entropy_with_logits: max(z,0) - z*l + log1p(exp(-|z|))."""
    return torch.clamp(logits, min=0) - logits * labels + torch.log1p(torch.exp(-logits.abs()))


def gan_loss(mode, d_real, d_fake):
    """trainer._gan_loss (trainer.py:217-252) without the gradient penalty term."""
    if mode in ("wgan", "wgan-gp"):
        return -d_fake.mean(), d_fake.mean() - d_real.mean()
    if mode == "dcgan":
        g = sigmoid_ce(d_fake, torch.ones_like(d_fake)).mean()
        d = sigmoid_ce(d_fake, torch.zeros_like(d_fake)).mean() + sigmoid_ce(d_real, torch.ones_like(d_real)).mean()
        return g, d / 2.0
    if mode == "lsgan":
        return ((d_fake - 1) ** 2).mean(), (((d_real - 1) ** 2).mean() + (d_fake ** 2).mean()) / 2.0
    raise Exception()


def gradient_penalty(disc_fn, real, fake, alpha):
    """trainer.py:226-236 for an image-space critic: alpha per sample, slopes = L2 norm of
    d D(xhat)/d xhat over all non-batch axes, gp = mean((slopes-1)^2).  (As written the reference
    reduces over axis 1 of 2-D inputs; for 4-D images the per-sample norm is the north-star variant,
    SURVEY.md q8.)  Returns gp (differentiable w.r.t. the critic parameters)."""
    a = alpha.reshape(-1, *([1] * (real.dim() - 1)))
    xhat = (real + a * (fake - real)).detach().requires_grad_(True)
    out = disc_fn(xhat)
    grad, = torch.autograd.grad(out.sum(), xhat, create_graph=True)
    slopes = torch.sqrt((grad ** 2).flatten(1).sum(dim=1))
    return ((slopes - 1.0) ** 2).mean(), slopes, grad


# ----------------------------------------------------------------------------- optimisers
def adam_step(p, g, m, v, lr, t, beta1=0.5, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer (trainer.py:130-140): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr_t*m/(sqrt(v)+eps).  In place, first step t=1."""
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    p.sub_(lr_t * m / (v.sqrt() + eps))


def rmsprop_step(p, g, ms, lr, decay=0.9, eps=1e-10, clip=None):
    """tf.train.RMSPropOptimizer(lr) (trainer.py:119-128): ms starts at ones; ms = d*ms + (1-d) g^2;
    p -= lr*g/sqrt(ms+eps); then the WGAN weight clip to +-0.01 (trainer.py:124-128)."""
    ms.mul_(decay).addcmul_(g, g, value=1.0 - decay)
    p.sub_(lr * g / torch.sqrt(ms + eps))
    if clip is not None:
        p.clamp_(-clip, clip)


# ----------------------------------------------------------------------------- pose maps
def pose_rasterize(rcv, H=128, W=64, radius=4):
    """coord2channel_simple_rcv (utils.py:259-287, is_normalized=False) followed by tf_poseInflate
    (utils.py:289-318): tf.to_int32 truncates the coordinates; a visible keypoint lights the radius-4
    disc {dr^2+dc^2 <= 16} (the 49 offsets enumerated at utils.py:300-314); result in {-1,+1}."""
    n, k = rcv.shape[0], rcv.shape[1]
    r0 = rcv[:, :, 0].to(torch.int64)  # trunc toward zero == tf.to_int32 for the non-negative inputs
    c0 = rcv[:, :, 1].to(torch.int64)
    vis = rcv[:, :, 2]
    yy = torch.arange(H).view(1, H, 1, 1)
    xx = torch.arange(W).view(1, 1, W, 1)
    d2 = (yy - r0.view(n, 1, 1, k)) ** 2 + (xx - c0.view(n, 1, 1, k)) ** 2
    inside = (d2 <= radius * radius).to(rcv.dtype) * vis.view(n, 1, 1, k).clamp(max=1.0)
    return inside * 2.0 - 1.0


def denorm_img(g):
    """utils.denorm_img (utils.py:88-89)."""
    return torch.clamp((g + 1.0) * 127.5, 0, 255)
