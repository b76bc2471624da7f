"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the reference's hot-path graphs built on oracle/tf_ops.py:

  encoder_fgbg      <- models.GeneratorCNN_ID_Encoder_BodyROIVis_FgBgFeaTwoBranch   models.py:390-471
  encoder_roi       <- models.GeneratorCNN_ID_Encoder_BodyROIVis (DeepFashion)      models.py:328-388
  unet_generator    <- models.GeneratorCNN_ID_UAEAfterResidual                      models.py:518-576
  dcgan_discriminator <- WGAN_GP.DCGANDiscriminator                                 wgan_gp.py:407-440
  fc_discriminator  <- WGAN_GP.FCDiscriminator                                      wgan_gp.py:399-405
  gaussian_fc_res   <- models.GaussianFCRes                                         models.py:474-486
  Stage-I model (--model=1) forward, losses and the g_optim / d_optim updates
                    <- trainer.DPIG_Encoder_GAN_BodyROI_FgBg.build_model / train    trainer.py:567-625, 336-347
  DeepFashion 256x256 Stage-I model (--model=101): NetConfig.deepfashion()
                    <- trainer_256.DPIG_Encoder_GAN_BodyROI_256.build_model         trainer_256.py:31-93

Parameters live in a dict keyed by the reference's TensorFlow variable names (slim auto-numbering
`Conv`, `Conv_1`, ... / `fully_connected`, `fully_connected_1` inside `Encoder/G_encoder` and
`ID_AE/G`; `Discriminator.N.Filters` etc. from tflib), weights HWIO / [in,out] exactly as TF stores them.
Parity is unpinned (no TensorFlow here); see oracle/tf_ops.py.
"""
import math
from collections import OrderedDict

import numpy as np
import torch

from . import tf_ops as T


class NetConfig:
    """Shapes of the Stage-I graphs.

    Market-1501 128x64, --model=1 (config.py:23-25, trainer.py:74-75, 576-582): Fg/Bg two-branch encoder,
    ROI pyramid and U-Net both `repeat_num` levels deep, D applied to x and G in separate calls.
    DeepFashion 256x256, --model=101 (trainer_256.py:31-68): `fgbg=False` (models.GeneratorCNN_ID_Encoder_BodyROIVis,
    no mask, no background branch), ROI pyramid `repeat_num+1` levels on 64x64 crops, U-Net `repeat_num-1` levels,
    D applied once to concat([x, G]) (`d_joint`: joint batch statistics, logits split in half afterwards)."""

    def __init__(self, img_h=128, img_w=64, hidden=128, z_num=64, roi_size=48, n_parts=7, part_z=32,
                 keypoints=18, d_dim=64, repeat_num=None, fgbg=True, enc_repeat=None, unet_repeat=None, d_joint=False,
                 use_vis=True):
        self.img_h, self.img_w, self.hidden, self.z_num = img_h, img_w, hidden, z_num
        # use_vis=False: models.GeneratorCNN_ID_Encoder_BodyROI (models.py:275-325; --model=103 / 1002): no visibility gating
        self.use_vis = use_vis
        self.roi_size, self.n_parts, self.part_z, self.keypoints, self.d_dim = roi_size, n_parts, part_z, keypoints, d_dim
        self.repeat_num = repeat_num if repeat_num is not None else int(math.log2(img_h)) - 2  # trainer.py:75
        self.fgbg, self.d_joint = fgbg, d_joint
        self.enc_repeat = enc_repeat if enc_repeat is not None else self.repeat_num
        self.unet_repeat = unet_repeat if unet_repeat is not None else self.repeat_num
        # 7*32 + 128 = 352 (models.py:464-468); 7*32 = 224 without the background branch (models.py:384)
        self.emb_dim = n_parts * part_z + (4 * part_z if fgbg else 0)
        # D's Linear reads rows of 8*4*8*dim features: `tf.reshape(output, [-1, 8*4*8*dim])` (wgan_gp.py:433) is
        # hard-wired to the 128x64 geometry, so a 256x256 image yields 8 rows = 8 logits.  Reduced test
        # geometries whose whole map is smaller than that keep one row per image.
        self.d_row = min(8 * 4 * 8 * d_dim, (img_h // 16) * (img_w // 16) * 8 * d_dim)

    @classmethod
    def deepfashion(cls, img_h=256, img_w=256, hidden=128, roi_size=64, **kw):
        """--model=101 (trainer_256.py:40-55): encoder repeat_num+1 levels, roi 64, U-Net repeat_num-1 levels."""
        rn = int(math.log2(img_h)) - 2
        return cls(img_h=img_h, img_w=img_w, hidden=hidden, roi_size=roi_size, repeat_num=rn, fgbg=False,
                   enc_repeat=rn + 1, unet_repeat=rn - 1, d_joint=True, **kw)


# --------------------------------------------------------------------------------- parameters
def _xavier(rng, shape, fan_in, fan_out):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


class _Scope:
    """Reproduces slim's variable auto-naming inside one variable_scope."""

    def __init__(self, prefix, params, rng):
        self.prefix, self.params, self.rng = prefix, params, rng
        self.nconv = 0
        self.nfc = 0

    def conv(self, k, cin, cout):
        name = "%s/Conv%s" % (self.prefix, "" if self.nconv == 0 else "_%d" % self.nconv)
        self.nconv += 1
        self.params[name + "/weights"] = _xavier(self.rng, (k, k, cin, cout), k * k * cin, k * k * cout)
        self.params[name + "/biases"] = np.zeros(cout, np.float32)
        return name

    def fc(self, cin, cout):
        name = "%s/fully_connected%s" % (self.prefix, "" if self.nfc == 0 else "_%d" % self.nfc)
        self.nfc += 1
        self.params[name + "/weights"] = _xavier(self.rng, (cin, cout), cin, cout)
        self.params[name + "/biases"] = np.zeros(cout, np.float32)
        return name


def init_params(cfg, seed=1234, bias_noise=0.0):
    """Synthetic parameters with the reference's initialisers (slim xavier_uniform + zero bias;
    D: U(+-0.02*sqrt(3)), tflib/ops/conv2d.py:56-80, wgan_gp.py:411-413; norm scale 1 / offset 0).
    bias_noise > 0 perturbs biases / norm params so tests exercise them."""
    rng = np.random.default_rng(seed)
    p = OrderedDict()
    hn, ern, rn = cfg.hidden, cfg.enc_repeat, cfg.unet_repeat
    # ---- Encoder/G_encoder (models.py:390-471 / 328-388), creation order = slim numbering
    s = _Scope("Encoder/G_encoder", p, rng)
    s.conv(3, 3, hn)
    s.conv(3, hn, hn)
    s.conv(3, hn, hn)
    for idx in range(ern):
        c = hn * (idx + 1)
        s.conv(3, c, c)
        s.conv(3, c, c)
        if idx < ern - 1:
            s.conv(3, c, hn * (idx + 2))
    roi_final = cfg.roi_size
    for _ in range(ern - 1):          # stride-2 SAME: ceil(size / 2) per level (48 -> ... -> 3 -> 2 -> 1 at 7 levels)
        roi_final = -(-roi_final // 2)
    s.fc(roi_final * roi_final * hn * ern, cfg.part_z)
    if cfg.fgbg:
        for idx in range(ern):
            c = hn * (idx + 1)
            s.conv(3, c, c)
            s.conv(3, c, c)
            if idx < ern - 1:
                s.conv(3, c, hn * (idx + 2))
        bh, bw = cfg.img_h >> (ern - 1), cfg.img_w >> (ern - 1)
        s.fc(bh * bw * hn * ern, cfg.part_z * 4)
    fh, fw = cfg.img_h >> (rn - 1), cfg.img_w >> (rn - 1)
    # ---- ID_AE/G (models.py:518-576)
    s = _Scope("ID_AE/G", p, rng)
    s.conv(3, cfg.emb_dim + cfg.keypoints, hn)
    for idx in range(rn):
        c = hn * (idx + 1)
        s.conv(3, c, c)
        s.conv(3, c, c)
        if idx < rn - 1:
            s.conv(3, c, hn * (idx + 2))
    s.fc(fh * fw * hn * rn, cfg.z_num)
    s.fc(cfg.z_num, fh * fw * hn)
    x_c = hn
    for idx in range(rn):
        c = x_c + hn * (rn - idx)
        s.conv(3, c, c)
        s.conv(3, c, c)
        if idx < rn - 1:
            x_c = hn * (rn - idx - 1)
            s.conv(1, c, x_c)
        else:
            x_c = c
    s.conv(3, x_c, 3)
    # ---- Discriminator (wgan_gp.py:407-440)
    d = cfg.d_dim
    lim = 0.02 * math.sqrt(3.0)
    chans = [3, d, 2 * d, 4 * d, 8 * d]
    for i in range(4):
        p["Discriminator.%d.Filters" % (i + 1)] = rng.uniform(-lim, lim, size=(5, 5, chans[i], chans[i + 1])).astype(np.float32)
        p["Discriminator.%d.Biases" % (i + 1)] = np.zeros(chans[i + 1], np.float32)
        if i >= 1:
            p["Discriminator.BN%d.offset" % (i + 1)] = np.zeros(chans[i + 1], np.float32)
            p["Discriminator.BN%d.scale" % (i + 1)] = np.ones(chans[i + 1], np.float32)
    d_in = cfg.d_row  # 8*4*8*dim (wgan_gp.py:433-434)
    p["Discriminator.Output.W"] = rng.uniform(-lim, lim, size=(d_in, 1)).astype(np.float32)
    p["Discriminator.Output.b"] = np.zeros(1, np.float32)
    if bias_noise > 0:
        for k in p:
            if k.endswith(("biases", "Biases", ".b", ".offset")):
                p[k] = (p[k] + rng.normal(0, bias_noise, size=p[k].shape)).astype(np.float32)
            if k.endswith(".scale"):
                p[k] = (p[k] + rng.normal(0, bias_noise, size=p[k].shape)).astype(np.float32)
    return p


def to_torch(params, dtype=torch.float64, requires_grad=False):
    return OrderedDict((k, torch.tensor(v, dtype=dtype).requires_grad_(requires_grad)) for k, v in params.items())


def is_generator_param(name):
    return name.startswith("Encoder/") or name.startswith("ID_AE/")


def is_disc_param(name):
    return name.startswith("Discriminator.")


# --------------------------------------------------------------------------------- graphs
class _Walker:
    """Hands out the slim layer names in creation order while a graph function runs."""

    def __init__(self, prefix, p, branches=None):
        """branches (test aid, see dcgan_discriminator): {prefix: [bool NHWC tensor per activated conv, in creation
        order]} -- `conv output > 0` as the implementation under test decided it; ReLU then gates with those bits."""
        self.prefix, self.p, self.nconv, self.nfc = prefix, p, 0, 0
        self.signs = iter(branches[prefix]) if branches and prefix in branches else None
        # branches["record"] (a dict): filled with the decisions this run takes, in the same layout
        self.rec = branches["record"].setdefault(prefix, []) if branches and "record" in branches else None

    def conv(self, x, stride=1, act=True):
        name = "%s/Conv%s" % (self.prefix, "" if self.nconv == 0 else "_%d" % self.nconv)
        self.nconv += 1
        y = T.conv2d_same(x, self.p[name + "/weights"], self.p[name + "/biases"], stride)
        if not act:
            return y
        if self.rec is not None:
            self.rec.append(y.detach() > 0)
        if self.signs is None:
            return torch.relu(y)  # trainers pass activation_fn=tf.nn.relu (trainer.py:581, 595)
        sign = next(self.signs)
        if sign.shape[1] * 2 == y.shape[1]:     # bits taken before a nearest-neighbour x2 upsample (1x1 conv after it)
            sign = sign.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
        assert sign.shape == y.shape, (name, sign.shape, y.shape)
        return torch.where(sign, y, torch.zeros_like(y))

    def fc(self, x):
        name = "%s/fully_connected%s" % (self.prefix, "" if self.nfc == 0 else "_%d" % self.nfc)
        self.nfc += 1
        return x @ self.p[name + "/weights"] + self.p[name + "/biases"]


def _tap(taps, name, t):
    """Debug hook: record an intermediate (and keep its gradient) under `name`."""
    if taps is not None:
        if t.requires_grad:
            t.retain_grad()
        taps[name] = t
    return t


def _pyramid(w, x, hn, rn, taps=None, tag=""):
    for idx in range(rn):
        res = x
        x = _tap(taps, "%s_a%d" % (tag, idx), w.conv(x))
        x = w.conv(x)
        x = _tap(taps, "%s_y%d" % (tag, idx), x + res)
        if idx < rn - 1:
            x = _tap(taps, "%s_x%d" % (tag, idx + 1), w.conv(x, stride=2))
    return x


def encoder_fgbg(p, cfg, x, fg_mask, roi_bbox, roi_vis, taps=None, branches=None):
    """models.py:390-471.  x [B,H,W,3]; fg_mask [B,H,W,1]; roi_bbox int [B,7,4] (y1,x1,y2,x2 pixels);
    roi_vis [B,7].  Returns the [B,352] embedding."""
    w = _Walker("Encoder/G_encoder", p, branches)
    B, H, W, _ = x.shape
    x = w.conv(x)
    res = x
    x = w.conv(x)
    x = w.conv(x)
    x = _tap(taps, "xs", x + res)
    x_fg = x * fg_mask
    x_bg = _tap(taps, "x_bg", x * (1.0 - fg_mask))
    rois = []
    for i in range(cfg.n_parts):
        bb = roi_bbox[:, i, :].to(x.dtype)
        boxes = torch.stack([bb[:, 0] / float(H), bb[:, 1] / float(W), bb[:, 2] / float(H), bb[:, 3] / float(W)], dim=1)
        rois.append(T.crop_and_resize(x_fg, boxes, torch.arange(B), (cfg.roi_size, cfg.roi_size)))
    body = _tap(taps, "rois", torch.cat(rois, dim=0))
    body = _pyramid(w, body, cfg.hidden, cfg.enc_repeat, taps, "roi")
    body = w.fc(body.reshape(body.shape[0], -1))
    feats = list(torch.split(body, B, dim=0))
    for i in range(cfg.n_parts):
        feats[i] = feats[i] * roi_vis[:, i:i + 1].to(x.dtype)
    bg = _pyramid(w, x_bg, cfg.hidden, cfg.enc_repeat, taps, "bg")
    bg = w.fc(bg.reshape(B, -1))
    feats.append(bg)
    return torch.cat(feats, dim=-1)


def encoder_roi(p, cfg, x, roi_bbox, roi_vis, taps=None, branches=None):
    """models.GeneratorCNN_ID_Encoder_BodyROIVis (models.py:328-388; the DeepFashion encoder, trainer_256.py:40-41):
    same stem / residual block / 7 ROI crops / shared pyramid / FC / visibility gating as the two-branch encoder,
    but the crops are taken from the unmasked feature map and there is no background branch.
    Returns the [B, n_parts*part_z] embedding."""
    w = _Walker("Encoder/G_encoder", p, branches)
    B, H, W, _ = x.shape
    x = w.conv(x)
    res = x
    x = w.conv(x)
    x = w.conv(x)
    x = _tap(taps, "xs", x + res)
    rois = []
    for i in range(cfg.n_parts):
        bb = roi_bbox[:, i, :].to(x.dtype)
        boxes = torch.stack([bb[:, 0] / float(H), bb[:, 1] / float(W), bb[:, 2] / float(H), bb[:, 3] / float(W)], dim=1)
        rois.append(T.crop_and_resize(x, boxes, torch.arange(B), (cfg.roi_size, cfg.roi_size)))
    body = _tap(taps, "rois", torch.cat(rois, dim=0))
    body = _pyramid(w, body, cfg.hidden, cfg.enc_repeat, taps, "roi")
    body = w.fc(body.reshape(body.shape[0], -1))
    feats = list(torch.split(body, B, dim=0))
    if getattr(cfg, "use_vis", True):
        for i in range(cfg.n_parts):
            feats[i] = feats[i] * roi_vis[:, i:i + 1].to(x.dtype)
    return torch.cat(feats, dim=-1)


def encoder(p, cfg, batch, taps=None, branches=None):
    """The appearance encoder the config selects (trainer.py:581 / trainer_256.py:40)."""
    if cfg.fgbg:
        return encoder_fgbg(p, cfg, batch["x"], batch["mask"], batch["part_bbox"], batch["part_vis"], taps, branches)
    return encoder_roi(p, cfg, batch["x"], batch["part_bbox"], batch["part_vis"], taps, branches)


def unet_generator(p, cfg, emb, pose, taps=None, branches=None):
    """trainer.py:588-590 (spatial broadcast of the embedding) + models.py:518-576.
    emb [B,352]; pose [B,H,W,18].  Returns (G [B,H,W,3], z [B,z_num])."""
    w = _Walker("ID_AE/G", p, branches)
    B = emb.shape[0]
    H, W = pose.shape[1], pose.shape[2]
    hn, rn = cfg.hidden, cfg.unet_repeat
    emb_rep = emb[:, None, None, :].expand(B, H, W, emb.shape[1])
    x = torch.cat([emb_rep, pose], dim=3)
    x = _tap(taps, "g0", w.conv(x))
    skips = []
    for idx in range(rn):
        res = x
        x = _tap(taps, "genc_a%d" % idx, w.conv(x))
        x = w.conv(x)
        x = _tap(taps, "genc_y%d" % idx, x + res)
        skips.append(x)
        if idx < rn - 1:
            x = _tap(taps, "genc_x%d" % (idx + 1), w.conv(x, stride=2))
    sh = x.shape
    z = x = w.fc(x.reshape(B, -1))
    x = w.fc(z).reshape(B, sh[1], sh[2], hn)
    for idx in range(rn):
        x = _tap(taps, "cat%d" % idx, torch.cat([x, skips[rn - 1 - idx]], dim=-1))
        res = x
        x = _tap(taps, "dec_a%d" % idx, w.conv(x))
        x = w.conv(x)
        x = _tap(taps, "dec_y%d" % idx, x + res)
        if idx < rn - 1:
            x = T.upscale2(x)
            x = w.conv(x)  # 1x1 (the walker reads the kernel size off the weights)
    out = _tap(taps, "G", w.conv(x, act=False))
    return out, z


def dcgan_discriminator(p, cfg, x_nhwc, mode="dcgan", lrelu_signs=None, record=None):
    """wgan_gp.py:407-440 on an NHWC image (the reference transposes to NCHW first, trainer.py:601-602;
    the only place the layout matters is the flatten before the Linear, done C-major below).
    lrelu_signs (test aid): four bool NHWC tensors, `pre-activation > 0` per layer as the implementation under test
    decided it.  LeakyReLU then takes its branch from them, so that a gradient comparison differentiates the same
    piecewise-linear function on both sides; a pre-activation within fp32 rounding of zero otherwise flips the branch
    in one of the two and moves the gradients by O(1/sqrt(#activations)) (tests/probe_grad_flake.py)."""
    norm = T.layernorm if mode == "wgan-gp" else T.batchnorm_train  # wgan_gp.py:34-40

    def lrelu(z, i):
        if record is not None:     # a list: receives this run's own decisions
            record.append(z.detach() > 0)
        return T.leaky_relu(z) if lrelu_signs is None else torch.where(lrelu_signs[i], z, 0.2 * z)
    h = T.conv2d_same(x_nhwc, p["Discriminator.1.Filters"], p["Discriminator.1.Biases"], 2)
    h = lrelu(h, 0)
    for i in (2, 3, 4):
        h = T.conv2d_same(h, p["Discriminator.%d.Filters" % i], p["Discriminator.%d.Biases" % i], 2)
        h = norm(h, p["Discriminator.BN%d.scale" % i], p["Discriminator.BN%d.offset" % i])
        h = lrelu(h, i - 1)
    # NCHW flatten (index = c*(h*w) + y*w + x), then tf.reshape(output, [-1, 8*4*8*dim]) (wgan_gp.py:433): one row per
    # image at 128x64; at 256x256 the 16x16x512 map becomes 8 rows per image (64 channels each), i.e. 8 logits (q5)
    flat = h.permute(0, 3, 1, 2).reshape(-1, cfg.d_row)
    out = flat @ p["Discriminator.Output.W"] + p["Discriminator.Output.b"]
    return out.reshape(-1)


def fc_discriminator(p, x, n_layers=3, name=""):
    """wgan_gp.py:399-405."""
    h = T.leaky_relu(x @ p[name + "Discriminator.Input.Linear.W"] + p[name + "Discriminator.Input.Linear.b"])
    for i in range(n_layers):
        h = T.leaky_relu(h @ p[name + "Discriminator.%d.Linear.W" % i] + p[name + "Discriminator.%d.Linear.b" % i])
    return (h @ p[name + "Discriminator.Out.W"] + p[name + "Discriminator.Out.b"]).reshape(-1)


def gaussian_fc_res(p, z, repeat_num=4, prefix="G_FC", act=torch.relu):
    """models.py:474-486 with the noise z supplied by the caller."""
    w = _Walker(prefix, p)
    z = act(w.fc(z))
    for _ in range(repeat_num):
        res = z
        z = act(w.fc(z))
        z = act(w.fc(z))
        z = res + z
    return w.fc(z)


# --------------------------------------------------------------------------------- Stage-I model
def stage1_forward(p, cfg, batch, mode="dcgan", gp_alpha=None, lam=10.0, taps=None, branches=None):
    """build_model of --model=1 (trainer.py:568-625).  batch: dict x, pose, mask, part_bbox, part_vis.
    Returns dict with emb, z, G, D_real, D_fake, g_loss (incl. 20*L1), d_loss, L1.
    branches (test aid): the ReLU / LeakyReLU decisions of the implementation under test -- keys "Encoder/G_encoder",
    "ID_AE/G" (see _Walker) and "D_real", "D_fake" (or "D_pair"), "D_hat" (see dcgan_discriminator).  The graph is
    piecewise linear in its activations; a gradient comparison means something only on the same piece."""
    x = batch["x"]
    br = branches or {}
    rec = (lambda k: br["record"].setdefault(k, [])) if "record" in br else (lambda k: None)
    emb = encoder(p, cfg, batch, taps, branches)
    G, z = unet_generator(p, cfg, emb, batch["pose"], taps, branches)
    if cfg.d_joint:   # trainer_256.py:61-66: one call on concat([x, G]) (joint batch statistics), then tf.split(D_z, 2)
        d_both = dcgan_discriminator(p, cfg, torch.cat([x, G], dim=0), mode, br.get("D_pair"), rec("D_pair"))
        d_real, d_fake = d_both[:d_both.shape[0] // 2], d_both[d_both.shape[0] // 2:]
    else:             # trainer.py:601-602: two calls
        d_real = dcgan_discriminator(p, cfg, x, mode, br.get("D_real"), rec("D_real"))
        d_fake = dcgan_discriminator(p, cfg, G, mode, br.get("D_fake"), rec("D_fake"))
    g_gan, d_loss = T.gan_loss(mode, d_real, d_fake)
    out = dict(emb=emb, z=z, G=G, D_real=d_real, D_fake=d_fake)
    if mode == "wgan-gp" and gp_alpha is not None:
        gp, slopes, _ = T.gradient_penalty(lambda t: dcgan_discriminator(p, cfg, t, mode, br.get("D_hat"), rec("D_hat")),
                                           x, G, gp_alpha)
        d_loss = d_loss + lam * gp
        out.update(gp=gp, slopes=slopes)
    l1 = (G - x).abs().mean()
    out.update(L1=l1, g_loss_only=g_gan, g_loss=g_gan + 20.0 * l1, d_loss=d_loss)
    return out


def stage1_grads(p, cfg, batch, which, mode="dcgan", gp_alpha=None, branches=None):
    """Gradients of g_loss w.r.t. Encoder+G params (which='g') or of d_loss w.r.t. D params (which='d');
    what Optimizer.minimize(var_list=...) differentiates (trainer.py:622-625)."""
    out = stage1_forward(p, cfg, batch, mode, gp_alpha, branches=branches)
    if which == "g":
        names = [k for k in p if is_generator_param(k)]
        loss = out["g_loss"]
    else:
        names = [k for k in p if is_disc_param(k)]
        loss = out["d_loss"]
    grads = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
    return out, OrderedDict((k, g) for k, g in zip(names, grads))


class Stage1Trainer:
    """The optimiser half of the step (trainer.py:116-149, 336-347): Adam(b1=.5) in dcgan mode,
    Adam(b1=.5,b2=.9) in wgan-gp, RMSProp + clip in wgan."""

    def __init__(self, params, cfg, mode="dcgan", g_lr=2e-5, d_lr=2e-5, dtype=torch.float32):
        self.cfg, self.mode, self.g_lr, self.d_lr = cfg, mode, g_lr, d_lr
        self.p = to_torch(params, dtype, requires_grad=True)
        self.m = {k: torch.zeros_like(v) for k, v in self.p.items()}
        init_v = torch.ones_like if mode in ("wgan", "lsgan") else torch.zeros_like
        self.v = {k: init_v(v) for k, v in self.p.items()}
        self.t = {"g": 0, "d": 0}

    def _apply(self, grads, which):
        self.t[which] += 1
        lr = self.g_lr if which == "g" else self.d_lr
        with torch.no_grad():
            for k, g in grads.items():
                if g is None:
                    continue
                if self.mode in ("wgan", "lsgan"):
                    clip = 0.01 if (self.mode == "wgan" and which == "d") else None
                    T.rmsprop_step(self.p[k], g, self.v[k], lr, clip=clip)
                else:
                    b2 = 0.9 if self.mode == "wgan-gp" else 0.999
                    T.adam_step(self.p[k], g, self.m[k], self.v[k], lr, self.t[which], 0.5, b2)

    def g_step(self, batch, gp_alpha=None):
        out, grads = stage1_grads(self.p, self.cfg, batch, "g", self.mode, gp_alpha)
        self._apply(grads, "g")
        return out, grads

    def d_step(self, batch, gp_alpha=None):
        out, grads = stage1_grads(self.p, self.cfg, batch, "d", self.mode, gp_alpha)
        self._apply(grads, "d")
        return out, grads


# --------------------------------------------------------------------------------- Stage-II (--model=3)
def init_stage2_params(fg_dim=224, bg_dim=128, seed=4321, bias_noise=0.0):
    """Parameters of the two GaussianFCRes nets (scopes Gaussian_FC_Fg / Gaussian_FC_Bg, trainer.py:752-758;
    slim xavier_uniform) and the two FC critics (names 'Fg_FCDis_' / 'Bg_FCDis_', trainer.py:764-775; tflib Linear
    'he' for the LeakyReLULayers wgan_gp.py:30-32, glorot for .Out, tflib/ops/linear.py:36-66)."""
    rng = np.random.default_rng(seed)
    p = OrderedDict()
    for scope, dim, hid in (("Gaussian_FC_Fg/G_FC", fg_dim, 512), ("Gaussian_FC_Bg/G_FC", bg_dim, 256)):
        s = _Scope(scope, p, rng)
        s.fc(dim, hid)
        for _ in range(8):
            s.fc(hid, hid)
        s.fc(hid, dim)
    for name, dim in (("Fg_FCDis_", fg_dim), ("Bg_FCDis_", bg_dim)):
        dims = [("Input", dim, 512)] + [(str(i), 512, 512) for i in range(3)]
        for tag, a, b in dims:
            std = math.sqrt(2.0 / a)
            p[name + "Discriminator.%s.Linear.W" % tag] = rng.uniform(-std * math.sqrt(3), std * math.sqrt(3), size=(a, b)).astype(np.float32)
            p[name + "Discriminator.%s.Linear.b" % tag] = np.zeros(b, np.float32)
        std = math.sqrt(2.0 / (512 + 1))
        p[name + "Discriminator.Out.W"] = rng.uniform(-std * math.sqrt(3), std * math.sqrt(3), size=(512, 1)).astype(np.float32)
        p[name + "Discriminator.Out.b"] = np.zeros(1, np.float32)
    if bias_noise > 0:
        for k in p:
            if k.endswith(("biases", ".b")):
                p[k] = (p[k] + rng.normal(0, bias_noise, size=p[k].shape)).astype(np.float32)
    return p


def stage2_losses(p, factor, real, z, mode="wgan"):
    """g_loss_embs / d_loss_embs of one factor (trainer.py:752-775): fake = GaussianFCRes(z) with
    activation_fn=LeakyReLU; critic = FCDiscriminator on real and fake embeddings."""
    if factor == "pose":     # --model=4: PoseGaussian sampler, critic 'Pose_emb_' (trainer.py:893-910)
        scope, name = "PoseGaussian/G_FC", "Pose_emb_"
    elif factor == "app":    # --model=103: one appearance sampler, critic 'FCDis_' (trainer_256.py:324-334)
        scope, name = "Gaussian_FC/G_FC", "FCDis_"
    else:
        scope = "Gaussian_FC_%s/G_FC" % ("Fg" if factor == "fg" else "Bg")
        name = "Fg_FCDis_" if factor == "fg" else "Bg_FCDis_"
    fake = gaussian_fc_res(p, z, repeat_num=4, prefix=scope, act=lambda t: T.leaky_relu(t, 0.2))
    d_real = fc_discriminator(p, real, name=name)
    d_fake = fc_discriminator(p, fake, name=name)
    g_loss, d_loss = T.gan_loss(mode, d_real, d_fake)
    return dict(fake=fake, d_real=d_real, d_fake=d_fake, g_loss=g_loss, d_loss=d_loss)


# --------------------------------------------------------------------------------- sampling (--model=13)
def pose_encoder_fc_res(p, pose_rcv_norm, prefix="PoseAE/G_Pose_Encoder", repeat_num=4, act=None):
    """models.PoseEncoderFCRes (models.py:488-499): residual MLP 54 -> 512 -> ... -> 32."""
    act = act or (lambda t: T.leaky_relu(t, 0.2))
    w = _Walker(prefix, p)
    x = act(w.fc(pose_rcv_norm))
    for _ in range(repeat_num):
        res = x
        x = act(w.fc(x))
        x = act(w.fc(x))
        x = res + x
    return w.fc(x)


def pose_decoder_fc_res(p, z, prefix="PoseAE/G_Pose_Decoder", repeat_num=4, act=None):
    """models.PoseDecoderFCRes (models.py:501-515): first FC without activation, residual blocks, then the
    coordinate head (no activation) and the visibility head sigmoid + binaryRound (models.py:97-108)."""
    act = act or (lambda t: T.leaky_relu(t, 0.2))
    w = _Walker(prefix, p)
    x = w.fc(z)
    for _ in range(repeat_num):
        res = x
        x = act(w.fc(x))
        x = act(w.fc(x))
        x = res + x
    coord = w.fc(x)
    vis = torch.round(torch.sigmoid(w.fc(x)))
    return coord, vis


def pose_ae_loss(p, pose_rcv_norm):
    """reconstruct_loss of --model=2 (trainer.py:638-658): encode, decode, G_pose_rcv = (coord, binaryRound(sigmoid))
    with the straight-through gradient of binaryRound (models.py:97-108: Round -> Identity), mean squared error over
    [B,18,3].  Returns (loss, G_pose_rcv)."""
    B, K = pose_rcv_norm.shape[0], pose_rcv_norm.shape[1]
    z = pose_encoder_fc_res(p, pose_rcv_norm.reshape(B, -1))
    act = lambda t: T.leaky_relu(t, 0.2)  # noqa: E731
    w = _Walker("PoseAE/G_Pose_Decoder", p)
    x = w.fc(z)
    for _ in range(4):
        res = x
        x = act(w.fc(x))
        x = act(w.fc(x))
        x = res + x
    coord = w.fc(x)
    sg = torch.sigmoid(w.fc(x))
    vis = sg + (torch.round(sg) - sg).detach()
    g = torch.cat([coord.reshape(B, K, 2), vis[:, :, None]], dim=-1)
    return ((pose_rcv_norm - g) ** 2).mean(), g


def init_pose_params(keypoints=18, seed=777, bias_noise=0.0):
    rng = np.random.default_rng(seed)
    p = OrderedDict()
    s = _Scope("PoseAE/G_Pose_Encoder", p, rng)
    s.fc(keypoints * 3, 512)
    for _ in range(8):
        s.fc(512, 512)
    s.fc(512, 32)
    s = _Scope("PoseAE/G_Pose_Decoder", p, rng)
    s.fc(32, 512)
    for _ in range(8):
        s.fc(512, 512)
    s.fc(512, keypoints * 2)
    s.fc(512, keypoints)
    if bias_noise > 0:
        for k in p:
            if k.endswith("biases"):
                p[k] = (p[k] + rng.normal(0, bias_noise, size=p[k].shape)).astype(np.float32)
    return p


def sample_factor_forward(p, cfg, batch, z_fg, z_bg, sample_fg, sample_bg, sample_pose, mode="dcgan"):
    """DPIG_FourNetsFgBg_testOnlySampleFactor.build_model (tester.py:473-571): sample-or-hold each factor, decode the
    (encoded real) pose when sample_pose (quirk q7), inflate, generate, denorm, score with D."""
    x = batch["x"]
    B, H, W = x.shape[0], cfg.img_h, cfg.img_w
    rcv = batch["pose_rcv"]
    norm = torch.stack([rcv[:, :, 0] / float(H) * 2.0 - 1, rcv[:, :, 1] / float(W) * 2.0 - 1, rcv[:, :, 2]], dim=-1)
    pose_embs = pose_encoder_fc_res(p, norm.reshape(B, -1))
    coord, vis = pose_decoder_fc_res(p, pose_embs)
    if sample_pose:
        g_rcv = torch.cat([coord.reshape(B, cfg.keypoints, 2), vis[:, :, None]], dim=-1)
    else:
        g_rcv = norm[:1].expand(B, -1, -1)
    # coord2channel_simple_rcv(is_normalized=True) (utils.py:259-273): back to pixels, clamped into the image
    R = torch.clamp((g_rcv[:, :, 0] + 1) / 2.0 * H, 0, H - 1)
    C = torch.clamp((g_rcv[:, :, 1] + 1) / 2.0 * W, 0, W - 1)
    pix = torch.stack([R, C, g_rcv[:, :, 2]], dim=-1)
    pose_maps = T.pose_rasterize(pix, H, W, 4)
    emb = encoder(p, cfg, batch)
    nfg = cfg.n_parts * cfg.part_z
    app_fg = gaussian_fc_res(p, z_fg, 4, "Gaussian_FC_Fg/G_FC", lambda t: T.leaky_relu(t, 0.2))
    app_bg = gaussian_fc_res(p, z_bg, 4, "Gaussian_FC_Bg/G_FC", lambda t: T.leaky_relu(t, 0.2))
    e_fg = app_fg if sample_fg else emb[:1, :nfg].expand(B, -1)
    e_bg = app_bg if sample_bg else emb[:1, nfg:].expand(B, -1)
    G, _ = unet_generator(p, cfg, torch.cat([e_fg, e_bg], dim=-1), pose_maps)
    score = dcgan_discriminator(p, cfg, G, mode)
    return dict(G=T.denorm_img(G), G_raw=G, score=score, pose_pix=pix, pose_maps=pose_maps)


def four_nets_forward(p, cfg, batch, z_fg, z_bg, sample_app, one_app_per_batch, sample_pose, mode="dcgan"):
    """DPIG_FourNetsFgBg_testOnly.build_model (tester.py:323-417, --model=11): `sample_app` replaces the whole
    appearance embedding by the two GaussianFCRes outputs (the Fg row of sample 0 tiled when `one_app_per_batch`),
    otherwise the encoder embedding is kept (first sample's Fg part tiled when `one_app_per_batch`); without
    `sample_pose` every sample keeps its own real pose (tester.py:349-351)."""
    x = batch["x"]
    B, H, W = x.shape[0], cfg.img_h, cfg.img_w
    rcv = batch["pose_rcv"]
    norm = torch.stack([rcv[:, :, 0] / float(H) * 2.0 - 1, rcv[:, :, 1] / float(W) * 2.0 - 1, rcv[:, :, 2]], dim=-1)
    coord, vis = pose_decoder_fc_res(p, pose_encoder_fc_res(p, norm.reshape(B, -1)))
    g_rcv = torch.cat([coord.reshape(B, cfg.keypoints, 2), vis[:, :, None]], dim=-1) if sample_pose else norm
    R = torch.clamp((g_rcv[:, :, 0] + 1) / 2.0 * H, 0, H - 1)
    C = torch.clamp((g_rcv[:, :, 1] + 1) / 2.0 * W, 0, W - 1)
    pix = torch.stack([R, C, g_rcv[:, :, 2]], dim=-1)
    pose_maps = T.pose_rasterize(pix, H, W, 4)
    emb = encoder(p, cfg, batch)
    nfg = cfg.n_parts * cfg.part_z
    lrelu = lambda t: T.leaky_relu(t, 0.2)
    if sample_app:
        fg = gaussian_fc_res(p, z_fg, 4, "Gaussian_FC_Fg/G_FC", lrelu)
        bg = gaussian_fc_res(p, z_bg, 4, "Gaussian_FC_Bg/G_FC", lrelu)
    else:
        fg, bg = emb[:, :nfg], emb[:, nfg:]
    if one_app_per_batch:
        fg = fg[:1].expand(B, -1)
    G, _ = unet_generator(p, cfg, torch.cat([fg, bg], dim=-1), pose_maps)
    score = dcgan_discriminator(p, cfg, G, mode)
    return dict(G=T.denorm_img(G), G_raw=G, score=score, pose_pix=pix, pose_maps=pose_maps)


def condition_forward(p, cfg, batch, pose_rcv_target, mode="dcgan"):
    """DPIG_FourNetsFgBg_testOnlyCondition.build_model (tester.py:657-686, --model=12; _256 form tester.py:815-836):
    appearance of x, TARGET pose maps (rasterised + inflated keypoints of the second image of the pair)."""
    pose_t = T.pose_rasterize(pose_rcv_target, cfg.img_h, cfg.img_w, 4)
    emb = encoder(p, cfg, batch)
    G, _ = unet_generator(p, cfg, emb, pose_t)
    score = None if cfg.d_joint else dcgan_discriminator(p, cfg, G, mode)
    return dict(G=T.denorm_img(G), G_raw=G, score=score, pose_maps=pose_t)
