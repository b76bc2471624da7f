"""Sampling / inference call surface of the reference (--model=13):

    class DPIG_FourNetsFgBg_testOnlySampleFactor      reference tester.py:419-613

Forward only: appearance encoder -> sample-or-hold the Fg / Bg embeddings (GaussianFCRes nets) -> pose branch
(PoseEncoderFCRes -> PoseDecoderFCRes; the decoder is fed the ENCODED REAL pose, quirk q7) -> keypoints to maps +
radius-4 inflation -> U-Net -> denorm -> DCGANDiscriminator score.  Everything numeric runs through the C ABI, incl. the
per-sample SSIM of generate() (dpig_ssim_gray_u8) and the uint8 conversion; test() writes the reference's result
directories (x, x_target, G, pose, pose_target, G_pose, mask, mask_target -- the input of score.py) through
outputs.ResultWriter, PNG encoding batched over host threads."""
import os

import numpy as np
import torch

from . import _lib, engine, outputs, stage2, synth, tf_checkpoint
from ._lib import ACT_LRELU, ACT_NONE
from .tensor import ptr
from .trainer import SyntheticLoader, make_loader


def pose_specs(keypoints=18):
    specs = stage2.fc_res_specs("PoseAE/G_Pose_Encoder", keypoints * 3, 512, 32)
    dec = [("PoseAE/G_Pose_Decoder", 32, 512)] + [("PoseAE/G_Pose_Decoder", 512, 512)] * 8 + \
          [("PoseAE/G_Pose_Decoder", 512, keypoints * 2), ("PoseAE/G_Pose_Decoder", 512, keypoints)]
    for i, (pre, a, b) in enumerate(dec):
        name = "%s/fully_connected%s" % (pre, "" if i == 0 else "_%d" % i)
        specs += [(name + "/weights", (a, b)), (name + "/biases", (b,))]
    return specs


class DPIG_FourNetsFgBg_testOnlySampleFactor(object):
    def __init__(self, config, loader=None):
        self.config = config
        self.batch_size = config.batch_size
        self.img_H, self.img_W = config.img_H, config.img_W
        self.conv_hidden_num, self.z_num = config.conv_hidden_num, config.z_num
        self.sample_fg, self.sample_bg, self.sample_pose = config.sample_fg, config.sample_bg, config.sample_pose
        self.encode_unused = os.environ.get("DPIG_TESTER_ENCODE_UNUSED", "0") == "1"
        self.keypoint_num = 18
        self.model_dir = config.model_dir or os.path.join(config.log_dir, "dpig_model%d" % config.model)
        self.pretrained_path = config.pretrained_path
        self.test_batch_num = 400        # tester.py:475
        self.loader = loader or make_loader(config, self.batch_size, self.img_H, self.img_W)
        self.net_cfg = None

    def init_net(self, net_cfg=None):
        self.ctx = _lib.Context(0)
        cfg = net_cfg or engine.NetConfig(img_h=self.img_H, img_w=self.img_W, hidden=self.conv_hidden_num, z_num=self.z_num)
        self.cfg = cfg
        dev = torch.device("cuda", 0)
        B = self.batch_size
        self.s1 = engine.Stage1Engine(self.ctx, cfg, B, mode="dcgan", inference=True)
        self.s1.load_params(engine.init_params(cfg, seed=self.config.random_seed))
        self.s2 = stage2.Stage2Engine(self.s1, mode="wgan")
        self.s2.load_params(stage2.init_stage2_params())
        self._build_pose_tape(B, dev)
        # the four partial restores of tester.py:423-472 (Encoder+ID_AE, Gaussian_FC_*, PoseAE, Discriminator.): any of
        # --pretrained_path, --pretrained_appSample_path, --pretrained_poseAE_path (TensorFlow V2 checkpoints or .npz)
        for path in (self.pretrained_path, getattr(self.config, "pretrained_appSample_path", None),
                     getattr(self.config, "pretrained_poseAE_path", None)):
            if path:
                self.load_params(tf_checkpoint.load_any(path))

    def _build_pose_tape(self, B, dev):
        # pose auto-encoder tapes (models.py:488-515), activation_fn=LeakyReLU (tester.py:487, 495)
        self.pp = engine.ParamGroup(pose_specs(self.keypoint_num), dev)
        names = list(self.pp.specs)
        lay = [(names[2 * i], names[2 * i + 1]) for i in range(len(names) // 2)]
        enc, dec = lay[:10], lay[10:]
        t = self.pose_tape = stage2.FCTape(self.ctx, self.pp, B, dev)
        self.pose_in = stage2._Node(B, self.keypoint_num * 3, dev)
        h = t.linear(self.pose_in, *enc[0], act=ACT_LRELU)
        for r in range(4):
            a = t.linear(h, *enc[1 + 2 * r], act=ACT_LRELU)
            b = t.linear(a, *enc[2 + 2 * r], act=ACT_LRELU)
            h = t.add(h, b)
        self.pose_emb = t.linear(h, *enc[9])
        h = t.linear(self.pose_emb, *dec[0])
        for r in range(4):
            a = t.linear(h, *dec[1 + 2 * r], act=ACT_LRELU)
            b = t.linear(a, *dec[2 + 2 * r], act=ACT_LRELU)
            h = t.add(h, b)
        self.pose_coord = t.linear(h, *dec[9])
        self.pose_vis_logit = t.linear(h, *dec[10])
        self.p_pose = t.forward_program()

    def load_params(self, params):
        """Parameters by TF variable name: Encoder/..., ID_AE/..., Discriminator.*, Gaussian_FC_Fg/..., Gaussian_FC_Bg/...,
        PoseAE/... (the four partial checkpoints of tester.py:423-472 merged into one dict)."""
        self.s1.load_params(params)
        self.s2.load_params(params)
        for name in self.pp.specs:
            if name in params:
                self.pp.view(name).copy_(torch.as_tensor(np.asarray(params[name]), dtype=torch.float32).cuda())

    def generate(self, x_fixed, pose_fixed, pose_rcv_fixed, part_bbox_fixed, part_vis_fixed, root_path=None, path=None,
                 idx=None, save=False, mask=None, z_fg=None, z_bg=None):
        """Returns (G in [0,255] NHWC float32, G_pose_inflated_img, G_dis_score) like tester.py:573-613."""
        s1, s2, cfg, B = self.s1, self.s2, self.cfg, self.batch_size
        H, W = cfg.img_h, cfg.img_w
        st = torch.cuda.current_stream().cuda_stream
        if mask is None:
            mask = np.ones((B, H, W, 1), np.float32)
        s1.set_batch(dict(x=x_fixed, pose_rcv=pose_rcv_fixed, mask=mask, part_bbox=part_bbox_fixed, part_vis=part_vis_fixed))
        G, pose_img, score, ssim = self.generate_on_device(z_fg, z_bg)
        self.last_ssim = ssim.cpu().numpy()
        G_np, pose_np = G.cpu().numpy(), pose_img.cpu().numpy()
        if save and root_path is not None:                       # tester.py:242-252
            ssim_mean = float(np.mean(self.last_ssim))
            outputs.save_image(G_np, path or os.path.join(root_path, "%s_G_ssim%s.png" % (idx, ssim_mean)))
            outputs.save_image(pose_np, os.path.join(root_path, "%s_G_pose_inflated_reLoss%s.png" % (idx, 0.0)))
        return G_np, pose_np, score.cpu().numpy()

    def generate_on_device(self, z_fg=None, z_bg=None):
        """The device half of generate() for the batch already in HBM (Stage1Engine.set_batch): returns DEVICE tensors
        (G in [0,255] NHWC float32, inflated-pose image, critic score [B], per-sample SSIM(G, x) [B])."""
        s1, s2, cfg, B = self.s1, self.s2, self.cfg, self.batch_size
        H, W = cfg.img_h, cfg.img_w
        st = torch.cuda.current_stream().cuda_stream
        # ---- pose branch (tester.py:477-505)
        rcv = s1.pose_rcv
        norm = torch.stack([rcv[:, :, 0] / float(H) * 2.0 - 1, rcv[:, :, 1] / float(W) * 2.0 - 1, rcv[:, :, 2]], dim=-1)
        self.pose_in.data.copy_(norm.reshape(B, -1))
        self.p_pose.run(st)
        if self.sample_pose:
            vis = torch.round(torch.sigmoid(self.pose_vis_logit.data))            # binaryRound (models.py:97-108)
            g = torch.cat([self.pose_coord.data.reshape(B, self.keypoint_num, 2), vis[:, :, None]], dim=-1)
        else:
            g = self._held_pose(norm)
        R = torch.clamp((g[:, :, 0] + 1) / 2.0 * H, 0, H - 1)                      # utils.py:266-271
        Cc = torch.clamp((g[:, :, 1] + 1) / 2.0 * W, 0, W - 1)
        s1.pose_rcv.copy_(torch.stack([R, Cc, g[:, :, 2]], dim=-1))
        # ---- appearance branch (tester.py:509-554)
        self._appearance_branch(st, z_fg, z_bg)
        self._fill_embedding()
        # ---- U-Net, denorm, critic score (tester.py:561-571)
        s1.run_unet(st)
        score = self._score(st)
        G = torch.clamp((s1.G + 1.0) * 127.5, 0, 255)
        pose_maps = s1.gin.slice(0, cfg.keypoints).hi.float()
        pose_img = (pose_maps.amax(dim=-1, keepdim=True).expand(-1, -1, -1, 3) + 1) * 127.5
        # ---- SSIM(G, x) per sample on the uint8 images (tester.py:236-241), on the device
        return G, pose_img, score, self.ssim_G_x(st, to_host=False)

    def _held_pose(self, norm):
        """sample_pose=False: the first sample's real pose for the whole batch (tester.py:500-503)."""
        return norm[:1].expand(self.batch_size, -1, -1)

    def _appearance_branch(self, st, z_fg, z_bg):
        """tester.py:509-554.  The encoder is an ancestor of the fetched G only through a HELD factor (the first sample's
        embedding tiled over the batch, tester.py:539, 550): with --sample_fg and --sample_bg both set (run_market_test.sh:
        64-80) tf.Session.run never executes it, and neither does this.  One held factor needs its own pyramid only
        (Stage2Engine.prune).  encode_unused=True (DPIG_TESTER_ENCODE_UNUSED=1) runs the whole encoder regardless --
        BASELINE.json's configs[4] counts it in the sampling pass, bench.py measures both."""
        s2 = self.s2
        held = self._held_factors()
        if self.encode_unused or len(held) == 2:
            s2.encode_real()
        elif held:
            s2.encode_real(held[0])
        for factor, z in (("fg", z_fg), ("bg", z_bg)):
            s2.sample_noise(factor, z)
            s2.f[factor].p_g_fwd.run(st)

    def _held_factors(self):
        """The appearance factors whose embedding comes from the encoder rather than from a sampler."""
        return [f for f, sampled in (("fg", self.sample_fg), ("bg", self.sample_bg)) if not sampled]

    def _score(self, st):
        return self.s1.score_generated(st)

    def _fill_embedding(self):
        """Which appearance factors reach the U-Net (tester.py:540-554): a sampled factor takes the GaussianFCRes
        output, a held one the first sample's encoder embedding tiled over the batch."""
        s1, s2, B = self.s1, self.s2, self.batch_size
        nfg = s2.fg_dim
        for factor, sample in (("fg", self.sample_fg), ("bg", self.sample_bg)):
            f = s2.f[factor]
            sl = slice(0, nfg) if factor == "fg" else slice(nfg, None)
            if sample:
                s1.emb[:, sl].copy_(f.fake.data)
            else:
                s1.emb[:, sl].copy_(f.real.data[:1].expand(B, -1))

    def ssim_G_x(self, stream=None, to_host=True):
        """skimage-style SSIM between the generated and the input images of the current batch, per sample [B]."""
        s1, B = self.s1, self.batch_size
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        H, W = self.cfg.img_h, self.cfg.img_w
        g8 = torch.empty((B, H, W, 3), dtype=torch.uint8, device=s1.device)
        x8 = torch.empty((B, H, W, 3), dtype=torch.uint8, device=s1.device)
        out = torch.empty((B,), dtype=torch.float32, device=s1.device)
        self.ctx.denorm_u8(ptr(s1.G), s1.G.numel(), ptr(g8), st)
        self.ctx.denorm_u8(ptr(s1.x), s1.x.numel(), ptr(x8), st)
        self.ctx.ssim_gray_u8(ptr(g8), ptr(x8), B, H, W, ptr(out), st)
        return out.cpu().numpy() if to_host else out

    def _pose_max_img(self, pose_rcv):
        """(amax over the 18 inflated keypoint maps + 1) * 127.5 (tester.py:175-176) for a [B,18,3] keypoint array."""
        B, H, W = self.batch_size, self.cfg.img_h, self.cfg.img_w
        rcv = torch.as_tensor(np.asarray(pose_rcv, np.float32)).to(self.s1.device)
        maps = torch.empty((B, H, W, self.keypoint_num), dtype=torch.float32, device=self.s1.device)
        self.ctx.pose_rasterize(ptr(rcv), B, self.keypoint_num, H, W, 4, None, ptr(maps),
                                torch.cuda.current_stream().cuda_stream)
        return ((maps.amax(dim=-1) + 1.0) * 127.5).cpu().numpy()

    def test(self, num_batches=None):
        """tester.py:138-202: the eight per-sample PNG directories + the sample sheets of the first batch."""
        out_dir = os.path.join(self.model_dir, "test_result_SampleFg%rSampleBg%rSamplePose%r" % (
            self.sample_fg, self.sample_bg, self.sample_pose))
        wr = outputs.ResultWriter(out_dir)
        B = self.batch_size
        for i in range(num_batches or self.test_batch_num):
            b = self.loader.next_batch()
            G, G_pose, score = self.generate(b["x"], None, b["pose_rcv"], b["part_bbox"], b["part_vis"], mask=b["mask"],
                                             root_path=out_dir, idx=i, save=i < 4)
            x255 = (np.asarray(b["x"]) + 1.0) * 127.5                         # unprocess_image (tester.py:207)
            xt255 = (np.asarray(b.get("x_target", b["x"])) + 1.0) * 127.5
            mask255 = np.asarray(b["mask"]) * 255.0
            maskt255 = np.asarray(b.get("mask_target", b["mask"])) * 255.0
            p = self._pose_max_img(b["pose_rcv"])
            pt = self._pose_max_img(b.get("pose_rcv_target", b["pose_rcv"]))
            wr.add_batch(i, B, x255, xt255, G, p, pt, G_pose, mask255, maskt255, np.asarray(score).reshape(-1))
            if i == 0:
                wr.add_grid(x255, "x_fixed.png")
                wr.add_grid(xt255, "x_target_fixed.png")
                wr.add_grid(mask255, "mask_fixed.png")
                wr.add_grid(maskt255, "mask_target_fixed.png")
                wr.add_grid(p[..., None], "pose_fixed.png")
                wr.add_grid(pt[..., None], "pose_target_fixed.png")
        self.files_written = wr.close()
        return out_dir


class DPIG_FourNetsFgBg_testOnly(DPIG_FourNetsFgBg_testOnlySampleFactor):
    """--model=11 (tester.py:256-417): the same four networks with the `--sample_app` / `--one_app_per_batch` switches
    instead of per-factor ones.
      sample_app:        embedding = [Gaussian_FC_Fg(z) (its first row tiled if one_app_per_batch), Gaussian_FC_Bg(z)]
      otherwise:         the encoder embedding; with one_app_per_batch the first sample's Fg part tiled, own Bg parts
      sample_pose=False: every sample keeps its OWN real pose (tester.py:349-351), unlike model 13."""

    def __init__(self, config, loader=None):
        super().__init__(config, loader=loader)
        self.sample_app = getattr(config, "sample_app", False)
        self.one_app_per_batch = getattr(config, "one_app_per_batch", False)
        self.test_batch_num = 400

    def _held_pose(self, norm):
        return norm

    def _held_factors(self):
        return [] if self.sample_app else ["fg", "bg"]

    def _fill_embedding(self):
        s1, s2, B = self.s1, self.s2, self.batch_size
        nfg = s2.fg_dim
        fg, bg = s2.f["fg"], s2.f["bg"]
        if self.sample_app:
            s1.emb[:, :nfg].copy_(fg.fake.data[:1].expand(B, -1) if self.one_app_per_batch else fg.fake.data)
            s1.emb[:, nfg:].copy_(bg.fake.data)
        else:
            s1.emb[:, :nfg].copy_(fg.real.data[:1].expand(B, -1) if self.one_app_per_batch else fg.real.data)
            s1.emb[:, nfg:].copy_(bg.real.data)

    def test(self, num_batches=None):
        out = super().test(num_batches)
        return out


class DPIG_FourNetsFgBg_testOnlyCondition(object):
    """--model=12 (tester.py:616-773): pose-conditioned generation.  Encoder(x, mask, part boxes) -> embedding -> U-Net
    driven by the TARGET pose maps -> G, critic score of G; SSIM against x_target; result directories without G_pose and
    with plain `%05d.png` names for G (tester.py:752)."""
    deepfashion = False

    def __init__(self, config, loader=None):
        self.config = config
        self.batch_size = config.batch_size
        self.img_H, self.img_W = config.img_H, config.img_W
        self.conv_hidden_num, self.z_num = config.conv_hidden_num, config.z_num
        self.keypoint_num = 18
        self.model_dir = config.model_dir or os.path.join(config.log_dir, "dpig_model%d" % config.model)
        self.pretrained_path, self.ckpt_path = config.pretrained_path, config.ckpt_path
        self.test_batch_num = 600        # tester.py:650
        self.test_dir_name = "test_samples_result_ROI7_Condition_TargetPose_%dx%d" % (self.test_batch_num, self.batch_size)
        self.loader = loader or make_loader(config, self.batch_size, self.img_H, self.img_W)

    def init_net(self, net_cfg=None):
        self.ctx = _lib.Context(0)
        if net_cfg is None:
            net_cfg = (engine.NetConfig.deepfashion(img_h=self.img_H, img_w=self.img_W, hidden=self.conv_hidden_num)
                       if self.deepfashion else
                       engine.NetConfig(img_h=self.img_H, img_w=self.img_W, hidden=self.conv_hidden_num, z_num=self.z_num))
        self.cfg = net_cfg
        self.s1 = engine.Stage1Engine(self.ctx, net_cfg, self.batch_size, mode="dcgan", inference=True)
        self.s1.load_params(engine.init_params(net_cfg, seed=self.config.random_seed))
        # saverPart = Encoder + ID_AE + Discriminator. (tester.py:619-623); --ckpt_path restores everything
        if self.pretrained_path:
            self.s1.load_params(tf_checkpoint.load_any(self.pretrained_path, scopes=["Encoder", "ID_AE", "Discriminator."]))
        if self.ckpt_path:
            self.s1.load_params(tf_checkpoint.load_any(self.ckpt_path))

    def load_params(self, params):
        self.s1.load_params(params)

    def generate(self, x_fixed, x_target_fixed, pose_rcv_target_fixed, mask_fixed, part_bbox_fixed, part_vis_fixed,
                 root_path=None, path=None, idx=None, save=True):
        """(G in [0,255] NHWC float32, critic score [B]) like tester.py:688-702; the target pose is given as keypoints
        (pose_peaks_1_rcv) and rasterised + inflated on the device (the reference feeds the ready maps)."""
        s1, B = self.s1, self.batch_size
        H, W = self.cfg.img_h, self.cfg.img_w
        st = torch.cuda.current_stream().cuda_stream
        if mask_fixed is None:
            mask_fixed = np.ones((B, H, W, 1), np.float32)
        s1.set_batch(dict(x=x_fixed, pose_rcv=pose_rcv_target_fixed, mask=mask_fixed, part_bbox=part_bbox_fixed,
                          part_vis=part_vis_fixed))
        s1.forward(with_disc=False)
        score = s1.score_generated(st) if not self.deepfashion else torch.zeros((B,), device=s1.device)
        G = torch.clamp((s1.G + 1.0) * 127.5, 0, 255)
        # SSIM against the TARGET image (tester.py:692-697)
        g8 = torch.empty((B, H, W, 3), dtype=torch.uint8, device=s1.device)
        t8 = torch.empty_like(g8)
        xt = torch.as_tensor(np.asarray(x_target_fixed, np.float32)).to(s1.device).contiguous()
        out = torch.empty((B,), dtype=torch.float32, device=s1.device)
        self.ctx.denorm_u8(ptr(s1.G), s1.G.numel(), ptr(g8), st)
        self.ctx.denorm_u8(ptr(xt), xt.numel(), ptr(t8), st)
        self.ctx.ssim_gray_u8(ptr(g8), ptr(t8), B, H, W, ptr(out), st)
        self.last_ssim = out.cpu().numpy()
        G_np = G.cpu().numpy()
        if save and (path is not None or root_path is not None):
            outputs.save_image(G_np, path or os.path.join(root_path, "%s_G_ssim%s.png" % (idx, float(self.last_ssim.mean()))))
        return G_np, score.reshape(-1).cpu().numpy()

    def _pose_max_img(self, pose_rcv):
        B, H, W = self.batch_size, self.cfg.img_h, self.cfg.img_w
        rcv = torch.as_tensor(np.asarray(pose_rcv, np.float32)).to(self.s1.device)
        maps = torch.empty((B, H, W, self.keypoint_num), dtype=torch.float32, device=self.s1.device)
        self.ctx.pose_rasterize(ptr(rcv), B, self.keypoint_num, H, W, 4, None, ptr(maps),
                                torch.cuda.current_stream().cuda_stream)
        return ((maps.amax(dim=-1) + 1.0) * 127.5).cpu().numpy()

    def test(self, num_batches=None):
        """tester.py:704-773."""
        out_dir = os.path.join(self.model_dir, self.test_dir_name)
        wr = outputs.ResultWriter(out_dir)
        B = self.batch_size
        for i in range(num_batches or self.test_batch_num):
            b = self.loader.next_batch()
            xt = b.get("x_target", b["x"])
            G, score = self.generate(b["x"], xt, b.get("pose_rcv_target", b["pose_rcv"]), b.get("mask"), b["part_bbox"],
                                     b["part_vis"], root_path=out_dir, idx=i, save=(i == 0))
            x255, xt255 = (np.asarray(b["x"]) + 1.0) * 127.5, (np.asarray(xt) + 1.0) * 127.5
            mask = np.asarray(b["mask"]) if "mask" in b else np.zeros((B, self.cfg.img_h, self.cfg.img_w, 1), np.float32)
            maskt = np.asarray(b.get("mask_target", mask))
            p, pt = self._pose_max_img(b["pose_rcv"]), self._pose_max_img(b.get("pose_rcv_target", b["pose_rcv"]))
            for j in range(B):
                idx = i * B + j
                for d, arr in (("x", x255[j]), ("x_target", xt255[j]), ("G", G[j]), ("pose", p[j]), ("pose_target", pt[j]),
                               ("mask", np.squeeze(mask[j] * 255.0)), ("mask_target", np.squeeze(maskt[j] * 255.0))):
                    wr._submit(arr, "%s/%s/%05d.png" % (out_dir, d, idx))
            if i == 0:
                wr.add_grid(x255, "x_fixed.png")
                wr.add_grid(xt255, "x_target_fixed.png")
                wr.add_grid(mask * 255.0, "mask_fixed.png")
                wr.add_grid(maskt * 255.0, "mask_target_fixed.png")
                wr.add_grid(p[..., None], "pose_fixed.png")
                wr.add_grid(pt[..., None], "pose_target_fixed.png")
        self.files_written = wr.close()
        return out_dir


class DPIG_ThreeNetsApp_testOnlyCondition_256(DPIG_FourNetsFgBg_testOnlyCondition):
    """--model=1001 (tester.py:775-915): the DeepFashion 256x256 form -- encoder without mask / background branch
    (GeneratorCNN_ID_Encoder_BodyROIVis, roi 64), 5-level U-Net, no critic in the graph."""
    deepfashion = True

    def __init__(self, config, loader=None):
        super().__init__(config, loader=loader)
        self.test_dir_name = "test_result_ROI7_Condition_TargetPose_%dx%d" % (self.test_batch_num, self.batch_size)


class DPIG_ThreeNetsApp_testOnlySampleFactor_256(DPIG_FourNetsFgBg_testOnlySampleFactor):
    """--model=1002 (trainer_256.py:845-1088): the DeepFashion sampler.  Three networks: the Stage-I graph of --model=101
    (GeneratorCNN_ID_Encoder_BodyROIVis on 64x64 crops, visibility-gated, + the 5-level U-Net), ONE appearance sampler
    (GaussianFCRes 224 -> 512 x 4 -> 224 in scope Gaussian_FC) and the pose auto-encoder.
      sample_app:   embedding = Gaussian_FC(z); otherwise the first sample's encoder embedding tiled over the batch
      sample_pose:  the decoded (encoded real) pose; otherwise the first sample's real pose for the whole batch
    There is no critic in this graph (trainer_256.py:846-853): the score is zero."""

    def __init__(self, config, loader=None):
        super().__init__(config, loader=loader)
        self.sample_app = getattr(config, "sample_app", False)
        self.test_batch_num = 100                                   # trainer_256.py:1036

    def init_net(self, net_cfg=None):
        self.ctx = _lib.Context(0)
        cfg = net_cfg or engine.NetConfig.deepfashion(img_h=self.img_H, img_w=self.img_W, hidden=self.conv_hidden_num,
                                                      z_num=self.z_num)
        self.cfg = cfg
        dev = torch.device("cuda", 0)
        B = self.batch_size
        self.s1 = engine.Stage1Engine(self.ctx, cfg, B, mode="dcgan", inference=True)
        self.s1.load_params(engine.init_params(cfg, seed=self.config.random_seed))
        self.factor = stage2._Factor(self.ctx, B, cfg.emb_dim, 512, "Gaussian_FC/G_FC", "FCDis_", dev)
        self.s2 = stage2.Stage2Engine(self.s1, mode="wgan", factors={"app": self.factor})
        self.s2.load_params(stage2.init_factor_params(self.factor, seed=self.config.random_seed))
        self._build_pose_tape(B, dev)
        for path in (self.pretrained_path, getattr(self.config, "pretrained_appSample_path", None),
                     getattr(self.config, "pretrained_poseAE_path", None)):
            if path:
                self.load_params(tf_checkpoint.load_any(path))

    def _fill_embedding(self):
        s1, B = self.s1, self.batch_size
        if self.sample_app:
            s1.emb.copy_(self.factor.fake.data)
        else:
            s1.emb.copy_(self.factor.real.data[:1].expand(B, -1))

    def _appearance_branch(self, st, z_fg, z_bg):
        if self.encode_unused or not self.sample_app:     # a sampled appearance does not read the encoder (see the base class)
            self.s2.encode_real()
        self.s2.sample_noise("app", z_fg)
        self.factor.p_g_fwd.run(st)

    def _score(self, st):
        return torch.zeros((self.batch_size,), device=self.s1.device)
