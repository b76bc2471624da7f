"""Trainer call surface of the reference's sub-network stages (run_market_train.sh runs them after --model=1):

    class DPIG_PoseRCV_AE_BodyROI                          --model=2   reference trainer.py:626-708
        pose auto-encoder: PoseEncoderFCRes -> PoseDecoderFCRes, loss = mean((pose_rcv_norm - G_pose_rcv)^2) * 20,
        Adam(beta1=.5) on the PoseAE variables only.
    class DPIG_Encoder_subSampleAppNetFgBg_GAN_BodyROI     --model=3   reference trainer.py:715-867
        embedding-space WGAN (RMSProp + clip) of the Fg and Bg GaussianFCRes samplers against the frozen Stage-I
        appearance encoder; previews decode [fixed Fg | varying Fg] x [varying Bg | fixed Bg] through the U-Net.
    class DPIG_subnetSamplePoseRCV_GAN_BodyROI             --model=4   reference trainer.py:870-1033
        embedding-space WGAN of the pose sampler (PoseGaussian/G_FC, critic 'Pose_emb_') against the frozen pose
        encoder; previews decode the sampled pose embedding to keypoints -> inflated maps -> U-Net.

`__init__(config)`, `init_net()`, `train()`, `test()`, `generate(...)` keep the reference's meaning; every FC layer runs
through the C ABI (dpig_linear_fwd/bwd, dpig_pose_ae_loss, dpig_loss_gan, dpig_rmsprop_step / dpig_adam_step).
Checkpoints are TensorFlow V2 bundles (tf_checkpoint.py) keyed by the TF variable names; restores are by name.
"""
import json
import os
import time

import numpy as np
import torch

from . import _lib, engine, stage2, tf_checkpoint
from .tensor import ptr
from .trainer import DPIG_Encoder_GAN_BodyROI_FgBg


def _load_npz(paths):
    """Variables of every given checkpoint (TensorFlow V2 prefix / directory or .npz), later paths winning."""
    out = {}
    for path in paths:
        if path:
            out.update(tf_checkpoint.load_any(path))
    return out


def _save(model_dir, step, tensors, trainer=None):
    """model.ckpt-<step> with everything the reference's full tf.train.Saver() writes for a resume: the variables, the
    optimiser slots (the engines' get_state) and the step / g_lr / d_lr / phase variables of _common_init / _define_input."""
    tensors = dict(tensors)
    tensors["step"] = np.int32(step)
    if trainer is not None:
        if trainer.dist is not None and trainer.dist.rank != 0:
            return None
        tensors["g_lr"], tensors["d_lr"] = np.float32(trainer.g_lr), np.float32(trainer.d_lr)
        tensors["phase"] = np.bool_(trainer.is_train)
    return tf_checkpoint.save_checkpoint(os.path.join(model_dir, "model.ckpt-%d" % step), tensors)


def _restore_lr(trainer, state):
    """--ckpt_path resume: the halved learning rates come back with the checkpoint (trainer.py:55-59, 211-213)."""
    for name in ("g_lr", "d_lr"):
        if name in state:
            setattr(trainer, name, float(np.asarray(state[name])))


class DPIG_PoseRCV_AE_BodyROI(DPIG_Encoder_GAN_BodyROI_FgBg):
    def __init__(self, config, loader=None, dist=None):
        super().__init__(config, loader=loader, dist=dist)
        self.sample_pose = getattr(config, "sample_pose", False)

    def init_net(self):
        os.makedirs(self.model_dir, exist_ok=True)
        self.ctx = _lib.Context(0)
        self.device = torch.device("cuda", 0)
        self.pose_ae = stage2.PoseAE(self.ctx, self.batch_size, self.device, self.keypoint_num)
        params = stage2.init_pose_ae_params(self.keypoint_num, seed=self.config.random_seed)
        params.update({k: v for k, v in _load_npz([self.pretrained_path]).items() if k in params})
        self.pose_ae.load_params(params)
        if self.ckpt_path:        # full restore: variables, Adam slots, step counter, learning rates
            state = tf_checkpoint.load_any(self.ckpt_path)
            self.pose_ae.load_state(state)
            _restore_lr(self, state)
        self._log = open(os.path.join(self.model_dir, "summary.jsonl"), "a")

    def _feed_pose(self, batch):
        rcv = torch.as_tensor(np.asarray(batch["pose_rcv"], np.float32)).to(self.device)
        self.pose_ae.pose_in.data.copy_(stage2.PoseAE.normalise(rcv, self.img_H, self.img_W).reshape(self.batch_size, -1))

    def train(self):
        """trainer.py:666-695: g_optim from step 1 on, summary every log_step, lr halving, checkpoint every 30 log_steps."""
        t0 = time.time()
        for step in range(self.start_step, self.max_step):
            if step > 0:
                self._feed_pose(self.loader.next_batch())
                self.pose_ae.step(self.g_lr)
            if step == 0 or step % self.log_step == self.log_step - 1:
                self._feed_pose(self.loader.next_batch())
                self.pose_ae.grads()
                rec = {"step": step, "loss/reconstruct_loss": float(self.pose_ae.loss.cpu()[0]), "wall_s": time.time() - t0}
                self._log.write(json.dumps(rec) + "\n")
                self._log.flush()
            if step % self.lr_update_step == self.lr_update_step - 1:
                self.g_lr *= 0.5
            if step % (self.log_step * 30) == (self.log_step * 30) - 1:
                self.save(step)
        torch.cuda.synchronize()

    def save(self, step):
        return _save(self.model_dir, step, self.pose_ae.get_state(), self)

    def generate(self, pose_rcv):
        """Reconstructed keypoints G_pose_rcv [B,18,3] (normalised r, c and the binary visibility)."""
        self._feed_pose({"pose_rcv": pose_rcv})
        self.pose_ae.grads()
        return self.pose_ae.g_rcv.cpu().numpy()

    def test(self):
        out_dir = os.path.join(self.model_dir, self.test_dir_name)
        os.makedirs(out_dir, exist_ok=True)
        for i in range(4):
            np.save(os.path.join(out_dir, "G_pose_rcv_%05d.npy" % i), self.generate(self.loader.next_batch()["pose_rcv"]))
        return out_dir


class DPIG_Encoder_subSampleAppNetFgBg_GAN_BodyROI(DPIG_Encoder_GAN_BodyROI_FgBg):
    def init_net(self):
        os.makedirs(self.model_dir, exist_ok=True)
        assert self.batch_size > 0 and self.batch_size % 2 == 0, "batch should be Even and >0"   # trainer.py:777
        device = self.dist.local_rank if self.dist is not None else 0
        self.ctx = _lib.Context(device)
        cfg = self._net_config()
        # the frozen Stage-I networks run forward only (trainer.py:737-741: no gradient reaches the Encoder / ID_AE scopes)
        self.net = engine.Stage1Engine(self.ctx, cfg, self.batch_size, mode="dcgan", inference=True,
                                       device="cuda:%d" % device)
        params = engine.init_params(cfg, seed=self.config.random_seed)
        loaded = _load_npz([self.pretrained_path, self.ckpt_path])        # Encoder + ID_AE restored, frozen (trainer.py:180-183)
        params.update({k: v for k, v in loaded.items() if k in params})
        self.net.load_params(params)
        self.s2 = stage2.Stage2Engine(self.net, mode="wgan", g_lr=self.g_lr, d_lr=self.d_lr,   # MODE='wgan' trainer.py:720-725
                                      dist=self.dist)
        p2 = stage2.init_stage2_params(seed=self.config.random_seed)
        p2.update({k: v for k, v in loaded.items() if k in p2})
        self.s2.load_params(p2)
        if self.ckpt_path:
            state = tf_checkpoint.load_any(self.ckpt_path)
            self.s2.load_state(state)
            _restore_lr(self, state)
        self._log = open(os.path.join(self.model_dir, "summary.jsonl"), "a")

    def train(self, on_step=None):
        """trainer.py:812-867: per step and factor one g_optim (step > 0) and CRITIC_ITERS x (d_optim + clip)."""
        t0 = time.time()
        for step in range(self.start_step, self.max_step):
            self.s2.g_lr, self.s2.d_lr = self.g_lr, self.d_lr
            self.s2.train_iteration(step, self.loader.next_batch)
            if on_step is not None:
                on_step(step, self)
            if step == 0 or step % self.log_step == self.log_step - 1:
                rec = {"step": step, "misc/g_lr": self.g_lr, "misc/d_lr": self.d_lr, "wall_s": time.time() - t0}
                for factor in ("fg", "bg"):
                    self.net.set_batch(self.loader.next_batch())
                    self.s2.encode_real()
                    self.s2.sample_noise(factor)
                    self.s2.d_grads(factor)       # fills both losses of the factor for the summary batch
                    lg = self.s2.f[factor].loss.cpu()
                    rec["loss/g_loss_embs_%s" % factor] = float(lg[0])
                    rec["loss/d_loss_embs_%s" % factor] = float(lg[1])
                self._log.write(json.dumps(rec) + "\n")
                self._log.flush()
            if step % self.lr_update_step == self.lr_update_step - 1:
                self.g_lr *= 0.5
                self.d_lr *= 0.5
            if step % (self.log_step * 5) == (self.log_step * 5) - 1:
                self.save(step)
        torch.cuda.synchronize()

    def save(self, step):
        d = self.net.get_params()
        d.update(self.s2.get_state())
        return _save(self.model_dir, step, d, self)

    def generate(self, x, x_target, pose, part_bbox, part_vis, root_path=None, path=None, idx=None, save=False,
                 mask=None, z_fg=None, z_bg=None):
        """Preview of trainer.py:777-793: sampled appearance [fixed Fg ; varying Fg] x [varying Bg ; fixed Bg] decoded
        by the U-Net at the given poses.  Returns NHWC uint8."""
        B, h = self.batch_size, self.batch_size // 2
        net, s2 = self.net, self.s2
        st = torch.cuda.current_stream().cuda_stream
        if mask is None:
            mask = np.ones((B, self.img_H, self.img_W, 1), np.float32)
        net.set_batch(dict(x=np.asarray(x, np.float32), pose_rcv=np.asarray(pose, np.float32), mask=mask,
                           part_bbox=np.asarray(part_bbox), part_vis=np.asarray(part_vis, np.float32)))
        for factor, z in (("fg", z_fg), ("bg", z_bg)):
            s2.sample_noise(factor, z)
            s2.f[factor].p_g_fwd.run(st)
        fg, bg = s2.f["fg"].fake.data, s2.f["bg"].fake.data
        nfg = s2.fg_dim
        net.emb[:h, :nfg].copy_(fg[:1].expand(h, -1))      # app_embs_fixFg
        net.emb[h:, :nfg].copy_(fg[h:])                    # app_embs_varyFg
        net.emb[:h, nfg:].copy_(bg[h:])                    # app_embs_varyBg
        net.emb[h:, nfg:].copy_(bg[:1].expand(h, -1))      # app_embs_fixBg
        net.run_unet(st)
        out = torch.empty((B, self.img_H, self.img_W, 3), dtype=torch.uint8, device=net.device)
        self.ctx.denorm_u8(ptr(net.G), net.G.numel(), ptr(out), st)
        return out.cpu().numpy()


class DPIG_subnetSamplePoseRCV_GAN_BodyROI(DPIG_PoseRCV_AE_BodyROI):
    def init_net(self):
        os.makedirs(self.model_dir, exist_ok=True)
        self.ctx = _lib.Context(0)
        self.device = dev = torch.device("cuda", 0)
        B = self.batch_size
        loaded = _load_npz([self.pretrained_path, getattr(self.config, "pretrained_poseAE_path", None), self.ckpt_path])
        # frozen pose auto-encoder (restored from --model=2, trainer.py:185-187)
        self.pose_ae = stage2.PoseAE(self.ctx, B, dev, self.keypoint_num)
        pa = stage2.init_pose_ae_params(self.keypoint_num, seed=self.config.random_seed)
        pa.update({k: v for k, v in loaded.items() if k in pa})
        self.pose_ae.load_params(pa)
        # pose sampler + critic (trainer.py:893-910): GaussianFCRes 32 -> 512 x 4 blocks -> 32, FCDiscriminator 'Pose_emb_'
        self.factor = stage2._Factor(self.ctx, B, 32, 512, "PoseGaussian/G_FC", "Pose_emb_", dev)
        self.s2 = stage2.Stage2Engine(None, mode="wgan", g_lr=self.g_lr, d_lr=self.d_lr, factors={"pose": self.factor})
        self.s2.load_params(stage2.init_factor_params(self.factor, seed=self.config.random_seed))
        self.s2.load_params(loaded)
        if self.ckpt_path:
            self.s2.load_state(loaded)
            _restore_lr(self, loaded)
        # decoder applied to the SAMPLED embedding (shares the PoseAE parameters, trainer.py:897-899)
        self.sample_dec = stage2.PoseAE(self.ctx, B, dev, self.keypoint_num, group=self.pose_ae.group, encoder=False,
                                        decoder_from=self.factor.fake)
        self.net = None      # Stage-I engine, built on first generate()
        self._log = open(os.path.join(self.model_dir, "summary.jsonl"), "a")

    def _encode_real(self, batch):
        self._feed_pose(batch)
        self.pose_ae.p_fwd.run(torch.cuda.current_stream().cuda_stream)
        self.factor.real.data.copy_(self.pose_ae.pose_emb.data)

    def train(self):
        """trainer.py:972-1008: g_optim_embs (step > 0), CRITIC_ITERS x (d_optim_embs + clip)."""
        t0 = time.time()
        s2 = self.s2
        for step in range(self.start_step, self.max_step):
            s2.g_lr, s2.d_lr = self.g_lr, self.d_lr
            if step > 0:
                s2.sample_noise("pose")
                s2.g_step("pose")
            for _ in range(5):                      # wgan: CRITIC_ITERS = 5 (wgan_gp.py:113)
                self._encode_real(self.loader.next_batch())
                s2.sample_noise("pose")
                s2.d_step("pose")
            if step == 0 or step % self.log_step == self.log_step - 1:
                self._encode_real(self.loader.next_batch())
                s2.sample_noise("pose")
                s2.d_grads("pose")
                lg = self.factor.loss.cpu()
                rec = {"step": step, "loss/g_loss_embs": float(lg[0]), "loss/d_loss_embs": float(lg[1]),
                       "misc/g_lr": self.g_lr, "misc/d_lr": self.d_lr, "wall_s": time.time() - t0}
                self._log.write(json.dumps(rec) + "\n")
                self._log.flush()
            if step % self.lr_update_step == self.lr_update_step - 1:
                self.g_lr *= 0.5
                self.d_lr *= 0.5
            if step % (self.log_step * 30) == (self.log_step * 30) - 1:
                self.save(step)
        torch.cuda.synchronize()

    def save(self, step):
        d = self.pose_ae.get_params()
        d.update(self.s2.get_state())
        return _save(self.model_dir, step, d, self)

    def sample_pose_rcv(self, z=None):
        """G_pose_rcv [B,18,3]: noise -> PoseGaussian -> PoseDecoderFCRes -> (r, c in [-1,1], binary visibility)."""
        st = torch.cuda.current_stream().cuda_stream
        self.s2.sample_noise("pose", z)
        self.factor.p_g_fwd.run(st)
        d = self.sample_dec
        d.p_fwd.run(st)
        vis = torch.round(torch.sigmoid(d.vis_logit.data))
        return torch.cat([d.coord.data.reshape(self.batch_size, self.keypoint_num, 2), vis[:, :, None]], dim=-1)

    def generate(self, x_fixed, x_target_fixed, pose_fixed, part_bbox_fixed, root_path=None, path=None, idx=None,
                 save=False, part_vis=None, mask=None, z=None):
        """trainer.py:1010-1033: sampled keypoints -> inflated pose maps (radius 4) -> U-Net with the real appearance
        embedding of x_fixed.  Returns NHWC uint8."""
        B, H, W = self.batch_size, self.img_H, self.img_W
        if self.net is None:
            cfg = self._net_config()
            self.net = engine.Stage1Engine(self.ctx, cfg, B, mode="dcgan", inference=True)
            p1 = engine.init_params(cfg, seed=self.config.random_seed)
            p1.update({k: v for k, v in _load_npz([self.pretrained_path]).items() if k in p1})
            self.net.load_params(p1)
        g = self.sample_pose_rcv(z)
        R = torch.clamp((g[:, :, 0] + 1) / 2.0 * H, 0, H - 1)         # coord2channel_simple_rcv(is_normalized=True)
        Cc = torch.clamp((g[:, :, 1] + 1) / 2.0 * W, 0, W - 1)
        if mask is None:
            mask = np.ones((B, H, W, 1), np.float32)
        if part_vis is None:
            part_vis = np.ones((B, 37), np.float32)
        self.net.set_batch(dict(x=np.asarray(x_fixed, np.float32), pose_rcv=np.zeros((B, self.keypoint_num, 3), np.float32),
                                mask=mask, part_bbox=np.asarray(part_bbox_fixed), part_vis=part_vis))
        self.net.pose_rcv.copy_(torch.stack([R, Cc, g[:, :, 2]], dim=-1))
        st = torch.cuda.current_stream().cuda_stream
        self.net.run_encoder(st)
        self.net.run_unet(st)
        out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=self.net.device)
        self.ctx.denorm_u8(ptr(self.net.G), self.net.G.numel(), ptr(out), st)
        return out.cpu().numpy()

    def test(self):
        out_dir = os.path.join(self.model_dir, self.test_dir_name)
        os.makedirs(out_dir, exist_ok=True)
        for i in range(4):
            np.save(os.path.join(out_dir, "G_pose_rcv_%05d.npy" % i), self.sample_pose_rcv().cpu().numpy())
        return out_dir
