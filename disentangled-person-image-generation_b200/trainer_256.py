"""Trainer call surface of the reference for the DeepFashion 256x256 Stage-I model (--model=101):

    class DPIG_Encoder_GAN_BodyROI_256    reference trainer_256.py:10-134

Same step loop as the Market-1501 trainer (trainer_256.py:95-134 repeats trainer.py:326-366); the graph differs
(trainer_256.py:31-68): `models.GeneratorCNN_ID_Encoder_BodyROIVis` (no Fg/Bg branch) with repeat_num+1 levels on
64x64 ROI crops, the U-Net with repeat_num-1 levels, and ONE DCGANDiscriminator call on concat([x, G]) whose
16384-wide reshape turns every 256x256 image into 8 logits -- engine.NetConfig.deepfashion().
MODE is hard-wired to 'dcgan' (trainer_256.py:28).
"""
import json
import os
import time

import numpy as np
import torch

from . import _lib, engine, stage2, tf_checkpoint
from .tensor import ptr
from .trainer import DPIG_Encoder_GAN_BodyROI_FgBg
from .trainer_sub import (DPIG_Encoder_subSampleAppNetFgBg_GAN_BodyROI, DPIG_PoseRCV_AE_BodyROI,
                          DPIG_subnetSamplePoseRCV_GAN_BodyROI, _load_npz, _restore_lr, _save)


def df_sampler_config(trainer):
    """The Stage-I graph inside the DeepFashion sampler stages (--model=103 / 104 / 1002, trainer_256.py:306-322, 599-614):
    models.GeneratorCNN_ID_Encoder_BodyROI -- repeat_num+1 levels on the default 48x48 crops (48 -> 24 -> 12 -> 6 -> 3 -> 2
    -> 1), part features NOT gated by the visibilities -- and the repeat_num-1 level U-Net.  Variable shapes equal those of
    --model=101 (its 64x64 crops also end at 1x1x896), so the Encoder / ID_AE scopes restore from a Stage-I checkpoint."""
    return engine.NetConfig.deepfashion(img_h=trainer.img_H, img_w=trainer.img_W, hidden=trainer.conv_hidden_num,
                                        z_num=trainer.z_num, roi_size=48, use_vis=False)


class DPIG_Encoder_GAN_BodyROI_256(DPIG_Encoder_GAN_BodyROI_FgBg):
    def __init__(self, config, loader=None, dist=None):
        super().__init__(config, loader=loader, dist=dist)
        self.gan_mode = "dcgan"

    def _net_config(self):
        return engine.NetConfig.deepfashion(img_h=self.img_H, img_w=self.img_W, hidden=self.conv_hidden_num,
                                            z_num=self.z_num)


class DPIG_PoseRCV_AE_BodyROI_256(DPIG_PoseRCV_AE_BodyROI):
    """--model=102 (trainer_256.py:404-509): the pose auto-encoder stage on DeepFashion keypoints.  The graph and the
    step are those of --model=2 (trainer.py:626-708; the two build_model / train bodies differ only in commented-out
    preview code): PoseEncoderFCRes -> PoseDecoderFCRes on (r / img_H, c / img_W, v) normalised to [-1, 1] with the
    256 x 256 image size of the flags, loss mean((pose - G_pose)^2) * 20, Adam(beta1 = .5) on the PoseAE variables."""


class DPIG_Encoder_subSampleAppNet_GAN_BodyROI_256(DPIG_Encoder_subSampleAppNetFgBg_GAN_BodyROI):
    """--model=103 (trainer_256.py:266-402): embedding-space WGAN (RMSProp + clip, MODE='wgan') of ONE appearance sampler
    -- GaussianFCRes 224 -> 512 x 4 blocks -> 224 in scope Gaussian_FC, critic 'FCDis_' -- against the frozen DeepFashion
    Stage-I encoder; the step is that of --model=3 with a single factor (g_optim_embs from step 1 on, CRITIC_ITERS x
    (d_optim_embs + clip)); previews decode the sampled embedding through the U-Net at the given poses."""

    def _net_config(self):
        return df_sampler_config(self)

    def init_net(self):
        os.makedirs(self.model_dir, exist_ok=True)
        device = self.dist.local_rank if self.dist is not None else 0
        self.ctx = _lib.Context(device)
        cfg = self._net_config()
        dev = "cuda:%d" % device
        self.net = engine.Stage1Engine(self.ctx, cfg, self.batch_size, mode="dcgan", inference=True, device=dev)
        params = engine.init_params(cfg, seed=self.config.random_seed)
        loaded = _load_npz([self.pretrained_path, self.ckpt_path])     # Encoder + ID_AE restored, frozen (trainer_256.py:269-272)
        params.update({k: v for k, v in loaded.items() if k in params})
        self.net.load_params(params)
        self.factor = stage2._Factor(self.ctx, self.batch_size, cfg.emb_dim, 512, "Gaussian_FC/G_FC", "FCDis_",
                                     torch.device(dev))
        self.s2 = stage2.Stage2Engine(self.net, mode="wgan", g_lr=self.g_lr, d_lr=self.d_lr, factors={"app": self.factor},
                                      dist=self.dist)
        self.s2.load_params(stage2.init_factor_params(self.factor, seed=self.config.random_seed))
        self.s2.load_params(loaded)
        if self.ckpt_path:
            self.s2.load_state(loaded)
            _restore_lr(self, loaded)
        self._log = open(os.path.join(self.model_dir, "summary.jsonl"), "a")

    def train(self, on_step=None):
        t0 = time.time()
        for step in range(self.start_step, self.max_step):
            self.s2.g_lr, self.s2.d_lr = self.g_lr, self.d_lr
            self.s2.train_iteration(step, self.loader.next_batch)
            if on_step is not None:
                on_step(step, self)
            if step == 0 or step % self.log_step == self.log_step - 1:
                self.net.set_batch(self.loader.next_batch())
                self.s2.encode_real()
                self.s2.sample_noise("app")
                self.s2.d_grads("app")
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

    def generate(self, x, x_target, pose, part_bbox, part_vis=None, root_path=None, path=None, idx=None, save=False,
                 mask=None, z=None):
        """trainer_256.py:332-340: G = U-Net(sampled appearance embedding, pose).  Returns NHWC uint8."""
        B = self.batch_size
        net, s2 = self.net, self.s2
        st = torch.cuda.current_stream().cuda_stream
        if part_vis is None:
            part_vis = np.ones((B, 37), np.float32)
        net.set_batch(dict(x=np.asarray(x, np.float32), pose_rcv=np.asarray(pose, np.float32),
                           mask=np.ones((B, self.img_H, self.img_W, 1), np.float32), part_bbox=np.asarray(part_bbox),
                           part_vis=np.asarray(part_vis, np.float32)))
        s2.sample_noise("app", z)
        self.factor.p_g_fwd.run(st)
        net.emb.copy_(self.factor.fake.data)
        net.run_unet(st)
        out = torch.empty((B, self.img_H, self.img_W, 3), dtype=torch.uint8, device=net.device)
        self.ctx.denorm_u8(ptr(net.G), net.G.numel(), ptr(out), st)
        return out.cpu().numpy()


class DPIG_subnetSamplePoseRCV_GAN_BodyROI_256(DPIG_subnetSamplePoseRCV_GAN_BodyROI):
    """--model=104 (trainer_256.py:511-700): the pose-sampler stage of --model=4 on DeepFashion keypoints (normalised with
    the 256 x 256 image size of the flags); previews run the DeepFashion Stage-I graph of the sampler stages."""

    def _net_config(self):
        return df_sampler_config(self)
