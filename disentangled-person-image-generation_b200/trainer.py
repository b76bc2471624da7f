"""Trainer call surface of the reference for the Stage-I Market-1501 model (--model=1):

    class DPIG_Encoder_GAN_BodyROI_FgBg   reference trainer.py:567-625 (build_model), :326-366 (train),
                                          :44-110 (_common_init), :498-526 (generate)

`__init__(config)`, `init_net()`, `train()`, `test()`, `generate(...)` keep the reference's meaning; the TF graph
is replaced by engine.Stage1Engine (static launch programs over the C ABI).  Scalar names of the summaries are
kept (`loss/L1Loss`, `loss/g_loss`, ...) and written as JSON lines instead of TF event files.

Checkpoints are TensorFlow V2 bundles written / read without TensorFlow (tf_checkpoint.py), keyed by the reference's
variable names, so `--pretrained_path` / `--ckpt_path` accept the reference's checkpoints and vice versa.
Input: `--synthetic_data=false` reads the reference's TFRecord pair files under `<data_dir>/<dataset>` through
datasets.TFRecordPairLoader (datasets/market1501.py + _load_batch_pair_pose, without TensorFlow); otherwise batches come
from synth.make_batch, or from any loader object with `next_batch()` that is supplied.
"""
import json
import os
import time

import numpy as np
import torch

from . import _lib, datasets, engine, outputs, synth, tf_checkpoint
from .tensor import ptr


class SyntheticLoader:
    """Stands in for `_load_batch_pair_pose` (trainer.py:537-564): every call yields a fresh batch, like the TF
    queue does on every sess.run (reference quirk q2)."""

    def __init__(self, batch_size, img_h, img_w, seed=123):
        self.batch_size, self.img_h, self.img_w, self.seed, self.i = batch_size, img_h, img_w, seed, 0

    def next_batch(self):
        self.i += 1
        return synth.make_batch(self.batch_size, self.img_h, self.img_w, seed=self.seed + self.i)


def make_loader(config, batch_size, img_h, img_w):
    """trainer.py:35-42 / 1049-1055: dataset name -> market1501 / deepfashion get_split('train' | 'test', data_path) with
    data_path = <data_dir>/<dataset> (utils.py:136).  `--synthetic_data=true` (default here: no dataset ships) draws
    synthetic batches of the same shapes instead."""
    if getattr(config, "synthetic_data", True):
        return SyntheticLoader(batch_size, img_h, img_w, config.random_seed)
    name = config.dataset.lower()
    if "market" in name:
        data_name = "Market1501"
    elif "deepfashion" in name or "df" in name:
        data_name = "DeepFashion"
    else:
        raise Exception("dataset %r: expected a Market-1501 or DeepFashion TFRecord directory" % config.dataset)
    data_path = getattr(config, "data_path", None) or os.path.join(config.data_dir, config.dataset)
    return datasets.get_split("train" if config.is_train else "test", data_path, data_name=data_name,
                              batch_size=batch_size, seed=config.random_seed)


class DPIG_Encoder_GAN_BodyROI_FgBg(object):
    def __init__(self, config, loader=None, dist=None):
        self._common_init(config)
        self.D_arch = config.D_arch
        self.part_num = 37
        self.keypoint_num = 18
        self.loader = loader or make_loader(config, self.batch_size, self.img_H, self.img_W)
        self.dist = dist
        self.net = None

    def _common_init(self, config):
        # field names follow trainer.py:44-110
        self.config = config
        self.dataset = config.dataset
        self.batch_size = config.batch_size
        self.g_lr, self.d_lr = config.g_lr, config.d_lr
        self.z_num = config.z_num
        self.conv_hidden_num = config.conv_hidden_num
        self.img_H, self.img_W = config.img_H, config.img_W
        self.model_dir = config.model_dir or os.path.join(config.log_dir, "dpig_model%d" % config.model)
        self.start_step, self.max_step = config.start_step, config.max_step
        self.log_step, self.lr_update_step = config.log_step, config.lr_update_step
        self.is_train = config.is_train
        self.ckpt_path = config.ckpt_path
        self.pretrained_path = config.pretrained_path
        self.repeat_num = int(np.log2(self.img_H)) - 2      # trainer.py:74-75
        self.gan_mode = getattr(config, "gan_mode", "dcgan")  # _define_input hard-codes MODE='dcgan' (trainer.py:257)
        self.test_dir_name = "test_result"

    # ------------------------------------------------------------------ build
    def _net_config(self):
        return engine.NetConfig(img_h=self.img_H, img_w=self.img_W, hidden=self.conv_hidden_num, z_num=self.z_num)

    def init_net(self):
        """build_model + session setup of the reference (trainer.py:177-215, 568-625)."""
        os.makedirs(self.model_dir, exist_ok=True)
        device = self.dist.local_rank if self.dist is not None else 0
        self.ctx = _lib.Context(device)
        cfg = self._net_config()
        self.net = engine.Stage1Engine(self.ctx, cfg, self.batch_size, mode=self.gan_mode, dist=self.dist,
                                       device="cuda:%d" % device)
        self.net.g_lr, self.net.d_lr = self.g_lr, self.d_lr
        self.net.load_params(engine.init_params(cfg, seed=self.config.random_seed))
        # tf.train.Saver restores (trainer.py:180-213): --pretrained_path = the Encoder + ID_AE scopes only,
        # --ckpt_path = everything incl. the optimiser slots.  TensorFlow V2 checkpoints (prefix or directory) or .npz.
        if self.pretrained_path:
            self.net.load_params(tf_checkpoint.load_any(self.pretrained_path, scopes=["Encoder", "ID_AE"]))
        if self.ckpt_path:
            self.net.load_state(tf_checkpoint.load_any(self.ckpt_path))
        self._log = open(os.path.join(self.model_dir, "summary.jsonl"), "a")

    # ------------------------------------------------------------------ train
    def train(self, on_step=None):
        """The step loop of trainer.py:336-366: G update (skipped at global step 0), then disc_ITERS critic updates,
        each on its own batch; lr halving every lr_update_step; parameter dump every 30*log_step.
        on_step(step, trainer): optional hook called after the optimiser calls of every step (bench.py reads the losses
        there; the reference's loop fetches nothing between summaries)."""
        net = self.net
        disc_iters = 1 if self.gan_mode in ("dcgan", "lsgan") else 5   # wgan_gp.CRITIC_ITERS = 5 (wgan_gp.py:113)
        t0 = time.time()
        for step in range(self.start_step, self.max_step):
            if step > 0:
                net.set_batch(self.loader.next_batch())
                net.g_step()
            for _ in range(disc_iters):
                net.set_batch(self.loader.next_batch())
                net.d_step()
            if on_step is not None:
                on_step(step, self)
            if step == 0 or step % self.log_step == self.log_step - 1:
                net.set_batch(self.loader.next_batch())
                net.forward(with_disc=True)
                g_gan, d_loss, l1 = net.losses()
                rec = {"step": step, "loss/L1Loss": l1, "loss/g_loss_only": g_gan, "loss/g_loss": g_gan + 20.0 * l1,
                       "loss/d_loss": d_loss, "misc/g_lr": net.g_lr, "misc/d_lr": net.d_lr, "wall_s": time.time() - t0}
                self._log.write(json.dumps(rec) + "\n")
                self._log.flush()
            if step % self.lr_update_step == self.lr_update_step - 1:
                net.g_lr *= 0.5
                net.d_lr *= 0.5
            if step % (self.log_step * 30) == (self.log_step * 30) - 1:
                self.save(step)
        torch.cuda.synchronize()

    def save(self, step):
        """saver.save(sess, model_dir/model.ckpt, global_step=step) (trainer.py:365-366): a TensorFlow V2 checkpoint
        (model.ckpt-<step>.index / .data-00000-of-00001 + the `checkpoint` state file) readable by the reference."""
        state = self.net.get_state()
        state["step"] = np.int32(step)
        state["g_lr"], state["d_lr"] = np.float32(self.net.g_lr), np.float32(self.net.d_lr)
        return tf_checkpoint.save_checkpoint(os.path.join(self.model_dir, "model.ckpt-%d" % step), state)

    # ------------------------------------------------------------------ inference
    def generate(self, x, x_target, pose, part_bbox, part_vis, root_path=None, path=None, idx=None, save=False,
                 mask=None):
        """Reference generate() (trainer.py:498-526): returns the generated images as NHWC numpy in [0,255].
        Unlike the reference (quirk q3) the matching foreground mask is fed when given."""
        B = x.shape[0]
        if mask is None:
            mask = np.ones((B, self.img_H, self.img_W, 1), np.float32)
        self.net.set_batch(dict(x=np.asarray(x, np.float32), pose_rcv=np.asarray(pose, np.float32), mask=mask,
                                part_bbox=np.asarray(part_bbox), part_vis=np.asarray(part_vis, np.float32)))
        self.net.forward(with_disc=False)
        st = torch.cuda.current_stream().cuda_stream
        out = torch.empty((B, self.img_H, self.img_W, 3), dtype=torch.uint8, device=self.net.device)
        self.ctx.denorm_u8(ptr(self.net.G), self.net.G.numel(), ptr(out), st)
        # per-sample SSIM(G, x) on the uint8 images (trainer.py:516-521), computed on the device
        x8 = torch.empty_like(out)
        ssim = torch.empty((B,), dtype=torch.float32, device=self.net.device)
        self.ctx.denorm_u8(ptr(self.net.x), self.net.x.numel(), ptr(x8), st)
        self.ctx.ssim_gray_u8(ptr(out), ptr(x8), B, self.img_H, self.img_W, ptr(ssim), st)
        self.last_ssim = ssim.cpu().numpy()
        G = out.cpu().numpy()
        if save and (path is not None or root_path is not None):   # trainer.py:522-525
            outputs.save_image(G, path or os.path.join(root_path, "%s_G_ssim%s.png" % (idx, float(self.last_ssim.mean()))))
        return G

    def test(self):
        """Reconstruction pass over the loader (tester-style): writes G as .npy batches under model_dir/test_result."""
        out_dir = os.path.join(self.model_dir, self.test_dir_name)
        os.makedirs(out_dir, exist_ok=True)
        for i in range(4):
            b = self.loader.next_batch()
            g = self.generate(b["x"], b["x"], b["pose_rcv"], b["part_bbox"], b["part_vis"], mask=b["mask"])
            np.save(os.path.join(out_dir, "G_%05d.npy" % i), g)
        return out_dir
