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


def _dataset_files(config):
    """(data_name, data_path, TFRecord files of the active split) for `--dataset`, or (None, data_path, [])."""
    import glob
    name = (config.dataset or "").lower()
    data_name = "Market1501" if "market" in name else ("DeepFashion" if ("deepfashion" in name or "df" in name) else None)
    data_path = getattr(config, "data_path", None) or os.path.join(config.data_dir, config.dataset)
    if data_name is None:
        return None, data_path, []
    split = "train" if config.is_train else "test"
    return data_name, data_path, sorted(glob.glob(os.path.join(data_path, "%s_%s_*.tfrecord" % (data_name, split))))


def make_loader(config, batch_size, img_h, img_w, rank=0, world=1):
    """trainer.py:35-42 / 1049-1055: dataset name -> market1501 / deepfashion get_split('train' | 'test', data_path) with
    data_path = <data_dir>/<dataset> (utils.py:136).
    `--synthetic_data` unset (the default): the reference's TFRecords are read when they exist under data_path -- the
    run_market_*.sh command lines then train on the dataset, as they do in the reference -- and synthetic batches of the
    same shapes are drawn, with a loud warning, only when no record file is there (no dataset ships with this repo).
    `--synthetic_data=true / false` forces either; false without files raises.  The choice is recorded in
    config.synthetic_data_effective (params.json) and in every summary line.
    rank / world: data-parallel ranks read different data (seed offset; the record files are dealt round-robin)."""
    data_name, data_path, files = _dataset_files(config)
    want = getattr(config, "synthetic_data", None)
    synthetic = (not files) if want is None else bool(want)
    config.synthetic_data_effective = synthetic
    if synthetic:
        if want is None:
            import warnings
            warnings.warn("DPIG: no %s TFRecords under %r -- training / testing on SYNTHETIC batches (random pixels, "
                          "synthetic keypoints); pass --data_dir / --dataset of the converted dataset, or "
                          "--synthetic_data=true to silence this" % (data_name or config.dataset, data_path), stacklevel=2)
        return SyntheticLoader(batch_size, img_h, img_w, config.random_seed + 7919 * rank)
    if data_name is None:
        raise Exception("dataset %r: expected a Market-1501 or DeepFashion TFRecord directory" % config.dataset)
    if not files:
        raise IOError("--synthetic_data=false but no %s_%s_*.tfrecord under %s" % (
            data_name, "train" if config.is_train else "test", data_path))
    return datasets.get_split("train" if config.is_train else "test", data_path, data_name=data_name,
                              batch_size=batch_size, seed=config.random_seed + 7919 * rank)


class DPIG_Encoder_GAN_BodyROI_FgBg(object):
    def __init__(self, config, loader=None, dist=None):
        self._common_init(config)
        self.D_arch = config.D_arch
        self.part_num = 37
        self.keypoint_num = 18
        self.dist = dist
        self.loader = loader or make_loader(config, self.batch_size, self.img_H, self.img_W,
                                            rank=dist.rank if dist is not None else 0,
                                            world=dist.world_size if dist is not None else 1)
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
            state = tf_checkpoint.load_any(self.ckpt_path)
            self.net.load_state(state)
            # the learning rates are tf.Variables of the reference graph (trainer.py:55-59) and come back with the
            # checkpoint: a run resumed after an lr halving continues at the halved rate
            for name in ("g_lr", "d_lr"):
                if name in state:
                    setattr(self, name, float(np.asarray(state[name])))
                    setattr(self.net, name, float(np.asarray(state[name])))
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
        # trainer.py:327-334: one fixed batch for the periodic previews, its inputs saved as sample sheets once
        period = self.log_step * 3
        first = self.start_step + ((period - 1 - self.start_step) % period)       # first step with step % period == period-1
        previews = self.max_step > self.start_step and (self.start_step == 0 or first < self.max_step)
        fixed = None
        if previews and (self.dist is None or self.dist.rank == 0):
            fixed = self.loader.next_batch()
            outputs.save_image((np.asarray(fixed["x"]) + 1.0) * 127.5, os.path.join(self.model_dir, "x_fixed.png"))
            outputs.save_image(np.asarray(fixed["mask"]) * 255.0, os.path.join(self.model_dir, "mask_fixed.png"))
            outputs.save_image(self._pose_max_img(fixed["pose_rcv"])[..., None], os.path.join(self.model_dir, "pose_fixed.png"))
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
                rec = {"step": step, "synthetic_data": bool(getattr(self.config, "synthetic_data_effective", False)),
                       "loss/L1Loss": l1, "loss/g_loss_only": g_gan, "loss/g_loss": g_gan + 20.0 * l1,
                       "loss/d_loss": d_loss, "misc/g_lr": net.g_lr, "misc/d_lr": net.d_lr, "wall_s": time.time() - t0}
                self._log.write(json.dumps(rec) + "\n")
                self._log.flush()
            if fixed is not None and (step == 0 or step % (self.log_step * 3) == (self.log_step * 3) - 1):
                # trainer.py:356-359: generate() on the fixed batch -> <model_dir>/<step>_G_ssim<mean>.png
                self.generate(fixed["x"], fixed.get("x_target", fixed["x"]), fixed["pose_rcv"], fixed["part_bbox"],
                              fixed["part_vis"], self.model_dir, idx=step, save=True, mask=fixed["mask"])
            if step % self.lr_update_step == self.lr_update_step - 1:
                net.g_lr *= 0.5
                net.d_lr *= 0.5
            if step % (self.log_step * 30) == (self.log_step * 30) - 1:
                self.save(step)
        torch.cuda.synchronize()

    def save(self, step):
        """saver.save(sess, model_dir/model.ckpt, global_step=step) (trainer.py:365-366): a TensorFlow V2 checkpoint
        (model.ckpt-<step>.index / .data-00000-of-00001 + the `checkpoint` state file) readable by the reference."""
        if self.dist is not None and self.dist.rank != 0:      # every rank holds the same weights: rank 0 writes
            return None
        state = self.net.get_state()
        state["step"] = np.int32(step)
        state["g_lr"], state["d_lr"] = np.float32(self.net.g_lr), np.float32(self.net.d_lr)
        state["phase"] = np.bool_(self.is_train)               # tf.Variable(self.is_train, name='phase') of _define_input
        return tf_checkpoint.save_checkpoint(os.path.join(self.model_dir, "model.ckpt-%d" % step), state)

    # ------------------------------------------------------------------ inference
    def generate(self, x, x_target, pose, part_bbox, part_vis, root_path=None, path=None, idx=None, save=False,
                 mask=None):
        """Reference generate() (trainer.py:498-526): returns the generated images as NHWC numpy in [0,255].
        Unlike the reference (quirk q3) the matching foreground mask is fed when given.  `pose` is the KEYPOINT array
        pose_rcv [B,18,3] (row, col, visible): the reference feeds the ready 18-channel maps; here they are rasterised
        and inflated on the device (dpig_pose_rasterize), so the argument is one step earlier in the same pipeline."""
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

    def _pose_max_img(self, pose_rcv):
        """(amax over the 18 inflated keypoint maps + 1) * 127.5 (trainer.py:401-402) for a [B,18,3] keypoint array."""
        B = self.batch_size
        rcv = torch.as_tensor(np.asarray(pose_rcv, np.float32)).to(self.net.device)
        maps = torch.empty((B, self.img_H, self.img_W, self.keypoint_num), dtype=torch.float32, device=self.net.device)
        self.ctx.pose_rasterize(ptr(rcv), B, self.keypoint_num, self.img_H, self.img_W, 4, None, ptr(maps),
                                torch.cuda.current_stream().cuda_stream)
        return ((maps.amax(dim=-1) + 1.0) * 127.5).cpu().numpy()

    def test(self, num_batches=100):
        """Reference test() (trainer.py:368-427): 100 batches through generate(); per-sample PNGs under
        <model_dir>/test_result/{x, x_target, G, pose, pose_target, mask, mask_target}/%05d.png -- the directories
        score.py reads (score.py:33-36) -- plus the sample sheets of the first batch.  PNG encoding is batched over host
        threads (outputs.ResultWriter)."""
        out_dir = os.path.join(self.model_dir, self.test_dir_name)
        wr = outputs.ResultWriter(out_dir)
        B = self.batch_size
        for i in range(num_batches):
            b = self.loader.next_batch()
            xt = b.get("x_target", b["x"])
            G = self.generate(b["x"], xt, b["pose_rcv"], b["part_bbox"], b["part_vis"], out_dir, idx=self.start_step,
                              save=(i == 0), mask=b["mask"])
            x255, xt255 = (np.asarray(b["x"]) + 1.0) * 127.5, (np.asarray(xt) + 1.0) * 127.5
            mask, maskt = np.asarray(b["mask"]) * 255.0, np.asarray(b.get("mask_target", b["mask"])) * 255.0
            p, pt = self._pose_max_img(b["pose_rcv"]), self._pose_max_img(b.get("pose_rcv_target", b["pose_rcv"]))
            for j in range(B):
                idx = i * B + j
                for d, arr in (("x", x255[j]), ("x_target", xt255[j]), ("G", G[j]), ("pose", p[j]), ("pose_target", pt[j]),
                               ("mask", np.squeeze(mask[j])), ("mask_target", np.squeeze(maskt[j]))):
                    wr._submit(arr, "%s/%s/%05d.png" % (out_dir, d, idx))
            if i == 0:
                wr.add_grid(x255, "x_fixed.png")
                wr.add_grid(xt255, "x_target_fixed.png")
                wr.add_grid(mask, "mask_fixed.png")
                wr.add_grid(maskt, "mask_target_fixed.png")
                wr.add_grid(p[..., None], "pose_fixed.png")
                wr.add_grid(pt[..., None], "pose_target_fixed.png")
        self.files_written = wr.close()
        return out_dir
