"""Generates tests/golden/*.npz from the float64 CPU oracle (oracle/nets.py).

The reference (Python 2.7 + TensorFlow 1.4.1) cannot be imported or run in this environment and ships no
golden vectors of its own, so these fixtures pin the ORACLE, not TensorFlow: they make
oracle regressions visible and give the GPU tests a file to compare with that does not need the oracle at run
time.  Inputs / weights are regenerated from seeds (synth.make_batch seed 123, nets.init_params seed 1234).

    python tests/golden/make_golden.py            # Market-1501 graph (--model=1)
    python tests/golden/make_golden.py --df       # DeepFashion graph (--model=101), small + 256x256
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from dpig_b200 import synth  # noqa: E402
from oracle import nets  # noqa: E402
from oracle import tf_ops as T  # noqa: E402

SMALL = dict(img_h=32, img_w=16, hidden=64, roi_size=12, d_dim=64)
# DeepFashion graph (--model=101, trainer_256.py:31-68) at a reduced geometry: 6 ROI levels 32 -> 1, 4 U-Net levels,
# two D rows per image
DF_SMALL = dict(img_h=128, img_w=128, hidden=64, roi_size=32)


def oracle_batch(b, cfg, dt=torch.float64):
    return dict(x=torch.tensor(b["x"], dtype=dt), mask=torch.tensor(b["mask"], dtype=dt),
                pose=T.pose_rasterize(torch.tensor(b["pose_rcv"], dtype=dt), cfg.img_h, cfg.img_w),
                part_bbox=torch.tensor(b["part_bbox"][:, :7]), part_vis=torch.tensor(b["part_vis"][:, :7]))


def compute(kw, batch, mode="dcgan", df=False):
    cfg = nets.NetConfig.deepfashion(**kw) if df else nets.NetConfig(**kw)
    params = nets.init_params(cfg, seed=1234, bias_noise=0.05)
    b = synth.make_batch(batch, cfg.img_h, cfg.img_w, seed=123)
    p = nets.to_torch(params, torch.float64, requires_grad=True)
    out = nets.stage1_forward(p, cfg, oracle_batch(b, cfg), mode)
    res = {k: out[k].detach().numpy().astype(np.float32) for k in ("emb", "z", "G", "D_real", "D_fake")}
    res.update({k: np.float32(float(out[k])) for k in ("L1", "g_loss_only", "g_loss", "d_loss")})
    res["pose"] = T.pose_rasterize(torch.tensor(b["pose_rcv"]), cfg.img_h, cfg.img_w).numpy().astype(np.int8)
    return res


def main():
    if "--df" not in sys.argv:
        np.savez_compressed(os.path.join(HERE, "stage1_small_b2.npz"), **compute(SMALL, 2))
        full = compute({}, 1)
        keep = {k: full[k] for k in ("emb", "z", "D_real", "D_fake", "L1", "g_loss_only", "g_loss", "d_loss")}
        keep["G"] = full["G"]                      # 128x64x3 fp32 = 98 KB
        np.savez_compressed(os.path.join(HERE, "stage1_full_b1.npz"), **keep)
    if "--df" in sys.argv or "--all" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "df_small_b2.npz"), **compute(DF_SMALL, 2, df=True))
        full = compute({}, 1, df=True)            # 256x256, 7 ROI levels, 5 U-Net levels: minutes of float64 CPU time
        full.pop("pose")
        np.savez_compressed(os.path.join(HERE, "df_full_b1.npz"), **full)
    print("written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
