"""GPU parity of the sampling path (reference tester.py:419-613, --model=13) against the float64 oracle:
sample-or-hold Fg / Bg / pose, pose auto-encoder, inflation, U-Net, denorm, critic score."""
import argparse
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nets  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sample", [True, False, (True, False, True), (False, True, False)])
def test_sample_factor_forward(sample):
    """sample = all three switches, or (sample_fg, sample_bg, sample_pose).  The encoder runs as far as the fetched G needs
    it (tester._appearance_branch): not at all when both appearance factors are sampled, one pyramid when one is held."""
    from dpig_b200 import config as cfgmod
    s_fg, s_bg, s_pose = sample if isinstance(sample, tuple) else (sample, sample, sample)
    from dpig_b200 import engine, synth, tester
    B = 4
    kw = dict(img_h=32, img_w=16, hidden=64, roi_size=12, d_dim=64)
    conf, _ = cfgmod.get_config(["--model=13", "--is_train=False", "--batch_size=%d" % B, "--img_H=32", "--img_W=16",
                                 "--conv_hidden_num=64", "--sample_fg=%s" % s_fg, "--sample_bg=%s" % s_bg,
                                 "--sample_pose=%s" % s_pose])
    t = tester.DPIG_FourNetsFgBg_testOnlySampleFactor(conf)
    t.init_net(engine.NetConfig(**kw))
    ocfg = nets.NetConfig(**kw)
    params = dict(nets.init_params(ocfg, seed=11, bias_noise=0.05))
    params.update(nets.init_stage2_params(seed=12, bias_noise=0.05))
    params.update(nets.init_pose_params(seed=13, bias_noise=0.05))
    t.load_params(params)
    b = synth.make_batch(B, 32, 16, seed=21)
    rng = np.random.default_rng(5)
    z_fg = rng.normal(0, 0.2, size=(B, 224)).astype(np.float32)
    z_bg = rng.normal(0, 0.2, size=(B, 128)).astype(np.float32)
    t.s1.emb.fill_(float("nan"))          # nothing may read an embedding half the run did not produce
    t.s1.fea.fill_(float("nan"))
    t.s1.bg_fea.fill_(float("nan"))
    G, pose_img, score = t.generate(b["x"], None, b["pose_rcv"], b["part_bbox"], b["part_vis"], mask=b["mask"],
                                    z_fg=z_fg, z_bg=z_bg)
    assert np.isfinite(G).all() and np.isfinite(score).all()
    p = nets.to_torch(params, torch.float64)
    ob = dict(x=torch.tensor(b["x"], dtype=torch.float64), mask=torch.tensor(b["mask"], dtype=torch.float64),
              pose_rcv=torch.tensor(b["pose_rcv"], dtype=torch.float64),
              part_bbox=torch.tensor(b["part_bbox"][:, :7]), part_vis=torch.tensor(b["part_vis"][:, :7]))
    ref = nets.sample_factor_forward(p, ocfg, ob, torch.tensor(z_fg, dtype=torch.float64),
                                     torch.tensor(z_bg, dtype=torch.float64), s_fg, s_bg, s_pose)
    # keypoint pixels are truncated to ints by the rasteriser: compare the maps, then the image
    maps = t.s1.gin.slice(0, 18).hi.float().cpu().double()
    mism = float((maps != ref["pose_maps"]).double().mean())
    if mism >= 2e-3:
        print("engine pose_rcv[0:2]:", t.s1.pose_rcv[0:2, :4].cpu().tolist())
        print("oracle pose_pix[0:2]:", ref["pose_pix"][0:2, :4].tolist())
        print("per-sample mismatch:", (maps != ref["pose_maps"]).double().mean(dim=(1, 2, 3)).tolist())
    assert mism < 2e-3, mism     # a decoded coordinate within 1e-5 of an integer may truncate differently
    if mism == 0:
        assert float(np.abs(G - ref["G"].numpy()).max()) < 0.2            # 1e-3 on [-1,1] == 0.13 on [0,255]
        assert float(np.abs(score - ref["score"].numpy()).max()) < 1e-3
    assert G.shape == (B, 32, 16, 3) and pose_img.shape == (B, 32, 16, 3) and score.shape == (B,)
    assert G.min() >= 0 and G.max() <= 255


def test_sample_factor_forward_batch512_full_geometry():
    """BASELINE.json configs[4]: sampling (tester.py:573-613, --model=13) at 128x64, batch 512, through a forward-only
    engine (inference=True: no gradient buffers; a training engine at this batch would need ~150 GB).  The generator is
    per-sample, so the float64 oracle runs on a subset of rows; the critic score normalises with the statistics of the
    WHOLE batch (tflib/ops/batchnorm.py:29-30, quirk q4), so the oracle critic is run on all 512 generated images."""
    from dpig_b200 import config as cfgmod
    from dpig_b200 import synth, tester
    B, rows = 512, [0, 255, 511]
    conf, _ = cfgmod.get_config(["--model=13", "--is_train=False", "--batch_size=%d" % B, "--sample_fg=True",
                                 "--sample_bg=True", "--sample_pose=True"])
    t = tester.DPIG_FourNetsFgBg_testOnlySampleFactor(conf)
    t.init_net()
    assert not t.s1.training
    ocfg = nets.NetConfig()
    params = dict(nets.init_params(ocfg, seed=11, bias_noise=0.05))
    params.update(nets.init_stage2_params(seed=12, bias_noise=0.05))
    params.update(nets.init_pose_params(seed=13, bias_noise=0.05))
    t.load_params(params)
    b = synth.make_batch(B, ocfg.img_h, ocfg.img_w, seed=27)
    rng = np.random.default_rng(8)
    z_fg = rng.normal(0, 0.2, size=(B, 224)).astype(np.float32)
    z_bg = rng.normal(0, 0.2, size=(B, 128)).astype(np.float32)
    G, pose_img, score = t.generate(b["x"], None, b["pose_rcv"], b["part_bbox"], b["part_vis"], mask=b["mask"],
                                    z_fg=z_fg, z_bg=z_bg)
    assert G.shape == (B, 128, 64, 3) and score.shape == (B,) and np.isfinite(G).all() and np.isfinite(score).all()
    p = nets.to_torch(params, torch.float64)
    ob = dict(x=torch.tensor(b["x"][rows], dtype=torch.float64), mask=torch.tensor(b["mask"][rows], dtype=torch.float64),
              pose_rcv=torch.tensor(b["pose_rcv"][rows], dtype=torch.float64),
              part_bbox=torch.tensor(b["part_bbox"][rows][:, :7]), part_vis=torch.tensor(b["part_vis"][rows][:, :7]))
    with torch.no_grad():
        ref = nets.sample_factor_forward(p, ocfg, ob, torch.tensor(z_fg[rows], dtype=torch.float64),
                                         torch.tensor(z_bg[rows], dtype=torch.float64), True, True, True)
        maps = t.s1.gin.slice(0, 18).hi.float()[rows].cpu().double()
        same = (maps == ref["pose_maps"]).reshape(len(rows), -1).all(dim=1)       # see test_sample_factor_forward
        assert int(same.sum()) >= 2, same
        Ge = t.s1.G.detach().cpu().double()
        err = (Ge[rows] - ref["G_raw"] if "G_raw" in ref else (torch.clamp((Ge[rows] + 1) * 127.5, 0, 255) - ref["G"]) / 127.5)
        assert float(err[same].abs().max()) < 1e-3, float(err[same].abs().max())   # north-star bound, pre-denorm scale
        ref_score = nets.dcgan_discriminator(p, ocfg, Ge, "dcgan").reshape(-1)
    assert float(np.abs(score.reshape(-1) - ref_score.numpy()).max()) < 1e-3


def _four_nets(cls_name, argv, B=4):
    from dpig_b200 import config as cfgmod
    from dpig_b200 import engine, tester
    kw = dict(img_h=32, img_w=16, hidden=64, roi_size=12, d_dim=64)
    conf, _ = cfgmod.get_config(["--is_train=False", "--batch_size=%d" % B, "--img_H=32", "--img_W=16",
                                 "--conv_hidden_num=64"] + argv)
    t = getattr(tester, cls_name)(conf)
    t.init_net(engine.NetConfig(**kw))
    ocfg = nets.NetConfig(**kw)
    params = dict(nets.init_params(ocfg, seed=11, bias_noise=0.05))
    params.update(nets.init_stage2_params(seed=12, bias_noise=0.05))
    params.update(nets.init_pose_params(seed=13, bias_noise=0.05))
    t.load_params(params)
    return t, ocfg, nets.to_torch(params, torch.float64)


@pytest.mark.parametrize("sample_app,one_app", [(True, False), (True, True), (False, True), (False, False)])
def test_four_nets_test_only_model11(sample_app, one_app):
    """--model=11 (tester.py:256-417): sample_app / one_app_per_batch switches; every sample keeps its own pose."""
    from dpig_b200 import synth
    B = 4
    t, ocfg, p = _four_nets("DPIG_FourNetsFgBg_testOnly", ["--model=11", "--sample_app=%s" % sample_app,
                                                           "--one_app_per_batch=%s" % one_app, "--sample_pose=False"], B)
    b = synth.make_batch(B, 32, 16, seed=23)
    rng = np.random.default_rng(6)
    z_fg = rng.normal(0, 0.2, size=(B, 224)).astype(np.float32)
    z_bg = rng.normal(0, 0.2, size=(B, 128)).astype(np.float32)
    G, pose_img, score = t.generate(b["x"], None, b["pose_rcv"], b["part_bbox"], b["part_vis"], mask=b["mask"], z_fg=z_fg,
                                    z_bg=z_bg)
    ob = dict(x=torch.tensor(b["x"], dtype=torch.float64), mask=torch.tensor(b["mask"], dtype=torch.float64),
              pose_rcv=torch.tensor(b["pose_rcv"], dtype=torch.float64),
              part_bbox=torch.tensor(b["part_bbox"][:, :7]), part_vis=torch.tensor(b["part_vis"][:, :7]))
    ref = nets.four_nets_forward(p, ocfg, ob, torch.tensor(z_fg, dtype=torch.float64), torch.tensor(z_bg, dtype=torch.float64),
                                 sample_app, one_app, False)
    maps = t.s1.gin.slice(0, 18).hi.float().cpu().double()
    assert float((maps != ref["pose_maps"]).double().mean()) == 0.0      # integer keypoints: exact
    assert float(np.abs(G - ref["G"].numpy()).max()) < 0.2               # 1e-3 on [-1,1] == 0.13 on [0,255]
    assert float(np.abs(score - ref["score"].numpy()).max()) < 1e-3


def test_condition_model12_matches_oracle_and_writes_results(tmp_path):
    """--model=12 (tester.py:616-773): appearance of x, target pose; SSIM against x_target; result directories."""
    from dpig_b200 import synth
    B = 4
    t, ocfg, p = _four_nets("DPIG_FourNetsFgBg_testOnlyCondition", ["--model=12", "--model_dir=%s" % tmp_path], B)
    b = synth.make_batch(B, 32, 16, seed=31)
    bt = synth.make_batch(B, 32, 16, seed=32)
    G, score = t.generate(b["x"], bt["x"], bt["pose_rcv"], b["mask"], b["part_bbox"], b["part_vis"], save=False)
    ob = dict(x=torch.tensor(b["x"], dtype=torch.float64), mask=torch.tensor(b["mask"], dtype=torch.float64),
              part_bbox=torch.tensor(b["part_bbox"][:, :7]), part_vis=torch.tensor(b["part_vis"][:, :7]))
    ref = nets.condition_forward(p, ocfg, ob, torch.tensor(bt["pose_rcv"], dtype=torch.float64))
    assert float(np.abs(G - ref["G"].numpy()).max()) < 0.2
    assert float(np.abs(score - ref["score"].numpy()).max()) < 1e-3
    from oracle import image_metrics as im
    G8 = np.clip(G, 0, 255).astype(np.uint8)
    t8 = np.clip((bt["x"] + 1.0) * 127.5, 0, 255).astype(np.uint8)
    assert np.allclose(t.last_ssim, im.ssim_generate(G8, t8), atol=5e-6)

    class PairLoader:
        def next_batch(self):
            d = dict(b)
            d.update(x_target=bt["x"], pose_rcv_target=bt["pose_rcv"], mask_target=bt["mask"])
            return d
    t.loader = PairLoader()
    out_dir = t.test(num_batches=2)
    for d in ("x", "x_target", "G", "pose", "pose_target", "mask", "mask_target"):
        assert sorted(os.listdir(os.path.join(out_dir, d)))[:2] == ["00000.png", "00001.png"], d
        assert len(os.listdir(os.path.join(out_dir, d))) == 2 * B
    assert len(os.listdir(os.path.join(out_dir, "G_pose"))) == 0          # model 12 writes no G_pose files
    assert any(f.startswith("0_G_ssim") for f in os.listdir(out_dir))


def test_condition_256_model1001_matches_oracle():
    """--model=1001 (tester.py:775-915): DeepFashion form (no mask / Bg branch, no critic in the graph), small geometry."""
    from dpig_b200 import config as cfgmod
    from dpig_b200 import engine, synth, tester
    B = 2
    kw = dict(img_h=128, img_w=128, hidden=64, roi_size=32)          # DF_SMALL of tests/test_df256_gpu.py
    conf, _ = cfgmod.get_config(["--model=1001", "--is_train=False", "--batch_size=%d" % B, "--img_H=128", "--img_W=128",
                                 "--conv_hidden_num=64"])
    t = tester.DPIG_ThreeNetsApp_testOnlyCondition_256(conf)
    ecfg = engine.NetConfig.deepfashion(**kw)
    t.init_net(ecfg)
    ocfg = nets.NetConfig.deepfashion(**kw)
    params = nets.init_params(ocfg, seed=17, bias_noise=0.05)
    t.load_params(params)
    b = synth.make_batch(B, 128, 128, seed=41)
    bt = synth.make_batch(B, 128, 128, seed=42)
    G, score = t.generate(b["x"], bt["x"], bt["pose_rcv"], None, b["part_bbox"], b["part_vis"], save=False)
    ob = dict(x=torch.tensor(b["x"], dtype=torch.float64), mask=torch.tensor(b["mask"], dtype=torch.float64),
              part_bbox=torch.tensor(b["part_bbox"][:, :7]), part_vis=torch.tensor(b["part_vis"][:, :7]))
    ref = nets.condition_forward(nets.to_torch(params, torch.float64), ocfg, ob, torch.tensor(bt["pose_rcv"], dtype=torch.float64))
    assert float(np.abs(G - ref["G"].numpy()).max()) < 0.2
    assert (score == 0).all()
