"""GPU parity tests of the DeepFashion 256x256 Stage-I graph (reference --model=101, trainer_256.py:31-93):
encoder without Fg/Bg branch (models.py:328-388) with `repeat_num+1` ROI levels on 64x64 crops, U-Net with
`repeat_num-1` levels (models.py:518-576), and ONE discriminator call on concat([x, G]) whose hard-wired
`reshape(-1, 8*4*8*dim)` (wgan_gp.py:433) yields 8 logits per 256x256 image.

The engine (C ABI -> sm_100a kernels) is compared with the float64 CPU oracle (oracle/nets.py, NetConfig.deepfashion)
on identical injected weights and seeded inputs:
  small : 128x128, hidden 64, roi 32 (6 ROI levels 32 -> 1x1, 4 U-Net levels, 2 logits per image), batch 2 -- forward,
          generator VJP, discriminator VJP through the joint batch statistics, optimiser steps;
  full  : 256x256, hidden 128, roi 64, batch 1 -- forward against the committed golden file
          (tests/golden/df_full_b1.npz, made by tests/golden/make_golden.py --df; the oracle is not run on the GPU box).
Tolerances as in tests/test_stage1_gpu.py (forward 1e-3 max-abs; gradients relative L2, see there).
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import nets  # noqa: E402
from oracle import tf_ops as T  # noqa: E402

pytestmark = pytest.mark.gpu

TOL_ABS = 1e-3
# relative L2 per parameter tensor with the oracle on the engine's ReLU / LeakyReLU branches (see test_stage1_gpu.py):
# measured 2.5e-5 (generator) and 2.1e-5 (critic; 2.0e-4 on the scalar Output.b)
TOL_GVJP = 5e-4
TOL_DVJP = 1e-3
DF_SMALL = dict(img_h=128, img_w=128, hidden=64, roi_size=32)


def _setup(small, batch, seed=1234, oracle_side=True):
    import dpig_b200
    from dpig_b200 import engine, synth
    kw = DF_SMALL if small else {}
    ocfg = nets.NetConfig.deepfashion(**kw)
    ecfg = engine.NetConfig.deepfashion(**kw)
    params = nets.init_params(ocfg, seed=seed, bias_noise=0.05)
    ctx = dpig_b200.Context(0)
    eng = engine.Stage1Engine(ctx, ecfg, batch, mode="dcgan")
    assert set(eng.param_names()) == set(params.keys()), set(eng.param_names()) ^ set(params.keys())
    eng.load_params(params)
    b = synth.make_batch(batch, ocfg.img_h, ocfg.img_w, seed=123)
    eng.set_batch(b)
    if not oracle_side:
        return eng, ocfg, None, None
    ob = dict(x=torch.tensor(b["x"], dtype=torch.float64), mask=torch.tensor(b["mask"], dtype=torch.float64),
              pose=T.pose_rasterize(torch.tensor(b["pose_rcv"], dtype=torch.float64), ocfg.img_h, ocfg.img_w),
              part_bbox=torch.tensor(b["part_bbox"][:, :7]), part_vis=torch.tensor(b["part_vis"][:, :7]))
    p = nets.to_torch(params, torch.float64, requires_grad=True)
    return eng, ocfg, p, ob


def _maxabs(a, b):
    return float((torch.as_tensor(a).double().cpu() - torch.as_tensor(b).double()).abs().max())


def _metrics(got, ref):
    got = torch.as_tensor(got).double().cpu()
    ref = ref.double()
    return (float((got - ref).norm() / (ref.norm() + 1e-30)), float((got - ref).abs().max() / (ref.abs().max() + 1e-30)))


def _logits(eng, half):
    """Engine logits of one half in the reference's order: row index = image*rows_per_image + r
    (the engine keeps them [r][image] so that each half is contiguous; a pure permutation)."""
    R = eng.cfg.d_rows
    return half.logits.view(R, eng.B).t().reshape(-1)


def check_forward(small=True, batch=2):
    eng, cfg, p, ob = _setup(small, batch)
    eng.forward(with_disc=True)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = nets.stage1_forward(p, cfg, ob, "dcgan")
    rep = dict(emb=_maxabs(eng.emb, ref["emb"]), z=_maxabs(eng.z, ref["z"]), G=_maxabs(eng.G, ref["G"]),
               D_real=_maxabs(_logits(eng, eng.d_real), ref["D_real"]),
               D_fake=_maxabs(_logits(eng, eng.d_fake), ref["D_fake"]))
    g_gan, d_loss, l1 = eng.losses()
    rep.update(g_gan=abs(g_gan - float(ref["g_loss_only"])), d_loss=abs(d_loss - float(ref["d_loss"])),
               L1=abs(l1 - float(ref["L1"])))
    assert ref["D_real"].numel() == batch * eng.cfg.d_rows
    return rep


def check_golden(small):
    name = "df_small_b2.npz" if small else "df_full_b1.npz"
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))
    eng, cfg, _, _ = _setup(small, 2 if small else 1, oracle_side=False)
    eng.forward(with_disc=True)
    torch.cuda.synchronize()
    g_gan, d_loss, l1 = eng.losses()
    got = dict(emb=eng.emb, z=eng.z, G=eng.G, D_real=_logits(eng, eng.d_real), D_fake=_logits(eng, eng.d_fake))
    rep = {k: float(np.abs(v.detach().cpu().numpy() - ref[k]).max()) for k, v in got.items()}
    rep.update(L1=abs(l1 - float(ref["L1"])), g_gan=abs(g_gan - float(ref["g_loss_only"])),
               d_loss=abs(d_loss - float(ref["d_loss"])))
    return rep


def check_generator_vjp(small=True, batch=2):
    """Backward of Encoder (ROI branch only) + U-Net with the oracle's dL/dG injected as the cotangent."""
    eng, cfg, p, ob = _setup(small, batch)
    s = torch.cuda.current_stream().cuda_stream
    eng.forward(with_disc=False)
    taps = {}
    gen_bits = {k: v for k, v in eng.activation_bits().items() if "/" in k}     # the generator's ReLU decisions
    out = nets.stage1_forward(p, cfg, ob, "dcgan", taps=taps, branches=gen_bits)
    names = [k for k in p if nets.is_generator_param(k)]
    grads = torch.autograd.grad(out["g_loss"], [p[k] for k in names] + [taps["G"]])
    eng.gp.grad.zero_()
    eng.g_G.copy_(grads[-1].float().cuda())
    eng.p_bwd_gen.run(s)
    torch.cuda.synchronize()
    got = eng.get_params(grads=True)
    return {k: _metrics(got[k], g) for k, g in zip(names, grads[:-1])}


def check_disc_vjp(small=True, batch=2):
    """Backward of the joint discriminator call: the oracle D is fed concat([x, engine G]), so both sides
    differentiate the same function (joint batch statistics, 2 logits per image) at the same point."""
    eng, cfg, p, ob = _setup(small, batch)
    eng.d_grads()
    torch.cuda.synchronize()
    got = eng.get_params(grads=True)
    Gc = eng.G.detach().double().cpu()
    names = [k for k in p if nets.is_disc_param(k)]

    def both(g):      # on the LeakyReLU branches the engine's joint pass took (tests/probe_grad_flake.py)
        bits = [eng.d_pair.sign_bits(i).cpu() for i in range(4)]
        d = nets.dcgan_discriminator(p, cfg, torch.cat([ob["x"], g], dim=0), "dcgan", bits)
        return d[:d.shape[0] // 2], d[d.shape[0] // 2:]

    d_real, d_fake = both(Gc)
    _, d_loss = T.gan_loss("dcgan", d_real, d_fake)
    grads = torch.autograd.grad(d_loss, [p[k] for k in names])
    rep = {k: _metrics(got[k], g) for k, g in zip(names, grads) if float(g.abs().max()) > 1e-12}
    rep["d_loss"] = (abs(eng.losses()[1] - float(d_loss)), 0.0)
    # G step: gradient of g_loss w.r.t. G through the joint call (x shares the batch statistics)
    eng.g_grads()
    torch.cuda.synchronize()
    Gv = eng.G.detach().double().cpu().requires_grad_(True)
    d_real, d_fake = both(Gv)
    g_gan, _ = T.gan_loss("dcgan", d_real, d_fake)
    gx, = torch.autograd.grad(g_gan, Gv)
    rep["dL/dG (through D)"] = _metrics(eng.d_fake.g_x, gx)
    return rep


def test_df_small_forward():
    rep = check_forward(True)
    assert max(rep.values()) < TOL_ABS, rep


@pytest.mark.parametrize("small", [True, False])
def test_df_engine_matches_golden(small):
    rep = check_golden(small)
    assert max(rep.values()) < TOL_ABS, rep


def test_df_generator_vjp():
    rep = check_generator_vjp(True)
    bad = {k: v for k, v in rep.items() if not v[0] < TOL_GVJP}
    assert not bad, bad


def test_df_discriminator_vjp_joint_batch_statistics():
    rep = check_disc_vjp(True)
    assert rep.pop("d_loss")[0] < TOL_ABS
    bad = {k: v for k, v in rep.items() if not v[0] < TOL_DVJP}
    assert not bad, bad


def test_df_steps_move_parameters():
    eng, cfg, _, _ = _setup(True, 2, oracle_side=False)
    before = eng.get_params()
    eng.g_step()
    eng.d_step()
    torch.cuda.synchronize()
    after = eng.get_params()
    # (the deepest ROI levels work on 2x2 / 1x1 maps where most filter taps only ever see padding: zero gradient)
    for name in ("ID_AE/G/Conv_3/weights", "Encoder/G_encoder/Conv_5/weights", "Discriminator.3.Filters",
                 "Discriminator.Output.W"):
        d = np.abs(after[name] - before[name])
        assert 0.5e-5 < float(np.median(d)) < 2.5e-5, (name, float(np.median(d)))


def test_df_full_size_step_runs():
    """One g_optim + one d_optim update at 256x256 (batch 2): every kernel shape of the DF graph launches
    (896-channel 1x1-pixel ROI maps, 163840 -> 64 bottleneck FC, 8 logits per image) and the losses are finite."""
    eng, cfg, _, _ = _setup(False, 2, oracle_side=False)
    eng.g_step()
    eng.d_step()
    torch.cuda.synchronize()
    vals = eng.losses()
    assert all(np.isfinite(v) for v in vals), vals
    centre = eng.get_params(grads=False)["Encoder/G_encoder/Conv_22/weights"]      # 896 -> 896 on 1x1-pixel maps
    assert centre.shape == (3, 3, 896, 896)
    assert eng.d_real.logits.numel() == 2 * 8


if __name__ == "__main__":
    small = (sys.argv[1] if len(sys.argv) > 1 else "small") == "small"
    print("device:", torch.cuda.get_device_name(0), "DF", "small" if small else "full", flush=True)
    if small:
        print("forward max-abs errors:", check_forward(True), flush=True)
        for nm, fn in (("generator VJP", check_generator_vjp), ("discriminator VJP (joint)", check_disc_vjp)):
            rep = fn(True)
            worst = sorted(rep.items(), key=lambda kv: -kv[1][0])[:10]
            print("%s: worst (relL2, relMax):" % nm, flush=True)
            for k, v in worst:
                print("   %-45s L2 %.3e  max %.3e" % (k, v[0], v[1]), flush=True)
    print("golden:", check_golden(small), flush=True)


def test_df_sampler_stage_model103_and_tester_1002():
    """--model=103 (trainer_256.py:266-402) and --model=1002 (:845-1088) at a reduced geometry: the BodyROI encoder of the
    sampler stages (48x48 crops through 7 levels with TensorFlow's ceil halving 48 -> 24 -> 12 -> 6 -> 3 -> 2 -> 1, part
    features NOT gated by the visibilities), the single 'app' sampler / critic pair, and the tester's three-network
    forward -- against the float64 oracle."""
    import numpy as np
    from dpig_b200 import config as cfgmod
    from dpig_b200 import stage2, synth, tester, trainer_256
    from dpig_b200 import engine
    B = 2
    geo = dict(img_h=128, img_w=128, hidden=64)
    conf, _ = cfgmod.get_config(["--model=103", "--batch_size=%d" % B, "--img_H=128", "--img_W=128", "--conv_hidden_num=64",
                                 "--synthetic_data=true", "--model_dir=/tmp/dpig_test_103"])
    tr = trainer_256.DPIG_Encoder_subSampleAppNet_GAN_BodyROI_256(conf)
    tr.init_net()
    ecfg = tr.net.cfg
    assert (ecfg.roi_size, ecfg.use_vis, ecfg.enc_repeat) == (48, False, 6) and tr.net.roi_pyr.dims[-1][:2] == (2, 2)
    ocfg = nets.NetConfig.deepfashion(roi_size=48, use_vis=False, **geo)
    p1 = nets.init_params(ocfg, seed=31, bias_noise=0.05)
    tr.net.load_params(p1)
    p2 = {k: v + 0 for k, v in stage2.init_factor_params(tr.factor, seed=32).items()}
    tr.s2.load_params(p2)
    b = synth.make_batch(B, 128, 128, seed=43)
    b["part_vis"][:, 2] = 0.0                     # an "invisible" part must still contribute (no visibility gating here)
    tr.net.set_batch(b)
    tr.s2.encode_real()
    po = nets.to_torch(p1, torch.float64)
    ob = dict(x=torch.tensor(b["x"], dtype=torch.float64), mask=torch.tensor(b["mask"], dtype=torch.float64),
              part_bbox=torch.tensor(b["part_bbox"][:, :7]), part_vis=torch.tensor(b["part_vis"][:, :7]))
    emb = nets.encoder(po, ocfg, ob)
    real = tr.factor.real.data.detach().double().cpu()
    assert float((real - emb).abs().max()) < 1e-3
    z = np.random.default_rng(3).normal(0, 0.2, size=(B, ecfg.emb_dim)).astype(np.float32)
    tr.s2.sample_noise("app", z)
    p = nets.to_torch(p2, torch.float64, requires_grad=True)
    out = nets.stage2_losses(p, "app", real, torch.tensor(z, dtype=torch.float64), "wgan")
    tr.s2.d_grads("app")
    torch.cuda.synchronize()
    got = tr.s2.get_params(grads=True)
    dn = [k for k in p if k.startswith("FCDis_")]
    for k, g in zip(dn, torch.autograd.grad(out["d_loss"], [p[k] for k in dn], retain_graph=True)):
        assert float((torch.as_tensor(got[k]).double() - g).norm() / (g.norm() + 1e-30)) < 1e-4, k
    tr.s2.train_iteration(1, lambda: synth.make_batch(B, 128, 128, seed=44))      # one full step of the stage runs
    G = tr.generate(b["x"], b["x"], b["pose_rcv"], b["part_bbox"], b["part_vis"], z=z)
    assert G.shape == (B, 128, 128, 3) and G.dtype == np.uint8

    # ---- --model=1002: Stage-I graph of --model=101 (roi 64, visibility-gated) + Gaussian_FC sampler + pose AE
    conf2, _ = cfgmod.get_config(["--model=1002", "--is_train=False", "--batch_size=%d" % B, "--img_H=128", "--img_W=128",
                                  "--conv_hidden_num=64", "--sample_app=True", "--sample_pose=False", "--synthetic_data=true"])
    te = tester.DPIG_ThreeNetsApp_testOnlySampleFactor_256(conf2)
    ecfg2 = engine.NetConfig.deepfashion(roi_size=32, **geo)
    te.init_net(ecfg2)
    ocfg2 = nets.NetConfig.deepfashion(roi_size=32, **geo)
    q1 = nets.init_params(ocfg2, seed=33, bias_noise=0.05)
    q2 = stage2.init_factor_params(te.factor, seed=34)
    params = dict(q1)
    params.update(q2)
    params.update(nets.init_pose_params(seed=35, bias_noise=0.05))
    te.load_params(params)
    G, pose_img, score = te.generate(b["x"], None, b["pose_rcv"], b["part_bbox"], b["part_vis"], z_fg=z)
    assert (score == 0).all() and G.shape == (B, 128, 128, 3)
    pq = nets.to_torch(params, torch.float64)
    with torch.no_grad():
        app = nets.gaussian_fc_res(pq, torch.tensor(z, dtype=torch.float64), 4, "Gaussian_FC/G_FC", lambda t: T.leaky_relu(t, 0.2))
        rcv = torch.tensor(b["pose_rcv"], dtype=torch.float64)[:1].expand(B, -1, -1)     # sample_pose=False: first sample's pose
        Gref, _ = nets.unet_generator(pq, ocfg2, app, T.pose_rasterize(rcv, 128, 128))
    assert float(np.abs(G - T.denorm_img(Gref).numpy()).max()) < 0.2               # 1e-3 on [-1,1] == 0.13 on [0,255]
