"""GPU parity + call-surface tests of the sub-network stages (reference --model=2 / 3 / 4, trainer.py:626-1033):
pose auto-encoder loss and gradients (straight-through binaryRound), the pose-embedding WGAN factor, and the three
trainer classes driven through main.py."""
import json
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import nets  # noqa: E402

pytestmark = pytest.mark.gpu


def _rel(got, ref):
    got, ref = torch.as_tensor(got).double().cpu(), ref.double()
    return float((got - ref).norm() / (ref.norm() + 1e-30))


def test_pose_autoencoder_loss_and_gradients():
    import dpig_b200
    from dpig_b200 import stage2, synth
    B, H, W = 8, 128, 64
    ctx = dpig_b200.Context(0)
    ae = stage2.PoseAE(ctx, B, torch.device("cuda", 0))
    params = nets.init_pose_params(seed=777, bias_noise=0.05)
    assert set(params) == set(ae.group.specs)
    ae.load_params(params)
    rcv = torch.tensor(synth.make_batch(B, H, W, seed=5)["pose_rcv"])
    norm = stage2.PoseAE.normalise(rcv, H, W)
    ae.pose_in.data.copy_(norm.reshape(B, -1).cuda())
    ae.grads(weight=20.0)
    torch.cuda.synchronize()
    p = nets.to_torch(params, torch.float64, requires_grad=True)
    loss, g = nets.pose_ae_loss(p, norm.double())
    grads = torch.autograd.grad(loss * 20.0, list(p.values()))
    assert abs(float(ae.loss.cpu()[0]) - float(loss)) < 1e-5
    assert float((ae.g_rcv.cpu().double() - g.detach()).abs().max()) < 1e-4
    got = ae.get_params(grads=True)
    bad = {k: _rel(got[k], gr) for k, gr in zip(p, grads) if _rel(got[k], gr) > 1e-3}
    assert not bad, bad
    # one Adam step (TF formula): every weight with a gradient moves by ~lr
    before = ae.get_params()
    ae.step(1e-4)
    torch.cuda.synchronize()
    d = np.abs(ae.get_params()["PoseAE/G_Pose_Decoder/fully_connected_3/weights"] - before["PoseAE/G_Pose_Decoder/fully_connected_3/weights"])
    assert 0.5e-4 < float(np.median(d)) < 1.5e-4


@pytest.mark.parametrize("mode", ["wgan", "lsgan"])
def test_pose_embedding_gan_factor(mode):
    """--model=4 (trainer.py:893-910): PoseGaussian sampler vs the 'Pose_emb_' critic on 32-d pose embeddings."""
    import dpig_b200
    from dpig_b200 import stage2
    B = 8
    ctx = dpig_b200.Context(0)
    dev = torch.device("cuda", 0)
    f = stage2._Factor(ctx, B, 32, 512, "PoseGaussian/G_FC", "Pose_emb_", dev)
    s2 = stage2.Stage2Engine(None, mode=mode, factors={"pose": f})
    params = stage2.init_factor_params(f, seed=11)
    rng = np.random.default_rng(3)
    for k in params:
        if k.endswith(("biases", ".b")):
            params[k] = rng.normal(0, 0.05, size=params[k].shape).astype(np.float32)
    s2.load_params(params)
    real = rng.normal(0, 0.3, size=(B, 32)).astype(np.float32)
    z = rng.normal(0, 0.2, size=(B, 32)).astype(np.float32)
    f.real.data.copy_(torch.tensor(real).cuda())
    p = nets.to_torch(params, torch.float64, requires_grad=True)
    out = nets.stage2_losses(p, "pose", torch.tensor(real, dtype=torch.float64), torch.tensor(z, dtype=torch.float64), mode)
    for which, loss_key, names in (("d", "d_loss", [k for k in p if k.startswith("Pose_emb_")]),
                                   ("g", "g_loss", [k for k in p if k.startswith("PoseGaussian/")])):
        s2.sample_noise("pose", z)
        (s2.d_grads if which == "d" else s2.g_grads)("pose")
        torch.cuda.synchronize()
        got = s2.get_params(grads=True)
        ref = torch.autograd.grad(out[loss_key], [p[k] for k in names], retain_graph=True)
        bad = {k: _rel(got[k], g) for k, g in zip(names, ref) if float(g.abs().max()) > 1e-12 and _rel(got[k], g) > 2e-3}
        assert not bad, (which, bad)
    assert abs(float(f.loss.cpu()[0]) - float(out["g_loss"])) < 1e-4


def _main(tmp_path, model, extra=()):
    import dpig_b200  # noqa: F401
    from dpig_b200 import config as C
    from dpig_b200 import main as M
    argv = ["--model=%d" % model, "--is_train=True", "--batch_size=4", "--max_step=3", "--log_step=2", "--gpu=-1",
            "--model_dir=%s" % tmp_path, "--conv_hidden_num=64", "--img_H=32", "--img_W=16"] + list(extra)
    cfg, _ = C.get_config(argv)
    tr = M.main(cfg)
    recs = [json.loads(ln) for ln in open(os.path.join(str(tmp_path), "summary.jsonl"))]
    assert [r["step"] for r in recs] == [0, 1]
    return tr, recs


def test_main_model2_pose_autoencoder(tmp_path):
    tr, recs = _main(tmp_path, 2)
    assert type(tr).__name__ == "DPIG_PoseRCV_AE_BodyROI"
    assert all(np.isfinite(r["loss/reconstruct_loss"]) for r in recs)
    g = tr.generate(tr.loader.next_batch()["pose_rcv"])
    assert g.shape == (4, 18, 3) and set(np.unique(g[:, :, 2])) <= {0.0, 1.0}
    from dpig_b200 import tf_checkpoint
    tr.g_lr = tr.g_lr * 0.5                              # as after an lr halving (trainer.py:362-363)
    z = tf_checkpoint.load_checkpoint(tr.save(2))
    # 42 variables, their Adam slots (`<var>/Adam`, `<var>/Adam_1`), the beta powers + step counter, step / g_lr / d_lr / phase
    assert "PoseAE/G_Pose_Encoder/fully_connected/weights" in z and int(z["step"]) == 2
    assert len(z) == 42 * 3 + 3 + 4 and bool(z["phase"]) and "PoseAE/G_Pose_Decoder/fully_connected_3/biases/Adam_1" in z
    # --ckpt_path resume: weights, Adam moments, step counter and the (halved) learning rate come back
    import copy
    from dpig_b200 import main as M
    cfg2 = copy.copy(tr.config)
    cfg2.ckpt_path, cfg2.max_step, cfg2.model_dir = str(tmp_path), 0, str(tmp_path / "resumed")
    tr2 = M.main(cfg2)
    a, b = tr.pose_ae.get_state(), tr2.pose_ae.get_state()
    assert set(a) == set(b) and all(np.array_equal(np.asarray(a[k]), np.asarray(b[k])) for k in a)
    assert tr2.pose_ae.t == tr.pose_ae.t == 2 and tr2.g_lr == pytest.approx(tr.g_lr)


def test_main_model3_appearance_samplers(tmp_path):
    tr, recs = _main(tmp_path, 3)
    assert type(tr).__name__ == "DPIG_Encoder_subSampleAppNetFgBg_GAN_BodyROI"
    for k in ("loss/g_loss_embs_fg", "loss/d_loss_embs_fg", "loss/g_loss_embs_bg", "loss/d_loss_embs_bg"):
        assert all(np.isfinite(r[k]) for r in recs)
    # the frozen Stage-I weights did not move; the critic weights stay inside the clip box (trainer.py:124-128)
    p = tr.s2.get_params()
    assert max(float(np.abs(v).max()) for k, v in p.items() if k.startswith("Fg_FCDis_")) <= 0.01 + 1e-7
    b = tr.loader.next_batch()
    g = tr.generate(b["x"], b["x"], b["pose_rcv"], b["part_bbox"], b["part_vis"], mask=b["mask"])
    assert g.shape == (4, 32, 16, 3) and g.dtype == np.uint8
    # preview composition (trainer.py:778-782): rows 0-1 share the Fg code, rows 2-3 share the Bg code
    emb = tr.net.emb.cpu().numpy()
    nfg = tr.s2.fg_dim
    assert np.array_equal(emb[0, :nfg], emb[1, :nfg]) and not np.array_equal(emb[2, :nfg], emb[3, :nfg])
    assert np.array_equal(emb[2, nfg:], emb[3, nfg:]) and not np.array_equal(emb[0, nfg:], emb[1, nfg:])


def test_main_model4_pose_sampler(tmp_path):
    tr, recs = _main(tmp_path, 4)
    assert type(tr).__name__ == "DPIG_subnetSamplePoseRCV_GAN_BodyROI"
    assert all(np.isfinite(r["loss/g_loss_embs"]) and np.isfinite(r["loss/d_loss_embs"]) for r in recs)
    rcv = tr.sample_pose_rcv().cpu().numpy()
    assert rcv.shape == (4, 18, 3) and set(np.unique(rcv[:, :, 2])) <= {0.0, 1.0}
    b = tr.loader.next_batch()
    g = tr.generate(b["x"], b["x"], b["pose_rcv"], b["part_bbox"], part_vis=b["part_vis"], mask=b["mask"])
    assert g.shape == (4, 32, 16, 3) and g.dtype == np.uint8
    from dpig_b200 import tf_checkpoint
    z = tf_checkpoint.CheckpointReader(tr.save(2))
    assert z.has_tensor("PoseGaussian/G_FC/fully_connected/weights") and z.has_tensor("Pose_emb_Discriminator.Out.W")
