"""GPU parity: FC sampling nets + the Stage-II embedding-space WGAN step (reference trainer.py:715-845,
models.py:474-486, wgan_gp.py:399-405) through the C ABI against the float64 oracle.  Pure fp32 SGEMM path:
tolerances 1e-4 relative L2 on gradients, 1e-5 on activations."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nets  # noqa: E402
from oracle import tf_ops as T  # noqa: E402

pytestmark = pytest.mark.gpu


def _setup(batch=4, mode="wgan"):
    import dpig_b200
    from dpig_b200 import engine, stage2, synth
    kw = dict(img_h=32, img_w=16, hidden=64, roi_size=12, d_dim=64)
    cfg = engine.NetConfig(**kw)
    ctx = dpig_b200.Context(0)
    s1 = engine.Stage1Engine(ctx, cfg, batch, mode="dcgan")
    p1 = nets.init_params(nets.NetConfig(**kw), seed=1234, bias_noise=0.05)
    s1.load_params(p1)
    s2 = stage2.Stage2Engine(s1, mode=mode)
    p2 = nets.init_stage2_params(seed=4321, bias_noise=0.05)
    assert set(p2) == set(s2.get_params())
    s2.load_params(p2)
    b = synth.make_batch(batch, 32, 16, seed=77)
    s1.set_batch(b)
    return s1, s2, p1, p2, b, nets.NetConfig(**kw)


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("factor", ["fg", "bg"])
@pytest.mark.parametrize("mode", ["wgan", "lsgan"])
def test_stage2_grads(factor, mode):
    s1, s2, p1, p2, b, cfg = _setup(mode=mode)
    s2.encode_real()
    f = s2.f[factor]
    rng = np.random.default_rng(9)
    z = rng.normal(0, 0.2, size=(s2.B, f.dim)).astype(np.float32)
    s2.sample_noise(factor, z)
    p = nets.to_torch(p2, torch.float64, requires_grad=True)
    real = f.real.data.detach().double().cpu()
    # the real embedding is the encoder's output: check it against the oracle encoder too
    po = nets.to_torch(p1, torch.float64)
    ob = dict(x=torch.tensor(b["x"], dtype=torch.float64), mask=torch.tensor(b["mask"], dtype=torch.float64),
              part_bbox=torch.tensor(b["part_bbox"][:, :7]), part_vis=torch.tensor(b["part_vis"][:, :7]))
    emb = nets.encoder_fgbg(po, cfg, ob["x"], ob["mask"], ob["part_bbox"], ob["part_vis"])
    sl = slice(0, 224) if factor == "fg" else slice(224, 352)
    assert float((real - emb[:, sl]).abs().max()) < 1e-3
    out = nets.stage2_losses(p, factor, real, torch.tensor(z, dtype=torch.float64), mode)
    # generator step
    s2.g_grads(factor)
    torch.cuda.synchronize()
    got = s2.get_params(grads=True)
    assert _rel(f.fake.data, out["fake"].detach()) < 1e-5
    gnames = [k for k in p if k.startswith("Gaussian_FC_%s" % ("Fg" if factor == "fg" else "Bg"))]
    gg = torch.autograd.grad(out["g_loss"], [p[k] for k in gnames], retain_graph=True)
    for k, g in zip(gnames, gg):
        assert _rel(got[k], g) < 1e-4, k
    # critic step
    s2.d_grads(factor)
    torch.cuda.synchronize()
    got = s2.get_params(grads=True)
    dnames = [k for k in p if k.startswith("Fg_FCDis_" if factor == "fg" else "Bg_FCDis_")]
    dg = torch.autograd.grad(out["d_loss"], [p[k] for k in dnames])
    for k, g in zip(dnames, dg):
        assert _rel(got[k], g) < 1e-4, k
    assert abs(float(f.loss.cpu()[1]) - float(out["d_loss"])) < 1e-5


def test_stage2_rmsprop_and_clip_step():
    """One critic update in MODE='wgan': TF RMSProp (rms slot = ones) followed by the +-0.01 clip."""
    s1, s2, p1, p2, b, cfg = _setup()
    s2.encode_real()
    s2.sample_noise("fg", np.random.default_rng(3).normal(0, 0.2, size=(s2.B, 224)).astype(np.float32))
    before = s2.get_params()
    s2.d_grads("fg")
    torch.cuda.synchronize()
    grads = s2.get_params(grads=True)
    s2._optim(s2.f["fg"], "d", torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    after = s2.get_params()
    for k in ("Fg_FCDis_Discriminator.1.Linear.W", "Fg_FCDis_Discriminator.Out.b"):
        pr = torch.tensor(before[k], dtype=torch.float64)
        ms = torch.ones_like(pr)
        T.rmsprop_step(pr, torch.tensor(grads[k], dtype=torch.float64), ms, 2e-5, clip=0.01)
        assert float((torch.tensor(after[k]).double() - pr).abs().max()) < 1e-7, k
        assert float(np.abs(after[k]).max()) <= 0.01 + 1e-9


@pytest.mark.parametrize("factor", ["fg", "bg"])
def test_stage2_pruned_encoder_embedding(factor):
    """A critic call runs only the encoder pyramid its factor reads (Stage2Engine.prune; the reference's sess.run evaluates
    both because fg_embs / bg_embs are slices of the concatenated embedding, trainer.py:741-742): the factor's real
    embedding must be what the whole encoder gives, although the other half of the embedding is never computed."""
    s1, s2, p1, p2, b, cfg = _setup()
    assert s2.prune                                   # the default
    s2.prune = False
    s2.encode_real(factor)
    torch.cuda.synchronize()
    full = s2.f[factor].real.data.clone()
    other = "bg" if factor == "fg" else "fg"
    # poison everything the pruned run must not depend on / must recompute
    s1.emb.fill_(float("nan"))
    s2.f[factor].real.data.fill_(float("nan"))
    (s1.bg_fea if factor == "fg" else s1.fea).fill_(float("nan"))
    keep_other = s2.f[other].real.data.clone()
    s2.prune = True
    s2.encode_real(factor)
    torch.cuda.synchronize()
    got = s2.f[factor].real.data
    assert torch.isfinite(got).all()
    # same launches on the same inputs; the FC at the top of a pyramid sums its split-K partials with fp32 atomics
    assert float((got - full).abs().max()) <= 1e-6 * float(full.abs().max())
    assert torch.equal(s2.f[other].real.data, keep_other)        # the other factor's buffer is not touched


def test_stage2_pruned_iteration_matches_the_full_encoder():
    """One train_iteration (2 g_optim + 10 d_optim, trainer.py:822-845) with and without the pruning, same batches and
    the same noise: the updated samplers and critics agree to fp32-atomic noise."""
    from dpig_b200 import synth
    outs = []
    for prune in (False, True):
        s1, s2, p1, p2, b, cfg = _setup()
        s2.prune = prune
        torch.manual_seed(1234)
        it = iter(range(10 ** 6))
        s2.train_iteration(1, lambda: synth.make_batch(s2.B, 32, 16, seed=100 + next(it)))
        torch.cuda.synchronize()
        outs.append(s2.get_params())
    for k in outs[0]:
        d = float(np.abs(outs[0][k] - outs[1][k]).max())
        assert d <= 1e-6, (k, d)


def test_stage2_iteration_runs():
    from dpig_b200 import synth
    s1, s2, p1, p2, b, cfg = _setup()
    it = iter(range(10 ** 6))
    s2.train_iteration(1, lambda: synth.make_batch(s2.B, 32, 16, seed=100 + next(it)))
    torch.cuda.synchronize()
    assert all(np.isfinite(v).all() for v in s2.get_params().values())


@pytest.mark.parametrize("factor", ["fg", "bg"])
def test_stage2_wgan_gp_fc_critic(factor):
    """Gradient penalty on [B,F] embeddings through the FC critic (trainer.py:226-236 as written for 2-D inputs):
    the JVP/adjoint second pass vs torch double-backward."""
    s1, s2, p1, p2, b, cfg = _setup(mode="wgan-gp")
    s2.encode_real()
    f = s2.f[factor]
    rng = np.random.default_rng(19)
    s2.sample_noise(factor, rng.normal(0, 0.2, size=(s2.B, f.dim)).astype(np.float32))
    alpha = torch.tensor(rng.uniform(0, 1, size=s2.B))
    f.gp_alpha.copy_(alpha.float().cuda())
    f.gp_alpha_fixed = True
    s2.d_grads(factor)
    torch.cuda.synchronize()
    got = s2.get_params(grads=True)
    p = nets.to_torch(p2, torch.float64, requires_grad=True)
    name = "Fg_FCDis_" if factor == "fg" else "Bg_FCDis_"
    real = f.real.data.detach().double().cpu()
    fake = f.fake.data.detach().double().cpu()
    disc = lambda t: nets.fc_discriminator(p, t, name=name)  # noqa: E731
    _, d_loss = T.gan_loss("wgan-gp", disc(real), disc(fake))
    gp, slopes, _ = T.gradient_penalty(disc, real, fake, alpha)
    dnames = [k for k in p if k.startswith(name)]
    grads = torch.autograd.grad(d_loss + 10.0 * gp, [p[k] for k in dnames], allow_unused=True)
    assert _rel(f.gp_slopes, slopes.detach()) < 1e-5
    assert abs(float(f.gp_loss.cpu()[0]) - float(gp)) < 1e-5
    for k, g in zip(dnames, grads):
        if g is not None and float(g.abs().max()) > 1e-12:
            assert _rel(got[k], g) < 2e-4, k
