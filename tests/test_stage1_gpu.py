"""GPU parity tests, network level: the Stage-I graph (reference trainer.py:568-625) run through the engine
(C ABI -> sm_100a kernels) against the float64 CPU oracle (oracle/nets.py) with identical injected weights
and identical seeded inputs.

Tolerances
  forward   : the north-star bound, 1e-3 max-abs fp32 per pixel on G (pre-denorm, values in ~[-1,1]), on the
              embedding, z, the logits and the losses (measured on B200: <= 7e-5 at full size).
  gradients : relative L2 error per parameter tensor.  The split-bf16 operands carry 2^-18 relative precision
              and the tensor-core fp32 accumulation truncates, so gradients are ~100x coarser than IEEE fp32;
              the Stage-I gradient is also ill-conditioned at full size (the float32 and float64 runs of the
              ORACLE ITSELF differ by 1e-2 relative L2 there -- DESIGN.md "precision").  The module-level VJP
              tests therefore inject identical cotangents / inputs on both sides, and the bounds below are
              ~3x the values measured on B200 (small: G-VJP 1.2e-2, D-VJP 1.4e-4; full: 2.1e-2, 9e-3).
Run as a script for a verbose report:  python tests/test_stage1_gpu.py [small|full]
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

TOL_ABS = 1e-3       # BASELINE.json north_star: max-abs fp32 per pixel
# Gradient bounds, relative L2 per parameter tensor, {small, full}.  The graph is piecewise linear in its 40-odd ReLU /
# LeakyReLU layers, so the float64 oracle is evaluated on the engine's linear piece (Stage1Engine.activation_bits ->
# `branches` of oracle/nets.py): an activation within fp32 rounding of zero otherwise takes different branches on the two
# sides and a handful of such bits moved these numbers to 1e-2 .. 1e-1 (round 1's bounds; tests/probe_grad_flake.py).
# Measured on B200 with the branches aligned: generator VJP 2.0e-5 / 6.4e-5, end to end 2.1e-4 / 7.3e-5 (g), 1.4e-4 /
# 7.0e-5 (d).
TOL_GVJP = {True: 5e-4, False: 5e-4}    # generator VJP
# discriminator VJP, relative L2, with the oracle on the engine's LeakyReLU branches (check_disc_vjp): measured 1.3e-4
# small / 4.4e-5 full (dcgan), 1.3e-5 / 2.5e-5 (wgan-gp incl. the penalty's second-order pass).  Without the branch
# bits the same comparison read up to 9e-3: one sign bit within fp32 rounding of zero (tests/probe_grad_flake.py).
TOL_DVJP = {True: 5e-4, False: 5e-4}
TOL_E2E = {True: 1e-3, False: 1e-3}     # end-to-end parameter gradients


def _setup(small, batch, mode="dcgan", seed=1234):
    import dpig_b200
    from dpig_b200 import engine, synth
    if small:
        kw = dict(img_h=32, img_w=16, hidden=64, roi_size=12, d_dim=64)
    else:
        kw = dict()
    ocfg = nets.NetConfig(**kw)
    ecfg = engine.NetConfig(**kw)
    params = nets.init_params(ocfg, seed=seed, bias_noise=0.05)
    ctx = dpig_b200.Context(0)
    eng = engine.Stage1Engine(ctx, ecfg, batch, mode=mode)
    assert set(eng.param_names()) == set(params.keys()), set(eng.param_names()) ^ set(params.keys())
    eng.load_params(params)
    b = synth.make_batch(batch, ocfg.img_h, ocfg.img_w, seed=123)
    eng.set_batch(b)
    ob = dict(x=torch.tensor(b["x"], dtype=torch.float64), mask=torch.tensor(b["mask"], dtype=torch.float64),
              pose=T.pose_rasterize(torch.tensor(b["pose_rcv"], dtype=torch.float64), ocfg.img_h, ocfg.img_w),
              part_bbox=torch.tensor(b["part_bbox"][:, :7]), part_vis=torch.tensor(b["part_vis"][:, :7]))
    p = nets.to_torch(params, torch.float64, requires_grad=True)
    return eng, ocfg, p, ob


def _maxabs(a, b):
    return float((torch.as_tensor(a).double().cpu() - b.double()).abs().max())


def check_forward(small, batch=2, mode="dcgan"):
    eng, cfg, p, ob = _setup(small, batch, mode)
    eng.forward(with_disc=True)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = nets.stage1_forward(p, cfg, ob, mode)
    rep = dict(
        emb=_maxabs(eng.emb, ref["emb"]), z=_maxabs(eng.z, ref["z"]), G=_maxabs(eng.G, ref["G"]),
        D_real=_maxabs(eng.d_real.logits, ref["D_real"]), D_fake=_maxabs(eng.d_fake.logits, ref["D_fake"]))
    g_gan, d_loss, l1 = eng.losses()
    rep.update(g_gan=abs(g_gan - float(ref["g_loss_only"])), d_loss=abs(d_loss - float(ref["d_loss"])),
               L1=abs(l1 - float(ref["L1"])))
    rep["ref_scale"] = dict(G=float(ref["G"].abs().max()), emb=float(ref["emb"].abs().max()),
                            logits=float(ref["D_fake"].abs().max()))
    return rep


def check_forward_batch64(batch=64, rows=(0, 21, 42, 63)):
    """BASELINE.json configs[1] -- the configuration bench.py is quoted on: full 128x64 graph, batch 64.  The generator is
    per-image (models.py:390-576), so the float64 oracle runs the Encoder + U-Net on a subset of rows; the critic
    normalises with the statistics of the whole batch (tflib/ops/batchnorm.py:29-30), so the oracle critic is run on all
    64 real and all 64 generated images (the engine's G, so that both sides evaluate D at the same point)."""
    eng, cfg, p, ob = _setup(False, batch)
    eng.forward(with_disc=True)
    torch.cuda.synchronize()
    rows = list(rows)
    sub = {k: v[rows] for k, v in ob.items()}
    with torch.no_grad():
        ref = nets.stage1_forward(p, cfg, sub, "dcgan")
        Gc = eng.G.detach().double().cpu()
        d_real = nets.dcgan_discriminator(p, cfg, ob["x"], "dcgan")
        d_fake = nets.dcgan_discriminator(p, cfg, Gc, "dcgan")
        g_gan, d_loss = T.gan_loss("dcgan", d_real, d_fake)
        l1 = (Gc - ob["x"]).abs().mean()
    rep = dict(emb=_maxabs(eng.emb[rows], ref["emb"]), z=_maxabs(eng.z[rows], ref["z"]), G=_maxabs(eng.G[rows], ref["G"]),
               D_real=_maxabs(eng.d_real.logits, d_real.reshape(-1)), D_fake=_maxabs(eng.d_fake.logits, d_fake.reshape(-1)))
    e_gan, e_d, e_l1 = eng.losses()
    rep.update(g_gan=abs(e_gan - float(g_gan)), d_loss=abs(e_d - float(d_loss)), L1=abs(e_l1 - float(l1)))
    return rep


def test_stage1_forward_batch64():
    rep = check_forward_batch64()
    for k in ("emb", "z", "G", "D_real", "D_fake", "g_gan", "d_loss", "L1"):
        assert rep[k] < TOL_ABS, rep


def check_grads(small, which, batch=2, mode="dcgan"):
    eng, cfg, p, ob = _setup(small, batch, mode)
    (eng.g_grads if which == "g" else eng.d_grads)()
    torch.cuda.synchronize()
    got = eng.get_params(grads=True)
    # the oracle differentiates on the engine's linear piece: every ReLU / LeakyReLU takes the branch the engine took
    _, ref = nets.stage1_grads(p, cfg, ob, which, mode, branches=eng.activation_bits())
    rep = {}
    # conv biases feeding a Batch/LayerNorm have an exactly-zero gradient (the norm removes the mean):
    # skip tensors whose reference gradient is numerically zero
    top = max(float(g.abs().max()) for g in ref.values() if g is not None)
    for name, g in ref.items():
        if g is None or float(g.abs().max()) < 1e-9 * top:
            continue
        rep[name] = float((torch.as_tensor(got[name]).double() - g).norm() / g.norm())
    return rep


def check_golden(small):
    """Engine forward against the committed golden vectors (tests/golden/*.npz, generated by make_golden.py
    from the float64 oracle) -- the oracle is not run here."""
    name = "stage1_small_b2.npz" if small else "stage1_full_b1.npz"
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))
    eng, cfg, p, ob = _setup(small, 2 if small else 1)
    eng.forward(with_disc=True)
    torch.cuda.synchronize()
    g_gan, d_loss, l1 = eng.losses()
    rep = {k: float(np.abs(v.detach().cpu().numpy() - ref[k]).max()) for k, v in
           dict(emb=eng.emb, z=eng.z, G=eng.G, D_real=eng.d_real.logits, D_fake=eng.d_fake.logits).items()}
    rep.update(L1=abs(l1 - float(ref["L1"])), g_gan=abs(g_gan - float(ref["g_loss_only"])),
               d_loss=abs(d_loss - float(ref["d_loss"])))
    return rep


@pytest.mark.parametrize("small", [True, False])
def test_engine_matches_golden(small):
    rep = check_golden(small)
    assert max(rep.values()) < TOL_ABS, rep


def _metrics(got, ref):
    """(relative L2 error, max-abs error / max-abs ref) of one tensor."""
    got = torch.as_tensor(got).double().cpu()
    ref = ref.double()
    return (float((got - ref).norm() / (ref.norm() + 1e-30)), float((got - ref).abs().max() / (ref.abs().max() + 1e-30)))


def check_generator_vjp(small, batch=2):
    """Backward of Encoder+U-Net in isolation: the oracle's dL/dG is injected as the cotangent, so the
    comparison is free of the discriminator's (ill-conditioned) sensitivity to G."""
    eng, cfg, p, ob = _setup(small, batch)
    s = torch.cuda.current_stream().cuda_stream
    eng.forward(with_disc=False)
    taps = {}
    gen_bits = {k: v for k, v in eng.activation_bits().items() if "/" in k}     # the generator's ReLU decisions
    out = nets.stage1_forward(p, cfg, ob, "dcgan", taps=taps, branches=gen_bits)
    names = [k for k in p if nets.is_generator_param(k)]
    grads = torch.autograd.grad(out["g_loss"], [p[k] for k in names] + [taps["G"]])
    gG = grads[-1]
    eng.gp.grad.zero_()
    eng.g_G.copy_(gG.float().cuda())
    eng.p_bwd_gen.run(s)
    torch.cuda.synchronize()
    got = eng.get_params(grads=True)
    return {k: _metrics(got[k], g) for k, g in zip(names, grads[:-1])}


def _signs(dp):
    return [dp.sign_bits(i).cpu() for i in range(4)]


def check_disc_vjp(small, batch=2):
    """Backward of the discriminator in isolation: the oracle D is fed the ENGINE's generated image, so both
    sides differentiate the same function at the same point."""
    eng, cfg, p, ob = _setup(small, batch)
    eng.d_grads()
    torch.cuda.synchronize()
    got = eng.get_params(grads=True)
    Gc = eng.G.detach().double().cpu()
    names = [k for k in p if nets.is_disc_param(k)]
    # ... and the same branch of every LeakyReLU: the oracle takes the sign bits the engine saved, so a pre-activation
    # within fp32 rounding of zero cannot put the two sides on different linear pieces (tests/probe_grad_flake.py)
    sr, sf = _signs(eng.d_real), _signs(eng.d_fake)
    d_real = nets.dcgan_discriminator(p, cfg, ob["x"], "dcgan", sr)
    d_fake = nets.dcgan_discriminator(p, cfg, Gc, "dcgan", sf)
    _, d_loss = T.gan_loss("dcgan", d_real, d_fake)
    grads = torch.autograd.grad(d_loss, [p[k] for k in names])
    rep = {k: _metrics(got[k], g) for k, g in zip(names, grads) if float(g.abs().max()) > 1e-12}
    # data gradient of the generator loss through D (the G step's entry cotangent)
    eng.g_grads()
    torch.cuda.synchronize()
    Gv = eng.G.detach().double().cpu().requires_grad_(True)
    g_gan, _ = T.gan_loss("dcgan", d_real.detach(), nets.dcgan_discriminator(p, cfg, Gv, "dcgan", _signs(eng.d_fake)))
    gx, = torch.autograd.grad(g_gan, Gv)
    rep["dL/dG (through D)"] = _metrics(eng.d_fake.g_x, gx)
    return rep


def check_wgan_gp(small, batch=2):
    """wgan-gp critic step (trainer.py:222-236, LayerNorm critic wgan_gp.py:34-40): d_loss incl. lambda*GP and
    its gradient w.r.t. the critic parameters -- the hand-derived second backward pass (JVP of D + adjoint)
    against torch double-backward in the float64 oracle.  The oracle critic is fed the engine's G."""
    eng, cfg, p, ob = _setup(small, batch, mode="wgan-gp")
    alpha = torch.tensor(np.random.default_rng(4321).uniform(0, 1, size=batch), dtype=torch.float64)
    eng.gp_alpha.copy_(alpha.float().cuda())
    eng.gp_alpha_fixed = True
    eng.d_grads()
    torch.cuda.synchronize()
    got = eng.get_params(grads=True)
    Gc = eng.G.detach().double().cpu()
    names = [k for k in p if nets.is_disc_param(k)]
    # every critic application with the LeakyReLU branches its engine counterpart took (see check_disc_vjp)
    disc_on = lambda dp: (lambda t: nets.dcgan_discriminator(p, cfg, t, "wgan-gp", _signs(dp)))  # noqa: E731
    disc = disc_on(eng.d_hat)
    _, d_loss = T.gan_loss("wgan-gp", disc_on(eng.d_real)(ob["x"]), disc_on(eng.d_fake)(Gc))
    gp, slopes, _ = T.gradient_penalty(disc, ob["x"], Gc, alpha)
    total = d_loss + 10.0 * gp
    grads = torch.autograd.grad(total, [p[k] for k in names])
    rep = {k: _metrics(got[k], g) for k, g in zip(names, grads) if float(g.abs().max()) > 1e-12}
    rep["slopes"] = _metrics(eng.slopes, slopes.detach())
    rep["d_loss"] = (abs(eng.losses()[1] - float(total)), abs(float(total)))
    # the penalty term alone (isolates the second-order pass from the first-order critic gradient)
    g_gp = torch.autograd.grad(10.0 * T.gradient_penalty(disc, ob["x"], Gc, alpha)[0], [p[k] for k in names],
                               allow_unused=True)
    eng.dp.grad.zero_()
    eng.p_gp.run(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got2 = eng.get_params(grads=True)
    for k, g in zip(names, g_gp):
        if g is not None and float(g.abs().max()) > 1e-12:
            rep["GP-only " + k] = _metrics(got2[k], g)
    return rep


@pytest.mark.parametrize("small", [True, False])
def test_wgan_gp_critic_step(small):
    rep = check_wgan_gp(small)
    # d_loss = wgan term + 10 * mean((slope - 1)^2) (~12 at full size, slopes ~2.2): the penalty amplifies the error of a
    # slope by 2 * lambda * (s - 1) ~ 24.  With the oracle on the engine's LeakyReLU branches (check_wgan_gp) the slopes
    # agree to 7e-6 relative and d_loss to 1.3e-5 (small) / 3.9e-4 .. 4.0e-4 (full) absolute -- inside the north star's 1e-3
    # again (round 1 measured 0.6e-3 .. 1.07e-3: a slope is a data gradient, and one flipped sign bit moved it).  Asserted
    # with a factor of margin over the two runs measured: 2e-3 absolute = 1.7e-4 of |d_loss|.
    assert rep["d_loss"][0] < 2 * TOL_ABS and rep["slopes"][0] < 1e-4, rep
    bad = {k: v for k, v in rep.items() if k not in ("d_loss", "slopes") and not v[0] < TOL_DVJP[False]}
    assert not bad, bad


@pytest.mark.parametrize("small", [True, False])
def test_stage1_forward(small):
    rep = check_forward(small)
    for k in ("emb", "z", "G", "D_real", "D_fake", "g_gan", "d_loss", "L1"):
        assert rep[k] < TOL_ABS, rep


@pytest.mark.parametrize("small,which", [(True, "g"), (True, "d"), (False, "g"), (False, "d")])
def test_stage1_grads_end_to_end(small, which):
    rep = check_grads(small, which)
    bad = {k: v for k, v in rep.items() if not v < TOL_E2E[small]}
    assert not bad, bad


@pytest.mark.parametrize("small", [True, False])
def test_generator_vjp(small):
    rep = check_generator_vjp(small)
    bad = {k: v for k, v in rep.items() if not v[0] < TOL_GVJP[small]}
    assert not bad, bad


@pytest.mark.parametrize("small", [True, False])
def test_discriminator_vjp(small):
    rep = check_disc_vjp(small)
    bad = {k: v for k, v in rep.items() if not v[0] < TOL_DVJP[small]}
    assert not bad, bad


def test_stage1_steps_move_parameters():
    """One g_optim + one d_optim update (trainer.py:336-347): TF-Adam's first step moves every weight with
    a non-zero gradient by ~lr (sign step), in both groups."""
    eng, cfg, p, ob = _setup(True, 2)
    before = eng.get_params()
    eng.g_step()
    eng.d_step()
    torch.cuda.synchronize()
    after = eng.get_params()
    for name in ("ID_AE/G/Conv_3/weights", "Encoder/G_encoder/Conv_5/weights", "Discriminator.3.Filters"):
        d = np.abs(after[name] - before[name])
        assert 0.5e-5 < float(np.median(d)) < 2.5e-5, (name, float(np.median(d)))


@pytest.mark.parametrize("mode", ["dcgan", "wgan-gp", "wgan"])
def test_step_graphs_match_eager_steps(mode, monkeypatch):
    """Whole-step CUDA graphs (engine._step): five g_optim + d_optim updates, the last three of each kind replayed from
    the captured graphs, against the same five updates launched eagerly from the same weights and batches.
      * every update, replayed or not, moves the weights by about one Adam / RMSProp step (a replay that did nothing, or
        ran with a stale step size, would show here);
      * both runs end at nearby weights.  The filter gradients are summed with fp32 atomics and the Stage-I gradient is
        ill-conditioned, so two runs differ by sign flips of small gradients; the bound is on the median;
      * the step size is a run-time input of the replayed graph: with lr = 0 a replayed step moves nothing."""
    from dpig_b200 import synth
    lr = 2e-4
    watch = ("ID_AE/G/Conv_3/weights", "Encoder/G_encoder/Conv_5/weights", "Discriminator.3.Filters")

    def run(graphs):
        monkeypatch.setenv("DPIG_GRAPHS", "1" if graphs else "0")
        eng, cfg, p, ob = _setup(True, 2, mode=mode)
        assert eng.use_graphs == graphs
        eng.gp_alpha_fixed = True
        if mode == "wgan-gp":
            eng.gp_alpha.copy_(torch.linspace(0.2, 0.8, eng.gp_alpha.numel(), device=eng.gp_alpha.device).reshape(eng.gp_alpha.shape))
        eng.g_lr = eng.d_lr = lr
        init = prev = eng.get_params()
        for i in range(5):
            eng.set_batch(synth.make_batch(2, cfg.img_h, cfg.img_w, seed=500 + 2 * i))
            eng.g_step()
            eng.set_batch(synth.make_batch(2, cfg.img_h, cfg.img_w, seed=501 + 2 * i))
            eng.d_step()
            torch.cuda.synchronize()
            cur = eng.get_params()
            for name in watch:
                step = float(np.median(np.abs(cur[name] - prev[name])))
                # Adam: ~lr per step while t is small; RMSProp (slot initialised to ones): lr*|g|/sqrt(0.9 + 0.1 g^2), and
                # the first wgan critic update clips the U(+-0.035) initial weights to +-0.01 (trainer.py:124-128)
                lo_, hi_ = (0.08 * lr, 1.6 * lr) if mode != "wgan" else (1e-9, 3.5 * lr if "Discriminator" not in name else 0.05)
                assert lo_ < step < hi_, (graphs, i, name, step)
            prev = cur
        return eng, init

    e_graph, init = run(True)
    assert set(e_graph._graphs) == {"g", "d"} and e_graph.ctx.replayed_launches > 0
    e_eager, _ = run(False)
    assert not e_eager._graphs and e_eager.ctx.replayed_launches == 0
    assert e_graph.t == e_eager.t == {"g": 5, "d": 5}
    pg, pe = e_graph.get_params(), e_eager.get_params()
    for name in watch:
        d = np.abs(pg[name].astype(np.float64) - pe[name].astype(np.float64))
        moved = np.abs(pe[name].astype(np.float64) - init[name].astype(np.float64))
        assert float(np.median(d)) <= 0.5 * float(np.median(moved)), (name, float(np.median(d)), float(np.median(moved)))
    # lr is a run-time input of the replayed graph
    before = e_graph.get_params()
    e_graph.g_lr = e_graph.d_lr = 0.0
    e_graph.g_step()
    e_graph.d_step()
    torch.cuda.synchronize()
    after = e_graph.get_params()
    for name in before:
        assert np.array_equal(before[name], after[name]), name


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    small = which == "small"
    mode = sys.argv[2] if len(sys.argv) > 2 else "dcgan"
    print("device:", torch.cuda.get_device_name(0), "config:", which, mode, flush=True)
    only_gp = len(sys.argv) > 3
    if not only_gp:
        rep = check_forward(small, mode=mode)
        print("forward max-abs errors:", rep, flush=True)
    for w in ("g", "d"):
        if mode != "dcgan":
            break
        rep = check_grads(small, w, mode=mode)
        worst = sorted(rep.items(), key=lambda kv: -kv[1])[:6]
        print("end-to-end grads[%s] worst relative-L2 errors:" % w, flush=True)
        for k, v in worst:
            print("   %-45s %.3e" % (k, v), flush=True)
    checks = (("generator VJP", check_generator_vjp), ("discriminator VJP", check_disc_vjp),
              ("wgan-gp critic step", check_wgan_gp))
    for nm, fn in (checks[2:] if only_gp else checks):
        rep = fn(small)
        worst = sorted(rep.items(), key=lambda kv: -kv[1][0])[:14]
        print("%s: worst (relL2, relMax):" % nm, flush=True)
        for k, v in worst:
            print("   %-45s L2 %.3e  max %.3e" % (k, v[0], v[1]), flush=True)
