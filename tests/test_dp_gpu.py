"""Data-parallel equivalence on hardware: ONE engine on a global batch of B images against
k = 2 engines on B/k images each that exchange what the NCCL path exchanges -- the (sum x, sum x^2) BatchNorm sums of
the critic in the forward pass, the (sum dy, sum dy*xhat) sums in its backward pass (sync-BN) and the gradient arenas
before the optimiser -- must give the same D logits, d_loss, critic gradients and post-step weights, and both must
agree with the float64 oracle on the GLOBAL batch (reference trainer.py:601-605: D(x), D(G) with batch statistics,
tflib/ops/batchnorm.py:29-30).  The ranks run as threads on one GPU (ddp.LocalGroup) through the same `dist` hooks
torch.distributed uses; tests/ddp_check.py repeats the comparison under torchrun with NCCL.
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

SMALL = dict(img_h=32, img_w=16, hidden=64, roi_size=12, d_dim=64)
LR = 2e-4
WATCH = ("ID_AE/G/Conv_3/weights", "Encoder/G_encoder/Conv_5/weights", "Discriminator.3.Filters",
         "Discriminator.BN3.scale")


def _sequence(eng, batch_d, batch_g):
    """What both sides run: critic gradients on one batch (recorded), then one d_optim and one g_optim update."""
    eng.g_lr = eng.d_lr = LR
    eng.set_batch(batch_d)
    eng.d_grads()
    torch.cuda.current_stream().synchronize()
    rec = dict(logits_real=eng.d_real.logits.detach().clone(), logits_fake=eng.d_fake.logits.detach().clone(),
               G=eng.G.detach().clone(), d_loss=eng.losses()[1], dgrad=eng.dp.grad.detach().clone())
    for i in range(4):       # the LeakyReLU branch bits the backward pass used
        rec["sign_real%d" % i], rec["sign_fake%d" % i] = eng.d_real.sign_bits(i), eng.d_fake.sign_bits(i)
    eng.d_step()
    eng.set_batch(batch_g)
    eng.g_step()
    torch.cuda.current_stream().synchronize()
    rec["params"] = eng.get_params()
    return rec


def run_dp_equivalence(small, B=4, world=2):
    import dpig_b200
    from dpig_b200 import ddp, engine, synth
    kw = SMALL if small else {}
    ocfg, ecfg = nets.NetConfig(**kw), engine.NetConfig(**kw)
    params = nets.init_params(ocfg, seed=77, bias_noise=0.05)
    bd = synth.make_batch(B, ocfg.img_h, ocfg.img_w, seed=321)
    bg = synth.make_batch(B, ocfg.img_h, ocfg.img_w, seed=322)

    os.environ["DPIG_GRAPHS"] = "0"            # both sides eager: the comparison is about the exchange, not the replay
    try:
        single = engine.Stage1Engine(dpig_b200.Context(0), ecfg, B, mode="dcgan")
        single.load_params(params)
        init = single.get_params()
        one = _sequence(single, bd, bg)

        def rank_fn(dist):
            eng = engine.Stage1Engine(dpig_b200.Context(0), ecfg, B // world, mode="dcgan", dist=dist)
            eng.load_params(params)
            rec = _sequence(eng, ddp.shard(bd, dist.rank, world), ddp.shard(bg, dist.rank, world))
            rec["overlap"] = eng.overlap_comm
            return rec

        ranks = ddp.LocalGroup(world).run(rank_fn)
    finally:
        os.environ.pop("DPIG_GRAPHS", None)

    rep = {}
    cat = lambda k: torch.cat([r[k] for r in ranks])  # noqa: E731
    rep["logits_real"] = float((cat("logits_real") - one["logits_real"]).abs().max())
    rep["logits_fake"] = float((cat("logits_fake") - one["logits_fake"]).abs().max())
    rep["G"] = float((cat("G") - one["G"]).abs().max())
    rep["d_loss"] = abs(float(np.mean([r["d_loss"] for r in ranks])) - one["d_loss"])
    g_dp = sum(r["dgrad"] for r in ranks) / world            # what _optim hands to Adam: all-reduce sum * 1/world
    rep["dgrad_rel"] = float((g_dp - one["dgrad"]).norm() / one["dgrad"].norm())
    # the BatchNorm scale / offset gradients are where a missing backward exchange would show first
    for name in ("Discriminator.BN2.scale", "Discriminator.BN3.scale", "Discriminator.BN4.offset"):
        off, n, _ = single.dp.specs[name]
        a, b = g_dp[off:off + n], one["dgrad"][off:off + n]
        rep["grad " + name] = float((a - b).norm() / (b.norm() + 1e-30))
    # every rank ends with the same weights, bit for bit; and close to the single-engine weights
    for k, v in ranks[0]["params"].items():
        assert np.array_equal(v, ranks[1]["params"][k]), k
    for name in WATCH:
        d = np.abs(ranks[0]["params"][name].astype(np.float64) - one["params"][name])
        moved = np.abs(one["params"][name].astype(np.float64) - init[name])
        rep["step " + name] = (float(np.median(d)), float(np.median(moved)))
    rep["overlap"] = ranks[0]["overlap"]

    # ---- the oracle on the GLOBAL batch, fed the engines' generated images (both sides differentiate D at one point)
    p = nets.to_torch(params, torch.float64, requires_grad=True)
    x = torch.tensor(bd["x"], dtype=torch.float64)
    Gc = cat("G").double().cpu()
    # LeakyReLU branches as the ranks took them: a pre-activation within fp32 rounding of zero otherwise flips one
    # branch between fp32 and float64 and moves the gradients by ~6e-3 at this size (tests/probe_grad_flake.py)
    d_real = nets.dcgan_discriminator(p, ocfg, x, "dcgan", [cat("sign_real%d" % i).cpu() for i in range(4)])
    d_fake = nets.dcgan_discriminator(p, ocfg, Gc, "dcgan", [cat("sign_fake%d" % i).cpu() for i in range(4)])
    _, d_loss = T.gan_loss("dcgan", d_real, d_fake)
    names = [k for k in p if nets.is_disc_param(k)]
    grads = torch.autograd.grad(d_loss, [p[k] for k in names])
    rep["oracle logits_real"] = float((cat("logits_real").double().cpu() - d_real.detach()).abs().max())
    rep["oracle logits_fake"] = float((cat("logits_fake").double().cpu() - d_fake.detach()).abs().max())
    rep["oracle d_loss"] = abs(float(np.mean([r["d_loss"] for r in ranks])) - float(d_loss))
    worst = 0.0
    got = {}
    for name in names:
        off, n, shape = single.dp.specs[name]
        got[name] = g_dp[off:off + n].view(shape)
    got["Discriminator.Output.W"] = None      # kept NHWC-flattened internally; compared through get_params elsewhere
    for name, g in zip(names, grads):
        if got[name] is None or float(g.abs().max()) < 1e-12:
            continue
        worst = max(worst, float((got[name].double().cpu() - g).norm() / g.norm()))
    rep["oracle dgrad_rel_worst"] = worst
    return rep


@pytest.mark.parametrize("small", [True, False])
def test_two_ranks_equal_one_engine_and_the_oracle(small):
    rep = run_dp_equivalence(small)
    assert rep["overlap"], rep                    # the overlapped ID_AE all-reduce is the path under test
    # forward: the per-image generator is identical work; the critic sees the same global statistics
    # (measured on B200: G 7e-6 / 1.1e-5, logits 1e-5 / 1.1e-4, d_loss 6e-7 -- two tilings of the same sums)
    assert rep["G"] < 1e-4 and rep["logits_real"] < 5e-4 and rep["logits_fake"] < 5e-4 and rep["d_loss"] < 2e-4, rep
    # critic gradient after the exchange, engine against engine: measured 2e-5 small, 3.8e-3 full -- the two tilings round
    # differently and may sit on different LeakyReLU branches where a pre-activation is ~0; each agrees with the oracle
    # on its own branches to 3e-5 (below)
    # (one differing branch bit moves these by up to 6e-3 at the reduced geometry, tests/probe_grad_flake.py: the bound
    #  leaves room for a few; the sharp statement is the oracle comparison on the ranks' own branches at the end)
    assert rep["dgrad_rel"] < 3e-2, rep
    for k, v in rep.items():
        if k.startswith("grad "):
            assert v < 3e-2, (k, rep)
        if k.startswith("step "):
            assert v[0] <= 0.25 * v[1], (k, rep)
    # against the float64 oracle on the global batch (north-star bound on the logits; DVJP bound on the gradients)
    assert rep["oracle logits_real"] < 1e-3 and rep["oracle logits_fake"] < 1e-3 and rep["oracle d_loss"] < 1e-3, rep
    # (measured 2.9e-5 small / 2.6e-5 full with the oracle on the ranks' LeakyReLU branches; 6e-4 .. 6e-3 before that)
    assert rep["oracle dgrad_rel_worst"] < 5e-4, rep


if __name__ == "__main__":
    for small in (True, False):
        print("small" if small else "full", run_dp_equivalence(small), flush=True)
