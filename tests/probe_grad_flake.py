"""Not a pytest file.  Why does the critic-gradient comparison of tests/ddp_check.py / tests/test_dp_gpu.py land at
1e-5 on most runs and at ~6e-3 on some?  Hypothesis: a LeakyReLU sign bit of a pre-activation that sits within the
engine's forward error of zero (the engine saves `pre > 0` bit masks, the float64 oracle decides on its own value).
For every trial this prints the BatchNorm-scale gradient error against (a) the plain oracle and (b) the oracle whose
LeakyReLU uses the ENGINE's saved sign bits: if (b) stays at the 1e-5 level when (a) jumps, the jump is the sign bit
and not a race.       python tests/probe_grad_flake.py [trials]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dpig_b200  # noqa: E402
from dpig_b200 import engine, synth  # noqa: E402
from oracle import nets  # noqa: E402
from oracle import tf_ops as T  # noqa: E402


def unpack(mask, n, h, w, c):
    bits = (mask.unsqueeze(-1) >> torch.arange(32, device=mask.device, dtype=torch.int32)) & 1
    return bits.reshape(mask.shape[0], -1)[:, :c].reshape(n, h, w, c).bool().cpu()


def critic(p, x, masks=None, signs=None):
    """oracle/nets.py dcgan_discriminator with the LeakyReLU decisions optionally taken from `masks`."""
    def lrelu(z, i):
        if signs is not None:
            signs.append((z.detach() > 0))
        if masks is None:
            return T.leaky_relu(z)
        return torch.where(masks[i], z, 0.2 * z)
    h = lrelu(T.conv2d_same(x, p["Discriminator.1.Filters"], p["Discriminator.1.Biases"], 2), 0)
    for i in (2, 3, 4):
        h = T.conv2d_same(h, p["Discriminator.%d.Filters" % i], p["Discriminator.%d.Biases" % i], 2)
        h = lrelu(T.batchnorm_train(h, p["Discriminator.BN%d.scale" % i], p["Discriminator.BN%d.offset" % i]), i - 1)
    flat = h.permute(0, 3, 1, 2).reshape(h.shape[0], -1)
    return (flat @ p["Discriminator.Output.W"] + p["Discriminator.Output.b"]).reshape(-1)


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    kw = dict(img_h=32, img_w=16, hidden=64, roi_size=12, d_dim=64)
    ocfg, cfg = nets.NetConfig(**kw), engine.NetConfig(**kw)
    B = 16
    os.environ["DPIG_GRAPHS"] = "0"
    eng = engine.Stage1Engine(dpig_b200.Context(0), cfg, B, mode="dcgan")
    names = ["Discriminator.BN%d.scale" % i for i in (2, 3, 4)]
    for t in range(trials):
        params = nets.init_params(ocfg, seed=77 + t // 3, bias_noise=0.05)   # every parameter set three times
        eng.load_params(params)
        gb = synth.make_batch(B, 32, 16, seed=321 + t // 3)
        eng.set_batch(gb)
        eng.d_grads()
        torch.cuda.synchronize()
        grad = eng.dp.grad.clone()
        G = eng.G.double().cpu()
        x = torch.tensor(gb["x"], dtype=torch.float64)
        em = []
        for dp in (eng.d_real, eng.d_fake):
            em.append([unpack(dp.m[i], B, 32 >> (i + 1), 16 >> (i + 1), 64 << i) for i in range(4)])
        out = []
        flips = None
        for use_masks in (False, True):
            p = nets.to_torch(params, torch.float64, requires_grad=True)
            sr, sf = [], []
            d_real = critic(p, x, em[0] if use_masks else None, sr)
            d_fake = critic(p, G, em[1] if use_masks else None, sf)
            _, d_loss = T.gan_loss("dcgan", d_real, d_fake)
            gs = torch.autograd.grad(d_loss, [p[k] for k in names])
            e = 0.0
            for k, g in zip(names, gs):
                off, n, _ = eng.dp.specs[k]
                e = max(e, float((grad[off:off + n].double().cpu() - g).norm() / g.norm()))
            out.append(e)
            if flips is None:
                flips = [int((a != b).sum()) for a, b in zip(sr + sf, em[0] + em[1])]
        print("trial %2d (params %d): BN-scale grad error %.2e with the oracle's own signs, %.2e with the engine's bits; "
              "sign bits that differ per layer (real 1-4, fake 1-4): %s" % (t, 77 + t // 3, out[0], out[1], flips), flush=True)


if __name__ == "__main__":
    main()
