"""FC sampling nets and the Stage-II embedding-space adversarial step of the reference:

    models.GaussianFCRes            models.py:474-486   noise -> residual MLP -> fake embedding
    models.PoseEncoderFCRes         models.py:488-499   (same residual-MLP shape; built by FCStack)
    WGAN_GP.FCDiscriminator         wgan_gp.py:399-405  LeakyReLU MLP critic on embeddings
    DPIG_Encoder_subSampleAppNetFgBg_GAN_BodyROI        trainer.py:715-845 (--model=3): MODE='wgan',
        RMSProp + weight clipping, g_optim then 5x(d_optim + clip) for the Fg and then the Bg factor.

All layers are skinny fp32 GEMMs (dpig_linear_fwd/bwd); a tiny tape records forward calls and derives the
backward program (linear / activation / residual add only).  The real embeddings come from the frozen Stage-I
appearance encoder (engine.Stage1Engine.run_encoder()).
"""
import math
import os
from collections import OrderedDict

import numpy as np
import torch

from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU, GAN_MODES
from .engine import ParamGroup, Program
from .tensor import ptr


class _Node:
    def __init__(self, b, n, device):
        self.data = torch.zeros((b, n), device=device)
        self.grad = torch.zeros((b, n), device=device)
        self.n = n


class FCTape:
    """Records y = act(x W + b) / y = a + b on [B, n] fp32 matrices and emits forward / backward programs."""

    def __init__(self, ctx, group, batch, device):
        self.ctx, self.group, self.B, self.device = ctx, group, batch, device
        self.ops = []
        self.nodes = []
        self.scratch = {}

    def node(self, n):
        nd = _Node(self.B, n, self.device)
        self.nodes.append(nd)
        return nd

    def linear(self, x, name_w, name_b, act=ACT_NONE, alpha=0.2):
        w = self.group.view(name_w)
        y = self.node(w.shape[1])
        self.ops.append(("linear", x, y, name_w, name_b, act, alpha))
        return y

    def add(self, a, b):
        y = self.node(a.n)
        self.ops.append(("add", a, b, y))
        return y

    def forward_program(self):
        p = Program(self.ctx)
        p.keep.append((self.nodes, self.ops, self.group))   # programs hold raw pointers: pin the buffers' lifetime
        for op in self.ops:
            if op[0] == "linear":
                _, x, y, nw, nb, act, alpha = op
                w, b = self.group.view(nw), self.group.view(nb)
                p.add("linear_fwd", ptr(x.data), ptr(w), ptr(b), ptr(y.data), self.B, w.shape[0], w.shape[1], act, alpha)
            else:
                _, a, b, y = op
                p.add("add_f32", ptr(y.data), ptr(a.data), ptr(b.data), y.data.numel(), 1.0, 1.0)
        return p

    def backward_program(self, params=True, outs=None):
        """Expects the .grad of the output node(s) filled by the caller (`outs`, default: the last op's output); all
        other node grads are zeroed first.  Accumulates parameter gradients when params=True; leaves d(loss)/d(node)
        in every node's .grad."""
        p = Program(self.ctx)
        p.keep.append((self.nodes, self.ops, self.group, self.scratch))
        if outs is None:
            outs = [self.ops[-1][2] if self.ops[-1][0] == "linear" else self.ops[-1][3]]
        zero = [nd for nd in self.nodes if not any(nd is o for o in outs)]
        p.add_py(lambda s: [nd.grad.zero_() for nd in zero])
        for op in reversed(self.ops):
            if op[0] == "linear":
                _, x, y, nw, nb, act, alpha = op
                w = self.group.view(nw)
                if act != ACT_NONE:
                    p.add("act_bwd_f32", ptr(y.data), ptr(y.grad), y.grad.numel(), alpha if act == ACT_LRELU else 0.0)
                need_dx = any(x is nd for nd in self.nodes)   # inputs outside the tape (noise) need no gradient
                tmp = self.scratch.setdefault(x.n, torch.zeros((self.B, x.n), device=self.device))
                p.add("linear_bwd", ptr(x.data), ptr(w), ptr(y.grad), ptr(tmp) if need_dx else None,
                      ptr(self.group.gview(nw)) if params else None, ptr(self.group.gview(nb)) if params else None,
                      self.B, w.shape[0], w.shape[1])
                if need_dx:
                    p.add("add_f32", ptr(x.grad), ptr(x.grad), ptr(tmp), x.grad.numel(), 1.0, 1.0)
            else:
                _, a, b, y = op
                p.add("add_f32", ptr(a.grad), ptr(a.grad), ptr(y.grad), y.grad.numel(), 1.0, 1.0)
                p.add("add_f32", ptr(b.grad), ptr(b.grad), ptr(y.grad), y.grad.numel(), 1.0, 1.0)
        return p


def fc_res_specs(prefix, in_dim, hidden, out_dim, repeat_num=4):
    """slim variable names/shapes of GaussianFCRes / PoseEncoderFCRes inside scope `prefix`."""
    specs, dims = [], [(in_dim, hidden)] + [(hidden, hidden)] * (2 * repeat_num) + [(hidden, out_dim)]
    for i, (a, b) in enumerate(dims):
        name = "%s/fully_connected%s" % (prefix, "" if i == 0 else "_%d" % i)
        specs += [(name + "/weights", (a, b)), (name + "/biases", (b,))]
    return specs


def fc_critic_specs(name, in_dim, fc_dim=512, n_layers=3):
    specs = [(name + "Discriminator.Input.Linear.W", (in_dim, fc_dim)), (name + "Discriminator.Input.Linear.b", (fc_dim,))]
    for i in range(n_layers):
        specs += [(name + "Discriminator.%d.Linear.W" % i, (fc_dim, fc_dim)), (name + "Discriminator.%d.Linear.b" % i, (fc_dim,))]
    specs += [(name + "Discriminator.Out.W", (fc_dim, 1)), (name + "Discriminator.Out.b", (1,))]
    return specs


def _init_specs(rng, specs, p):
    """Reference initialisers by variable kind: slim xavier_uniform / zero bias (`.../weights`, `.../biases`); tflib Linear
    'he' for the critic's LeakyReLULayers (wgan_gp.py:30-32) and glorot for `.Out.` (tflib/ops/linear.py:36-66)."""
    for name, shape in specs:
        if name.endswith("weights"):
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            p[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        elif name.endswith(".W"):
            std = math.sqrt(2.0 / (shape[0] + shape[1])) if ".Out." in name else math.sqrt(2.0 / shape[0])
            p[name] = rng.uniform(-std * math.sqrt(3), std * math.sqrt(3), size=shape).astype(np.float32)
        else:
            p[name] = np.zeros(shape, np.float32)
    return p


def init_stage2_params(fg_dim=224, bg_dim=128, seed=4321):
    """Gaussian_FC_Fg / Gaussian_FC_Bg samplers and their critics 'Fg_FCDis_' / 'Bg_FCDis_' (trainer.py:752-775)."""
    rng = np.random.default_rng(seed)
    p = OrderedDict()
    for scope, dim, hid in (("Gaussian_FC_Fg/G_FC", fg_dim, 512), ("Gaussian_FC_Bg/G_FC", bg_dim, 256)):
        _init_specs(rng, fc_res_specs(scope, dim, hid, dim), p)
    for pre, dim in (("Fg_FCDis_", fg_dim), ("Bg_FCDis_", bg_dim)):
        _init_specs(rng, fc_critic_specs(pre, dim), p)
    return p


def init_factor_params(factor, seed=4321):
    """Random-init parameters of one sampler + critic pair (e.g. PoseGaussian/G_FC + 'Pose_emb_', trainer.py:893-905)."""
    rng = np.random.default_rng(seed)
    p = OrderedDict()
    for grp in (factor.gp, factor.dp):
        _init_specs(rng, [(n, sp[2]) for n, sp in grp.specs.items()], p)
    return p


class _Factor:
    """One factor (Fg or Bg): GaussianFCRes generator + FCDiscriminator critic with their optimiser state."""

    def __init__(self, ctx, batch, dim, hidden, g_scope, d_name, device):
        self.ctx, self.B, self.dim, self.device = ctx, batch, dim, device
        self.g_scope, self.d_name = g_scope, d_name
        self.gp = ParamGroup(fc_res_specs(g_scope, dim, hidden, dim), device)
        self.dp = ParamGroup(fc_critic_specs(d_name, dim), device)
        self.gp.v.fill_(1.0)   # TF RMSProp slot 'rms' starts at ones
        self.dp.v.fill_(1.0)
        # generator tape (models.py:474-486, activation_fn=LeakyReLU as passed by trainer.py:753-757)
        t = self.gt = FCTape(ctx, self.gp, batch, device)
        self.z = _Node(batch, dim, device)
        names = [n for n, _ in fc_res_specs(g_scope, dim, hidden, dim)]
        lay = [(names[2 * i], names[2 * i + 1]) for i in range(len(names) // 2)]
        h = t.linear(self.z, *lay[0], act=ACT_LRELU)
        for r in range(4):
            res = h
            a = t.linear(h, *lay[1 + 2 * r], act=ACT_LRELU)
            b = t.linear(a, *lay[2 + 2 * r], act=ACT_LRELU)
            h = t.add(res, b)
        self.fake = t.linear(h, *lay[-1])
        self.p_g_fwd = t.forward_program()
        self.p_g_bwd = t.backward_program(params=True)
        # critic tapes: one on the real embedding, one on the fake one (shared parameters)
        self.real = _Node(batch, dim, device)
        self.crit = {}
        for key, inp in (("real", self.real), ("fake", self.fake)):
            ct = FCTape(ctx, self.dp, batch, device)
            dn = [n for n, _ in fc_critic_specs(d_name, dim)]
            dl = [(dn[2 * i], dn[2 * i + 1]) for i in range(len(dn) // 2)]
            x = _Node(batch, dim, device)
            x.data = inp.data         # alias the input storage
            ct.nodes.append(x)
            hh = x
            for wn, bn in dl[:-1]:
                hh = ct.linear(hh, wn, bn, act=ACT_LRELU)
            out = ct.linear(hh, *dl[-1])
            self.crit[key] = dict(tape=ct, x=x, out=out, fwd=ct.forward_program(),
                                  bwd_par=ct.backward_program(params=True), bwd_data=ct.backward_program(params=False))
        self.loss = torch.zeros((2,), device=device)
        self.t = {"g": 0, "d": 0}
        self._build_gp(ctx, batch, dim, d_name, device)

    def _build_gp(self, ctx, batch, dim, d_name, device):
        """WGAN-GP on 2-D [B,F] inputs exactly as trainer.py:226-236 writes it: alpha [B,1], interpolates =
        real + alpha*(fake-real), slopes = sqrt(sum_axis1 grad^2), gp = mean((slopes-1)^2).  The critic is a LeakyReLU
        MLP (piecewise linear), so d(lambda*gp)/d(theta) is the parameter gradient of the JVP along
        v = d(lambda*gp)/d(grad): tangent pass hdot_i = mask_i * (hdot_{i-1} W_i), then its adjoint (linear_bwd on the
        tangent activations).  No primal-path term: the masks are locally constant."""
        dn = [n for n, _ in fc_critic_specs(d_name, dim)]
        self.gp_layers = [(dn[2 * i], dn[2 * i + 1]) for i in range(len(dn) // 2)]
        ct = FCTape(ctx, self.dp, batch, device)
        self.xhat = _Node(batch, dim, device)
        ct.nodes.append(self.xhat)
        hh = self.xhat
        self.gp_h = []
        for wn, bn in self.gp_layers[:-1]:
            hh = ct.linear(hh, wn, bn, act=ACT_LRELU)
            self.gp_h.append(hh)
        self.gp_out = ct.linear(hh, *self.gp_layers[-1])
        self.gp_tape = ct
        self.gp_fwd = ct.forward_program()
        self.gp_bwd_data = ct.backward_program(params=False)
        self.gp_alpha = torch.zeros((batch,), device=device)
        self.gp_alpha_fixed = False
        self.gp_slopes = torch.zeros((batch,), device=device)
        self.gp_loss = torch.zeros((1,), device=device)
        self.gp_v = torch.zeros((batch, dim), device=device)
        self.gp_ones = torch.ones((batch, 1), device=device)
        widths = [self.dp.view(wn).shape[1] for wn, _ in self.gp_layers[:-1]]
        self.gp_hd = [torch.zeros((batch, w), device=device) for w in widths]      # tangent activations
        self.gp_hb = [torch.zeros((batch, w), device=device) for w in widths]      # their adjoints
        self.gp_xb = torch.zeros((batch, dim), device=device)

    def gp_program(self, lam):
        """Program computing lambda*gp (into gp_loss) and accumulating its critic-parameter gradient."""
        p = Program(self.ctx)
        p.keep.append(self)
        B, F = self.B, self.dim
        p.add("gp_interpolate", ptr(self.real.data), ptr(self.fake.data), ptr(self.gp_alpha), B, F, ptr(self.xhat.data))
        p.calls += self.gp_fwd.calls
        p.add_py(lambda s: self.gp_out.grad.fill_(1.0))
        p.calls += self.gp_bwd_data.calls
        p.add("gp_penalty", ptr(self.xhat.grad), B, F, float(lam), ptr(self.gp_slopes), ptr(self.gp_loss), ptr(self.gp_v))
        # tangent forward along v
        prev = self.gp_v
        for i, (wn, bn) in enumerate(self.gp_layers[:-1]):
            w = self.dp.view(wn)
            p.add("linear_fwd", ptr(prev), ptr(w), None, ptr(self.gp_hd[i]), B, w.shape[0], w.shape[1], ACT_NONE, 0.0)
            p.add("act_bwd_f32", ptr(self.gp_h[i].data), ptr(self.gp_hd[i]), self.gp_hd[i].numel(), 0.2)
            prev = self.gp_hd[i]
        # adjoint of S = sum_n hdot_L(n) . W_out
        wn, bn = self.gp_layers[-1]
        w = self.dp.view(wn)
        L = len(self.gp_hd)
        p.add("linear_bwd", ptr(self.gp_hd[L - 1]), ptr(w), ptr(self.gp_ones), ptr(self.gp_hb[L - 1]),
              ptr(self.dp.gview(wn)), None, B, w.shape[0], 1)
        for i in range(L - 1, -1, -1):
            wn, bn = self.gp_layers[i]
            w = self.dp.view(wn)
            p.add("act_bwd_f32", ptr(self.gp_h[i].data), ptr(self.gp_hb[i]), self.gp_hb[i].numel(), 0.2)
            x_in = self.gp_hd[i - 1] if i > 0 else self.gp_v
            dx = self.gp_hb[i - 1] if i > 0 else None
            p.add("linear_bwd", ptr(x_in), ptr(w), ptr(self.gp_hb[i]), ptr(dx) if dx is not None else None,
                  ptr(self.dp.gview(wn)), None, B, w.shape[0], w.shape[1])
        return p


def pose_ae_specs(keypoints=18):
    """slim variable names / shapes of PoseEncoderFCRes + PoseDecoderFCRes inside scope PoseAE (models.py:488-515;
    trainer.py:637-651): encoder 54 -> 512 -> 4 residual blocks -> 32; decoder 32 -> 512 -> 4 blocks -> (36 | 18)."""
    specs = fc_res_specs("PoseAE/G_Pose_Encoder", keypoints * 3, 512, 32)
    dec = [(32, 512)] + [(512, 512)] * 8 + [(512, keypoints * 2), (512, keypoints)]
    for i, (a, b) in enumerate(dec):
        name = "PoseAE/G_Pose_Decoder/fully_connected%s" % ("" if i == 0 else "_%d" % i)
        specs += [(name + "/weights", (a, b)), (name + "/biases", (b,))]
    return specs


def init_pose_ae_params(keypoints=18, seed=777):
    rng = np.random.default_rng(seed)
    p = OrderedDict()
    for name, shape in pose_ae_specs(keypoints):
        if name.endswith("weights"):
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            p[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        else:
            p[name] = np.zeros(shape, np.float32)
    return p


class PoseAE:
    """The pose auto-encoder as one FC tape: pose_in [B,54] (normalised r,c,v) -> pose_emb [B,32] -> (coord [B,36],
    vis_logit [B,18]).  `decoder_in` lets the decoder read another [B,32] node (the sampled embedding of --model=4)."""

    def __init__(self, ctx, batch, device, keypoints=18, group=None, encoder=True, decoder_from=None):
        self.ctx, self.B, self.K = ctx, batch, keypoints
        self.group = group or ParamGroup(pose_ae_specs(keypoints), device)
        names = list(self.group.specs)
        lay = [(names[2 * i], names[2 * i + 1]) for i in range(len(names) // 2)]
        enc, dec = lay[:10], lay[10:]
        t = self.tape = FCTape(ctx, self.group, batch, device)
        if encoder:
            self.pose_in = _Node(batch, keypoints * 3, device)
            h = t.linear(self.pose_in, *enc[0], act=ACT_LRELU)
            for r in range(4):
                a = t.linear(h, *enc[1 + 2 * r], act=ACT_LRELU)
                b = t.linear(a, *enc[2 + 2 * r], act=ACT_LRELU)
                h = t.add(h, b)
            self.pose_emb = t.linear(h, *enc[9])
            src = self.pose_emb
        else:
            src = decoder_from
        h = t.linear(src, *dec[0])
        for r in range(4):
            a = t.linear(h, *dec[1 + 2 * r], act=ACT_LRELU)
            b = t.linear(a, *dec[2 + 2 * r], act=ACT_LRELU)
            h = t.add(h, b)
        self.coord = t.linear(h, *dec[9])
        self.vis_logit = t.linear(h, *dec[10])
        self.p_fwd = t.forward_program()
        self.p_bwd = t.backward_program(params=True, outs=[self.coord, self.vis_logit])
        self.loss = torch.zeros((1,), device=device)
        self.g_rcv = torch.zeros((batch, keypoints, 3), device=device)
        self.t = 0

    def load_params(self, params):
        for name in self.group.specs:
            if name in params:
                self.group.view(name).copy_(torch.as_tensor(np.asarray(params[name]), dtype=torch.float32).to(
                    self.group.value.device))

    def get_params(self, grads=False):
        return OrderedDict((n, (self.group.gview(n) if grads else self.group.view(n)).detach().cpu().numpy().copy())
                           for n in self.group.specs)

    def get_state(self):
        """Variables + Adam slots (`<var>/Adam`, `<var>/Adam_1`) + step counter of the --model=2 optimiser."""
        out = self.get_params()
        g = self.group
        for name, (off, n, shape) in g.specs.items():
            out[name + "/Adam"] = g.m[off:off + n].view(shape).detach().cpu().numpy().copy()
            out[name + "/Adam_1"] = g.v[off:off + n].view(shape).detach().cpu().numpy().copy()
        out["beta1_power"] = np.float32(0.5 ** (self.t + 1))
        out["beta2_power"] = np.float32(0.999 ** (self.t + 1))
        out["dpig_step_count"] = np.int64(self.t)
        return out

    def load_state(self, state):
        from .engine import Stage1Engine
        self.load_params(state)
        g = self.group
        for name, (off, n, shape) in g.specs.items():
            for slot, arena in (("/Adam", g.m), ("/Adam_1", g.v)):
                if name + slot in state:
                    arena[off:off + n].view(shape).copy_(torch.as_tensor(np.asarray(state[name + slot]),
                                                                         dtype=torch.float32).to(arena.device).reshape(shape))
        self.t = Stage1Engine._restored_step_count(state, "", 0.999, False, self.t)

    @staticmethod
    def normalise(pose_rcv, img_h, img_w):
        """(row, col, visible) pixels -> [-1,1] x [-1,1] x {0,1} (trainer.py:639-644)."""
        return torch.stack([pose_rcv[:, :, 0] / float(img_h) * 2.0 - 1, pose_rcv[:, :, 1] / float(img_w) * 2.0 - 1,
                            pose_rcv[:, :, 2]], dim=-1)

    def grads(self, weight=20.0):
        """Forward + reconstruct_loss (trainer.py:658) + backward of loss*weight w.r.t. all PoseAE parameters."""
        s = torch.cuda.current_stream().cuda_stream
        self.group.grad.zero_()
        self.p_fwd.run(s)
        self.ctx.pose_ae_loss(ptr(self.pose_in.data), ptr(self.coord.data), ptr(self.vis_logit.data), self.B, self.K,
                              float(weight), ptr(self.loss), ptr(self.coord.grad), ptr(self.vis_logit.grad),
                              ptr(self.g_rcv), s)
        self.p_bwd.run(s)

    def step(self, lr, weight=20.0):
        """g_optim of --model=2: Adam(lr, beta1=0.5) on reconstruct_loss*20 (trainer.py:662-664)."""
        self.grads(weight)
        self.t += 1
        g = self.group
        self.ctx.adam_step(ptr(g.value), ptr(g.grad), ptr(g.m), ptr(g.v), g.total, lr, 0.5, 0.999, 1e-8, self.t, 1.0,
                           torch.cuda.current_stream().cuda_stream)


class Stage2Engine:
    """--model=3 step on top of a (frozen) Stage-I engine: `g_step(factor)` / `d_step(factor)` with factor in
    {'fg','bg'}; MODE 'wgan' (as shipped), 'lsgan' or 'dcgan' losses."""

    def __init__(self, stage1, mode="wgan", g_lr=2e-5, d_lr=2e-5, factors=None, dist=None):
        """dist: data-parallel hooks (ddp.Dist): the FC gradients are all-reduced before every optimiser call; the FC
        nets carry no batch-coupled normalisation (models.py:474-486, wgan_gp.py:399-405), so that is the only exchange."""
        self.s1, self.mode, self.g_lr, self.d_lr = stage1, mode, g_lr, d_lr
        self.gan_mode = GAN_MODES[mode]
        self.lam = 10.0
        self.dist = dist
        self.world = dist.world_size if dist is not None else 1
        # whole-call CUDA graphs of train_iteration's optimiser calls (a critic call = frozen-encoder forward, ~100
        # launches, + ~70 small FC launches + the update): same switch as Stage1Engine (DPIG_GRAPHS)
        gmode = int(os.environ.get("DPIG_GRAPHS", "1"))
        self.use_graphs = gmode >= 1 and (dist is None or getattr(dist, "capturable", True))
        self._graphs, self._eager_calls, self._lr_dev = {}, {}, {}
        # a critic call runs only the encoder pyramid its factor reads (exact: see encode_real); DPIG_STAGE2_PRUNE=0 runs
        # the whole encoder in every call, as the reference's sess.run does
        self.prune = os.environ.get("DPIG_STAGE2_PRUNE", "1") != "0"
        if factors is not None:      # custom factor set: the pose sampler of --model=4 (no Stage-I engine), or the single
            self.f = factors         # 'app' factor of the DeepFashion samplers (--model=103 / 1002) on a Stage-I engine
            first = next(iter(factors.values()))
            self.ctx, self.B = first.ctx, first.B
        else:
            self.ctx = stage1.ctx
            cfg, B, dev = stage1.cfg, stage1.B, stage1.device
            self.fg_dim = cfg.n_parts * cfg.part_z
            self.bg_dim = cfg.part_z * 4
            self.f = {"fg": _Factor(self.ctx, B, self.fg_dim, 512, "Gaussian_FC_Fg/G_FC", "Fg_FCDis_", dev),
                      "bg": _Factor(self.ctx, B, self.bg_dim, 256, "Gaussian_FC_Bg/G_FC", "Bg_FCDis_", dev)}
            self.B = B
        if mode == "wgan-gp":
            for f in self.f.values():
                f.p_gp = f.gp_program(self.lam)
                f.gp.v.zero_()     # Adam slots start at zero (RMSProp's start at one)
                f.dp.v.zero_()

    def param_groups(self):
        return [g for f in self.f.values() for g in (f.gp, f.dp)]

    def load_params(self, params):
        for grp in self.param_groups():
            for name in grp.specs:
                if name in params:
                    grp.view(name).copy_(torch.as_tensor(np.asarray(params[name]), dtype=torch.float32).to(grp.value.device))

    def get_params(self, grads=False):
        out = OrderedDict()
        for grp in self.param_groups():
            for name in grp.specs:
                out[name] = (grp.gview(name) if grads else grp.view(name)).detach().cpu().numpy().copy()
        return out

    def get_state(self):
        """Variables plus the optimiser slots under TensorFlow's slot names (`<var>/RMSProp` for the shipped wgan mode,
        `<var>/Adam`, `<var>/Adam_1` otherwise) and the per-factor step counters: what a --ckpt_path resume needs
        (tf.train.Saver() covers the slots, trainer.py:177-179)."""
        out = self.get_params()
        rms = self.mode in ("wgan", "lsgan")
        for fname, f in self.f.items():
            for which, grp in (("g", f.gp), ("d", f.dp)):
                for name, (off, n, shape) in grp.specs.items():
                    if rms:
                        out[name + "/RMSProp"] = grp.v[off:off + n].view(shape).detach().cpu().numpy().copy()
                    else:
                        out[name + "/Adam"] = grp.m[off:off + n].view(shape).detach().cpu().numpy().copy()
                        out[name + "/Adam_1"] = grp.v[off:off + n].view(shape).detach().cpu().numpy().copy()
                out["dpig_step_count/%s/%s" % (fname, which)] = np.int64(f.t[which])
        return out

    def load_state(self, state):
        self.load_params(state)
        rms = self.mode in ("wgan", "lsgan")
        for fname, f in self.f.items():
            for which, grp in (("g", f.gp), ("d", f.dp)):
                for name, (off, n, shape) in grp.specs.items():
                    for slot, arena in ((("/RMSProp", grp.v),) if rms else (("/Adam", grp.m), ("/Adam_1", grp.v))):
                        if name + slot in state:
                            arena[off:off + n].view(shape).copy_(torch.as_tensor(
                                np.asarray(state[name + slot]), dtype=torch.float32).to(arena.device).reshape(shape))
                key = "dpig_step_count/%s/%s" % (fname, which)
                if key in state:
                    f.t[which] = int(np.asarray(state[key]))

    def sample_noise(self, factor, z=None):
        f = self.f[factor]
        if z is None:
            f.z.data.normal_(0.0, 0.2)      # tf.random_normal(z_shape, 0.0, 0.2)   models.py:477
        else:
            f.z.data.copy_(torch.as_tensor(z, dtype=torch.float32).to(f.z.data.device))

    def encode_real(self, factor=None):
        """Real embeddings of the current Stage-I batch (frozen encoder forward, trainer.py:737-741).
        factor = "fg" / "bg" with self.prune: only the pyramid that factor's critic reads is run.  The reference's
        sess.run(d_optim_embs_fg) evaluates the whole encoder because fg_embs is a tf.slice of the concatenated
        embedding (trainer.py:741-742), but the Bg pyramid's output reaches nothing that call updates (and vice versa):
        the launches that produce the trained factor's embedding are the same ones on the same inputs, so dropping the
        other pyramid changes no result (tests/test_stage2_gpu.py::test_stage2_pruned_encoder_*)."""
        s = torch.cuda.current_stream().cuda_stream
        if "app" in self.f:          # one factor over the whole embedding (trainer_256.py:306-330)
            self.s1.run_encoder(s)
            self.f["app"].real.data.copy_(self.s1.emb)
            return
        if factor is not None and self.prune:
            self.s1.run_encoder(s, only="roi" if factor == "fg" else "bg")
        else:
            factor = None
            self.s1.run_encoder(s)
        if factor in (None, "fg"):
            self.f["fg"].real.data.copy_(self.s1.emb[:, :self.fg_dim])
        if factor in (None, "bg"):
            self.f["bg"].real.data.copy_(self.s1.emb[:, self.fg_dim:])

    def _optim(self, f, which, s):
        grp = f.gp if which == "g" else f.dp
        lr = self.g_lr if which == "g" else self.d_lr
        f.t[which] += 1
        if self.dist is not None:
            self.dist.all_reduce_sum(grp.grad)
        gs = 1.0 / self.world
        if self.mode in ("wgan", "lsgan"):
            clip = 0.01 if (self.mode == "wgan" and which == "d") else 0.0
            self.ctx.rmsprop_step(ptr(grp.value), ptr(grp.grad), ptr(grp.v), grp.total, lr, 0.9, 1e-10, gs, clip, s)
        else:
            b2 = 0.9 if self.mode == "wgan-gp" else 0.999
            self.ctx.adam_step(ptr(grp.value), ptr(grp.grad), ptr(grp.m), ptr(grp.v), grp.total, lr, 0.5, b2, 1e-8,
                               f.t[which], gs, s)

    def _optim_dev(self, f, which, lr_dev, s):
        """_optim with the step size read from device memory (dpig_*_step_dev): what a captured graph replays."""
        grp = f.gp if which == "g" else f.dp
        if self.dist is not None:
            self.dist.all_reduce_sum(grp.grad)
        gs = 1.0 / self.world
        if self.mode in ("wgan", "lsgan"):
            clip = 0.01 if (self.mode == "wgan" and which == "d") else 0.0
            self.ctx.rmsprop_step_dev(ptr(grp.value), ptr(grp.grad), ptr(grp.v), grp.total, ptr(lr_dev), 0.9, 1e-10, gs,
                                      clip, s)
        else:
            b2 = 0.9 if self.mode == "wgan-gp" else 0.999
            self.ctx.adam_step_dev(ptr(grp.value), ptr(grp.grad), ptr(grp.m), ptr(grp.v), grp.total, ptr(lr_dev), 0.5,
                                   b2, 1e-8, gs, s)

    def _step_size(self, f, which):
        lr = float(np.float32(self.g_lr if which == "g" else self.d_lr))
        if self.mode in ("wgan", "lsgan"):
            return lr
        b2 = float(np.float32(0.9 if self.mode == "wgan-gp" else 0.999))
        t = f.t[which]
        return lr * math.sqrt(1.0 - b2 ** t) / (1.0 - 0.5 ** t)

    def _call(self, factor, which):
        """One optimiser call of train_iteration on the Stage-I batch already in HBM: fresh noise, gradients, update
        (a critic call first runs the frozen encoder on the batch).  The first two calls of a kind run eagerly, the
        third is captured into a CUDA graph while it runs, later ones replay it; the step size is a device scalar."""
        f = self.f[factor]
        key = (factor, which, bool(self.prune))
        lr_dev = self._lr_dev.get(key)
        if lr_dev is None:
            lr_dev = self._lr_dev[key] = torch.zeros(1, dtype=torch.float32, device=f.loss.device)
        f.t[which] += 1
        lr_dev.fill_(self._step_size(f, which))

        def body():
            s = torch.cuda.current_stream().cuda_stream
            if which == "d" and self.s1 is not None:
                self.encode_real(factor)
            self.sample_noise(factor)
            (self.g_grads if which == "g" else self.d_grads)(factor)
            self._optim_dev(f, which, lr_dev, s)

        if not self.use_graphs:
            return body()
        g = self._graphs.get(key)
        if g is None:
            if self._eager_calls.get(key, 0) < 2:
                self._eager_calls[key] = self._eager_calls.get(key, 0) + 1
                return body()
            graph = torch.cuda.CUDAGraph()
            n0 = self.ctx.launch_count()
            # (thread_local: the NCCL watchdog thread polls events while this thread captures; under the default
            #  "global" mode such a call from another thread can invalidate the capture)
            with torch.cuda.graph(graph, capture_error_mode="thread_local" if self.dist is not None else "global"):
                body()
            g = self._graphs[key] = (graph, self.ctx.launch_count() - n0)
        g[0].replay()
        self.ctx.replayed_launches += g[1]

    def g_grads(self, factor):
        f = self.f[factor]
        s = torch.cuda.current_stream().cuda_stream
        f.gp.grad.zero_()
        f.p_g_fwd.run(s)
        c = f.crit["fake"]
        c["fwd"].run(s)
        self.ctx.loss_gan(self.gan_mode, None, ptr(c["out"].data), self.B, ptr(f.loss), ptr(c["out"].grad), None, None, s)
        c["bwd_data"].run(s)
        f.fake.grad.copy_(c["x"].grad)
        f.p_g_bwd.run(s)

    def d_grads(self, factor):
        f = self.f[factor]
        s = torch.cuda.current_stream().cuda_stream
        f.dp.grad.zero_()
        f.p_g_fwd.run(s)
        cr, cf = f.crit["real"], f.crit["fake"]
        cr["fwd"].run(s)
        cf["fwd"].run(s)
        self.ctx.loss_gan(self.gan_mode, ptr(cr["out"].data), ptr(cf["out"].data), self.B, ptr(f.loss), None,
                          ptr(cr["out"].grad), ptr(cf["out"].grad), s)
        cr["bwd_par"].run(s)
        cf["bwd_par"].run(s)
        if self.mode == "wgan-gp":
            if not f.gp_alpha_fixed:
                f.gp_alpha.uniform_(0.0, 1.0)
            f.p_gp.run(s)

    def g_step(self, factor):
        self.g_grads(factor)
        self._optim(self.f[factor], "g", torch.cuda.current_stream().cuda_stream)

    def d_step(self, factor):
        self.d_grads(factor)
        self._optim(self.f[factor], "d", torch.cuda.current_stream().cuda_stream)

    def train_iteration(self, step, next_batch):
        """trainer.py:821-845: per factor, one generator update (skipped at step 0) then CRITIC_ITERS critic updates
        (+clip, fused into the RMSProp kernel), every optimiser call on a fresh batch / fresh noise."""
        iters = 1 if self.mode in ("dcgan", "lsgan") else 5
        for factor in self.f:
            if step > 0:
                self._call(factor, "g")
            for _ in range(iters):
                self.s1.set_batch(next_batch())
                self._call(factor, "d")
