"""Static launch programs for the Stage-I graph of the reference (--model=1):

    build_model   reference trainer.py:568-625   (Encoder -> broadcast -> U-Net -> D(x), D(G) -> losses)
    train() step  reference trainer.py:336-347   (one g_optim update, then disc_ITERS d_optim updates)

The graph is static, so the engine allocates every activation / gradient buffer once, records the
sequence of C-ABI calls (ctypes function + frozen argument tuple) once, and a step is a replay of that
list on the current CUDA stream -- no autograd, no tracing compiler, no per-step allocation.  Backward
programs are written by hand (dgrad / wgrad kernels with fused ReLU masks and residual adds).

Data layout in HBM: NHWC split-bf16 activations (tensor.SplitTensor); skip connections are written
straight into the decoder's concat buffers (models.py:560 `tf.concat([x, skip])` never materialises);
parameters, their gradients and the optimiser slots are flat fp32 arenas (one NCCL all-reduce per step).
"""
import ctypes as C
import math
import os
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU, GAN_MODES, NORM_BATCH, NORM_LAYER
from .tensor import SplitTensor, ptr


def _pad8(c):
    return (c + 7) // 8 * 8


def halved(size, times):
    """Spatial size after `times` stride-2 SAME convolutions: ceil(size / 2) each (48 -> 24 -> 12 -> 6 -> 3 -> 2 -> 1)."""
    for _ in range(times):
        size = -(-size // 2)
    return size


class NetConfig:
    """Shapes of the Stage-I graphs.

    Market-1501 128x64, --model=1 (reference config.py:23-25, trainer.py:74-75, 576-582): Fg/Bg two-branch encoder
    (models.py:390-471), ROI pyramid and U-Net `repeat_num` levels deep, D applied to x and to G in separate calls.
    DeepFashion 256x256, --model=101 (reference trainer_256.py:31-68) = NetConfig.deepfashion(): `fgbg=False`
    (models.GeneratorCNN_ID_Encoder_BodyROIVis, models.py:328-388: no mask, no background branch), ROI pyramid
    `repeat_num+1` levels on 64x64 crops, U-Net `repeat_num-1` levels, D applied once to concat([x, G]) (`d_joint`:
    joint batch statistics; the logits are split in half afterwards)."""

    def __init__(self, img_h=128, img_w=64, hidden=128, z_num=64, roi_size=48, n_parts=7, part_z=32,
                 keypoints=18, d_dim=64, repeat_num=None, fgbg=True, enc_repeat=None, unet_repeat=None, d_joint=False,
                 use_vis=True):
        self.img_h, self.img_w, self.hidden, self.z_num = img_h, img_w, hidden, z_num
        # use_vis=False: models.GeneratorCNN_ID_Encoder_BodyROI (models.py:275-325, the DeepFashion samplers --model=103 /
        # 1002): the part features are NOT multiplied by the part visibilities
        self.use_vis = use_vis
        self.roi_size, self.n_parts, self.part_z, self.keypoints, self.d_dim = roi_size, n_parts, part_z, keypoints, d_dim
        self.repeat_num = repeat_num if repeat_num is not None else int(math.log2(img_h)) - 2
        self.fgbg, self.d_joint = fgbg, d_joint
        self.enc_repeat = enc_repeat if enc_repeat is not None else self.repeat_num
        self.unet_repeat = unet_repeat if unet_repeat is not None else self.repeat_num
        self.emb_dim = n_parts * part_z + (4 * part_z if fgbg else 0)
        # rows of D's Linear: tf.reshape(output, [-1, 8*4*8*dim]) (wgan_gp.py:433) is hard-wired to the 128x64 geometry,
        # so a 256x256 image yields 8 rows (= 8 logits) of 64 channels each; reduced test geometries whose
        # whole map is smaller keep one row per image
        self.d_row = min(8 * 4 * 8 * d_dim, (img_h // 16) * (img_w // 16) * 8 * d_dim)
        self.d_rows = (img_h // 16) * (img_w // 16) * 8 * d_dim // self.d_row

    @classmethod
    def deepfashion(cls, img_h=256, img_w=256, hidden=128, roi_size=64, **kw):
        """--model=101 (trainer_256.py:40-55): encoder repeat_num+1 levels, roi 64, U-Net repeat_num-1 levels."""
        rn = int(math.log2(img_h)) - 2
        return cls(img_h=img_h, img_w=img_w, hidden=hidden, roi_size=roi_size, repeat_num=rn, fgbg=False,
                   enc_repeat=rn + 1, unet_repeat=rn - 1, d_joint=True, **kw)


# ------------------------------------------------------------------------------------- parameters
class ParamGroup:
    """Flat fp32 arenas (value, grad, Adam m / v) with named views."""

    def __init__(self, specs, device, slots=True):
        self.specs = OrderedDict()
        off = 0
        for name, shape in specs:
            n = int(np.prod(shape))
            self.specs[name] = (off, n, tuple(shape))
            off += (n + 63) // 64 * 64
        self.total = off
        self.value = torch.zeros(off, device=device)
        # forward-only engines keep no gradient / optimiser arenas (one shared dummy element keeps the views valid)
        n = off if slots else 1
        self.grad = torch.zeros(n, device=device)
        self.m = torch.zeros(n, device=device)
        self.v = torch.zeros(n, device=device)

    def view(self, name, arena=None):
        off, n, shape = self.specs[name]
        return (self.value if arena is None else arena)[off:off + n].view(shape)

    def gview(self, name):
        return self.view(name, self.grad)


class ConvLayer:
    """One convolution: fp32 HWIO master inside a ParamGroup + packed bf16 operand copies."""

    def __init__(self, group, wname, bname, k, stride, cin, cout, device, need_bwd=True, small=False, cin_pad=None):
        self.group, self.wname, self.bname = group, wname, bname
        self.k, self.stride, self.cin, self.cout = k, stride, cin, cout
        # cin_pad must equal the channel count of the activation buffer the layer reads (zero-padded)
        self.cin_pad, self.cout_pad = (cin_pad or _pad8(cin)), _pad8(cout)
        self.small = small  # CUDA-core path (3-channel input): no packed copies
        taps = k * k
        if not small:
            self.fwd = torch.zeros((2, taps, cout, self.cin_pad), dtype=torch.bfloat16, device=device)
            self.bwd = torch.zeros((2, taps, cin, self.cout_pad), dtype=torch.bfloat16, device=device) if need_bwd else None

    @property
    def w(self):
        return self.group.view(self.wname)

    @property
    def b(self):
        return self.group.view(self.bname)

    @property
    def dw(self):
        return self.group.gview(self.wname)

    @property
    def db(self):
        return self.group.gview(self.bname)


class PatchLayer:
    """A convolution whose 3-channel side has been folded into the channel dimension (csrc/patch.cu), seen as a 1x1
    layer by conv_fwd / conv_wgrad: `master` is the fp32 [cin][cout] matrix the bf16 operand copy is packed from --
    for a cin = 3 layer the HWIO filter itself ([k*k*3][cout]), for the cout = 3 layer a re-laid copy."""

    def __init__(self, base, cin, cout, device, master=None, cout_launch=None, cin_pad=None):
        self.base, self.group, self.wname, self.bname = base, base.group, base.wname, base.bname
        self.k, self.stride, self.cin, self.rows = 1, 1, cin, cout
        # cout_launch > cout: the conv is launched with zero-padded output columns so that its epilogue takes the
        # whole-32-channel (coalesced) path; the extra weight rows stay zero
        self.cout = cout_launch or cout
        self.flops_cout = cout
        self.cin_pad, self.cout_pad = cin_pad or _pad8(cin), _pad8(self.cout)
        self.small, self.bwd = False, None
        self.fwd = torch.zeros((2, 1, self.cout, self.cin_pad), dtype=torch.bfloat16, device=device)
        self.master = master        # None: base.w

    w = property(lambda self: self.base.w)
    b = property(lambda self: self.base.b)
    dw = property(lambda self: self.base.dw)
    db = property(lambda self: self.base.db)

    def pack(self, ctx, stream):
        src = self.master if self.master is not None else self.base.w
        ctx.weight_pack(ptr(src), 1, self.cin, self.rows, self.cin_pad, _pad8(self.rows), ptr(self.fwd[0]),
                        ptr(self.fwd[1]), None, None, stream)


class Program:
    """A recorded list of C-ABI calls; run() replays it on the given stream."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.calls = []
        self.keep = []  # keeps ctypes structs / tensors referenced by the frozen arguments alive
        self.side = set()          # indices of calls that may run on the side stream (see add)
        self.side_stream = None
        self.two_streams = int(os.environ.get("DPIG_SIDE_STREAM", "1")) != 0

    def add(self, name, *args, flops=0.0, tag="", side=False):
        """side=True: the call may run on the program's side stream, concurrently with the calls that follow it on the
        main stream, until the next join (a python hook, or the end of the program).  Used for the filter gradients: a
        layer's wgrad and its dgrad both only READ dy, nothing later in a backward program overwrites what a wgrad
        reads, and the wgrads of one stream stay ordered among themselves (several accumulate into one dw)."""
        fn = getattr(self.ctx.lib, "dpig_" + name)
        if side:
            self.side.add(len(self.calls))
        self.calls.append((name, fn, args, flops, tag))

    def add_py(self, fn):
        self.calls.append((None, fn, None, 0.0, ""))

    def add_join(self):
        """The main stream waits here for everything launched on the side stream so far."""
        self.calls.append((None, None, None, 0.0, "join"))

    def _run_two_streams(self, stream):
        """Main-stream calls in order on `stream` (= torch's current stream), side calls on self.side_stream after an
        event that covers everything enqueued on the main stream so far; joins before python hooks and at the end.
        Under CUDA-graph capture the side stream forks from / joins the capturing stream through the same events."""
        h = self.ctx.handle
        main = torch.cuda.current_stream()
        if self.side_stream is None:
            self.side_stream = torch.cuda.Stream(device=main.device)
        side = self.side_stream
        pending = False

        def join():
            ev = torch.cuda.Event()
            ev.record(side)
            main.wait_event(ev)

        for i, (name, fn, args, flops, tag) in enumerate(self.calls):
            if name is None:
                if pending:
                    join()
                    pending = False
                if fn is not None:
                    fn(stream)
                continue
            if i in self.side:
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                rc = fn(h, *args, side.cuda_stream)
                pending = True
            else:
                rc = fn(h, *args, stream)
            if rc != 0:
                raise _lib.DpigError("dpig_%s failed (%d): %s" % (name, rc, self.ctx.last_error()))
        if pending:
            join()

    def run(self, stream, timings=None):
        """timings: optional list; when given, every C-ABI call is bracketed by CUDA events on the
        launching stream and (name, algorithmic_flops, start_event, end_event) is appended."""
        if self.side and timings is None and self.two_streams:
            return self._run_two_streams(stream)
        h = self.ctx.handle
        for name, fn, args, flops, tag in self.calls:
            if name is None:
                if fn is not None:
                    fn(stream)
                continue
            if timings is not None:
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            rc = fn(h, *args, stream)
            if timings is not None:
                e1.record()
                timings.append((name, flops, e0, e1, tag))
            if rc != 0:
                raise _lib.DpigError("dpig_%s failed (%d): %s" % (name, rc, self.ctx.last_error()))


def _mask(pixels, c, device):
    return torch.zeros((pixels, (c + 31) // 32), dtype=torch.int32, device=device)


def unpack_bits(mask, n, h, w, c):
    """A saved activation bit mask (one int32 word per pixel and 32 channels) as a bool [n, h, w, c] tensor."""
    bits = (mask.unsqueeze(-1) >> torch.arange(32, device=mask.device, dtype=torch.int32)) & 1
    return bits.reshape(mask.shape[0], -1)[:, :c].reshape(n, h, w, c).bool()


class _Pyramid:
    """`repeat_num` levels of {conv, conv, +res, [conv/s2]} (models.py:421-429, 454-462, 530-539)."""

    def __init__(self, eng, prefix, layers, x_in, n, h, w, rn, y_slots=None, in_mask=None):
        cfg = eng.cfg
        hn = cfg.hidden
        dev = eng.device
        self.layers = layers  # 3*rn - 1 ConvLayers in creation order
        self.n, self.rn = n, rn
        self.x_in, self.a, self.y, self.ma, self.mb, self.md = [x_in], [], [], [], [], []
        self.dims = []
        for idx in range(rn):
            c = hn * (idx + 1)
            hh, ww = halved(h, idx), halved(w, idx)
            self.dims.append((hh, ww, c))
            self.a.append(SplitTensor(n, hh, ww, c, dev))
            self.y.append(y_slots[idx] if y_slots else SplitTensor(n, hh, ww, c, dev))
            self.ma.append(_mask(n * hh * ww, c, dev))
            self.mb.append(_mask(n * hh * ww, c, dev))
            if idx < rn - 1:
                self.x_in.append(SplitTensor(n, halved(hh, 1), halved(ww, 1), c + hn, dev))
                self.md.append(_mask(n * halved(hh, 1) * halved(ww, 1), c + hn, dev))
        self.in_mask = in_mask
        self.in_layer = None   # ConvLayer producing the pyramid input (set by the owner when in_mask is used)
        if not eng.training:      # forward-only engine (sampling, tester.py): no gradient buffers
            return
        # backward buffers
        self.g_y = [SplitTensor(n, d[0], d[1], d[2], dev) for d in self.dims]
        self.gb = [SplitTensor(n, d[0], d[1], d[2], dev) for d in self.dims]
        self.ga = [SplitTensor(n, d[0], d[1], d[2], dev) for d in self.dims]
        self.gd = [SplitTensor(n, self.dims[i + 1][0], self.dims[i + 1][1], self.dims[i + 1][2], dev)
                   for i in range(rn - 1)]
        self.g_in = SplitTensor(n, h, w, hn, dev)  # grad wrt the pyramid input (masked if in_mask)

    def sign_bits(self):
        """The ReLU bits of the 3*rn - 1 convolutions in creation order (diagnostics, see Stage1Engine.activation_bits)."""
        out = []
        for idx in range(self.rn):
            hh, ww, c = self.dims[idx]
            out += [unpack_bits(self.ma[idx], self.n, hh, ww, c), unpack_bits(self.mb[idx], self.n, hh, ww, c)]
            if idx < self.rn - 1:
                nx = self.x_in[idx + 1]
                out.append(unpack_bits(self.md[idx], self.n, nx.h, nx.w, nx.c))
        return out

    def forward(self, eng, prog, side=False):
        li = 0
        for idx in range(self.rn):
            eng.conv_fwd(prog, self.layers[li], self.x_in[idx], out=self.a[idx], mask_out=self.ma[idx], side=side)
            eng.conv_fwd(prog, self.layers[li + 1], self.a[idx], out=self.y[idx], addend=self.x_in[idx],
                         mask_out=self.mb[idx], side=side)
            li += 2
            if idx < self.rn - 1:
                eng.conv_fwd(prog, self.layers[li], self.y[idx], out=self.x_in[idx + 1], mask_out=self.md[idx], side=side)
                li += 1

    def backward(self, eng, prog, skip_grads=None, wgrad=True):
        """Expects g_y[rn-1] (unmasked grad wrt the top level output, all contributions summed) to be
        filled by the caller.  skip_grads[idx] (idx < rn-1): extra grad wrt y[idx] (decoder skip path)."""
        rn = self.rn
        # top level: gb = g_y * mask_b
        eng.ew_combine_db(prog, self.gb[rn - 1], self.layers[3 * (rn - 1) + 1], self.g_y[rn - 1].ref(), None, None, None, 0,
                          ptr(self.mb[rn - 1]), 0.0, 0)
        for idx in range(rn - 1, -1, -1):
            l1, l2 = self.layers[3 * idx], self.layers[3 * idx + 1]
            hh, ww, c = self.dims[idx]
            if wgrad:
                eng.conv_wgrad(prog, l2, self.a[idx], self.gb[idx])
            eng.conv_dgrad(prog, l2, self.gb[idx], hh, ww, out_masked=self.ga[idx], mask_in=self.ma[idx], db_of=l1)
            if wgrad:
                eng.conv_wgrad(prog, l1, self.x_in[idx], self.ga[idx])
            if idx > 0:
                # x_in[idx] = relu(conv_s2(y[idx-1])): emit the masked gradient directly
                ls = self.layers[3 * (idx - 1) + 2]
                eng.conv_dgrad(prog, l1, self.ga[idx], hh, ww, out_masked=self.gd[idx - 1], mask_in=self.md[idx - 1],
                               addend=self.g_y[idx], db_of=ls)
                if wgrad:
                    eng.conv_wgrad(prog, ls, self.y[idx - 1], self.gd[idx - 1])
                ph, pw, pc = self.dims[idx - 1]
                eng.conv_dgrad(prog, ls, self.gd[idx - 1], ph, pw, out=self.g_y[idx - 1], out_masked=self.gb[idx - 1],
                               mask_in=self.mb[idx - 1], addend=skip_grads[idx - 1] if skip_grads else None,
                               db_of=self.layers[3 * (idx - 1) + 1])
            else:
                if self.in_mask is not None:
                    eng.conv_dgrad(prog, l1, self.ga[0], hh, ww, out_masked=self.g_in, mask_in=self.in_mask,
                                   addend=self.g_y[0], db_of=self.in_layer)
                else:
                    eng.conv_dgrad(prog, l1, self.ga[0], hh, ww, out=self.g_in, addend=self.g_y[0])


class _DiscPass:
    """Activations of one DCGANDiscriminator application (wgan_gp.py:407-440)."""

    def __init__(self, eng, n, segs=1):
        """n images in `segs` equal groups (2 = concat([x, G]), trainer_256.py:61-63).  The Linear sees
        cfg.d_rows rows per image; row (group g, row r, image b) of flat / logits sits at (g*d_rows + r)*(n/segs) + b,
        so the logits of each group are contiguous (tf.split(D_z, 2), trainer_256.py:66)."""
        cfg, dev = eng.cfg, eng.device
        d = cfg.d_dim
        H, W = cfg.img_h, cfg.img_w
        self.n, self.segs = n, segs
        self.patch = SplitTensor(n, H >> 1, W >> 1, _pad8(5 * 5 * 3), dev)   # im2col of the image (layer 1 as a 1x1 GEMM)
        self.h = [SplitTensor(n, H >> (i + 1), W >> (i + 1), d << i, dev) for i in range(4)]   # activated outputs
        self.m = [_mask(n * (H >> (i + 1)) * (W >> (i + 1)), d << i, dev) for i in range(4)]
        self.pre = [None] + [torch.zeros((n, H >> (i + 1), W >> (i + 1), d << i), device=dev) for i in (1, 2, 3)]
        self.groups = [None] + [(n if eng.norm_mode == NORM_LAYER else (d << i)) for i in (1, 2, 3)]
        self.sums = [None] + [torch.zeros((2, g), dtype=torch.float64, device=dev) for g in self.groups[1:]]
        self.stats = [None] + [torch.zeros((2, g), device=dev) for g in self.groups[1:]]
        self.red = [None] + [torch.zeros((2, g), dtype=torch.float64, device=dev) for g in self.groups[1:]]
        rows = n * cfg.d_rows
        self.flat = torch.zeros((rows, cfg.d_row), device=dev)
        self.logits = torch.zeros((rows,), device=dev)
        if not eng.training:
            return
        # backward
        self.dlogits = torch.zeros((rows,), device=dev)
        self.g_flat = torch.zeros((rows, cfg.d_row), device=dev)
        self.g_h = [SplitTensor(n, H >> (i + 1), W >> (i + 1), d << i, dev) for i in range(4)]  # grad wrt activated
        self.g_pre = [SplitTensor(n, H >> (i + 1), W >> (i + 1), d << i, dev) for i in range(4)]  # grad wrt conv out
        self.g_x = torch.zeros((n, H, W, 3), device=dev)


    def sign_bits(self, i):
        """The saved `pre-activation > 0` bits of layer i (0-based) as a bool NHWC tensor -- what the backward pass of
        LeakyReLU branches on (diagnostics; tests hand them to the oracle, oracle/nets.py dcgan_discriminator)."""
        h = self.h[i]
        return unpack_bits(self.m[i], h.n, h.h, h.w, h.c)


class _DiscHalf:
    """One half of a joint discriminator pass (views of its logits / logit gradients / image gradient)."""

    def __init__(self, dp, half):
        rows = dp.logits.numel() // 2
        n = dp.n // 2
        self.n = n
        self.logits = dp.logits[half * rows:(half + 1) * rows]
        if hasattr(dp, "dlogits"):
            self.dlogits = dp.dlogits[half * rows:(half + 1) * rows]
            self.g_x = dp.g_x[half * n:(half + 1) * n]


class Stage1Engine:
    def __init__(self, ctx, cfg, batch, mode="dcgan", lam=10.0, dist=None, device="cuda", inference=False):
        """dist: optional object with all_reduce_sum(tensor) and world_size (data-parallel hooks).
        inference: forward-only engine (tester.py: sampling at batch 512) -- no gradient buffers, no backward programs,
        no optimiser slots: ~1/2 of the activation memory of a training engine, none of its parameter-sized slots."""
        self.ctx, self.cfg, self.B, self.mode, self.lam, self.dist = ctx, cfg, int(batch), mode, lam, dist
        self.training = not inference
        self.device = torch.device(device)
        self.gan_mode = GAN_MODES[mode]
        self.norm_mode = NORM_LAYER if mode == "wgan-gp" else NORM_BATCH  # wgan_gp.py:34-40
        self.world = dist.world_size if dist is not None else 1
        if mode == "wgan-gp" and (cfg.d_joint or cfg.d_rows != 1):
            # the reference's 256x256 trainers hard-wire MODE='dcgan' (trainer_256.py:28)
            raise _lib.DpigError("wgan-gp is built for the 128x64 graph only (one logit per image, separate D calls)")
        self._keep = []
        self._db_done = set()
        self.fuse_bias_grad = True
        self._build_params()
        self._build_buffers()
        if mode in ("wgan", "lsgan"):      # TF RMSProp's 'rms' slot is initialised to ones
            self.gp.v.fill_(1.0)
            self.dp.v.fill_(1.0)
        self.t = {"g": 0, "d": 0}
        # step size of the next optimiser call, in device memory (dpig_adam_step_dev / dpig_rmsprop_step_dev): lets one
        # captured CUDA graph serve every step although Adam's bias-corrected lr_t changes with t
        self.lr_dev = {"g": torch.zeros(1, dtype=torch.float32, device=self.device),
                       "d": torch.zeros(1, dtype=torch.float32, device=self.device)}
        # whole-step CUDA graphs (forward + backward + update + weight re-pack = ~220 launches replayed by one
        # cudaGraphLaunch): on by default on one GPU; with torch.distributed the NCCL exchanges would have to be
        # captured too, opt in with DPIG_GRAPHS=2.  DPIG_GRAPHS=0 switches them off.
        gmode = int(os.environ.get("DPIG_GRAPHS", "1"))
        # N > 1: the NCCL exchanges (gradient slices, sync-BN sums) are captured into the step graphs too; a `dist` that
        # cannot be captured (ddp.LocalGroup's in-process ranks) says so with `capturable = False`
        self.use_graphs = gmode >= 1 and (dist is None or getattr(dist, "capturable", True))
        self._graphs = {}
        self._eager_steps = {"g": 0, "d": 0}
        # N > 1: the ID_AE (U-Net) slice of the gradient arena is complete before the appearance encoder's backward
        # starts; its all-reduce (285 of the 474 MB) runs on a side stream under the remaining ~40 % of the backward
        # pass (DPIG_OVERLAP=0: one all-reduce of the whole arena after the backward pass)
        self.overlap_comm = dist is not None and int(os.environ.get("DPIG_OVERLAP", "1")) != 0
        self._comm_stream = None
        self._comm_done = None
        self._early_ranges = []
        self.gp_alpha_fixed = False   # tests pin alpha to compare with the oracle
        self.g_lr = 2e-5
        self.d_lr = 2e-5
        self._build_programs()

    # -------------------------------------------------------------------------------- parameters
    def _build_params(self):
        cfg, dev = self.cfg, self.device
        hn, rn, ern = cfg.hidden, cfg.unet_repeat, cfg.enc_repeat
        gspecs, dspecs = [], []
        self.layers = OrderedDict()

        def conv(group_specs, scope, counter, k, stride, cin, cout, small=False, need_bwd=True, cin_pad=None):
            i = counter[0]
            counter[0] += 1
            name = "%s/Conv%s" % (scope, "" if i == 0 else "_%d" % i)
            group_specs.append((name + "/weights", (k, k, cin, cout)))
            group_specs.append((name + "/biases", (cout,)))
            self.layers[name] = dict(k=k, stride=stride, cin=cin, cout=cout, small=small, need_bwd=need_bwd,
                                     cin_pad=cin_pad)
            return name

        def fc(group_specs, scope, counter, cin, cout):
            i = counter[0]
            counter[0] += 1
            name = "%s/fully_connected%s" % (scope, "" if i == 0 else "_%d" % i)
            group_specs.append((name + "/weights", (cin, cout)))
            group_specs.append((name + "/biases", (cout,)))
            return name

        def pyramid(scope, cc, levels):
            names = []
            for idx in range(levels):
                c = hn * (idx + 1)
                names.append(conv(gspecs, scope, cc, 3, 1, c, c))
                names.append(conv(gspecs, scope, cc, 3, 1, c, c))
                if idx < levels - 1:
                    names.append(conv(gspecs, scope, cc, 3, 2, c, hn * (idx + 2)))
            return names

        # Encoder/G_encoder (models.py:390-471)
        sc, cc, fc_c = "Encoder/G_encoder", [0], [0]
        self.n_e0 = conv(gspecs, sc, cc, 3, 1, 3, hn, need_bwd=False)
        self.n_e1 = conv(gspecs, sc, cc, 3, 1, hn, hn)
        self.n_e2 = conv(gspecs, sc, cc, 3, 1, hn, hn)
        self.n_roi = pyramid(sc, cc, ern)
        roi_f = halved(cfg.roi_size, ern - 1)
        self.roi_flat = roi_f * roi_f * hn * ern
        self.n_roi_fc = fc(gspecs, sc, fc_c, self.roi_flat, cfg.part_z)
        if cfg.fgbg:   # background branch of the two-branch encoder (models.py:454-464)
            self.n_bg = pyramid(sc, cc, ern)
            self.bg_flat = (cfg.img_h >> (ern - 1)) * (cfg.img_w >> (ern - 1)) * hn * ern
            self.n_bg_fc = fc(gspecs, sc, fc_c, self.bg_flat, cfg.part_z * 4)
        self.fh, self.fw = cfg.img_h >> (rn - 1), cfg.img_w >> (rn - 1)
        self.gtop_flat = self.fh * self.fw * hn * rn
        # ID_AE/G (models.py:518-576)
        sc, cc, fc_c = "ID_AE/G", [0], [0]
        self.gin_c = cfg.emb_dim + cfg.keypoints
        self.pose_cpad = _pad8(cfg.keypoints)
        self.n_gstem = conv(gspecs, sc, cc, 3, 1, self.gin_c, hn, need_bwd=False)
        self.n_genc = pyramid(sc, cc, rn)
        self.n_gfc1 = fc(gspecs, sc, fc_c, self.gtop_flat, cfg.z_num)
        self.n_gfc2 = fc(gspecs, sc, fc_c, cfg.z_num, self.fh * self.fw * hn)
        self.n_gdec = []
        self.dec_c = []
        x_c = hn
        for idx in range(rn):
            c = x_c + hn * (rn - idx)
            self.dec_c.append((x_c, c))
            names = [conv(gspecs, sc, cc, 3, 1, c, c), conv(gspecs, sc, cc, 3, 1, c, c)]
            if idx < rn - 1:
                x_c = hn * (rn - idx - 1)
                names.append(conv(gspecs, sc, cc, 1, 1, c, x_c))
            else:
                x_c = c
            self.n_gdec.append(names)
        self.n_gout = conv(gspecs, sc, cc, 3, 1, x_c, 3)
        # Discriminator (wgan_gp.py:407-440)
        d = cfg.d_dim
        chans = [3, d, 2 * d, 4 * d, 8 * d]
        self.n_d = []
        for i in range(4):
            name = "Discriminator.%d" % (i + 1)
            dspecs.append((name + ".Filters", (5, 5, chans[i], chans[i + 1])))
            dspecs.append((name + ".Biases", (chans[i + 1],)))
            if i >= 1:
                dspecs.append(("Discriminator.BN%d.offset" % (i + 1), (chans[i + 1],)))
                dspecs.append(("Discriminator.BN%d.scale" % (i + 1), (chans[i + 1],)))
            self.layers[name] = dict(k=5, stride=2, cin=chans[i], cout=chans[i + 1], small=False, need_bwd=True,
                                     cin_pad=None)
            self.n_d.append(name)
        self.d_flat = cfg.d_row
        dspecs.append(("Discriminator.Output.W", (self.d_flat, 1)))
        dspecs.append(("Discriminator.Output.b", (1,)))

        self.gp = ParamGroup(gspecs, dev, slots=self.training)
        self.dp = ParamGroup(dspecs, dev, slots=self.training)
        self.conv = OrderedDict()
        for name, s in self.layers.items():
            if name == self.n_gstem:
                # stem shortcut: the 352 embedding rows never reach the tensor cores (dpig_stem_class_bias);
                # the packed operand holds the 18 pose rows only
                lay = ConvLayer(self.gp, name + "/weights", name + "/biases", 3, 1, self.gin_c, hn, dev, False, True)
                lay.stem = True                 # no packed copies of its own: the pose rows run in patch form (below)
                self.conv[name] = lay
                continue
            if name.startswith("Discriminator"):
                self.conv[name] = ConvLayer(self.dp, name + ".Filters", name + ".Biases", s["k"], s["stride"], s["cin"],
                                            s["cout"], dev, s["need_bwd"] and self.training, s["small"], s["cin_pad"])
            else:
                self.conv[name] = ConvLayer(self.gp, name + "/weights", name + "/biases", s["k"], s["stride"], s["cin"],
                                            s["cout"], dev, s["need_bwd"] and self.training, s["small"], s["cin_pad"])
        # 3-channel ends as 1x1 contractions over (tap, channel) patches (csrc/patch.cu)
        e0, d1, go = self.conv[self.n_e0], self.conv[self.n_d[0]], self.conv[self.n_gout]
        self.e0_patch = PatchLayer(e0, 9 * 3, hn, dev)
        self.d1_patch = PatchLayer(d1, 25 * 3, d, dev)
        self.gout_wf = torch.zeros((go.cin, 27), device=dev)      # [ci][tap*3+co]  forward: 1x1 conv to 27 tap channels
        self.gout_wd = torch.zeros((27, go.cin), device=dev)      # [tap*3+co][ci]  data gradient: 1x1 conv from dy patches
        self.gout_dwv = torch.zeros((go.cin, 27), device=dev)     # filter gradient in the forward layout
        self.gout_f = PatchLayer(go, go.cin, 27, dev, master=self.gout_wf, cout_launch=32)
        self.gout_d = PatchLayer(go, 27, go.cin, dev, master=self.gout_wd)
        # pose rows of the U-Net stem (18 of its 370 input channels; the 352 embedding rows are the class-bias shortcut):
        # as a 3x3 conv every filter tap spent a whole 64-deep K chunk on 18 channels (45 TFLOP/s); with the 9 taps folded
        # into the channel dimension it is a 1x1 contraction over K = 9*18 = 162 (3 chunks instead of 9) and its filter
        # gradient one M = 2x128-row GEMM instead of nine
        kp = 9 * cfg.keypoints
        self.stem_wp = torch.zeros((kp, hn), device=dev)       # [tap*18 + c][co] = W[tap][352 + c][co]
        self.stem_dwp = torch.zeros((kp, hn), device=dev)
        self.stem_patch = PatchLayer(self.conv[self.n_gstem], kp, hn, dev, master=self.stem_wp, cin_pad=(kp + 63) // 64 * 64)
        # BN / LN scale defaults to one
        for i in (2, 3, 4):
            self.dp.view("Discriminator.BN%d.scale" % i).fill_(1.0)
        # the D output weight is kept NHWC-flattened internally; TF order (c-major) at the API boundary
        hw = (cfg.img_h // 16) * (cfg.img_w // 16)
        self._dperm = torch.arange(self.d_flat, device=dev).view(8 * d // cfg.d_rows, hw).t().reshape(-1)  # nhwc idx -> nchw idx

    def param_names(self):
        return list(self.gp.specs) + list(self.dp.specs)

    def load_params(self, params):
        """params: dict TF-variable-name -> array (TF layouts). Missing names keep their current value."""
        for grp in (self.gp, self.dp):
            for name in grp.specs:
                if name in params:
                    t = torch.as_tensor(np.asarray(params[name]), dtype=torch.float32).to(self.device)
                    if name == "Discriminator.Output.W":
                        t = t.reshape(-1)[self._dperm].reshape(-1, 1)
                    grp.view(name).copy_(t.reshape(grp.specs[name][2]))
        self.pack_weights("g")
        self.pack_weights("d")
        torch.cuda.synchronize()

    def get_params(self, grads=False):
        out = OrderedDict()
        for grp in (self.gp, self.dp):
            for name in grp.specs:
                t = (grp.gview(name) if grads else grp.view(name)).detach().clone()
                if name == "Discriminator.Output.W":
                    inv = torch.empty_like(self._dperm)
                    inv[self._dperm] = torch.arange(self.d_flat, device=self.device)
                    t = t.reshape(-1)[inv].reshape(-1, 1)
                out[name] = t.cpu().numpy()
        return out

    def get_state(self):
        """Everything tf.train.Saver would write for this graph (trainer.py:365-366): the variables plus the optimiser
        slots under TensorFlow's slot names (`<var>/Adam` = m, `<var>/Adam_1` = v; RMSProp: `<var>/RMSProp` = ms) and the
        optimisers' step counters as beta-power accumulators (`beta1_power`, `beta2_power` for g_optim; `_1` for d_optim).
        `Discriminator.BNk.moving_mean / moving_variance` are exported at their never-updated initial values (q4)."""
        out = self.get_params()
        rms = self.mode in ("wgan", "lsgan")
        b2 = 0.9 if self.mode == "wgan-gp" else 0.999
        for which, grp, suffix in (("g", self.gp, ""), ("d", self.dp, "_1")):
            for name in grp.specs:
                off, n, shape = grp.specs[name]
                m, v = grp.m[off:off + n].view(shape), grp.v[off:off + n].view(shape)
                if name == "Discriminator.Output.W":
                    inv = torch.empty_like(self._dperm)
                    inv[self._dperm] = torch.arange(self.d_flat, device=self.device)
                    m, v = m.reshape(-1)[inv].reshape(-1, 1), v.reshape(-1)[inv].reshape(-1, 1)
                if rms:
                    out[name + "/RMSProp"] = v.cpu().numpy().copy()
                else:
                    out[name + "/Adam"] = m.cpu().numpy().copy()
                    out[name + "/Adam_1"] = v.cpu().numpy().copy()
            if not rms:
                out["beta1_power" + suffix] = np.float32(0.5 ** (self.t[which] + 1))
                out["beta2_power" + suffix] = np.float32(b2 ** (self.t[which] + 1))
            # the float32 beta powers underflow (0.999^t at t ~ 104k, 0.5^t at t ~ 150): the step counter itself travels
            # too, under a name no TensorFlow graph reads (a reference restore ignores it)
            out["dpig_step_count" + suffix] = np.int64(self.t[which])
        for i in (2, 3, 4):
            c = self.cfg.d_dim << (i - 1)
            out["Discriminator.BN%d.moving_mean" % i] = np.zeros(c, np.float32)
            out["Discriminator.BN%d.moving_variance" % i] = np.ones(c, np.float32)
        return out

    def load_state(self, state):
        """Inverse of get_state(): variables by name (missing names keep their value), optimiser slots and step counters
        when present (a full `--ckpt_path` restore resumes Adam exactly, trainer.py:211-213)."""
        self.load_params(state)
        rms = self.mode in ("wgan", "lsgan")
        b2 = 0.9 if self.mode == "wgan-gp" else 0.999
        for which, grp, suffix in (("g", self.gp, ""), ("d", self.dp, "_1")):
            for name in grp.specs:
                off, n, shape = grp.specs[name]
                for slot, arena in ((("/RMSProp", grp.v),) if rms else (("/Adam", grp.m), ("/Adam_1", grp.v))):
                    if name + slot in state:
                        t = torch.as_tensor(np.asarray(state[name + slot]), dtype=torch.float32).to(self.device)
                        if name == "Discriminator.Output.W":
                            t = t.reshape(-1)[self._dperm].reshape(-1, 1)
                        arena[off:off + n].view(shape).copy_(t.reshape(shape))
            self.t[which] = self._restored_step_count(state, suffix, b2, rms, self.t[which])

    @staticmethod
    def _restored_step_count(state, suffix, b2, rms, default):
        """Adam's t of a restored optimiser: the explicit counter when this implementation wrote the checkpoint, else
        recovered from TensorFlow's beta-power accumulators in float64 (beta2_power = b2^(t+1) while it has not
        underflowed in float32, then beta1_power); powers that have underflowed to 0 mean t is so large that the bias
        correction is 1 to float32 precision, which any large t reproduces."""
        if "dpig_step_count" + suffix in state:
            return int(np.asarray(state["dpig_step_count" + suffix]))
        if rms:
            return default
        for key, base in (("beta2_power" + suffix, b2), ("beta1_power" + suffix, 0.5)):
            if key in state:
                v = float(np.asarray(state[key], dtype=np.float64))
                if 0.0 < v < 1.0:
                    return max(0, int(round(math.log(v) / math.log(float(base)))) - 1)
                if v == 0.0:
                    return 1 << 24
        return default

    def pack_weights(self, which, stream=None):
        s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        for name, layer in self.conv.items():
            if layer.small or (name.startswith("Discriminator") != (which == "d")):
                continue
            if getattr(layer, "stem", False):
                continue
            self.ctx.weight_pack(ptr(layer.w), layer.k * layer.k, layer.cin, layer.cout, layer.cin_pad, layer.cout_pad,
                                 ptr(layer.fwd[0]), ptr(layer.fwd[1]),
                                 ptr(layer.bwd[0]) if layer.bwd is not None else None,
                                 ptr(layer.bwd[1]) if layer.bwd is not None else None, s)
        if which == "d":
            self.d1_patch.pack(self.ctx, s)
        else:
            self.e0_patch.pack(self.ctx, s)
            stem = self.conv[self.n_gstem]
            # (a torch copy on the current stream: every caller passes the current stream, like the add_py steps)
            self.stem_wp.copy_(stem.w.view(9, self.gin_c, self.cfg.hidden)[:, self.cfg.emb_dim:, :]
                               .reshape(self.stem_wp.shape))
            self.stem_patch.pack(self.ctx, s)
            go = self.gout_f.base
            self.ctx.permute_taps(ptr(go.w), ptr(self.gout_wf), 9, go.cin, 3, 0, s)
            self.ctx.permute_taps(ptr(go.w), ptr(self.gout_wd), 9, go.cin, 3, 1, s)
            self.gout_f.pack(self.ctx, s)
            self.gout_d.pack(self.ctx, s)

    # -------------------------------------------------------------------------------- buffers
    def _build_buffers(self):
        cfg, dev, B = self.cfg, self.device, self.B
        H, W, hn, rn, ern = cfg.img_h, cfg.img_w, cfg.hidden, cfg.unet_repeat, cfg.enc_repeat
        P = cfg.n_parts
        # inputs (device copies of one batch)
        self.x = torch.zeros((B, H, W, 3), device=dev)
        if cfg.d_joint:   # concat([x, G]) (trainer_256.py:61): x and G are the two halves of one buffer
            self.pair8 = SplitTensor(2 * B, H, W, 8, dev, zero=True)
            self.x8 = self.pair8.batch_slice(0, B)
        else:
            self.x8 = SplitTensor(B, H, W, 8, dev, zero=True)   # image as a channel-padded split tensor (TMA operand)
        self.pose_rcv = torch.zeros((B, cfg.keypoints, 3), device=dev)
        self.fg_mask = torch.zeros((B, H, W), device=dev)
        self.boxes = torch.zeros((P * B, 4), device=dev)
        self.box_ind = (torch.arange(P * B, device=dev, dtype=torch.int32) % B).contiguous()
        self.vis = torch.zeros((B, P), device=dev)
        # encoder
        self.x_patch = SplitTensor(B, H, W, _pad8(27), dev)     # im2col of the image: the stem as a 1x1 GEMM
        self.e0 = SplitTensor(B, H, W, hn, dev)
        self.me0 = _mask(B * H * W, hn, dev)
        self.e1 = SplitTensor(B, H, W, hn, dev)
        self.me1 = _mask(B * H * W, hn, dev)
        self.xs = SplitTensor(B, H, W, hn, dev)
        self.me2 = _mask(B * H * W, hn, dev)
        self.rois = SplitTensor(P * B, cfg.roi_size, cfg.roi_size, hn, dev)
        self.roi_pyr = _Pyramid(self, "roi", [self.conv[n] for n in self.n_roi], self.rois, P * B, cfg.roi_size,
                                cfg.roi_size, ern)
        self.roi_flat_f32 = torch.zeros((P * B, self.roi_flat), device=dev)
        self.fea = torch.zeros((P * B, cfg.part_z), device=dev)
        self.bg_z = cfg.part_z * 4 if cfg.fgbg else 0
        if cfg.fgbg:
            self.x_bg = SplitTensor(B, H, W, hn, dev)
            self.bg_pyr = _Pyramid(self, "bg", [self.conv[n] for n in self.n_bg], self.x_bg, B, H, W, ern)
            self.bg_flat_f32 = torch.zeros((B, self.bg_flat), device=dev)
        self.bg_fea = torch.zeros((B, max(self.bg_z, 1)), device=dev)
        self.emb = torch.zeros((B, cfg.emb_dim), device=dev)
        # generator
        self.gin = SplitTensor(B, H, W, self.pose_cpad, dev, zero=True)      # pose maps only (channels 0..17)
        self.pose_patch = SplitTensor(B, H, W, self.stem_patch.cin_pad, dev, zero=True)   # their 3x3 patches [tap*18 + c]
        self.stem_e = torch.zeros((9, B, hn), device=dev)                    # E[tap][n][co] = emb . W[tap, :352]
        self.stem_cb = torch.zeros((B, 9, hn), device=dev)                   # per-image border-class bias
        self.stem_cls = torch.zeros((B, 9, hn), device=dev)
        self.stem_ts = torch.zeros((9, B, hn), device=dev)
        self.stem_tmp = torch.zeros((B, cfg.emb_dim), device=dev)
        self.g0 = SplitTensor(B, H, W, hn, dev)
        self.mg0 = _mask(B * H * W, hn, dev)
        # decoder concat buffers; the encoder skip outputs are slices of them
        self.cat = []
        for idx in range(rn):
            xc, c = self.dec_c[idx]
            lvl = rn - 1 - idx
            self.cat.append(SplitTensor(B, H >> lvl, W >> lvl, c, dev))
        y_slots = [None] * rn
        for idx in range(rn):
            xc, c = self.dec_c[idx]
            y_slots[rn - 1 - idx] = self.cat[idx].slice(xc, c - xc)
        self.genc = _Pyramid(self, "genc", [self.conv[n] for n in self.n_genc], self.g0, B, H, W, rn, y_slots=y_slots,
                             in_mask=self.mg0)
        self.genc.in_layer = self.conv[self.n_gstem]
        self.gtop_f32 = torch.zeros((B, self.gtop_flat), device=dev)
        self.z = torch.zeros((B, cfg.z_num), device=dev)
        self.dec_in_f32 = torch.zeros((B, self.fh * self.fw * hn), device=dev)
        self.dec_a, self.dec_y, self.dec_ma, self.dec_mb, self.dec_mu = [], [], [], [], []
        for idx in range(rn):
            xc, c = self.dec_c[idx]
            lvl = rn - 1 - idx
            hh, ww = H >> lvl, W >> lvl
            self.dec_a.append(SplitTensor(B, hh, ww, c, dev))
            self.dec_y.append(SplitTensor(B, hh, ww, c, dev))
            self.dec_ma.append(_mask(B * hh * ww, c, dev))
            self.dec_mb.append(_mask(B * hh * ww, c, dev))
            if idx < rn - 1:
                self.dec_mu.append(_mask(B * hh * ww, self.dec_c[idx + 1][0], dev))
        self.G = torch.zeros((B, H, W, 3), device=dev)
        self.gout_y = torch.zeros((B, H, W, 32), device=dev)    # per-tap partial outputs of the 256 -> 3 conv
        self.G8 = self.pair8.batch_slice(B, B) if cfg.d_joint else SplitTensor(B, H, W, 8, dev, zero=True)
        if self.training:
            self._build_backward_buffers()
        # discriminator passes
        if cfg.d_joint:
            self.d_pair = _DiscPass(self, 2 * B, segs=2)
            self.d_real, self.d_fake = _DiscHalf(self.d_pair, 0), _DiscHalf(self.d_pair, 1)
        else:
            self.d_real = _DiscPass(self, B)
            self.d_fake = _DiscPass(self, B)
        if self.mode == "wgan-gp" and self.training:
            self._build_gp_buffers()
        # losses: [g_gan, d_loss], [L1], gp
        self.loss_gan = torch.zeros((2,), device=dev)
        self.loss_l1 = torch.zeros((1,), device=dev)
        self.loss_gp = torch.zeros((1,), device=dev)

    def _build_backward_buffers(self):
        cfg, dev, B = self.cfg, self.device, self.B
        H, W, hn, rn, ern = cfg.img_h, cfg.img_w, cfg.hidden, cfg.unet_repeat, cfg.enc_repeat
        P = cfg.n_parts
        # generator backward
        self.gG_patch = SplitTensor(B, H, W, 32, dev)           # transposed patches of dL/dG
        self.g_G = torch.zeros((B, H, W, 3), device=dev)
        self.g_G8 = SplitTensor(B, H, W, 8, dev, zero=True)
        self.g_cat, self.dec_gy, self.dec_gb, self.dec_ga, self.dec_gu = [], [], [], [], []
        for idx in range(rn):
            xc, c = self.dec_c[idx]
            lvl = rn - 1 - idx
            hh, ww = H >> lvl, W >> lvl
            self.g_cat.append(SplitTensor(B, hh, ww, c, dev))
            self.dec_gy.append(SplitTensor(B, hh, ww, c, dev))
            self.dec_gb.append(SplitTensor(B, hh, ww, c, dev))
            self.dec_ga.append(SplitTensor(B, hh, ww, c, dev))
            if idx < rn - 1:
                self.dec_gu.append(SplitTensor(B, hh, ww, self.dec_c[idx + 1][0], dev))
        self.g_dec_in_f32 = torch.zeros((B, self.fh * self.fw * hn), device=dev)
        self.g_z = torch.zeros((B, cfg.z_num), device=dev)
        self.g_gtop_f32 = torch.zeros((B, self.gtop_flat), device=dev)
        self.g_emb = torch.zeros((B, cfg.emb_dim), device=dev)
        # encoder backward
        self.g_fea = torch.zeros((P * B, cfg.part_z), device=dev)
        self.g_bg_fea = torch.zeros((B, max(self.bg_z, 1)), device=dev)
        self.g_roi_flat = torch.zeros((P * B, self.roi_flat), device=dev)
        self.g_crop = torch.zeros((B, H, W, hn), device=dev)
        if cfg.fgbg:
            self.g_bg_flat = torch.zeros((B, self.bg_flat), device=dev)
            self.g_xbg_s = SplitTensor(B, H, W, hn, dev)
        self.g_xs = SplitTensor(B, H, W, hn, dev)
        self.g_xs_m = SplitTensor(B, H, W, hn, dev)
        self.g_e1 = SplitTensor(B, H, W, hn, dev)
        self.g_e0 = SplitTensor(B, H, W, hn, dev)

    def _build_gp_buffers(self):
        """Interpolates, critic pass on them, tangent (JVP) and adjoint buffers of the gradient penalty."""
        cfg, dev, B = self.cfg, self.device, self.B
        H, W, d = cfg.img_h, cfg.img_w, cfg.d_dim
        self.gp_alpha = torch.zeros((B,), device=dev)
        self.xhat = torch.zeros((B, H, W, 3), device=dev)
        self.xhat8 = SplitTensor(B, H, W, 8, dev, zero=True)
        self.gp_v = torch.zeros((B, H, W, 3), device=dev)
        self.gp_v8 = SplitTensor(B, H, W, 8, dev, zero=True)
        self.gp_v_patch = SplitTensor(B, H >> 1, W >> 1, _pad8(75), dev)
        self.slopes = torch.zeros((B,), device=dev)
        self.ones_b = torch.ones((B,), device=dev)
        dh = self.d_hat = _DiscPass(self, B)
        dh.dlogits.fill_(1.0)

        def st(i):
            return SplitTensor(B, H >> (i + 1), W >> (i + 1), d << i, dev)

        dh.hd = [st(i) for i in range(4)]        # tangent activations  hdot_i
        dh.hdbar = [st(i) for i in range(4)]     # adjoints of hdot_i
        dh.pdbar = [st(i) for i in range(4)]     # adjoints of the tangent pre-norm conv outputs
        dh.pbar = [st(i) for i in range(4)]      # adjoints of the primal conv outputs (total)
        dh.pbarP = [st(i) for i in range(4)]     # ... part through the primal chain
        dh.hbar = [st(i) for i in range(4)]      # adjoints of the primal activations
        dh.pd = [None] + [torch.zeros((B, H >> (i + 1), W >> (i + 1), d << i), device=dev) for i in (1, 2, 3)]
        dh.pbarT = [None] + [torch.zeros((B, H >> (i + 1), W >> (i + 1), d << i), device=dev) for i in (1, 2, 3)]
        dh.tsums = [None] + [torch.zeros((2, B), dtype=torch.float64, device=dev) for _ in range(3)]
        dh.asums = [None] + [torch.zeros((3, B), dtype=torch.float64, device=dev) for _ in range(3)]
        dh.hd_flat = torch.zeros((B, self.d_flat), device=dev)
        dh.hbar_flat = torch.zeros((B, self.d_flat), device=dev)

    # -------------------------------------------------------------------------------- call helpers
    def _epilogue(self, prog, layer_bias, act, alpha, addend, mask_in, mask_neg, mask_out, out, out_masked, out_f32,
                  out_f32_ps, upsample, class_bias=None, colsum=None, stat_sums=None, stat_mode=0):
        ep = _lib.ConvEpilogue()
        if stat_sums is not None:
            ep.stat_sums = stat_sums.data_ptr()
            ep.stat_mode = stat_mode
        if class_bias is not None:
            ep.class_bias = class_bias.data_ptr()
        if colsum is not None:
            ep.colsum_masked = colsum.data_ptr()
        ep.bias = layer_bias.data_ptr() if layer_bias is not None else None
        ep.act = act
        ep.alpha = alpha
        if addend is not None:
            ep.addend = C.pointer(addend.struct())
        if mask_in is not None:
            ep.mask_in = mask_in.data_ptr()
        ep.mask_neg = mask_neg
        if mask_out is not None:
            ep.mask_out = mask_out.data_ptr()
        if out is not None:
            ep.out = C.pointer(out.struct())
        if out_masked is not None:
            ep.out_masked = C.pointer(out_masked.struct())
        if out_f32 is not None:
            ep.out_f32 = out_f32.data_ptr()
            ep.out_f32_pix_stride = out_f32_ps
        ep.upsample = upsample
        prog.keep.append((ep, addend, out, out_masked))
        return C.byref(ep)

    def conv_fwd(self, prog, layer, x, out=None, act=ACT_RELU, alpha=0.2, addend=None, mask_out=None, out_f32=None,
                 out_f32_ps=0, upsample=1, bias=True, out_masked=None, mask_in=None, mask_neg=0.0, class_bias=None,
                 stat_sums=None, stat_mode=0, side=False):
        ep = self._epilogue(prog, layer.b if bias else None, act, alpha, addend, mask_in, mask_neg, mask_out, out,
                            out_masked, out_f32, out_f32_ps, upsample, class_bias, stat_sums=stat_sums,
                            stat_mode=stat_mode)
        assert x.c == layer.cin_pad, (layer.wname, x.c, layer.cin_pad)
        oh, ow = -(-x.h // layer.stride), -(-x.w // layer.stride)
        prog.add("conv2d_fwd", x.ref(), ptr(layer.fwd[0]), ptr(layer.fwd[1]), layer.k, layer.k, layer.stride, layer.cout, ep,
                 flops=2.0 * x.n * oh * ow * getattr(layer, "flops_cout", layer.cout) * layer.k * layer.k *
                 getattr(layer, "flops_cin", layer.cin),
                 tag="%s %dx%dx%dx%d->%d k%ds%d" % (layer.wname, x.n, x.h, x.w, getattr(layer, "flops_cin", layer.cin),
                                                    layer.cout, layer.k, layer.stride), side=side)

    def conv_dgrad(self, prog, layer, dy, in_h, in_w, out=None, out_masked=None, mask_in=None, mask_neg=0.0, addend=None,
                   out_f32=None, out_f32_ps=0, db_of=None):
        """db_of: the ConvLayer whose output-gradient `out_masked` is; its bias gradient (column sums of out_masked)
        is then accumulated by this kernel's epilogue and the later conv_wgrad(db_of, ., out_masked) skips bias_grad."""
        colsum = None
        if db_of is not None and out_masked is not None and self.fuse_bias_grad:
            colsum = db_of.db
            self._db_done.add((id(prog), id(out_masked), db_of.wname))
        ep = self._epilogue(prog, None, ACT_NONE, 0.0, addend, mask_in, mask_neg, None, out, out_masked, out_f32,
                            out_f32_ps, 1, colsum=colsum)
        assert dy.c == layer.cout_pad, (layer.wname, dy.c, layer.cout_pad)
        prog.add("conv2d_bwd_data", dy.ref(), ptr(layer.bwd[0]), ptr(layer.bwd[1]), layer.k, layer.k, layer.stride,
                 in_h, in_w, layer.cin, ep, flops=2.0 * dy.n * dy.h * dy.w * layer.cout * layer.k * layer.k * layer.cin,
                 tag="%s dy %dx%dx%dx%d->%d k%ds%d" % (layer.wname, dy.n, dy.h, dy.w, layer.cout, layer.cin, layer.k, layer.stride))

    def ew_combine_db(self, prog, out, db_of, *args):
        """ew_combine whose output is the output-gradient of conv `db_of`: the bias gradient (column sums) is taken in
        the same pass, and the later conv_wgrad(db_of, ., out) skips its dpig_bias_grad read of the tensor."""
        if db_of is not None and self.fuse_bias_grad and out.c == db_of.cout and out.c % 8 == 0 and out.c // 8 <= 256:
            prog.add("ew_combine_colsum", out.ref(), *args, ptr(db_of.db))
            self._db_done.add((id(prog), id(out), db_of.wname))
        else:
            prog.add("ew_combine", out.ref(), *args)

    def conv_wgrad(self, prog, layer, x, dy, bias=True):
        prog.add("conv2d_bwd_filter", x.ref(), dy.ref(), layer.k, layer.k, layer.stride, layer.cin, layer.cout,
                 ptr(layer.dw), flops=2.0 * dy.n * dy.h * dy.w * layer.cout * layer.k * layer.k * layer.cin,
                 tag="%s %dx%dx%dx%d->%d k%ds%d" % (layer.wname, x.n, x.h, x.w, layer.cin, layer.cout, layer.k, layer.stride),
                 side=True)
        if not bias or (id(prog), id(dy), layer.wname) in self._db_done:
            return
        if dy.c != layer.cout:  # channel-padded gradient (e.g. the 3-channel image gradient held in 8)
            dy = dy.slice(0, layer.cout)
            prog.keep.append(dy)
        prog.add("bias_grad", dy.ref(), ptr(layer.db))

    def _linear(self, grp, name):
        if not self.training:
            return grp.view(name + "/weights"), grp.view(name + "/biases"), None, None
        return grp.view(name + "/weights"), grp.view(name + "/biases"), grp.gview(name + "/weights"), grp.gview(name + "/biases")

    # -------------------------------------------------------------------------------- programs
    def _build_programs(self):
        self.p_fwd_gen = Program(self.ctx)     # Encoder + U-Net forward (also the sampling path)
        self._prog_forward_generator(self.p_fwd_gen)
        self.p_fwd_enc = Program(self.ctx)     # appearance encoder only (Stage-II real embeddings)
        self._prog_forward_encoder(self.p_fwd_enc)
        self.p_fwd_enc_only = {}
        if self.cfg.fgbg:                      # ... and with one of the two pyramids only (Stage2Engine.prune)
            for which in ("roi", "bg"):
                self.p_fwd_enc_only[which] = Program(self.ctx)
                self._prog_forward_encoder(self.p_fwd_enc_only[which], branches=(which,))
        self.p_fwd_unet = Program(self.ctx)    # U-Net only, from self.emb / self.pose_rcv (sampling path, tester.py)
        self._prog_unet_forward(self.p_fwd_unet)
        if self.training:
            self.p_bwd_gen = Program(self.ctx)     # g_G -> all Encoder+G parameter gradients
            self._prog_backward_generator(self.p_bwd_gen)
        if self.cfg.d_joint:
            self.p_d_pair_fwd = Program(self.ctx)
            self._prog_disc_forward(self.p_d_pair_fwd, self.d_pair, self.pair8)
            if not self.training:
                return
            self.p_d_pair_bwd_data = Program(self.ctx)   # G step: gradient w.r.t. the image halves (the G half is used)
            self._prog_disc_backward(self.p_d_pair_bwd_data, self.d_pair, self.pair8, params=False, data=True)
            self.p_d_pair_bwd_par = Program(self.ctx)    # D step
            self._prog_disc_backward(self.p_d_pair_bwd_par, self.d_pair, self.pair8, params=True, data=False)
            return
        self.p_d_fake_fwd = Program(self.ctx)
        self._prog_disc_forward(self.p_d_fake_fwd, self.d_fake, self.G8)
        self.p_d_real_fwd = Program(self.ctx)
        self._prog_disc_forward(self.p_d_real_fwd, self.d_real, self.x8)
        if not self.training:
            return
        self.p_d_fake_bwd_data = Program(self.ctx)   # G step: gradient w.r.t. the generated image only
        self._prog_disc_backward(self.p_d_fake_bwd_data, self.d_fake, self.G8, params=False, data=True)
        self.p_d_fake_bwd_par = Program(self.ctx)    # D step
        self._prog_disc_backward(self.p_d_fake_bwd_par, self.d_fake, self.G8, params=True, data=False)
        self.p_d_real_bwd_par = Program(self.ctx)
        self._prog_disc_backward(self.p_d_real_bwd_par, self.d_real, self.x8, params=True, data=False)

        if self.mode == "wgan-gp":
            self.p_gp = Program(self.ctx)
            self._prog_gradient_penalty(self.p_gp)

    def _prog_gradient_penalty(self, p):
        """lambda * mean((||dD(xhat)/dxhat|| - 1)^2) and its gradient w.r.t. the critic parameters
        (reference trainer.py:226-236; 4-D image variant: alpha per sample, norm over H*W*C).
        Second backward pass = parameter gradient of the JVP of D along v = d(lambda*gp)/d(grad)."""
        cfg, B = self.cfg, self.B
        H, W, d = cfg.img_h, cfg.img_w, cfg.d_dim
        dh = self.d_hat
        per = H * W * 3
        L = [self.conv[n] for n in self.n_d]
        p.add("gp_interpolate", ptr(self.x), ptr(self.G), ptr(self.gp_alpha), B, per, ptr(self.xhat))
        p.add("pack_f32", ptr(self.xhat), 3, 3, self.xhat8.ref())
        self._prog_disc_forward(p, dh, self.xhat8)
        self._prog_disc_backward(p, dh, self.xhat8, params=False, data=True)       # dlogits == 1 -> g_x = grad
        p.add("gp_penalty", ptr(dh.g_x), B, per, float(self.lam), ptr(self.slopes), ptr(self.loss_gp), ptr(self.gp_v))
        p.add("pack_f32", ptr(self.gp_v), 3, 3, self.gp_v8.ref())
        # ---- tangent (JVP) forward along v
        p.add("im2col_small", self.gp_v8.ref(), 3, 5, 5, 2, 0, self.gp_v_patch.ref())
        self.conv_fwd(p, self.d1_patch, self.gp_v_patch, act=ACT_NONE, bias=False, out_masked=dh.hd[0], mask_in=dh.m[0],
                      mask_neg=0.2)
        for i in (1, 2, 3):
            hh, ww, c = H >> (i + 1), W >> (i + 1), d << i
            sc = self.dp.view("Discriminator.BN%d.scale" % (i + 1))
            self.conv_fwd(p, L[i], dh.hd[i - 1], act=ACT_NONE, bias=False, out_f32=dh.pd[i], out_f32_ps=c)
            p.add("layernorm_jvp_fwd", ptr(dh.pre[i]), ptr(dh.pd[i]), B, hh, ww, c, ptr(dh.stats[i]), ptr(sc),
                  ptr(dh.m[i]), 0.2, ptr(dh.tsums[i]), dh.hd[i].ref())
        # ---- adjoint of S = sum_n flat(hdot4_n) . W_out
        wo = self.dp.view("Discriminator.Output.W")
        dwo = self.dp.gview("Discriminator.Output.W")
        p.add("unpack_f32", dh.hd[3].ref(), ptr(dh.hd_flat), d * 8)
        p.add("linear_bwd", ptr(dh.hd_flat), ptr(wo), ptr(self.ones_b), ptr(dh.hbar_flat), ptr(dwo), None, B,
              self.d_flat, 1)
        p.add("pack_f32", ptr(dh.hbar_flat), d * 8, d * 8, dh.hdbar[3].ref())
        for i in (3, 2, 1):
            hh, ww, c = H >> (i + 1), W >> (i + 1), d << i
            sc = self.dp.view("Discriminator.BN%d.scale" % (i + 1))
            dsc = self.dp.gview("Discriminator.BN%d.scale" % (i + 1))
            dof = self.dp.gview("Discriminator.BN%d.offset" % (i + 1))
            p.add("layernorm_jvp_bwd", dh.hdbar[i].ref(), ptr(dh.m[i]), 0.2, ptr(dh.pre[i]), ptr(dh.pd[i]),
                  ptr(dh.stats[i]), ptr(sc), ptr(dh.tsums[i]), ptr(dh.asums[i]), ptr(dsc), dh.pdbar[i].ref(),
                  ptr(dh.pbarT[i]))
            # tangent-path conv adjoint
            self.conv_wgrad(p, L[i], dh.hd[i - 1], dh.pdbar[i], bias=False)
            if i > 1:
                self.conv_dgrad(p, L[i], dh.pdbar[i], hh * 2, ww * 2, out=dh.hdbar[i - 1])
            else:
                self.conv_dgrad(p, L[i], dh.pdbar[i], hh * 2, ww * 2, out_masked=dh.pdbar[0], mask_in=dh.m[0], mask_neg=0.2)
            # primal-path adjoint
            if i == 3:
                p.add("pack_f32", ptr(dh.pbarT[i]), c, c, dh.pbar[i].ref())
            else:
                p.add("norm_act_bwd_reduce", dh.hbar[i].ref(), ptr(dh.pre[i]), ptr(dh.stats[i]), ptr(dh.m[i]), 0.2,
                      self.norm_mode, ptr(sc), ptr(dh.red[i]), ptr(dsc), ptr(dof))
                p.add("norm_act_bwd_apply", dh.hbar[i].ref(), ptr(dh.pre[i]), ptr(dh.stats[i]), ptr(dh.m[i]), 0.2,
                      self.norm_mode, ptr(sc), ptr(dh.red[i]), float(hh * ww * c), dh.pbarP[i].ref())
                p.add("ew_combine", dh.pbar[i].ref(), dh.pbarP[i].ref(), None, None, ptr(dh.pbarT[i]), c, None, 0.0, 0)
            self.conv_wgrad(p, L[i], dh.h[i - 1], dh.pbar[i])
            if i > 1:
                self.conv_dgrad(p, L[i], dh.pbar[i], hh * 2, ww * 2, out=dh.hbar[i - 1])
            else:
                self.conv_dgrad(p, L[i], dh.pbar[i], hh * 2, ww * 2, out_masked=dh.pbar[0], mask_in=dh.m[0], mask_neg=0.2)
        self.conv_wgrad(p, self.d1_patch, dh.patch, dh.pbar[0])
        self.conv_wgrad(p, self.d1_patch, self.gp_v_patch, dh.pdbar[0], bias=False)

    def _prog_forward_generator(self, p):
        self._prog_forward_encoder(p)
        # ---- generator (trainer.py:588-590, models.py:518-576)
        self._prog_unet_forward(p)

    def run_encoder(self, stream=None, only=None):
        """Appearance-encoder forward only: fills self.emb [B, 352] for the current batch.  only = "roi" / "bg" (two-branch
        encoder): that pyramid alone -- the other half of self.emb keeps whatever an earlier call left there."""
        prog = self.p_fwd_enc_only[only] if only else self.p_fwd_enc
        prog.run(stream if stream is not None else torch.cuda.current_stream().cuda_stream)

    def run_unet(self, stream=None):
        """U-Net forward from the embedding in self.emb and the keypoints in self.pose_rcv; fills self.G / self.G8."""
        self.p_fwd_unet.run(stream if stream is not None else torch.cuda.current_stream().cuda_stream)

    def score_generated(self, stream=None):
        """DCGANDiscriminator logits of the current self.G (tester.py:568-571)."""
        if self.cfg.d_joint:
            raise _lib.DpigError("score_generated: the joint-D (256x256) graph scores concat([x, G]) only")
        self.p_d_fake_fwd.run(stream if stream is not None else torch.cuda.current_stream().cuda_stream)
        return self.d_fake.logits

    def _prog_forward_encoder(self, p, branches=("roi", "bg")):
        """branches: which of the two pyramids run (both by default).  Stage-II trains one factor per optimiser call
        (trainer.py:822-845); the other factor's pyramid feeds nothing in that call, see Stage2Engine.prune."""
        cfg, B = self.cfg, self.B
        H, W, hn, rn, P = cfg.img_h, cfg.img_w, cfg.hidden, cfg.enc_repeat, cfg.n_parts
        e0, e1, e2 = self.conv[self.n_e0], self.conv[self.n_e1], self.conv[self.n_e2]
        # ---- encoder (models.py:396-471; without the mask / background branch: models.py:334-384)
        p.add("pack_f32", ptr(self.x), 3, 3, self.x8.ref())
        p.add("im2col_small", self.x8.ref(), 3, 3, 3, 1, 0, self.x_patch.ref())
        self.conv_fwd(p, self.e0_patch, self.x_patch, out=self.e0, mask_out=self.me0)
        self.conv_fwd(p, e1, self.e0, out=self.e1, mask_out=self.me1)
        self.conv_fwd(p, e2, self.e1, out=self.xs, addend=self.e0, mask_out=self.me2)
        if cfg.fgbg and "bg" in branches:
            # the background branch (models.py:454-464) depends on xs only: it runs on the side stream, concurrently with
            # the ROI branch -- the upper pyramid levels of either branch (8x4 / 3x3 maps) leave half of the SMs idle
            side = "roi" in branches
            p.add("mask_split", self.xs.ref(), ptr(self.fg_mask), None, self.x_bg.ref())
            self.bg_pyr.forward(self, p, side=side)
            p.add("unpack_f32", self.bg_pyr.y[rn - 1].ref(), ptr(self.bg_flat_f32), hn * rn, side=side)
            w, b, _, _ = self._linear(self.gp, self.n_bg_fc)
            p.add("linear_fwd", ptr(self.bg_flat_f32), ptr(w), ptr(b), ptr(self.bg_fea), B, self.bg_flat, self.bg_z,
                  ACT_NONE, 0.0, side=side)
        if "roi" in branches:
            p.add("crop_and_resize_fwd", self.xs.ref(), ptr(self.fg_mask) if cfg.fgbg else None, ptr(self.boxes),
                  ptr(self.box_ind), P * B, self.rois.ref())
            self.roi_pyr.forward(self, p)
            p.add("unpack_f32", self.roi_pyr.y[rn - 1].ref(), ptr(self.roi_flat_f32), hn * rn)
            w, b, _, _ = self._linear(self.gp, self.n_roi_fc)
            p.add("linear_fwd", ptr(self.roi_flat_f32), ptr(w), ptr(b), ptr(self.fea), P * B, self.roi_flat, cfg.part_z,
                  ACT_NONE, 0.0)
        p.add_join()
        p.add("embedding_assemble", ptr(self.fea), ptr(self.bg_fea), ptr(self.vis), B, P, cfg.part_z, self.bg_z,
              ptr(self.emb), 0)

    def _prog_unet_forward(self, p):
        cfg, B = self.cfg, self.B
        H, W, hn, rn = cfg.img_h, cfg.img_w, cfg.hidden, cfg.unet_repeat
        # stem (models.py:528) on concat(tiled embedding, pose): the embedding channels are constant over space
        # (trainer.py:588-590), so their 3x3 contribution is a per-image, per-border-class bias
        stem = self.conv[self.n_gstem]
        wt = stem.w.view(9, self.gin_c, hn)
        for tap in range(9):
            p.add("linear_fwd", ptr(self.emb), ptr(wt[tap]), None, ptr(self.stem_e[tap]), B, cfg.emb_dim, hn,
                  ACT_NONE, 0.0)
        p.add("stem_class_bias", ptr(self.stem_e), B, hn, H, W, ptr(self.stem_cb))
        pose_slice = self.gin.slice(0, cfg.keypoints)
        self._keep.append(pose_slice)
        p.add("pose_rasterize", ptr(self.pose_rcv), B, cfg.keypoints, H, W, 4, pose_slice.ref(), None)
        p.add("pose_patch", ptr(self.pose_rcv), B, cfg.keypoints, H, W, 4, 3, 3, self.pose_patch.ref())
        self.conv_fwd(p, self.stem_patch, self.pose_patch, out=self.g0, mask_out=self.mg0, class_bias=self.stem_cb)
        self.genc.forward(self, p)
        top = self.genc.y[rn - 1]
        p.add("unpack_f32", top.ref(), ptr(self.gtop_f32), hn * rn)
        w, b, _, _ = self._linear(self.gp, self.n_gfc1)
        p.add("linear_fwd", ptr(self.gtop_f32), ptr(w), ptr(b), ptr(self.z), B, self.gtop_flat, cfg.z_num, ACT_NONE, 0.0)
        w, b, _, _ = self._linear(self.gp, self.n_gfc2)
        p.add("linear_fwd", ptr(self.z), ptr(w), ptr(b), ptr(self.dec_in_f32), B, cfg.z_num, self.fh * self.fw * hn,
              ACT_NONE, 0.0)
        x_slice = self.cat[0].slice(0, hn)
        self._keep.append(x_slice)
        p.add("pack_f32", ptr(self.dec_in_f32), hn, hn, x_slice.ref())
        for idx in range(rn):
            names = self.n_gdec[idx]
            self.conv_fwd(p, self.conv[names[0]], self.cat[idx], out=self.dec_a[idx], mask_out=self.dec_ma[idx])
            self.conv_fwd(p, self.conv[names[1]], self.dec_a[idx], out=self.dec_y[idx], addend=self.cat[idx],
                          mask_out=self.dec_mb[idx])
            if idx < rn - 1:
                up = self.cat[idx + 1].slice(0, self.dec_c[idx + 1][0])
                self._keep.append(up)
                self.conv_fwd(p, self.conv[names[2]], self.dec_y[idx], out=up, mask_out=self.dec_mu[idx], upsample=2)
        # 256 -> 3 output conv (models.py:573): 1x1 conv to 27 tap channels, then the col2im gather adds the bias
        self.conv_fwd(p, self.gout_f, self.dec_y[rn - 1], act=ACT_NONE, bias=False, out_f32=self.gout_y, out_f32_ps=32)
        p.add("col2im_small", ptr(self.gout_y), 32, B, H, W, 3, 3, 3, ptr(self.conv[self.n_gout].b), ptr(self.G), 3,
              self.G8.ref())

    def _prog_backward_generator(self, p):
        """Consumes self.g_G (fp32 grad wrt the generated image) and accumulates all Encoder+G param grads."""
        cfg, B = self.cfg, self.B
        H, W, hn, rn, ern, P = cfg.img_h, cfg.img_w, cfg.hidden, cfg.unet_repeat, cfg.enc_repeat, cfg.n_parts
        p.add("pack_f32", ptr(self.g_G), 3, 3, self.g_G8.ref())
        # ---- decoder
        lo = self.conv[self.n_gout]
        p.add("im2col_small", self.g_G8.ref(), 3, 3, 3, 1, 1, self.gG_patch.ref())
        fl = 2.0 * B * H * W * 27 * lo.cin
        tag = "%s %dx%dx%dx%d->3 k3s1 (patch form)" % (lo.wname, B, H, W, lo.cin)
        # filter gradient in the [ci][tap*3+co] layout, folded back into HWIO; bias gradient from dL/dG itself
        p.add_py(lambda s: self.gout_dwv.zero_())
        p.add("conv2d_bwd_filter", self.dec_y[rn - 1].ref(), self.gG_patch.ref(), 1, 1, 1, lo.cin, 27, ptr(self.gout_dwv),
              flops=fl, tag=tag)
        p.add("permute_taps", ptr(self.gout_dwv), ptr(lo.dw), 9, lo.cin, 3, 2)
        g3 = self.g_G8.slice(0, 3)
        self._keep.append(g3)
        p.add("bias_grad", g3.ref(), ptr(lo.db))
        # data gradient: 1x1 conv from the 27 patch channels of dL/dG
        db_of = self.conv[self.n_gdec[rn - 1][1]]
        colsum = db_of.db if self.fuse_bias_grad else None
        if colsum is not None:
            self._db_done.add((id(p), id(self.dec_gb[rn - 1]), db_of.wname))
        ep = self._epilogue(p, None, ACT_NONE, 0.0, None, self.dec_mb[rn - 1], 0.0, None, self.dec_gy[rn - 1],
                            self.dec_gb[rn - 1], None, 0, 1, colsum=colsum)
        p.add("conv2d_fwd", self.gG_patch.ref(), ptr(self.gout_d.fwd[0]), ptr(self.gout_d.fwd[1]), 1, 1, 1, lo.cin, ep,
              flops=fl, tag="dgrad " + tag)
        for idx in range(rn - 1, -1, -1):
            names = self.n_gdec[idx]
            lvl = rn - 1 - idx
            hh, ww = H >> lvl, W >> lvl
            l1, l2 = self.conv[names[0]], self.conv[names[1]]
            if idx < rn - 1:
                # gradient of the upsampled 1x1-conv output = x-part of the next level's concat gradient
                lu = self.conv[names[2]]
                gup = self.g_cat[idx + 1].slice(0, self.dec_c[idx + 1][0])
                self._keep.append(gup)
                self.ew_combine_db(p, self.dec_gu[idx], lu, gup.ref(), None, None, None, 0, ptr(self.dec_mu[idx]), 0.0, 1)
                self.conv_wgrad(p, lu, self.dec_y[idx], self.dec_gu[idx])
                self.conv_dgrad(p, lu, self.dec_gu[idx], hh, ww, out=self.dec_gy[idx], out_masked=self.dec_gb[idx],
                                mask_in=self.dec_mb[idx], db_of=l2)
            self.conv_wgrad(p, l2, self.dec_a[idx], self.dec_gb[idx])
            self.conv_dgrad(p, l2, self.dec_gb[idx], hh, ww, out_masked=self.dec_ga[idx], mask_in=self.dec_ma[idx], db_of=l1)
            self.conv_wgrad(p, l1, self.cat[idx], self.dec_ga[idx])
            self.conv_dgrad(p, l1, self.dec_ga[idx], hh, ww, out=self.g_cat[idx], addend=self.dec_gy[idx])
        # ---- bottleneck FCs (models.py:543-555)
        gx0 = self.g_cat[0].slice(0, hn)
        self._keep.append(gx0)
        p.add("unpack_f32", gx0.ref(), ptr(self.g_dec_in_f32), hn)
        w, b, dw, db = self._linear(self.gp, self.n_gfc2)
        p.add("linear_bwd", ptr(self.z), ptr(w), ptr(self.g_dec_in_f32), ptr(self.g_z), ptr(dw), ptr(db), B, cfg.z_num,
              self.fh * self.fw * hn)
        w, b, dw, db = self._linear(self.gp, self.n_gfc1)
        p.add("linear_bwd", ptr(self.gtop_f32), ptr(w), ptr(self.g_z), ptr(self.g_gtop_f32), ptr(dw), ptr(db), B,
              self.gtop_flat, cfg.z_num)
        # ---- U-Net encoder: top gradient = FC path + skip path
        skip = []
        for lvl in range(rn):
            idx = rn - 1 - lvl
            xc, c = self.dec_c[idx]
            s = self.g_cat[idx].slice(xc, c - xc)
            self._keep.append(s)
            skip.append(s)
        p.add("ew_combine", self.genc.g_y[rn - 1].ref(), skip[rn - 1].ref(), None, None, ptr(self.g_gtop_f32), hn * rn,
              None, 0.0, 0)
        self.genc.backward(self, p, skip_grads=skip)
        ls = self.conv[self.n_gstem]
        gi = self.genc.g_in
        e = cfg.emb_dim
        wt = ls.w.view(9, self.gin_c, hn)
        dwt = ls.dw.view(9, self.gin_c, hn)
        # pose rows of the filter gradient on the tensor cores; embedding rows and d(emb) from per-tap sums of g
        p.add_py(lambda s: self.stem_dwp.zero_())
        p.add("conv2d_bwd_filter", self.pose_patch.ref(), gi.ref(), 1, 1, 1, 9 * cfg.keypoints, hn, ptr(self.stem_dwp),
              flops=2.0 * B * H * W * hn * 9 * cfg.keypoints, tag="%s pose rows (patch form)" % ls.wname)
        p.add_py(lambda s: dwt[:, e:, :].add_(self.stem_dwp.view(9, cfg.keypoints, hn)))
        if (id(p), id(gi), ls.wname) not in self._db_done:
            p.add("bias_grad", gi.ref(), ptr(ls.db))
        p.add("stem_tap_sums", gi.ref(), ptr(self.stem_cls), ptr(self.stem_ts))
        for tap in range(9):
            p.add("linear_bwd", ptr(self.emb), ptr(wt[tap]), ptr(self.stem_ts[tap]), ptr(self.stem_tmp), ptr(dwt[tap]),
                  None, B, e, hn)
            p.add("add_f32", ptr(self.g_emb), ptr(self.stem_tmp), ptr(self.g_emb) if tap else None, B * e, 1.0, 1.0)
        # ---- appearance encoder (every ID_AE gradient is final here: start its all-reduce when data-parallel)
        if self.dist is not None:
            p.add_py(lambda s: self._early_allreduce("idae"))
        p.add("embedding_assemble", ptr(self.g_fea), ptr(self.g_bg_fea), ptr(self.vis), B, P, cfg.part_z, self.bg_z,
              ptr(self.g_emb), 1)
        w, b, dw, db = self._linear(self.gp, self.n_roi_fc)
        p.add("linear_bwd", ptr(self.roi_flat_f32), ptr(w), ptr(self.g_fea), ptr(self.g_roi_flat), ptr(dw), ptr(db), P * B,
              self.roi_flat, cfg.part_z)
        p.add("pack_f32", ptr(self.g_roi_flat), hn * ern, hn * ern, self.roi_pyr.g_y[ern - 1].ref())
        self.roi_pyr.backward(self, p)
        if self.dist is not None:     # the ROI pyramid's gradients are final: their all-reduce runs under the Bg pyramid's backward
            p.add_py(lambda s: self._early_allreduce("roi"))
        if int(os.environ.get("DPIG_CROP_GATHER", "1")) == 0:      # the atomic scatter form accumulates; the gather form overwrites
            p.add_py(lambda s: self.g_crop.zero_())
        p.add("crop_and_resize_bwd", self.roi_pyr.g_in.ref(), ptr(self.fg_mask) if cfg.fgbg else None, ptr(self.boxes),
              ptr(self.box_ind), P * B, ptr(self.g_crop), B, H, W, hn)
        if cfg.fgbg:
            w, b, dw, db = self._linear(self.gp, self.n_bg_fc)
            p.add("linear_bwd", ptr(self.bg_flat_f32), ptr(w), ptr(self.g_bg_fea), ptr(self.g_bg_flat), ptr(dw), ptr(db), B,
                  self.bg_flat, self.bg_z)
            p.add("pack_f32", ptr(self.g_bg_flat), hn * ern, hn * ern, self.bg_pyr.g_y[ern - 1].ref())
            self.bg_pyr.backward(self, p)
            if self.dist is not None:     # ... and the Bg pyramid's under the stem's backward
                p.add_py(lambda s: self._early_allreduce("bg"))
            # g_xs = g_crop (already * m) + g_xbg * (1 - m)      (models.py:402-403)
            p.add("mask_split", self.bg_pyr.g_in.ref(), ptr(self.fg_mask), None, self.g_xbg_s.ref())
            p.add("ew_combine", self.g_xs.ref(), self.g_xbg_s.ref(), None, None, ptr(self.g_crop), hn, None, 0.0, 0)
        else:
            p.add("pack_f32", ptr(self.g_crop), hn, hn, self.g_xs.ref())
        e0, e1, e2 = self.conv[self.n_e0], self.conv[self.n_e1], self.conv[self.n_e2]
        self.ew_combine_db(p, self.g_xs_m, e2, self.g_xs.ref(), None, None, None, 0, ptr(self.me2), 0.0, 0)
        self.conv_wgrad(p, e2, self.e1, self.g_xs_m)
        self.conv_dgrad(p, e2, self.g_xs_m, H, W, out_masked=self.g_e1, mask_in=self.me1, db_of=e1)
        self.conv_wgrad(p, e1, self.e0, self.g_e1)
        self.conv_dgrad(p, e1, self.g_e1, H, W, out_masked=self.g_e0, mask_in=self.me0, addend=self.g_xs, db_of=e0)
        self.conv_wgrad(p, self.e0_patch, self.x_patch, self.g_e0)

    def _prog_disc_forward(self, p, dp, img):
        cfg = self.cfg
        H, W, d, n = cfg.img_h, cfg.img_w, cfg.d_dim, dp.n
        p.add("im2col_small", img.ref(), 3, 5, 5, 2, 0, dp.patch.ref())
        self.conv_fwd(p, self.d1_patch, dp.patch, out=dp.h[0], act=ACT_LRELU, alpha=0.2, mask_out=dp.m[0])
        for i in (1, 2, 3):
            layer = self.conv[self.n_d[i]]
            hh, ww, c = H >> (i + 1), W >> (i + 1), d << i
            # conv + bias with the raw normalisation sums (sum x, sum x^2 per channel / per sample) emitted by the conv
            # epilogue, then ONE normalise + LeakyReLU pass: 2 launches per block (wgan_gp.py:417-431)
            self.conv_fwd(p, layer, dp.h[i - 1], act=ACT_NONE, out_f32=dp.pre[i], out_f32_ps=c,
                          stat_sums=dp.sums[i], stat_mode=self.norm_mode)
            count = float(hh * ww * c) if self.norm_mode == NORM_LAYER else float(n * hh * ww * self.world)
            if self.norm_mode == NORM_BATCH and self.dist is not None:
                p.add_py(lambda s, t=dp.sums[i]: self.dist.all_reduce_sum(t))
            sc = self.dp.view("Discriminator.BN%d.scale" % (i + 1))
            of = self.dp.view("Discriminator.BN%d.offset" % (i + 1))
            p.add("norm_act_fwd", ptr(dp.pre[i]), n, hh, ww, c, self.norm_mode, 1e-5, ptr(dp.sums[i]), count, ptr(sc),
                  ptr(of), ACT_LRELU, 0.2, ptr(dp.stats[i]), dp.h[i].ref(), ptr(dp.m[i]))
        for view, rows in self._disc_rows(dp, dp.h[3], dp.flat):
            p.add("unpack_f32", view.ref(), ptr(rows), view.c)
        w = self.dp.view("Discriminator.Output.W")
        b = self.dp.view("Discriminator.Output.b")
        p.add("linear_fwd", ptr(dp.flat), ptr(w), ptr(b), ptr(dp.logits), n * cfg.d_rows, self.d_flat, 1, ACT_NONE, 0.0)

    def _disc_rows(self, dp, t, flat):
        """(view of the last feature map, rows of the Linear input it fills): the whole map when every image is one
        row; else, per image group g and row r, the channel block [r*C/R, (r+1)*C/R) of that group's images
        (NCHW flatten + reshape [-1, 16384], wgan_gp.py:433) -> rows (g*R + r)*n_g .. +n_g."""
        R = self.cfg.d_rows
        if R == 1:
            return [(t, flat)]
        out = []
        ng, cpr = dp.n // dp.segs, t.c // R
        for g in range(dp.segs):
            for r in range(R):
                v = t.batch_slice(g * ng, ng).slice(r * cpr, cpr)
                self._keep.append(v)
                out.append((v, flat[(g * R + r) * ng:(g * R + r + 1) * ng]))
        return out

    def _prog_disc_backward(self, p, dp, img, params, data):
        """dp.dlogits -> parameter grads (params=True) and/or the gradient w.r.t. the input image (data=True)."""
        cfg = self.cfg
        H, W, d, n = cfg.img_h, cfg.img_w, cfg.d_dim, dp.n
        w = self.dp.view("Discriminator.Output.W")
        dw = self.dp.gview("Discriminator.Output.W")
        db = self.dp.gview("Discriminator.Output.b")
        p.add("linear_bwd", ptr(dp.flat), ptr(w), ptr(dp.dlogits), ptr(dp.g_flat), ptr(dw) if params else None,
              ptr(db) if params else None, n * cfg.d_rows, self.d_flat, 1)
        for view, rows in self._disc_rows(dp, dp.g_h[3], dp.g_flat):
            p.add("pack_f32", ptr(rows), view.c, view.c, view.ref())
        for i in (3, 2, 1):
            layer = self.conv[self.n_d[i]]
            hh, ww, c = H >> (i + 1), W >> (i + 1), d << i
            sc = self.dp.view("Discriminator.BN%d.scale" % (i + 1))
            dsc = self.dp.gview("Discriminator.BN%d.scale" % (i + 1))
            dof = self.dp.gview("Discriminator.BN%d.offset" % (i + 1))
            count = float(hh * ww * c) if self.norm_mode == NORM_LAYER else float(n * hh * ww * self.world)
            p.add("norm_act_bwd_reduce", dp.g_h[i].ref(), ptr(dp.pre[i]), ptr(dp.stats[i]), ptr(dp.m[i]), 0.2,
                  self.norm_mode, ptr(sc), ptr(dp.red[i]), ptr(dsc) if params else None, ptr(dof) if params else None)
            if self.norm_mode == NORM_BATCH and self.dist is not None:
                p.add_py(lambda s, t=dp.red[i]: self.dist.all_reduce_sum(t))
            p.add("norm_act_bwd_apply", dp.g_h[i].ref(), ptr(dp.pre[i]), ptr(dp.stats[i]), ptr(dp.m[i]), 0.2,
                  self.norm_mode, ptr(sc), ptr(dp.red[i]), count, dp.g_pre[i].ref())
            if params:
                self.conv_wgrad(p, layer, dp.h[i - 1], dp.g_pre[i])
            if i > 1:
                self.conv_dgrad(p, layer, dp.g_pre[i], hh * 2, ww * 2, out=dp.g_h[i - 1])
            else:
                # layer-1 output is a plain LeakyReLU: emit the masked gradient wrt its conv output
                self.conv_dgrad(p, layer, dp.g_pre[i], hh * 2, ww * 2, out_masked=dp.g_pre[0], mask_in=dp.m[0],
                                mask_neg=0.2, db_of=self.conv[self.n_d[0]] if params else None)
        l1 = self.conv[self.n_d[0]]
        if params:
            self.conv_wgrad(p, self.d1_patch, dp.patch, dp.g_pre[0])
        if data:
            self.conv_dgrad(p, l1, dp.g_pre[0], H, W, out_f32=dp.g_x, out_f32_ps=3)

    # -------------------------------------------------------------------------------- stepping
    def set_batch(self, batch, non_blocking=True):
        """Host -> device copy of one batch (dict of numpy arrays / pinned torch tensors as produced by
        synth.make_batch / the input pipeline): x, pose_rcv, mask, part_bbox, part_vis."""
        cfg, B, P = self.cfg, self.B, self.cfg.n_parts

        def dev(a, dtype=torch.float32):
            t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
            return t.to(self.device, dtype, non_blocking=non_blocking)

        self.x.copy_(dev(batch["x"]))
        self.pose_rcv.copy_(dev(batch["pose_rcv"]))
        if cfg.fgbg:
            self.fg_mask.copy_(dev(batch["mask"]).reshape(B, cfg.img_h, cfg.img_w))
        bb = dev(batch["part_bbox"])[:, :P, :]                      # [B,P,4] pixels (y1,x1,y2,x2)
        scale = torch.tensor([cfg.img_h, cfg.img_w, cfg.img_h, cfg.img_w], dtype=torch.float32, device=self.device)
        self.boxes.copy_((bb / scale).permute(1, 0, 2).reshape(P * B, 4))   # ROI i of image b -> row i*B+b
        if cfg.use_vis:
            self.vis.copy_(dev(batch["part_vis"])[:, :P])
        else:
            self.vis.fill_(1.0)

    def forward(self, with_disc=True):
        """Encoder + U-Net (+ D on x and G) forward; returns nothing (results stay in HBM)."""
        s = torch.cuda.current_stream().cuda_stream
        self.p_fwd_gen.run(s)
        if with_disc:
            self._disc_fwd(s, None, real=True)
            self.ctx.loss_gan(self.gan_mode, ptr(self.d_real.logits), ptr(self.d_fake.logits), self.n_logits,
                              ptr(self.loss_gan), None, None, None, s)
            self.ctx.loss_l1(ptr(self.G), ptr(self.x), self.G.numel(), 20.0, ptr(self.loss_l1), None, s)

    @property
    def n_logits(self):
        """logits per D application on B images (8 per image on the 256x256 graph)."""
        return self.B * self.cfg.d_rows

    def _disc_fwd(self, s, timings, real):
        """D on G (and on x when `real`); the joint graph always sees concat([x, G]) (trainer_256.py:61-63)."""
        if self.cfg.d_joint:
            self.p_d_pair_fwd.run(s, timings)
            return
        if real:
            self.p_d_real_fwd.run(s, timings)
        self.p_d_fake_fwd.run(s, timings)

    def _step_size(self, which):
        """The scalar the update kernel multiplies with: TF Adam's lr_t = lr*sqrt(1-b2^t)/(1-b1^t) (float64 on the
        host, like dpig_adam_step) or the plain RMSProp lr."""
        lr = float(np.float32(self.g_lr if which == "g" else self.d_lr))
        if self.mode in ("wgan", "lsgan"):
            return lr
        b2 = float(np.float32(0.9 if self.mode == "wgan-gp" else 0.999))   # the betas reach the C entry as fp32
        t = self.t[which]
        return lr * math.sqrt(1.0 - b2 ** t) / (1.0 - 0.5 ** t)

    def _slice_range(self, which):
        """[lo, hi) of a parameter slice of the generator arena whose gradients become final together, if it is one
        contiguous block, else None: 'idae' = every ID_AE/* parameter (final before the appearance encoder's backward),
        'roi' / 'bg' = the ROI / background pyramid convolutions and their FC (final after that pyramid's backward)."""
        if which == "idae":
            member = lambda name: name.startswith("ID_AE/")  # noqa: E731
        else:
            convs = self.n_roi if which == "roi" else getattr(self, "n_bg", [])
            fc = self.n_roi_fc if which == "roi" else getattr(self, "n_bg_fc", None)
            names = set()
            for n in list(convs) + ([fc] if fc else []):
                names.update((n + "/weights", n + "/biases"))
            member = lambda name: name in names  # noqa: E731
        inside = [(o, (n + 63) // 64 * 64) for name, (o, n, _) in self.gp.specs.items() if member(name)]
        if not inside:
            return None
        lo, hi = min(o for o, _ in inside), max(o + n for o, n in inside)
        for name, (o, n, _) in self.gp.specs.items():
            if not member(name) and lo <= o < hi:
                return None
        return lo, hi

    def _idae_range(self):
        return self._slice_range("idae")

    def _early_allreduce(self, which="idae"):
        """Called from the backward program once every gradient of a slice is final: all-reduce that slice on the
        communication stream while the main stream keeps computing (ID_AE: 285 of the 474 MB under the appearance
        encoder's backward; ROI pyramid under the Bg pyramid's backward; Bg pyramid under the stem's)."""
        if not self.overlap_comm:
            return
        rng = self._slice_range(which)
        if rng is None:
            return
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.device)
        ready = torch.cuda.Event()
        ready.record()
        self._comm_stream.wait_event(ready)
        with torch.cuda.stream(self._comm_stream):
            self.dist.all_reduce_sum(self.gp.grad[rng[0]:rng[1]])
            self._comm_done = torch.cuda.Event()
            self._comm_done.record()
        self._early_ranges.append(rng)

    def _optim(self, which, s):
        """All-reduce (N > 1), update, re-pack of the bf16 operand copies.  Reads the step size from self.lr_dev."""
        grp = self.gp if which == "g" else self.dp
        if self.dist is not None:
            if which == "g" and self._early_ranges:
                done, self._early_ranges = sorted(self._early_ranges), []
                pos = 0
                for lo, hi in done + [(grp.total, grp.total)]:     # what the early all-reduces left: the gaps
                    if lo > pos:
                        self.dist.all_reduce_sum(grp.grad[pos:lo])
                    pos = max(pos, hi)
                torch.cuda.current_stream().wait_event(self._comm_done)
            else:
                self.dist.all_reduce_sum(grp.grad)
        gs = 1.0 / self.world
        if self.mode in ("wgan", "lsgan"):
            clip = 0.01 if (self.mode == "wgan" and which == "d") else 0.0
            self.ctx.rmsprop_step_dev(ptr(grp.value), ptr(grp.grad), ptr(grp.v), grp.total, ptr(self.lr_dev[which]), 0.9,
                                      1e-10, gs, clip, s)
        else:
            b2 = 0.9 if self.mode == "wgan-gp" else 0.999
            self.ctx.adam_step_dev(ptr(grp.value), ptr(grp.grad), ptr(grp.m), ptr(grp.v), grp.total,
                                   ptr(self.lr_dev[which]), 0.5, b2, 1e-8, gs, s)
        self.pack_weights(which, s)

    def _step(self, which, timings):
        if not self.training:
            raise _lib.DpigError("this engine was built forward-only (inference=True): no optimiser steps")
        return self._step_impl(which, timings)

    def _step_impl(self, which, timings):
        """One optimiser call.  The first two calls of each kind run eagerly (they also set the kernels' launch
        attributes and warm NCCL); the third is captured into a CUDA graph while it runs, later calls replay it.
        Per-launch timing (timings is not None) always runs eagerly."""
        grads = self.g_grads if which == "g" else self.d_grads
        self.t[which] += 1
        self.lr_dev[which].fill_(self._step_size(which))      # a fill kernel on the current stream: ordered before the step
        if not self.use_graphs or timings is not None:
            grads(timings)
            self._optim(which, torch.cuda.current_stream().cuda_stream)
            return
        g = self._graphs.get(which)
        if g is None:
            if self._eager_steps[which] < 2:
                self._eager_steps[which] += 1
                grads(None)
                self._optim(which, torch.cuda.current_stream().cuda_stream)
                return
            graph = torch.cuda.CUDAGraph()
            n0 = self.ctx.launch_count()
            # (thread_local: the NCCL watchdog thread polls events while this thread captures; under the default
            #  "global" mode such a call from another thread can invalidate the capture)
            with torch.cuda.graph(graph, capture_error_mode="thread_local" if self.dist is not None else "global"):
                grads(None)
                self._optim(which, torch.cuda.current_stream().cuda_stream)
            g = self._graphs[which] = (graph, self.ctx.launch_count() - n0)
        g[0].replay()
        self.ctx.replayed_launches += g[1]

    def drop_graphs(self):
        """Forget the captured step graphs (after anything that changes what a step launches: fast mode, pair mode,
        a pinned GP alpha ...); the next steps re-capture."""
        self._graphs = {}

    def g_grads(self, timings=None):
        """Forward + backward of g_loss = gan(D(G)) + 20*L1 w.r.t. Encoder+G (trainer.py:605-607, 622-624)."""
        s = torch.cuda.current_stream().cuda_stream
        self.gp.grad.zero_()
        self.p_fwd_gen.run(s, timings)
        self._disc_fwd(s, timings, real=False)
        self.ctx.loss_gan(self.gan_mode, None, ptr(self.d_fake.logits), self.n_logits, ptr(self.loss_gan),
                          ptr(self.d_fake.dlogits), None, None, s)
        # batch-mean losses: per-rank grads are means over the local shard; summed then scaled by 1/world in _optim
        if self.cfg.d_joint:
            self.d_real.dlogits.zero_()   # g_loss does not depend on D(x), but x shares the batch statistics with G
            self.p_d_pair_bwd_data.run(s, timings)
        else:
            self.p_d_fake_bwd_data.run(s, timings)
        self.g_G.copy_(self.d_fake.g_x)
        self.ctx.loss_l1(ptr(self.G), ptr(self.x), self.G.numel(), 20.0, ptr(self.loss_l1), ptr(self.g_G), s)
        self.p_bwd_gen.run(s, timings)

    def activation_bits(self):
        """Every ReLU / LeakyReLU decision of the last forward pass, keyed and ordered as oracle/nets.py takes them
        (`branches` of stage1_forward): the backward programs gate gradients with exactly these bits.  Diagnostics --
        a device->host sized read; tests use it to put the float64 oracle on the engine's linear piece."""
        cfg = self.cfg
        B, H, W, hn, rn = self.B, cfg.img_h, cfg.img_w, cfg.hidden, cfg.unet_repeat
        enc = [unpack_bits(m, B, H, W, hn) for m in (self.me0, self.me1, self.me2)] + self.roi_pyr.sign_bits()
        if cfg.fgbg:
            enc += self.bg_pyr.sign_bits()
        gen = [unpack_bits(self.mg0, B, H, W, hn)] + self.genc.sign_bits()
        for idx in range(rn):
            c = self.dec_c[idx][1]
            lvl = rn - 1 - idx
            hh, ww = H >> lvl, W >> lvl
            gen += [unpack_bits(self.dec_ma[idx], B, hh, ww, c), unpack_bits(self.dec_mb[idx], B, hh, ww, c)]
            if idx < rn - 1:      # the 1x1 conv runs before the x2 upsample here (after it in models.py:569-570)
                gen.append(unpack_bits(self.dec_mu[idx], B, hh, ww, self.dec_c[idx + 1][0]))
        out = {"Encoder/G_encoder": enc, "ID_AE/G": gen}
        if cfg.d_joint:
            out["D_pair"] = [self.d_pair.sign_bits(i) for i in range(4)]
        else:
            out["D_real"] = [self.d_real.sign_bits(i) for i in range(4)]
            out["D_fake"] = [self.d_fake.sign_bits(i) for i in range(4)]
        if hasattr(self, "d_hat"):
            out["D_hat"] = [self.d_hat.sign_bits(i) for i in range(4)]
        return {k: [t.cpu() for t in v] for k, v in out.items()}

    def d_grads(self, timings=None):
        """Forward + backward of d_loss w.r.t. the discriminator (trainer.py:601-605, 625)."""
        s = torch.cuda.current_stream().cuda_stream
        self.dp.grad.zero_()
        self.p_fwd_gen.run(s, timings)
        self._disc_fwd(s, timings, real=True)
        self.ctx.loss_gan(self.gan_mode, ptr(self.d_real.logits), ptr(self.d_fake.logits), self.n_logits,
                          ptr(self.loss_gan), None, ptr(self.d_real.dlogits), ptr(self.d_fake.dlogits), s)
        if self.cfg.d_joint:
            self.p_d_pair_bwd_par.run(s, timings)
        else:
            self.p_d_real_bwd_par.run(s, timings)
            self.p_d_fake_bwd_par.run(s, timings)
        if self.mode == "wgan-gp":
            if not self.gp_alpha_fixed:
                self.gp_alpha.uniform_(0.0, 1.0)     # alpha ~ U[0,1] per sample (trainer.py:226-230)
            self.p_gp.run(s, timings)

    def g_step(self, timings=None):
        self._step("g", timings)

    def d_step(self, timings=None):
        self._step("d", timings)

    def losses(self):
        """(g_gan, d_loss, L1) as Python floats -- a device->host read.  In wgan-gp mode d_loss includes
        lambda * gradient_penalty of the last d_grads()."""
        lg = self.loss_gan.cpu()
        d_loss = float(lg[1])
        if self.mode == "wgan-gp":
            d_loss += self.lam * float(self.loss_gp.cpu()[0])
        return float(lg[0]), d_loss, float(self.loss_l1.cpu()[0])


def init_params(cfg, seed=1234):
    """Random-init parameters with the reference's initialisers, keyed by TF variable name:
    slim.conv2d / fully_connected -> xavier_uniform weights, zero biases (models.py:396 ff.);
    discriminator convs / linear -> U(+-0.02*sqrt(3)) (wgan_gp.py:411-413, tflib/ops/conv2d.py:56-80),
    zero biases, norm scale 1 / offset 0 (tflib/ops/batchnorm.py:23-24)."""
    rng = np.random.default_rng(seed)
    hn, rn, ern = cfg.hidden, cfg.unet_repeat, cfg.enc_repeat
    p = OrderedDict()

    def xavier(shape, fan_in, fan_out):
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        return rng.uniform(-lim, lim, size=shape).astype(np.float32)

    class Scope:
        def __init__(self, prefix):
            self.prefix, self.nc, self.nf = prefix, 0, 0

        def conv(self, k, cin, cout):
            name = "%s/Conv%s" % (self.prefix, "" if self.nc == 0 else "_%d" % self.nc)
            self.nc += 1
            p[name + "/weights"] = xavier((k, k, cin, cout), k * k * cin, k * k * cout)
            p[name + "/biases"] = np.zeros(cout, np.float32)

        def fc(self, cin, cout):
            name = "%s/fully_connected%s" % (self.prefix, "" if self.nf == 0 else "_%d" % self.nf)
            self.nf += 1
            p[name + "/weights"] = xavier((cin, cout), cin, cout)
            p[name + "/biases"] = np.zeros(cout, np.float32)

        def pyramid(self, levels):
            for idx in range(levels):
                c = hn * (idx + 1)
                self.conv(3, c, c)
                self.conv(3, c, c)
                if idx < levels - 1:
                    self.conv(3, c, hn * (idx + 2))

    fh, fw = cfg.img_h >> (rn - 1), cfg.img_w >> (rn - 1)
    s = Scope("Encoder/G_encoder")
    s.conv(3, 3, hn)
    s.conv(3, hn, hn)
    s.conv(3, hn, hn)
    s.pyramid(ern)
    rf = halved(cfg.roi_size, ern - 1)
    s.fc(rf * rf * hn * ern, cfg.part_z)
    if cfg.fgbg:
        s.pyramid(ern)
        s.fc((cfg.img_h >> (ern - 1)) * (cfg.img_w >> (ern - 1)) * hn * ern, cfg.part_z * 4)
    s = Scope("ID_AE/G")
    s.conv(3, cfg.emb_dim + cfg.keypoints, hn)
    s.pyramid(rn)
    s.fc(fh * fw * hn * rn, cfg.z_num)
    s.fc(cfg.z_num, fh * fw * hn)
    x_c = hn
    for idx in range(rn):
        c = x_c + hn * (rn - idx)
        s.conv(3, c, c)
        s.conv(3, c, c)
        if idx < rn - 1:
            x_c = hn * (rn - idx - 1)
            s.conv(1, c, x_c)
        else:
            x_c = c
    s.conv(3, x_c, 3)
    d = cfg.d_dim
    lim = 0.02 * math.sqrt(3.0)
    chans = [3, d, 2 * d, 4 * d, 8 * d]
    for i in range(4):
        p["Discriminator.%d.Filters" % (i + 1)] = rng.uniform(-lim, lim, size=(5, 5, chans[i], chans[i + 1])).astype(np.float32)
        p["Discriminator.%d.Biases" % (i + 1)] = np.zeros(chans[i + 1], np.float32)
        if i >= 1:
            p["Discriminator.BN%d.offset" % (i + 1)] = np.zeros(chans[i + 1], np.float32)
            p["Discriminator.BN%d.scale" % (i + 1)] = np.ones(chans[i + 1], np.float32)
    d_in = cfg.d_row
    p["Discriminator.Output.W"] = rng.uniform(-lim, lim, size=(d_in, 1)).astype(np.float32)
    p["Discriminator.Output.b"] = np.zeros(1, np.float32)
    return p
