"""GPU parity of the reference-surface mirrors (models.py / wgan_gp.py / tflib ops called eagerly, op by op)
against the float64 oracle, plus the remaining small kernels (denorm_u8, InstanceNorm mode)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nets  # noqa: E402
from oracle import tf_ops as T  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["dcgan", "wgan-gp"])
def test_dcgan_discriminator_mirror(mode):
    """WGAN_GP(...).DCGANDiscriminator(NCHW) composed from lib.ops.{conv2d,batchnorm|layernorm,linear} with the
    shared-by-name parameter registry, vs oracle (wgan_gp.py:407-440)."""
    from dpig_b200 import tflib as lib
    from dpig_b200 import wgan_gp
    lib.delete_all_params()
    cfg = nets.NetConfig()
    p = {k: v for k, v in nets.init_params(cfg, seed=5, bias_noise=0.05).items() if k.startswith("Discriminator")}
    for k, v in p.items():
        lib.param(k, v)
    g = torch.Generator().manual_seed(0)
    x = torch.rand((3, 128, 64, 3), generator=g) * 2 - 1
    net = wgan_gp.WGAN_GP(DATA_DIR="", MODE=mode, DIM=64, BATCH_SIZE=3)
    out = net.DCGANDiscriminator(x.permute(0, 3, 1, 2).cuda(), input_dim=3)
    out2 = net.DCGANDiscriminator(x.permute(0, 3, 1, 2).cuda(), input_dim=3)          # same names -> same weights
    ref = nets.dcgan_discriminator(nets.to_torch(p), cfg, x.double(), mode)
    # same names -> same weights; split-K fp32 atomics make the two runs equal only to rounding
    assert out.shape == (3,) and torch.allclose(out, out2, atol=1e-5)
    assert float((out.double().cpu() - ref).abs().max()) < 1e-3
    assert len(lib.params_with_name("Discriminator.")) == len(p) + (6 if mode == "dcgan" else 0)  # + moving stats
    lib.delete_all_params()


def test_models_mirror_encoder_and_generator():
    from dpig_b200 import models, synth
    kw = dict(img_h=32, img_w=16, hidden=64, roi_size=12)
    cfg = nets.NetConfig(**kw)
    b = synth.make_batch(2, 32, 16, seed=3)
    x = torch.tensor(b["x"])
    fea_all, fea_list, conv_fea, var_e = models.GeneratorCNN_ID_Encoder_BodyROIVis_FgBgFeaTwoBranch(
        x, torch.tensor(b["mask"]), torch.tensor(b["part_bbox"][:, :7]), torch.tensor(b["part_vis"][:, :7]), 7, 32,
        cfg.repeat_num, 64, "NHWC", activation_fn=models.relu, keep_part_prob=1.0, roi_size=12)
    assert fea_all.shape == (2, 352) and len(fea_list) == 8 and len(conv_fea) == 8
    pose = T.pose_rasterize(torch.tensor(b["pose_rcv"]), 32, 16)
    emb_rep = fea_all[:, None, None, :].expand(2, 32, 16, 352)
    G, z, var_g = models.GeneratorCNN_ID_UAEAfterResidual(emb_rep, pose, 3, 64, cfg.repeat_num, 64, "NHWC",
                                                          activation_fn=models.relu)
    G2, _, _ = models.GeneratorCNN_ID_UAEAfterResidual(emb_rep, None, 3, 64, cfg.repeat_num, 64, "NHWC",
                                                       activation_fn=models.relu, reuse=True, pose_rcv=b["pose_rcv"])
    assert float((G - G2).abs().max()) < 1e-5          # maps given == maps rasterised on the GPU
    p = {k: v.detach().double().cpu() for k, v in {**var_e, **var_g}.items()}
    emb = nets.encoder_fgbg(p, cfg, x.double(), torch.tensor(b["mask"]).double(), torch.tensor(b["part_bbox"][:, :7]),
                            torch.tensor(b["part_vis"][:, :7]))
    Gr, zr = nets.unet_generator(p, cfg, emb, pose.double())
    assert float((fea_all.double().cpu() - emb).abs().max()) < 1e-3
    assert float((G.double().cpu() - Gr).abs().max()) < 1e-3 and float((z.double().cpu() - zr).abs().max()) < 1e-3


def test_gaussian_fc_res_mirror():
    from dpig_b200 import models
    from dpig_b200.wgan_gp import LeakyReLU
    z = np.random.default_rng(1).normal(0, 0.2, size=(4, 224)).astype(np.float32)
    out, var = models.GaussianFCRes([4, 224], 224, repeat_num=4, hidden_num=512, data_format="NHWC",
                                    activation_fn=LeakyReLU, z=z, scope="Gaussian_FC_Fg/G_FC")
    p = {k: v.detach().double().cpu() for k, v in var.items()}
    ref = nets.gaussian_fc_res(p, torch.tensor(z).double(), 4, "Gaussian_FC_Fg/G_FC", lambda t: T.leaky_relu(t, 0.2))
    assert float((out.double().cpu() - ref).abs().max()) < 1e-5


def test_denorm_and_instance_norm():
    import dpig_b200
    from dpig_b200 import _lib
    from dpig_b200.tensor import SplitTensor, ptr
    ctx = dpig_b200.Context(0)
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(2)
    img = (torch.rand((2, 8, 4, 3), generator=g) * 3 - 1.5).cuda()
    u8 = torch.zeros((2, 8, 4, 3), dtype=torch.uint8, device="cuda")
    ctx.denorm_u8(ptr(img), img.numel(), ptr(u8), s)
    assert torch.equal(u8.cpu(), T.denorm_img(img.cpu()).to(torch.uint8))          # utils.py:88-89
    n, h, w, c = 2, 4, 4, 64
    x = torch.randn((n, h, w, c), generator=g)
    sc, of = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    xd, sd, od = x.cuda(), sc.cuda(), of.cuda()
    sums = torch.zeros((2, n * c), dtype=torch.float64, device="cuda")
    stats = torch.zeros((2, n * c), device="cuda")
    out = SplitTensor(n, h, w, c)
    ctx.norm_stats(ptr(xd), n, h, w, c, _lib.NORM_INSTANCE, ptr(sums), s)
    ctx.norm_act_fwd(ptr(xd), n, h, w, c, _lib.NORM_INSTANCE, 1e-3, ptr(sums), float(h * w), ptr(sd), ptr(od), 0, 0.0,
                     ptr(stats), out.ref(), None, s)
    ref = T.instance_norm(x.double(), sc.double(), of.double(), 1e-3)              # models.py:154-166
    assert float((out.float().double().cpu() - ref).abs().max()) < 1e-4
