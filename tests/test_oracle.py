"""CPU tests of the oracle itself: the vectorised torch restatement (oracle/tf_ops.py, oracle/nets.py) against
the independent scalar-loop restatement (oracle/np_ref.py) and against closed-form cases of the
TensorFlow semantics listed in SURVEY.md §8c.  (Parity with real TF is unpinned: TF is not installable here.)
"""
import math

import numpy as np
import pytest
import torch

from oracle import nets, np_ref
from oracle import tf_ops as T


def test_same_padding_rule():
    # k=3,s=2 on even size: pad (0,1); k=5,s=2: pad (1,2); k=3,s=1: (1,1)   
    assert T.same_pads(128, 3, 2) == (64, 0, 1)
    assert T.same_pads(128, 5, 2) == (64, 1, 2)
    assert T.same_pads(48, 3, 1) == (48, 1, 1)
    assert T.same_pads(3, 3, 2) == (2, 1, 1)
    assert T.same_pads(6, 1, 1) == (6, 0, 0)


def test_conv_matches_scalar_loops():
    rng = np.random.default_rng(0)
    for (h, w, ci, co, k, s) in [(6, 6, 3, 4, 3, 2), (5, 7, 2, 3, 3, 1), (8, 4, 3, 2, 5, 2), (4, 4, 5, 2, 1, 1)]:
        x = rng.normal(size=(2, h, w, ci))
        wt = rng.normal(size=(k, k, ci, co))
        b = rng.normal(size=(co,))
        ref = np_ref.conv2d_same(x, wt, b, s)
        got = T.conv2d_same(torch.tensor(x), torch.tensor(wt), torch.tensor(b), s).numpy()
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() < 1e-12


def test_crop_and_resize_matches_scalar_loops():
    rng = np.random.default_rng(1)
    img = rng.normal(size=(2, 9, 7, 3))
    boxes = np.array([[0.1, 0.2, 0.8, 0.9], [0.0, 0.0, 1.0, 1.0], [-0.2, 0.1, 0.5, 1.3], [0.3, 0.3, 0.3, 0.3],
                      [0.0, 0.0, 1 / 9.0, 1 / 7.0]])
    ind = np.array([0, 1, 1, 0, 1])
    for cs in [(4, 4), (1, 3), (6, 5)]:
        ref = np_ref.crop_and_resize(img, boxes, ind, cs)
        got = T.crop_and_resize(torch.tensor(img), torch.tensor(boxes), torch.tensor(ind), cs).numpy()
        assert np.abs(got - ref).max() < 1e-12


def test_crop_and_resize_identity_box():
    # box [0,0,1,1] with crop == image size reproduces the image (in = i exactly)
    img = torch.arange(2 * 5 * 4 * 1, dtype=torch.float64).reshape(2, 5, 4, 1)
    out = T.crop_and_resize(img, torch.tensor([[0., 0., 1., 1.]]), torch.tensor([1]), (5, 4))
    assert torch.equal(out[0], img[1])


def test_upscale_and_norms():
    rng = np.random.default_rng(2)
    x = rng.normal(size=(2, 3, 2, 4))
    assert np.array_equal(T.upscale2(torch.tensor(x)).numpy(), np_ref.resize_nn2(x))
    sc, of = rng.normal(size=4), rng.normal(size=4)
    assert np.abs(T.batchnorm_train(torch.tensor(x), torch.tensor(sc), torch.tensor(of)).numpy()
                  - np_ref.batchnorm_train(x, sc, of)).max() < 1e-12
    assert np.abs(T.layernorm(torch.tensor(x), torch.tensor(sc), torch.tensor(of)).numpy()
                  - np_ref.layernorm(x, sc, of)).max() < 1e-12


def test_losses_and_adam():
    z = np.array([-3.0, -0.1, 0.0, 0.4, 7.0])
    l = np.array([0.0, 1.0, 1.0, 0.0, 1.0])
    assert np.abs(T.sigmoid_ce(torch.tensor(z), torch.tensor(l)).numpy() - np_ref.sigmoid_ce(z, l)).max() < 1e-12
    # against the textbook definition -l*log(s) - (1-l)*log(1-s)
    s = 1 / (1 + np.exp(-z))
    assert np.abs(np_ref.sigmoid_ce(z, l) - (-l * np.log(s) - (1 - l) * np.log(1 - s))).max() < 1e-9
    rng = np.random.default_rng(3)
    p = rng.normal(size=7)
    grads = [rng.normal(size=7) for _ in range(3)]
    ref = np_ref.adam_steps(p, grads, 2e-5)
    pt = torch.tensor(p)
    m, v = torch.zeros(7, dtype=torch.float64), torch.zeros(7, dtype=torch.float64)
    for t, g in enumerate(grads, 1):
        T.adam_step(pt, torch.tensor(g), m, v, 2e-5, t)
    assert np.abs(pt.numpy() - ref).max() < 1e-15
    # first TF-Adam step moves every weight by ~lr*sign(g) (lr_t*m/(sqrt(v)+eps) with t=1)
    p1 = np_ref.adam_steps(p, grads[:1], 2e-5)
    assert np.allclose(p1 - p, -2e-5 * np.sign(grads[0]), rtol=1e-3)


def test_pose_rasterize_matches_reference_enumeration():
    from dpig_b200 import synth
    batch = synth.make_batch(2, 128, 64, seed=11)
    rcv = batch["pose_rcv"]
    got = T.pose_rasterize(torch.tensor(rcv), 128, 64, 4).numpy()
    ref = np_ref.pose_inflate(rcv, 128, 64)
    assert np.array_equal(got, ref)
    assert set(np.unique(got)) <= {-1.0, 1.0}


def test_network_shapes_and_names():
    cfg = nets.NetConfig()
    p = nets.init_params(cfg)
    n_gen = sum(v.size for k, v in p.items() if nets.is_generator_param(k))
    n_enc = sum(v.size for k, v in p.items() if k.startswith("Encoder/"))
    n_d = sum(v.size for k, v in p.items() if nets.is_disc_param(k))
    # Enc 47.3 M + G 71.2 M = 118.5 M; D 4.3 M
    assert abs(n_enc / 1e6 - 47.3) < 0.1 and abs((n_gen - n_enc) / 1e6 - 71.2) < 0.1 and abs(n_d / 1e6 - 4.3) < 0.1
    assert p["Encoder/G_encoder/Conv/weights"].shape == (3, 3, 3, 128)
    assert p["Encoder/G_encoder/fully_connected/weights"].shape == (5760, 32)
    assert p["Encoder/G_encoder/fully_connected_1/weights"].shape == (20480, 128)
    assert p["ID_AE/G/Conv/weights"].shape == (3, 3, 370, 128)
    assert p["ID_AE/G/fully_connected_1/weights"].shape == (64, 4096)
    assert p["ID_AE/G/Conv_29/weights"].shape == (3, 3, 256, 3)
    assert p["Discriminator.Output.W"].shape == (16384, 1)


def test_small_stage1_forward_and_grads():
    """A reduced graph (32x16, hidden 64) end to end in float64: shapes, finite losses, grads exist."""
    from dpig_b200 import synth
    cfg = nets.NetConfig(img_h=32, img_w=16, hidden=64, roi_size=12, d_dim=64)
    p = nets.to_torch(nets.init_params(cfg, bias_noise=0.05), torch.float64, requires_grad=True)
    b = synth.make_batch(2, 32, 16, seed=5)
    batch = dict(x=torch.tensor(b["x"], dtype=torch.float64), mask=torch.tensor(b["mask"], dtype=torch.float64),
                 pose=T.pose_rasterize(torch.tensor(b["pose_rcv"], dtype=torch.float64), 32, 16),
                 part_bbox=torch.tensor(b["part_bbox"][:, :7]), part_vis=torch.tensor(b["part_vis"][:, :7]))
    out, grads = nets.stage1_grads(p, cfg, batch, "g")
    assert out["emb"].shape == (2, 352) and out["G"].shape == (2, 32, 16, 3) and out["D_fake"].shape == (2,)
    assert all(g is not None and torch.isfinite(g).all() for g in grads.values())
    out, grads = nets.stage1_grads(p, cfg, batch, "d", mode="wgan-gp", gp_alpha=torch.tensor([0.3, 0.8], dtype=torch.float64))
    assert math.isfinite(float(out["d_loss"])) and float(out["gp"]) >= 0
    assert all(g is not None for g in grads.values())


def test_deepfashion_encoder_is_the_fg_branch_without_mask():
    """models.py:328-388 vs 390-471: with an all-ones mask the Fg branch of the two-branch encoder sees the same
    feature map as the DeepFashion encoder, and both number their shared layers identically (slim auto-naming)."""
    from dpig_b200 import synth
    kw = dict(img_h=32, img_w=16, hidden=8, roi_size=12)
    c2 = nets.NetConfig(**kw)
    c1 = nets.NetConfig(fgbg=False, **kw)
    assert c1.emb_dim == 7 * 32 and c2.emb_dim == 7 * 32 + 128
    p = nets.to_torch(nets.init_params(c2, seed=5, bias_noise=0.05))
    b = synth.make_batch(2, 32, 16, seed=9)
    x = torch.tensor(b["x"], dtype=torch.float64)
    bb, vis = torch.tensor(b["part_bbox"][:, :7]), torch.tensor(b["part_vis"][:, :7])
    e2 = nets.encoder_fgbg(p, c2, x, torch.ones(2, 32, 16, 1, dtype=torch.float64), bb, vis)
    e1 = nets.encoder_roi(p, c1, x, bb, vis)
    assert e1.shape == (2, 224)
    assert (e1 - e2[:, :224]).abs().max() < 1e-12


def test_discriminator_row_quirk_on_large_images():
    """wgan_gp.py:433 `tf.reshape(output, [-1, 8*4*8*dim])` after the NCHW transpose: a map with R times the
    128x64 element count becomes R rows per image, row r = channels [r*C/R, (r+1)*C/R) over all pixels."""
    cfg = nets.NetConfig.deepfashion(img_h=128, img_w=128, hidden=8, roi_size=32, d_dim=8)
    assert cfg.d_row == 8 * 4 * 8 * 8 and (128 // 16) * (128 // 16) * 64 // cfg.d_row == 2
    p = nets.to_torch(nets.init_params(cfg, seed=3, bias_noise=0.1))
    g = torch.Generator().manual_seed(1)
    x = torch.rand((3, 128, 128, 3), generator=g, dtype=torch.float64) * 2 - 1
    out = nets.dcgan_discriminator(p, cfg, x, "dcgan")
    assert out.shape == (6,)
    # explicit restatement of the last two lines with numpy indexing
    h = T.leaky_relu(T.conv2d_same(x, p["Discriminator.1.Filters"], p["Discriminator.1.Biases"], 2))
    for i in (2, 3, 4):
        h = T.conv2d_same(h, p["Discriminator.%d.Filters" % i], p["Discriminator.%d.Biases" % i], 2)
        h = T.leaky_relu(T.batchnorm_train(h, p["Discriminator.BN%d.scale" % i], p["Discriminator.BN%d.offset" % i]))
    hn = h.numpy()                                            # [3, 8, 8, 64]
    w = p["Discriminator.Output.W"].numpy().reshape(32, 64)   # row index = c' * 64 + (y*8 + x)
    ref = np.zeros(6)
    for n in range(3):
        for r in range(2):
            blk = hn[n, :, :, r * 32:(r + 1) * 32].reshape(64, 32)     # [pixel, c']
            ref[n * 2 + r] = (blk.T * w).sum() + float(p["Discriminator.Output.b"][0])
    assert np.abs(out.numpy() - ref).max() < 1e-10


def test_joint_discriminator_call_shares_batch_statistics():
    """trainer_256.py:61-66: D(concat([x, G])) then tf.split -- differs from two separate calls (trainer.py:601-602)
    exactly by the BatchNorm statistics."""
    cfg = nets.NetConfig.deepfashion(img_h=64, img_w=64, hidden=8, roi_size=16, d_dim=8)
    p = nets.to_torch(nets.init_params(cfg, seed=3, bias_noise=0.1))
    g = torch.Generator().manual_seed(2)
    x = torch.rand((2, 64, 64, 3), generator=g, dtype=torch.float64) * 2 - 1
    y = torch.rand((2, 64, 64, 3), generator=g, dtype=torch.float64) - 0.5
    joint = nets.dcgan_discriminator(p, cfg, torch.cat([x, y]), "dcgan")
    sep = torch.cat([nets.dcgan_discriminator(p, cfg, x, "dcgan"), nets.dcgan_discriminator(p, cfg, y, "dcgan")])
    assert joint.shape == sep.shape and (joint - sep).abs().max() > 1e-3
    # LayerNorm (wgan-gp) is per sample: there the two agree
    j2 = nets.dcgan_discriminator(p, cfg, torch.cat([x, y]), "wgan-gp")
    s2 = torch.cat([nets.dcgan_discriminator(p, cfg, x, "wgan-gp"), nets.dcgan_discriminator(p, cfg, y, "wgan-gp")])
    assert (j2 - s2).abs().max() < 1e-12


def test_ops_against_scipy_as_a_third_implementation():
    """The oracle's TF-semantics ops against scipy routines that were written by neither of us (the reference's
    TensorFlow is not installable, so independent agreement is the available evidence):
      * SAME stride-1 / stride-2 cross-correlation == scipy.signal.correlate2d on the explicitly padded image
        (pad_before = pad_total // 2, the extra row / column at the bottom / right), subsampled by the stride;
      * crop_and_resize == scipy.ndimage.map_coordinates(order=1) at TF's sample positions
        y1*(H-1) + i*(y2-y1)*(H-1)/(c-1), with zeros outside the image;
      * sigmoid cross-entropy == -l*log(sigmoid z) - (1-l)*log(1 - sigmoid z) via scipy.special.expit / log1p."""
    from scipy import ndimage, signal, special
    g = torch.Generator().manual_seed(3)
    # --- convolution
    for (H, W, k, s) in ((9, 7, 3, 1), (8, 6, 3, 2), (9, 7, 5, 2), (6, 6, 1, 1)):
        x = torch.randn((1, H, W, 2), generator=g, dtype=torch.float64)
        w = torch.randn((k, k, 2, 3), generator=g, dtype=torch.float64)
        y = T.conv2d_same(x, w, None, s)[0].numpy()
        oh, ow = -(-H // s), -(-W // s)
        ph, pw = max((oh - 1) * s + k - H, 0), max((ow - 1) * s + k - W, 0)
        xp = np.pad(x[0].numpy(), ((ph // 2, ph - ph // 2), (pw // 2, pw - pw // 2), (0, 0)))
        ref = np.zeros((oh, ow, 3))
        for co in range(3):
            for ci in range(2):
                full = signal.correlate2d(xp[:, :, ci], w[:, :, ci, co].numpy(), mode="valid")
                ref[:, :, co] += full[::s, ::s][:oh, :ow]
        assert np.abs(y - ref).max() < 1e-12, (H, W, k, s)
    # --- crop_and_resize
    H, W, c = 12, 9, 5
    img = torch.randn((2, H, W, 3), generator=g, dtype=torch.float64)
    boxes = torch.tensor([[0.1, 0.2, 0.8, 0.9], [0.0, 0.0, 1.0, 1.0], [-0.2, 0.3, 1.3, 0.6]], dtype=torch.float64)
    ind = torch.tensor([0, 1, 1])
    out = T.crop_and_resize(img, boxes, ind, (c, c)).numpy()
    for b in range(3):
        y1, x1, y2, x2 = boxes[b].tolist()
        ys = y1 * (H - 1) + np.arange(c) * (y2 - y1) * (H - 1) / (c - 1)
        xs = x1 * (W - 1) + np.arange(c) * (x2 - x1) * (W - 1) / (c - 1)
        yy, xx = np.meshgrid(ys, xs, indexing="ij")
        inside = (yy >= 0) & (yy <= H - 1) & (xx >= 0) & (xx <= W - 1)
        for ch in range(3):
            ref = ndimage.map_coordinates(img[int(ind[b]), :, :, ch].numpy(), [yy, xx], order=1, mode="nearest")
            assert np.abs(out[b, :, :, ch] - np.where(inside, ref, 0.0)).max() < 1e-12, b
    # --- sigmoid cross-entropy
    z = torch.tensor([-30.0, -2.0, 0.0, 0.5, 40.0], dtype=torch.float64)
    for lab in (0.0, 1.0):
        with np.errstate(divide="ignore", invalid="ignore"):      # expit saturates at +-40: those entries are skipped
            ref = -lab * np.log(special.expit(z.numpy())) - (1 - lab) * np.log1p(-special.expit(z.numpy()))
        got = T.sigmoid_ce(z, torch.full_like(z, lab)).numpy()
        ok = np.isfinite(ref)
        assert np.allclose(got[ok], ref[ok], rtol=1e-10, atol=1e-12)


def _small_batch(cfg, n=2, seed=5):
    from dpig_b200 import synth
    b = synth.make_batch(n, cfg.img_h, cfg.img_w, seed=seed)
    return dict(x=torch.tensor(b["x"], dtype=torch.float64), mask=torch.tensor(b["mask"], dtype=torch.float64),
                pose=T.pose_rasterize(torch.tensor(b["pose_rcv"], dtype=torch.float64), cfg.img_h, cfg.img_w),
                part_bbox=torch.tensor(b["part_bbox"][:, :7]), part_vis=torch.tensor(b["part_vis"][:, :7]))


def test_forced_branches_reproduce_the_free_run_and_move_gradients_when_flipped():
    """`branches` (the test aid the GPU gradient tests rest on): a run that is handed its own recorded ReLU / LeakyReLU
    decisions is the free run bit for bit -- values and gradients; flipping ONE late critic bit leaves the loss where it
    was to first order and moves a BatchNorm-scale gradient visibly (what tests/probe_grad_flake.py measures on the GPU)."""
    cfg = nets.NetConfig(img_h=32, img_w=16, hidden=64, roi_size=12, d_dim=64)
    p = nets.to_torch(nets.init_params(cfg, bias_noise=0.05), torch.float64, requires_grad=True)
    batch = _small_batch(cfg)
    rec = {}
    out0, g0 = nets.stage1_grads(p, cfg, batch, "g", branches={"record": rec})
    assert set(rec) == {"Encoder/G_encoder", "ID_AE/G", "D_real", "D_fake"}
    assert len(rec["Encoder/G_encoder"]) == 3 + 2 * (3 * cfg.enc_repeat - 1) and len(rec["D_fake"]) == 4
    assert len(rec["ID_AE/G"]) == 1 + (3 * cfg.unet_repeat - 1) + (3 * cfg.unet_repeat - 1)
    out1, g1 = nets.stage1_grads(p, cfg, batch, "g", branches=rec)
    assert torch.equal(out0["G"], out1["G"]) and torch.equal(out0["D_fake"], out1["D_fake"])
    for k in g0:
        assert torch.equal(g0[k], g1[k]), k
    # one flipped bit in the third critic layer of the D(G) pass
    _, d0 = nets.stage1_grads(p, cfg, batch, "d", branches=rec)
    flipped = dict(rec)
    flipped["D_fake"] = [t.clone() for t in rec["D_fake"]]
    flipped["D_fake"][2][0, 0, 0, 0] ^= True
    _, d1 = nets.stage1_grads(p, cfg, batch, "d", branches=flipped)
    k = "Discriminator.BN2.scale"
    rel = float((d1[k] - d0[k]).norm() / d0[k].norm())
    assert 1e-5 < rel < 0.5, rel


def test_engine_activation_bits_have_the_oracles_layout():
    """Stage1Engine.activation_bits() (what the GPU gradient tests hand to the oracle) against the oracle's own record:
    same keys, same number of activated layers in the same order, same shapes -- except the 1x1 convs of the decoder,
    which the engine runs BEFORE the x2 upsample (half the height and width; the oracle repeats the bits).  Buffers on the
    CPU, nothing launched."""
    from test_engine_dryrun import DryContext
    from dpig_b200 import engine
    cases = [("market", dict(img_h=32, img_w=16, hidden=64, roi_size=12, d_dim=64), nets.NetConfig, engine.NetConfig, "wgan-gp"),
             ("deepfashion", dict(img_h=64, img_w=64, hidden=64, roi_size=16), nets.NetConfig.deepfashion,
              engine.NetConfig.deepfashion, "dcgan")]
    for name, kw, omk, emk, mode in cases:
        ocfg, ecfg = omk(**kw), emk(**kw)
        p = nets.to_torch(nets.init_params(ocfg), torch.float64)
        rec = {}
        with torch.no_grad():
            alpha = torch.tensor([0.3, 0.8], dtype=torch.float64) if mode == "wgan-gp" else None
            if alpha is None:
                nets.stage1_forward(p, ocfg, _small_batch(ocfg), mode, branches={"record": rec})
        if alpha is not None:
            nets.stage1_forward(p, ocfg, _small_batch(ocfg), mode, gp_alpha=alpha, branches={"record": rec})
        eng = engine.Stage1Engine(DryContext(), ecfg, 2, mode=mode, device="cpu")
        bits = eng.activation_bits()
        assert set(bits) == set(rec), (name, set(bits) ^ set(rec))
        for key in rec:
            assert len(bits[key]) == len(rec[key]), (name, key, len(bits[key]), len(rec[key]))
            for i, (a, b) in enumerate(zip(bits[key], rec[key])):
                same = tuple(a.shape) == tuple(b.shape)
                half = (a.shape[0], 2 * a.shape[1], 2 * a.shape[2], a.shape[3]) == tuple(b.shape)
                assert same or (key == "ID_AE/G" and half), (name, key, i, tuple(a.shape), tuple(b.shape))


@pytest.mark.parametrize("seed", [1234, 77])
def test_float32_run_matches_float64_on_its_own_branches(seed):
    """The statement the GPU gradient bounds rest on, checked without a GPU: an fp32 evaluation of the graph and the
    float64 one agree on every parameter gradient to fp32 rounding (< 2e-4 relative L2) when the float64 run takes the
    ReLU / LeakyReLU branches the fp32 run took -- whatever they do when each decides for itself (with this
    torch build seed 77 has ONE activation whose sign differs between the two precisions, and the free-running g_loss
    gradients then differ by 3.6e-3 instead of 1.7e-5)."""
    cfg = nets.NetConfig(img_h=32, img_w=16, hidden=64, roi_size=12, d_dim=64)
    params = nets.init_params(cfg, seed=seed, bias_noise=0.05)
    b64 = _small_batch(cfg, seed=123)
    b32 = {k: (v.float() if v.is_floating_point() else v) for k, v in b64.items()}
    p32 = nets.to_torch(params, torch.float32, requires_grad=True)
    p64 = nets.to_torch(params, torch.float64, requires_grad=True)
    for which in ("g", "d"):
        rec = {}
        _, g32 = nets.stage1_grads(p32, cfg, b32, which, branches={"record": rec})
        _, g64 = nets.stage1_grads(p64, cfg, b64, which, branches=rec)
        top = max(float(g.abs().max()) for g in g64.values())
        worst = 0.0
        for k, g in g64.items():
            if float(g.abs().max()) < 1e-9 * top:        # conv biases under a BatchNorm: exactly-zero gradient
                continue
            worst = max(worst, float((g32[k].double() - g).norm() / g.norm()))
        assert worst < 2e-4, (which, worst)
