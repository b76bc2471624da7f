"""Output path of the reference's test() / generate(): the SSIM oracle (CPU), the sample sheets
and per-sample PNG directories (CPU), and -- marked gpu -- the device SSIM kernel and tester.test() end to end."""
import importlib
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import image_metrics as im  # noqa: E402

outputs = importlib.import_module("disentangled-person-image-generation_b200.outputs")


def _pair(seed, h=24, w=16):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(h // 4, w // 4, 3))
    x = np.kron(base, np.ones((4, 4, 1))).astype(np.uint8)
    g = np.clip(x.astype(np.int64) + rng.integers(-30, 31, size=x.shape), 0, 255).astype(np.uint8)
    return g, x


def test_ssim_oracle_against_bruteforce_and_closed_forms():
    for seed in range(3):
        g, x = _pair(seed)
        gg, xg = im.rgb2gray_u8(g), im.rgb2gray_u8(x)
        dr = xg.max() - xg.min()
        assert abs(im.ssim_gray(gg, xg, dr) - im.ssim_gray_bruteforce(gg, xg, dr)) < 1e-12
        assert abs(im.ssim_gray(xg, xg, dr) - 1.0) < 1e-12                      # identical images
    # rgb2gray weights sum to 1: a grey uint8 image maps to v/255
    grey = np.full((8, 8, 3), 51, np.uint8)
    assert np.allclose(im.rgb2gray_u8(grey), 0.2, atol=1e-15)
    # constant images a, b: variances vanish, S = (2ab + C1)/(a^2 + b^2 + C1) everywhere
    a, b, dr = 0.25, 0.5, 1.0
    s = im.ssim_gray(np.full((9, 9), a), np.full((9, 9), b), dr)
    assert abs(s - (2 * a * b + 1e-4) / (a * a + b * b + 1e-4)) < 1e-12
    assert im.ssim_generate(np.stack([g, x]), np.stack([x, x])).shape == (2,)


def test_make_grid_matches_reference_layout():
    # utils.py:157-175: cell = image + padding, origin 1 + padding//2, sheet = cells + 1 + padding//2
    t = np.arange(10 * 4 * 3 * 3, dtype=np.uint8).reshape(10, 4, 3, 3)
    g = outputs.make_grid(t, nrow=8, padding=2)
    assert g.shape == (6 * 2 + 2, 5 * 8 + 2, 3) and g.dtype == np.uint8
    assert np.array_equal(g[2:6, 2:5], t[0]) and np.array_equal(g[2:6, 7:10], t[1]) and np.array_equal(g[8:12, 2:5], t[8])
    assert g[0].sum() == 0 and g[:, 0].sum() == 0 and g[6:8].sum() == 0            # borders stay black
    one = outputs.make_grid(np.full((3, 4, 3), 7, np.uint8))                       # [n,h,w] masks broadcast to RGB
    assert one.shape[2] == 3 and (one[2:6, 2:5] == 7).all()


def test_result_writer_files_and_names(tmp_path):
    from PIL import Image
    B, H, W = 3, 16, 8
    rng = np.random.default_rng(1)
    img = lambda: rng.integers(0, 256, size=(B, H, W, 3)).astype(np.float32)
    plane = lambda: rng.integers(0, 256, size=(B, H, W)).astype(np.float32)
    wr = outputs.ResultWriter(str(tmp_path / "test_result"))
    G = img()
    wr.add_batch(2, B, img(), img(), G, plane(), plane(), img(), plane()[..., None], plane()[..., None], [0.5, -1.25, 3.0])
    wr.add_grid(G, "x_fixed.png")
    assert wr.close() == 8 * B + 1
    root = tmp_path / "test_result"
    assert sorted(os.listdir(root)) == sorted(list(outputs.DIRS) + ["x_fixed.png"])
    assert sorted(os.listdir(root / "x")) == ["00006.png", "00007.png", "00008.png"]            # idx = i*B + j
    assert sorted(os.listdir(root / "G_pose")) == ["0002_0000.png", "0002_0001.png", "0002_0002.png"]
    assert "0002_c1s1_000001_00007_-1.250000.png" in os.listdir(root / "G")                     # score.py parses these
    back = np.asarray(Image.open(root / "G" / "0002_c1s1_000000_00006_0.500000.png"))
    assert np.array_equal(back, G[0].astype(np.uint8))                                          # PNG is lossless
    assert np.asarray(Image.open(root / "mask" / "00006.png")).shape == (H, W)


@pytest.mark.gpu
def test_ssim_kernel_matches_oracle():
    import torch
    import dpig_b200
    from dpig_b200.tensor import ptr
    ctx = dpig_b200.Context(0)
    for (n, h, w) in ((5, 128, 64), (2, 256, 256), (3, 7, 9)):
        rng = np.random.default_rng(n)
        x = rng.integers(0, 256, size=(n, h, w, 3)).astype(np.uint8)
        x[0] = np.kron(rng.integers(0, 256, size=(1, 1, 3)), np.ones((h, w, 1))).astype(np.uint8) if n > 2 else x[0]
        g = np.clip(x.astype(np.int64) + rng.integers(-40, 41, size=x.shape), 0, 255).astype(np.uint8)
        g[-1] = x[-1]                                                              # identical pair -> 1
        out = torch.empty((n,), dtype=torch.float32, device="cuda")
        gd, xd = torch.from_numpy(g).cuda(), torch.from_numpy(x).cuda()
        ctx.ssim_gray_u8(ptr(gd), ptr(xd), n, h, w, ptr(out), torch.cuda.current_stream().cuda_stream)
        ref = im.ssim_generate(g, x)
        got = out.cpu().numpy().astype(np.float64)
        ok = np.isfinite(ref)                       # a constant input image has data_range 0: 0/0 on both sides
        assert np.allclose(got[ok], ref[ok], atol=2e-6), (got, ref)
        assert abs(got[-1] - 1.0) < 1e-6


@pytest.mark.gpu
def test_tester_writes_reference_result_directories(tmp_path):
    from PIL import Image
    from dpig_b200 import config as cfgmod
    from dpig_b200 import engine, tester
    B = 4
    conf, _ = cfgmod.get_config(["--model=13", "--is_train=False", "--batch_size=%d" % B, "--img_H=32", "--img_W=16",
                                 "--conv_hidden_num=64", "--sample_fg=True", "--model_dir=%s" % tmp_path])
    t = tester.DPIG_FourNetsFgBg_testOnlySampleFactor(conf)
    t.init_net(engine.NetConfig(img_h=32, img_w=16, hidden=64, roi_size=12, d_dim=64))
    out_dir = t.test(num_batches=2)
    assert t.files_written == 2 * 8 * B + 6
    for d in ("x", "x_target", "G", "pose", "pose_target", "G_pose", "mask", "mask_target"):
        assert len(os.listdir(os.path.join(out_dir, d))) == 2 * B, d
    sheets = [f for f in os.listdir(out_dir) if f.endswith(".png")]
    assert any(f.startswith("0_G_ssim") for f in sheets) and any(f.startswith("1_G_ssim") for f in sheets)
    assert {"x_fixed.png", "mask_fixed.png", "pose_fixed.png"} <= set(sheets)
    g = np.asarray(Image.open(os.path.join(out_dir, "G", sorted(os.listdir(os.path.join(out_dir, "G")))[0])))
    assert g.shape == (32, 16, 3) and g.dtype == np.uint8
    # the SSIM in the sheet's name is the device value of the last generate(): compare it with the oracle on the same images
    s = t.last_ssim
    G8 = np.clip((t.s1.G.cpu().numpy() + 1.0) * 127.5, 0, 255).astype(np.uint8)
    x8 = np.clip((t.s1.x.cpu().numpy() + 1.0) * 127.5, 0, 255).astype(np.uint8)
    assert np.allclose(s, im.ssim_generate(G8, x8), atol=5e-6)
