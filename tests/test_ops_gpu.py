"""GPU parity tests, op level: every kernel reached through the C ABI (ctypes) against the CPU oracle
(oracle/tf_ops.py in float64) on the same seeded inputs.  Tolerances are written next to each check;
the north-star bound is 1e-3 max-abs on O(1) activations, the kernels are held to much tighter ones.

Run as a script (`python tests/test_ops_gpu.py [filter]`) for a verbose bring-up report.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import tf_ops as T  # noqa: E402

pytestmark = pytest.mark.gpu

_CTX = None


def ctx():
    global _CTX
    if _CTX is None:
        import dpig_b200
        _CTX = dpig_b200.Context(0)
    return _CTX


def _imports():
    from dpig_b200 import _lib
    from dpig_b200.tensor import SplitTensor, ptr, split_ref
    return _lib, SplitTensor, ptr, split_ref


def stream():
    return torch.cuda.current_stream().cuda_stream


def rel_err(a, b):
    a = a.double().cpu()
    b = b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def pack_weights(w_hwio, cin_pad=None, cout_pad=None):
    """fp32 HWIO (torch cuda) -> packed operand copies via dpig_weight_pack."""
    _lib, SplitTensor, ptr, _ = _imports()
    kh, kw, cin, cout = w_hwio.shape
    cin_pad = cin_pad or (cin + 7) // 8 * 8
    cout_pad = cout_pad or (cout + 7) // 8 * 8
    taps = kh * kw
    f = torch.zeros((2, taps, cout, cin_pad), dtype=torch.bfloat16, device="cuda")
    b = torch.zeros((2, taps, cin, cout_pad), dtype=torch.bfloat16, device="cuda")
    ctx().weight_pack(ptr(w_hwio), taps, cin, cout, cin_pad, cout_pad, ptr(f[0]), ptr(f[1]), ptr(b[0]), ptr(b[1]),
                      stream())
    return f, b


def padded_split(x, cpad=None):
    """Split tensor whose channel count is zero-padded up to cpad (default: next multiple of 8)."""
    _lib, SplitTensor, ptr, split_ref = _imports()
    c = x.shape[-1]
    cpad = cpad or (c + 7) // 8 * 8
    if cpad != c:
        x = torch.cat([x, torch.zeros(x.shape[:-1] + (cpad - c,))], dim=-1)
    return SplitTensor.from_float(x.cuda())


def run_conv_fwd(n, h, w, cin, cout, k, stride, act="relu", residual=False, upsample=1, f32_out=False, seed=0,
                 c_alloc_in=None):
    _lib, SplitTensor, ptr, split_ref = _imports()
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((n, h, w, cin), generator=g)
    wt = torch.randn((k, k, cin, cout), generator=g) / np.sqrt(k * k * cin)
    bias = torch.randn((cout,), generator=g) * 0.1
    oh, ow = -(-h // stride), -(-w // stride)
    res = torch.randn((n, oh, ow, cout), generator=g) if residual else None

    xs = padded_split(x, c_alloc_in)
    wd = wt.cuda().contiguous()
    f, _ = pack_weights(wd, cin_pad=xs.c)
    bd = bias.cuda()
    out = SplitTensor(n, oh * upsample, ow * upsample, cout, zero=True)
    out32 = torch.zeros((n, oh * upsample, ow * upsample, cout), device="cuda") if f32_out else None
    words = (cout + 31) // 32
    mask = torch.zeros((n * oh * ow, words), dtype=torch.int32, device="cuda")
    import ctypes as C
    ep = _lib.ConvEpilogue()
    ep.bias = bd.data_ptr()
    ep.act = {"none": 0, "relu": 1, "lrelu": 2}[act]
    ep.alpha = 0.2
    rs = None
    if residual:
        rs = SplitTensor.from_float(res.cuda())
        ep.addend = C.pointer(rs.struct())
    ep.mask_out = mask.data_ptr()
    ep.out = C.pointer(out.struct())
    if f32_out:
        ep.out_f32 = out32.data_ptr()
        ep.out_f32_pix_stride = cout
    ep.upsample = upsample
    ctx().conv2d_fwd(xs.ref(), ptr(f[0]), ptr(f[1]), k, k, stride, cout, C.byref(ep), stream())
    torch.cuda.synchronize()

    # oracle on the values the kernel actually saw (split-rounded inputs), float64
    xr = split_ref(x).double()
    wr = split_ref(wt).double()
    pre = T.conv2d_same(xr, wr, bias.double(), stride)
    y = pre
    if act == "relu":
        y = torch.relu(pre)
    elif act == "lrelu":
        y = T.leaky_relu(pre, 0.2)
    if residual:
        y = y + split_ref(res).double()
    if upsample == 2:
        y = T.upscale2(y)
    got = out.float().cpu()
    err = rel_err(got, y)
    info = {"err": err}
    if f32_out:
        info["err_f32"] = rel_err(out32.cpu(), y)
    # mask bits: compare where |pre| is not tiny
    bits = mask.cpu().numpy().astype(np.uint32).reshape(n, oh, ow, words)
    exp = (pre > 0).numpy()
    got_bits = np.zeros_like(exp)
    for c in range(cout):
        got_bits[..., c] = (bits[..., c // 32] >> (c % 32)) & 1
    sure = (pre.abs() > 1e-4).numpy()
    info["mask_mismatch"] = int(((got_bits != exp) & sure).sum())
    return info


CONV_FWD_CASES = [
    # n, h, w, cin, cout, k, stride, act, residual, upsample, f32
    dict(n=2, h=16, w=8, cin=128, cout=128, k=3, stride=1, act="relu", residual=True),
    dict(n=2, h=16, w=8, cin=64, cout=128, k=3, stride=1, act="none"),
    dict(n=3, h=12, w=12, cin=128, cout=256, k=3, stride=2, act="relu"),
    dict(n=2, h=16, w=8, cin=256, cout=384, k=3, stride=1, act="relu", residual=True),
    dict(n=2, h=32, w=16, cin=64, cout=128, k=5, stride=2, act="none", f32_out=True),
    dict(n=2, h=8, w=4, cin=192, cout=64, k=1, stride=1, act="relu", upsample=2),
    dict(n=5, h=3, w=3, cin=640, cout=640, k=3, stride=1, act="relu", residual=True),
    dict(n=1, h=128, w=64, cin=128, cout=128, k=3, stride=1, act="relu"),
    dict(n=2, h=16, w=8, cin=256, cout=3, k=3, stride=1, act="none", f32_out=True),
    dict(n=2, h=16, w=8, cin=370, cout=128, k=3, stride=1, act="relu", c_alloc_in=384),
    # CTA-pair kernel: odd pixel-tile count (one all-padding tile), several units per cluster, N=192 halves
    dict(n=3, h=16, w=8, cin=128, cout=128, k=3, stride=1, act="relu", residual=True),
    dict(n=21, h=32, w=32, cin=64, cout=256, k=3, stride=1, act="relu", residual=True),
    dict(n=7, h=16, w=8, cin=64, cout=384, k=3, stride=1, act="lrelu"),
]


# single-CTA kernel only / 2-CTA cluster kernel forced wherever the shape allows
TILINGS = [0, 2]


@pytest.mark.parametrize("tiling", TILINGS)
@pytest.mark.parametrize("case", CONV_FWD_CASES)
def test_conv2d_fwd(case, tiling):
    ctx().set_pair_mode(tiling)
    try:
        info = run_conv_fwd(**case)
    finally:
        ctx().set_pair_mode(1)
    assert info["err"] < 5e-5, info  # 3-pass split-bf16 vs float64 on identical (split-rounded) inputs
    assert info["mask_mismatch"] == 0, info
    if "err_f32" in info:
        assert info["err_f32"] < 5e-5, info


CONV_STAT_CASES = [
    # the discriminator's conv -> norm blocks (wgan_gp.py:417-431) and shapes that stress the tiling: several images per
    # tile (8x4 maps: a warp's rows span images), ragged pixel tiles, two channel blocks / CTA pairs
    dict(n=4, h=32, w=16, cin=64, cout=128, k=5, stride=2),
    dict(n=6, h=16, w=8, cin=128, cout=256, k=5, stride=2),
    dict(n=5, h=8, w=4, cin=256, cout=512, k=5, stride=2),
    dict(n=3, h=12, w=10, cin=64, cout=96, k=3, stride=1),
]


@pytest.mark.parametrize("tiling", TILINGS)
@pytest.mark.parametrize("mode", ["batch", "layer"])
@pytest.mark.parametrize("case", CONV_STAT_CASES)
def test_conv2d_fwd_stat_emission(case, mode, tiling):
    """dpig_conv_epilogue::stat_sums: the raw sums (sum x, sum x^2 of conv + bias) per channel (Batchnorm,
    tflib/ops/batchnorm.py:29-30) or per sample (Layernorm, layernorm.py:6-20) leave the conv epilogue; checked against
    the float64 sums of the oracle convolution and against dpig_norm_stats on the kernel's own fp32 output."""
    _lib, SplitTensor, ptr, split_ref = _imports()
    import ctypes as C
    n, h, w, cin, cout, k, stride = (case[x] for x in ("n", "h", "w", "cin", "cout", "k", "stride"))
    g = torch.Generator().manual_seed(11)
    x = torch.randn((n, h, w, cin), generator=g)
    wt = torch.randn((k, k, cin, cout), generator=g) / np.sqrt(k * k * cin)
    bias = torch.randn((cout,), generator=g) * 0.5
    oh, ow = -(-h // stride), -(-w // stride)
    xs = padded_split(x)
    f, _ = pack_weights(wt.cuda().contiguous(), cin_pad=xs.c)
    bd = bias.cuda()
    md = {"layer": 0, "batch": 1}[mode]
    groups = n if mode == "layer" else cout
    out32 = torch.zeros((n, oh, ow, cout), device="cuda")
    sums = torch.full((2, groups), 123.0, dtype=torch.float64, device="cuda")    # the call zeroes it
    ep = _lib.ConvEpilogue()
    ep.bias = bd.data_ptr()
    ep.act = 0
    ep.out_f32 = out32.data_ptr()
    ep.out_f32_pix_stride = cout
    ep.upsample = 1
    ep.stat_sums = sums.data_ptr()
    ep.stat_mode = md
    ctx().set_pair_mode(tiling)
    try:
        ctx().conv2d_fwd(xs.ref(), ptr(f[0]), ptr(f[1]), k, k, stride, cout, C.byref(ep), stream())
    finally:
        ctx().set_pair_mode(1)
    torch.cuda.synchronize()
    pre = T.conv2d_same(split_ref(x).double(), split_ref(wt).double(), bias.double(), stride)
    dims = (1, 2, 3) if mode == "layer" else (0, 1, 2)
    ref = torch.stack([pre.sum(dim=dims), (pre * pre).sum(dim=dims)])
    assert rel_err(out32, pre) < 5e-5
    assert rel_err(sums[0], ref[0]) < 2e-5 and rel_err(sums[1], ref[1]) < 2e-5, (sums.cpu(), ref)
    if cout % 32 == 0:
        sums2 = torch.zeros_like(sums)
        ctx().norm_stats(ptr(out32), n, oh, ow, cout, md, ptr(sums2), stream())
        torch.cuda.synchronize()
        assert rel_err(sums, sums2) < 2e-6


def run_conv_bwd_data(n, h, w, cin, cout, k, stride, addend=False, masked=False, seed=1):
    _lib, SplitTensor, ptr, split_ref = _imports()
    import ctypes as C
    g = torch.Generator().manual_seed(seed)
    oh, ow = -(-h // stride), -(-w // stride)
    dy = torch.randn((n, oh, ow, cout), generator=g)
    wt = torch.randn((k, k, cin, cout), generator=g) / np.sqrt(k * k * cout)
    add = torch.randn((n, h, w, cin), generator=g) if addend else None
    mbits = torch.randint(0, 2, (n, h, w, cin), generator=g) if masked else None

    dys = padded_split(dy)
    _, b = pack_weights(wt.cuda().contiguous(), cout_pad=dys.c)
    out = SplitTensor(n, h, w, cin, zero=True)
    out2 = SplitTensor(n, h, w, cin, zero=True) if masked else None
    ep = _lib.ConvEpilogue()
    ep.act = 0
    ep.out = C.pointer(out.struct())
    if addend:
        adds = SplitTensor.from_float(add.cuda())
        ep.addend = C.pointer(adds.struct())
    if masked:
        words = (cin + 31) // 32
        mw = np.zeros((n, h, w, words), np.uint32)
        mb = mbits.numpy().astype(np.uint32)
        for c in range(cin):
            mw[..., c // 32] |= mb[..., c] << np.uint32(c % 32)
        mask_dev = torch.from_numpy(mw.view(np.int32)).cuda()
        ep.mask_in = mask_dev.data_ptr()
        ep.mask_neg = 0.2
        ep.out_masked = C.pointer(out2.struct())
        # bias gradient of the layer below = column sums of the masked gradient, accumulated (+=) by the epilogue
        colsum = torch.full((cin,), 0.5, device="cuda")
        ep.colsum_masked = colsum.data_ptr()
    ep.upsample = 1
    ctx().conv2d_bwd_data(dys.ref(), ptr(b[0]), ptr(b[1]), k, k, stride, h, w, cin, C.byref(ep), stream())
    torch.cuda.synchronize()

    x = torch.zeros((n, h, w, cin), dtype=torch.float64, requires_grad=True)
    y = T.conv2d_same(x, split_ref(wt).double(), None, stride)
    gx, = torch.autograd.grad(y, x, split_ref(dy).double())
    if addend:
        gx = gx + split_ref(add).double()
    info = {"err": rel_err(out.float(), gx)}
    if masked:
        gm = gx * torch.where(mbits.bool(), 1.0, 0.2)
        info["err_masked"] = rel_err(out2.float(), gm)
        info["err_colsum"] = rel_err(colsum.cpu() - 0.5, gm.sum(dim=(0, 1, 2)))
    return info


CONV_BWD_DATA_CASES = [
    dict(n=2, h=16, w=8, cin=128, cout=128, k=3, stride=1, addend=True, masked=True),
    dict(n=2, h=16, w=8, cin=128, cout=256, k=3, stride=2, addend=True, masked=True),
    dict(n=2, h=32, w=16, cin=64, cout=128, k=5, stride=2),
    dict(n=2, h=8, w=4, cin=192, cout=64, k=1, stride=1),
    dict(n=3, h=12, w=12, cin=256, cout=384, k=3, stride=2, masked=True),
    dict(n=2, h=16, w=8, cin=256, cout=3, k=3, stride=1),
    dict(n=21, h=32, w=32, cin=256, cout=64, k=3, stride=1, addend=True, masked=True),
    dict(n=3, h=16, w=8, cin=128, cout=128, k=3, stride=2, addend=True, masked=True),
    # odd input size: the four parity classes have different grids -> one launch per class (no merged launch)
    dict(n=2, h=15, w=9, cin=64, cout=128, k=3, stride=2, addend=True, masked=True),
    dict(n=40, h=32, w=16, cin=128, cout=256, k=3, stride=2, masked=True),   # many units per CTA, merged classes
]


@pytest.mark.parametrize("tiling", TILINGS)
@pytest.mark.parametrize("case", CONV_BWD_DATA_CASES)
def test_conv2d_bwd_data(case, tiling):
    ctx().set_pair_mode(tiling)
    try:
        info = run_conv_bwd_data(**case)
    finally:
        ctx().set_pair_mode(1)
    assert info["err"] < 5e-5, info
    if "err_masked" in info:
        assert info["err_masked"] < 5e-5, info
        assert info["err_colsum"] < 2e-4, info   # fp32 sums over up to 21k pixels, minus the 0.5 it started from


def test_conv2d_bwd_data_merged_classes_equal_separate_launches(monkeypatch):
    """Stride-2 data gradient: one launch running the four parity classes per pixel tile (default) against four
    launches (DPIG_DGRAD_MERGE=0) -- same taps, same accumulation order per output, so the results are bit-identical."""
    import ctypes as C
    import dpig_b200
    _lib, SplitTensor, ptr, split_ref = _imports()
    outs = []
    for merge in ("1", "0"):
        monkeypatch.setenv("DPIG_DGRAD_MERGE", merge)
        c = dpig_b200.Context(0)
        g = torch.Generator().manual_seed(11)
        n, h, w, cin, cout, k = 5, 32, 16, 128, 256, 5
        dy = torch.randn((n, h // 2, w // 2, cout), generator=g)
        wt = torch.randn((k, k, cin, cout), generator=g) * 0.05
        dys = padded_split(dy)
        _, b = pack_weights(wt.cuda().contiguous(), cout_pad=dys.c)
        out = SplitTensor(n, h, w, cin, zero=True)
        ep = _lib.ConvEpilogue()
        ep.act = 0
        ep.out = C.pointer(out.struct())
        ep.upsample = 1
        n0 = c.launch_count()
        c.conv2d_bwd_data(dys.ref(), ptr(b[0]), ptr(b[1]), k, k, 2, h, w, cin, C.byref(ep), stream())
        torch.cuda.synchronize()
        outs.append((out.float().cpu(), c.launch_count() - n0))
    assert outs[0][1] == 1 and outs[1][1] == 4
    assert torch.equal(outs[0][0], outs[1][0])
    assert float(outs[0][0].abs().max()) > 0.1


def run_conv_bwd_filter(n, h, w, cin, cout, k, stride, seed=2):
    _lib, SplitTensor, ptr, split_ref = _imports()
    g = torch.Generator().manual_seed(seed)
    oh, ow = -(-h // stride), -(-w // stride)
    x = torch.randn((n, h, w, cin), generator=g)
    dy = torch.randn((n, oh, ow, cout), generator=g)
    xs = padded_split(x)
    dys = padded_split(dy)
    dw = torch.zeros((k, k, cin, cout), device="cuda")
    ctx().conv2d_bwd_filter(xs.ref(), dys.ref(), k, k, stride, cin, cout, ptr(dw), stream())
    torch.cuda.synchronize()
    wv = torch.zeros((k, k, cin, cout), dtype=torch.float64, requires_grad=True)
    y = T.conv2d_same(split_ref(x).double(), wv, None, stride)
    gw, = torch.autograd.grad(y, wv, split_ref(dy).double())
    return {"err": rel_err(dw, gw)}


CONV_BWD_FILTER_CASES = [
    dict(n=2, h=16, w=8, cin=128, cout=128, k=3, stride=1),
    dict(n=3, h=12, w=12, cin=128, cout=256, k=3, stride=2),
    dict(n=2, h=32, w=16, cin=64, cout=128, k=5, stride=2),
    dict(n=2, h=8, w=4, cin=192, cout=64, k=1, stride=1),
    dict(n=5, h=3, w=3, cin=640, cout=640, k=3, stride=1),
    dict(n=2, h=16, w=8, cin=256, cout=3, k=3, stride=1),
    dict(n=2, h=16, w=8, cin=370, cout=128, k=3, stride=1),
    dict(n=1, h=128, w=64, cin=128, cout=128, k=3, stride=1),
    # CTA-pair kernel: even / odd numbers of (tap, channel tile) units, two output-channel tiles, split-K over many tiles
    dict(n=2, h=16, w=8, cin=256, cout=256, k=3, stride=1),
    dict(n=3, h=12, w=12, cin=384, cout=512, k=3, stride=1),
    dict(n=9, h=32, w=32, cin=256, cout=128, k=3, stride=1),
    dict(n=4, h=16, w=16, cin=128, cout=256, k=1, stride=1),
    # output widths no 256-wide block divides: 256-wide column segments + remainder, one launch each (640 is case 5)
    dict(n=3, h=12, w=12, cin=384, cout=384, k=3, stride=1),
    dict(n=2, h=16, w=8, cin=256, cout=384, k=3, stride=2),
    dict(n=2, h=8, w=4, cin=128, cout=896, k=3, stride=1),
]


@pytest.mark.parametrize("case", CONV_BWD_FILTER_CASES)
def test_conv2d_bwd_filter(case):
    info = run_conv_bwd_filter(**case)
    assert info["err"] < 2e-5, info


def test_conv_small():
    """CUDA-core 3-channel convs (encoder stem, D layer 1) forward and gradients."""
    _lib, SplitTensor, ptr, split_ref = _imports()
    g = torch.Generator().manual_seed(3)
    for (k, stride, cout, act) in [(3, 1, 128, 1), (5, 2, 64, 2)]:
        n, h, w, cin = 2, 32, 16, 3
        x = torch.randn((n, h, w, cin), generator=g)
        wt = torch.randn((k, k, cin, cout), generator=g) * 0.2
        bias = torch.randn((cout,), generator=g) * 0.1
        oh, ow = -(-h // stride), -(-w // stride)
        out = SplitTensor(n, oh, ow, cout, zero=True)
        o32 = torch.zeros((n, oh, ow, cout), device="cuda")
        mask = torch.zeros((n * oh * ow, cout // 32), dtype=torch.int32, device="cuda")
        xd, wd, bd = x.cuda(), wt.cuda(), bias.cuda()
        ctx().conv2d_small_fwd(ptr(xd), n, h, w, cin, ptr(wd), ptr(bd), k, k, stride, cout, act, 0.2, out.ref(),
                               ptr(o32), ptr(mask), stream())
        pre = T.conv2d_same(x.double(), wt.double(), bias.double(), stride)
        y = torch.relu(pre) if act == 1 else T.leaky_relu(pre, 0.2)
        assert rel_err(o32, y) < 1e-5
        assert rel_err(out.float(), y) < 2e-5
        dy = torch.randn((n, oh, ow, cout), generator=g)
        dyd = dy.cuda()
        dx = torch.zeros((n, h, w, cin), device="cuda")
        dw = torch.zeros((k, k, cin, cout), device="cuda")
        ctx().conv2d_small_bwd_data(ptr(dyd), n, oh, ow, cout, ptr(wd), k, k, stride, h, w, cin, ptr(dx), stream())
        ctx().conv2d_small_bwd_filter(ptr(xd), n, h, w, cin, ptr(dyd), k, k, stride, cout, ptr(dw), stream())
        xv = x.double().requires_grad_(True)
        wv = wt.double().requires_grad_(True)
        yy = T.conv2d_same(xv, wv, None, stride)
        gx, gw = torch.autograd.grad(yy, [xv, wv], dy.double())
        assert rel_err(dx, gx) < 1e-5
        assert rel_err(dw, gw) < 1e-5


def test_crop_and_resize():
    _lib, SplitTensor, ptr, split_ref = _imports()
    g = torch.Generator().manual_seed(4)
    n, h, w, c, cs = 3, 32, 16, 64, 12
    img = torch.randn((n, h, w, c), generator=g)
    m = (torch.rand((n, h, w), generator=g) > 0.4).float()
    px = torch.tensor([[2, 1, 20, 12], [0, 0, 1, 1], [5, 3, 31, 15], [0, 0, 31, 15], [10, 2, 12, 9], [3, 3, 30, 6]],
                      dtype=torch.float32)
    boxes = torch.stack([px[:, 0] / h, px[:, 1] / w, px[:, 2] / h, px[:, 3] / w], dim=1)
    ind = torch.tensor([0, 1, 2, 0, 1, 2], dtype=torch.int32)
    nb = boxes.shape[0]
    ims = SplitTensor.from_float(img.cuda())
    out = SplitTensor(nb, cs, cs, c, zero=True)
    md, bd, idd = m.cuda(), boxes.cuda(), ind.cuda()
    ctx().crop_and_resize_fwd(ims.ref(), ptr(md), ptr(bd), ptr(idd), nb, out.ref(), stream())
    iv = (split_ref(img).double() * m[..., None].double()).requires_grad_(True)
    ref = T.crop_and_resize(iv, boxes.double(), ind, (cs, cs))
    assert rel_err(out.float(), ref) < 2e-5
    # gradient w.r.t. the un-masked image
    gy = torch.randn((nb, cs, cs, c), generator=g)
    gys = SplitTensor.from_float(gy.cuda())
    gimg = torch.zeros((n, h, w, c), device="cuda")
    ctx().crop_and_resize_bwd(gys.ref(), ptr(md), ptr(bd), ptr(idd), nb, ptr(gimg), n, h, w, c, stream())
    gi, = torch.autograd.grad(ref, iv, split_ref(gy).double())
    gi = gi * m[..., None].double()
    assert rel_err(gimg, gi) < 2e-5


def test_crop_and_resize_bwd_gather_many_boxes():
    """Gather form of CropAndResizeGradImage: more boxes on one image than one shared-memory list holds (256), boxes
    that up- and down-sample, degenerate (invisible-part) boxes [0,0,1,1] px and boxes reaching outside the image
    (extrapolated samples carry no gradient); against autograd through the float64 oracle."""
    _lib, SplitTensor, ptr, split_ref = _imports()
    g = torch.Generator().manual_seed(41)
    n, h, w, c, cs = 3, 24, 16, 16, 5
    nb = 700
    # fractional box corners: with integer corners whole rows of samples land EXACTLY on the image border (in_y == 0 or
    # H-1), where one rounding of in_y -- float32 here, float64 in the oracle -- decides valid vs extrapolated
    y1 = torch.rand((nb,), generator=g) * h - 2
    x1 = torch.rand((nb,), generator=g) * w - 2
    y2 = y1 + torch.rand((nb,), generator=g) * h
    x2 = x1 + torch.rand((nb,), generator=g) * w
    px = torch.stack([y1, x1, y2, x2], dim=1)
    px[::17] = torch.tensor([0.0, 0.0, 1.0, 1.0])
    px[5] = torch.tensor([18.3, 12.1, 4.2, 1.4])        # flipped box (y2 < y1, x2 < x1)
    px[6] = torch.tensor([7.3, 2.1, 7.3, 11.4])         # zero height: every crop row samples the same image row
    boxes = torch.stack([px[:, 0] / h, px[:, 1] / w, px[:, 2] / h, px[:, 3] / w], dim=1)
    ind = torch.randint(0, n, (nb,), generator=g).to(torch.int32)
    ind[:400] = 1                                   # > 256 boxes on image 1
    m = (torch.rand((n, h, w), generator=g) > 0.3).float()
    iv = torch.zeros((n, h, w, c), dtype=torch.float64, requires_grad=True)
    ref = T.crop_and_resize(iv, boxes.double(), ind, (cs, cs))
    gy = torch.randn((nb, cs, cs, c), generator=g)
    gys = SplitTensor.from_float(gy.cuda())
    gimg = torch.full((n, h, w, c), 7.0, device="cuda")          # overwritten, not accumulated into
    md, bd, idd = m.cuda(), boxes.cuda(), ind.cuda()          # (kept alive: the kernel reads them after this line)
    ctx().crop_and_resize_bwd(gys.ref(), ptr(md), ptr(bd), ptr(idd), nb, ptr(gimg), n, h, w, c, stream())
    gi, = torch.autograd.grad(ref, iv, split_ref(gy).double())
    gi = gi * m[..., None].double()
    assert rel_err(gimg, gi) < 2e-5
    again = torch.zeros_like(gimg)
    ctx().crop_and_resize_bwd(gys.ref(), ptr(md), ptr(bd), ptr(idd), nb, ptr(again), n, h, w, c, stream())
    assert torch.equal(gimg, again)                 # fixed summation order: bit-identical run to run


def test_pose_patch_matches_rasterised_maps():
    """dpig_pose_patch: the 3x3 SAME patches of the inflated keypoint maps (utils.py:259-318) built straight from the
    keypoints == the oracle's rasterised maps, zero-padded and unfolded (bit-exact: the values are -1 / +1 / 0)."""
    _lib, SplitTensor, ptr, split_ref = _imports()
    rng = np.random.default_rng(9)
    n, k, h, w = 3, 18, 32, 16
    rcv = np.zeros((n, k, 3), np.float32)
    rcv[:, :, 0] = rng.uniform(0, h - 1, size=(n, k))
    rcv[:, :, 1] = rng.uniform(0, w - 1, size=(n, k))
    rcv[:, :, 2] = (rng.uniform(size=(n, k)) < 0.8)
    rcv[0, 0] = [0.0, 0.0, 1.0]
    rcv[0, 1] = [h - 1, w - 1, 1.0]
    rd = torch.from_numpy(rcv).cuda()
    out = SplitTensor(n, h, w, 192)
    out.buf.fill_(3.0)                                   # every element is written, pad channels included
    ctx().pose_patch(ptr(rd), n, k, h, w, 4, 3, 3, out.ref(), stream())
    torch.cuda.synchronize()
    maps = T.pose_rasterize(torch.from_numpy(rcv).double(), h, w)            # [n,h,w,18] in {-1,+1}
    pad = torch.nn.functional.pad(maps, (0, 0, 1, 1, 1, 1))                   # zero padding = the conv's SAME padding
    ref = torch.cat([pad[:, i:i + h, j:j + w, :] for i in range(3) for j in range(3)], dim=-1)
    assert torch.equal(out.hi.float().cpu()[..., :162].double(), ref)
    assert float(out.hi.float().abs().cpu()[..., 162:].max()) == 0.0 and float(out.lo.float().abs().max()) == 0.0


def test_kernels_match_tf_known_answers():
    """The CUDA kernels against TensorFlow's own unit-test vectors (tests/golden/tf_known_answers.py: conv_ops_test SAME
    cases incl. the bottom/right-padded stride-3 one, crop_and_resize_op_test bilinear cases).  Small integers: exact."""
    import ctypes as C
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import tf_known_answers as KA
    _lib, SplitTensor, ptr, split_ref = _imports()
    for case in KA.CONV2D:
        name, _, _, stride, padding, expected = case
        if padding != "SAME" or stride > 2:        # the kernels implement SAME, stride 1 / 2 (all the reference uses)
            continue
        x, w = KA.conv2d_inputs(case)
        xs = padded_split(torch.tensor(x, dtype=torch.float32))
        wt = torch.zeros((w.shape[0], w.shape[1], xs.c, w.shape[3]))
        wt[:, :, :w.shape[2], :] = torch.tensor(w, dtype=torch.float32)
        f, _ = pack_weights(wt.cuda().contiguous(), cin_pad=xs.c)
        n, h, wd, _ = x.shape
        oh, ow, co = -(-h // stride), -(-wd // stride), w.shape[3]
        out32 = torch.zeros((n, oh, ow, co), device="cuda")
        ep = _lib.ConvEpilogue()
        ep.act = 0
        ep.out_f32 = out32.data_ptr()
        ep.out_f32_pix_stride = co
        ep.upsample = 1
        ctx().conv2d_fwd(xs.ref(), ptr(f[0]), ptr(f[1]), w.shape[0], w.shape[1], stride, co, C.byref(ep), stream())
        torch.cuda.synchronize()
        assert out32.cpu().reshape(-1).tolist() == expected, name
    for name, img, boxes, ind, crop, ev, expected in KA.CROP_AND_RESIZE:
        image = torch.tensor(img, dtype=torch.float32)[None, :, :, None].repeat(1, 1, 1, 8)     # 8 identical channels
        ims = SplitTensor.from_float(image.cuda())
        bd = torch.tensor(boxes, dtype=torch.float32).cuda()
        idd = torch.tensor(ind, dtype=torch.int32).cuda()
        out = SplitTensor(len(boxes), crop[0], crop[1], 8, zero=True)
        ctx().crop_and_resize_fwd(ims.ref(), None, ptr(bd), ptr(idd), len(boxes), out.ref(), stream())
        torch.cuda.synchronize()
        got = out.float().cpu()
        want = torch.tensor(expected, dtype=torch.float32)
        for ch in (0, 7):
            assert torch.allclose(got[..., ch], want, atol=1e-6, rtol=0), (name, got[..., ch])


def test_linear():
    _lib, SplitTensor, ptr, _ = _imports()
    g = torch.Generator().manual_seed(5)
    for (m, k, n, act) in [(14, 5760, 32, 0), (4, 20480, 128, 0), (4, 64, 4096, 0), (8, 512, 512, 1)]:
        x = torch.randn((m, k), generator=g)
        w = torch.randn((k, n), generator=g) / np.sqrt(k)
        b = torch.randn((n,), generator=g)
        xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
        y = torch.zeros((m, n), device="cuda")
        ctx().linear_fwd(ptr(xd), ptr(wd), ptr(bd), ptr(y), m, k, n, act, 0.2, stream())
        ref = x.double() @ w.double() + b.double()
        if act == 1:
            ref = torch.relu(ref)
        assert rel_err(y, ref) < 1e-5
        dy = torch.randn((m, n), generator=g)
        dyd = dy.cuda()
        dx = torch.zeros((m, k), device="cuda")
        dw = torch.zeros((k, n), device="cuda")
        db = torch.zeros((n,), device="cuda")
        ctx().linear_bwd(ptr(xd), ptr(wd), ptr(dyd), ptr(dx), ptr(dw), ptr(db), m, k, n, stream())
        assert rel_err(dx, dy.double() @ w.double().T) < 1e-5
        assert rel_err(dw, x.double().T @ dy.double()) < 1e-5
        assert rel_err(db, dy.double().sum(0)) < 1e-5


@pytest.mark.parametrize("mode", ["batch", "layer"])
def test_norm_act(mode):
    _lib, SplitTensor, ptr, split_ref = _imports()
    g = torch.Generator().manual_seed(6)
    n, h, w, c = 4, 8, 4, 128
    x = torch.randn((n, h, w, c), generator=g) * 1.5 + 0.3
    scale = 1 + 0.1 * torch.randn((c,), generator=g)
    offset = 0.1 * torch.randn((c,), generator=g)
    md = {"layer": 0, "batch": 1}[mode]
    groups = n if mode == "layer" else c
    count = float(h * w * c) if mode == "layer" else float(n * h * w)
    xd, sd, od = x.cuda(), scale.cuda(), offset.cuda()
    sums = torch.zeros((2, groups), dtype=torch.float64, device="cuda")
    stats = torch.zeros((2, groups), device="cuda")
    out = SplitTensor(n, h, w, c, zero=True)
    mask = torch.zeros((n * h * w, c // 32), dtype=torch.int32, device="cuda")
    ctx().norm_stats(ptr(xd), n, h, w, c, md, ptr(sums), stream())
    ctx().norm_act_fwd(ptr(xd), n, h, w, c, md, 1e-5, ptr(sums), count, ptr(sd), ptr(od), 2, 0.2, ptr(stats),
                       out.ref(), ptr(mask), stream())
    xv = x.double().requires_grad_(True)
    sv = scale.double().requires_grad_(True)
    ov = offset.double().requires_grad_(True)
    fn = T.layernorm if mode == "layer" else T.batchnorm_train
    y = T.leaky_relu(fn(xv, sv, ov), 0.2)
    assert rel_err(out.float(), y) < 2e-5
    dy = torch.randn((n, h, w, c), generator=g)
    dys = SplitTensor.from_float(dy.cuda())
    red = torch.zeros((2, groups), dtype=torch.float64, device="cuda")
    dsc = torch.zeros((c,), device="cuda")
    dof = torch.zeros((c,), device="cuda")
    dx = SplitTensor(n, h, w, c, zero=True)
    ctx().norm_act_bwd_reduce(dys.ref(), ptr(xd), ptr(stats), ptr(mask), 0.2, md, ptr(sd), ptr(red), ptr(dsc),
                              ptr(dof), stream())
    ctx().norm_act_bwd_apply(dys.ref(), ptr(xd), ptr(stats), ptr(mask), 0.2, md, ptr(sd), ptr(red), count,
                             dx.ref(), stream())
    gx, gs, go = torch.autograd.grad(y, [xv, sv, ov], split_ref(dy).double())
    assert rel_err(dx.float(), gx) < 5e-5
    assert rel_err(dsc, gs) < 1e-5
    assert rel_err(dof, go) < 1e-5


def test_elementwise_and_losses():
    _lib, SplitTensor, ptr, split_ref = _imports()
    g = torch.Generator().manual_seed(7)
    n, h, w, c = 2, 8, 4, 64
    a = torch.randn((n, 2 * h, 2 * w, c), generator=g)
    b = torch.randn((n, 2 * h, 2 * w, c), generator=g)
    as_, bs = SplitTensor.from_float(a.cuda()), SplitTensor.from_float(b.cuda())
    mbits = torch.randint(0, 2, (n, h, w, c), generator=g)
    mw = np.zeros((n, h, w, c // 32), np.uint32)
    for ch in range(c):
        mw[..., ch // 32] |= mbits.numpy().astype(np.uint32)[..., ch] << np.uint32(ch % 32)
    md = torch.from_numpy(mw.view(np.int32)).cuda()
    out = SplitTensor(n, h, w, c, zero=True)
    ctx().ew_combine(out.ref(), as_.ref(), bs.ref(), None, None, 0, ptr(md), 0.0, 1, stream())
    s = split_ref(a).double() + split_ref(b).double()
    pooled = s.reshape(n, h, 2, w, 2, c).sum(dim=(2, 4)) * mbits.double()
    assert rel_err(out.float(), pooled) < 2e-5
    # L1 + GAN losses
    G = torch.randn((n, h, w, 3), generator=g)
    X = torch.randn((n, h, w, 3), generator=g)
    Gd, Xd = G.cuda(), X.cuda()
    o = torch.zeros((2,), device="cuda")
    dG = torch.zeros_like(Gd)
    ctx().loss_l1(ptr(Gd), ptr(Xd), G.numel(), 20.0, ptr(o), ptr(dG), stream())
    Gv = G.double().requires_grad_(True)
    l1 = (Gv - X.double()).abs().mean()
    gl, = torch.autograd.grad(20.0 * l1, Gv)
    assert abs(float(o[0]) - float(l1)) < 1e-6
    assert rel_err(dG, gl) < 1e-5
    for mode, mid in [("dcgan", 0), ("wgan", 1), ("lsgan", 3)]:
        dr = torch.randn((6,), generator=g)
        df = torch.randn((6,), generator=g)
        drd, dfd = dr.cuda(), df.cuda()
        o = torch.zeros((2,), device="cuda")
        g1, g2, g3 = (torch.zeros((6,), device="cuda") for _ in range(3))
        ctx().loss_gan(mid, ptr(drd), ptr(dfd), 6, ptr(o), ptr(g1), ptr(g2), ptr(g3), stream())
        rv = dr.double().requires_grad_(True)
        fv = df.double().requires_grad_(True)
        gl_, dl_ = T.gan_loss(mode, rv, fv)
        assert abs(float(o[0]) - float(gl_)) < 1e-5 and abs(float(o[1]) - float(dl_)) < 1e-5
        e1, = torch.autograd.grad(gl_, fv, retain_graph=True)
        e2, e3 = torch.autograd.grad(dl_, [rv, fv])
        assert rel_err(g1, e1) < 1e-5 and rel_err(g2, e2) < 1e-5 and rel_err(g3, e3) < 1e-5


def test_adam_and_pose():
    _lib, SplitTensor, ptr, _ = _imports()
    g = torch.Generator().manual_seed(8)
    p = torch.randn((1000,), generator=g)
    pd = p.cuda()
    m = torch.zeros_like(pd)
    v = torch.zeros_like(pd)
    pr, mr, vr = p.double().clone(), torch.zeros(1000, dtype=torch.float64), torch.zeros(1000, dtype=torch.float64)
    for t in (1, 2, 3):
        gr = torch.randn((1000,), generator=g)
        gd = gr.cuda()
        ctx().adam_step(ptr(pd), ptr(gd), ptr(m), ptr(v), 1000, 2e-5, 0.5, 0.999, 1e-8, t, 1.0, stream())
        T.adam_step(pr, gr.double(), mr, vr, 2e-5, t)
    # fp32 parameters of magnitude ~3 quantise at 2.4e-7; the three updates are 2e-5 each
    assert float((pd.cpu().double() - pr).abs().max()) < 5e-7
    # graph-replayable form: step size from device memory; vector body + scalar tail (1003 = 4*250 + 3) and an
    # unaligned view (scalar loop only) give the bits of the scalar-argument entry
    import math
    for off in (0, 1):
        q = torch.randn((1004,), generator=g)[off:off + 1003]
        gq = torch.randn((1004,), generator=g)[off:off + 1003]
        a = [t_.cuda() for t_ in (q.clone(), torch.zeros(1003), torch.zeros(1003))]
        bfull = [torch.zeros(1004, device="cuda") for _ in range(3)]
        bv = [t_[off:off + 1003] for t_ in bfull]
        bv[0].copy_(q)
        gfull = torch.zeros(1004, device="cuda")
        gv = gfull[off:off + 1003]
        gv.copy_(gq)
        ga = gq.cuda()
        lr_t = torch.tensor([float(np.float32(2e-5)) * math.sqrt(1.0 - float(np.float32(0.999)) ** 2) / (1.0 - 0.5 ** 2)], dtype=torch.float32).cuda()
        ctx().adam_step(ptr(a[0]), ptr(ga), ptr(a[1]), ptr(a[2]), 1003, 2e-5, 0.5, 0.999, 1e-8, 2, 0.5, stream())
        ctx().adam_step_dev(ptr(bv[0]), ptr(gv), ptr(bv[1]), ptr(bv[2]), 1003, ptr(lr_t), 0.5, 0.999, 1e-8, 0.5, stream())
        torch.cuda.synchronize()
        for x_, y_ in zip(a, bv):
            assert torch.equal(x_, y_)
    r1 = torch.randn((777,), generator=g).cuda()
    r2 = r1.clone()
    gr_ = torch.randn((777,), generator=g).cuda()
    ms1, ms2 = torch.ones(777, device="cuda"), torch.ones(777, device="cuda")
    lr_d = torch.tensor([5e-5], dtype=torch.float32).cuda()
    ctx().rmsprop_step(ptr(r1), ptr(gr_), ptr(ms1), 777, 5e-5, 0.9, 1e-10, 1.0, 0.01, stream())
    ctx().rmsprop_step_dev(ptr(r2), ptr(gr_), ptr(ms2), 777, ptr(lr_d), 0.9, 1e-10, 1.0, 0.01, stream())
    torch.cuda.synchronize()
    assert torch.equal(r1, r2) and torch.equal(ms1, ms2)
    assert rel_err(pd.cpu().double() - p.double(), pr - p.double()) < 2e-2
    from dpig_b200 import synth
    batch = synth.make_batch(3, 128, 64, seed=5)
    rcv = torch.from_numpy(batch["pose_rcv"])
    rd = rcv.cuda()
    o32 = torch.zeros((3, 128, 64, 18), device="cuda")
    ctx().pose_rasterize(ptr(rd), 3, 18, 128, 64, 4, None, ptr(o32), stream())
    assert torch.equal(o32.cpu(), T.pose_rasterize(rcv, 128, 64, 4))


if __name__ == "__main__":
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    torch.cuda.init()
    print("device:", torch.cuda.get_device_name(0), flush=True)
    groups = [("conv2d_fwd", run_conv_fwd, CONV_FWD_CASES), ("conv2d_bwd_data", run_conv_bwd_data, CONV_BWD_DATA_CASES),
              ("conv2d_bwd_filter", run_conv_bwd_filter, CONV_BWD_FILTER_CASES)]
    for name, fn, cases in groups:
        if flt and flt not in name:
            continue
        for case in cases:
            try:
                info = fn(**case)
                print("%-18s %s -> %s" % (name, case, info), flush=True)
            except Exception as e:  # noqa: BLE001
                print("%-18s %s -> EXC %r" % (name, case, e), flush=True)
                try:
                    torch.cuda.synchronize()
                except Exception as e2:  # noqa: BLE001
                    print("CUDA context is dead:", e2, flush=True)
                    sys.exit(3)
    for tname in ["test_conv_small", "test_crop_and_resize", "test_linear", "test_elementwise_and_losses",
                  "test_adam_and_pose"]:
        if flt and flt not in tname:
            continue
        try:
            globals()[tname]()
            print(tname, "OK", flush=True)
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            print(tname, "FAILED", repr(e)[:300], flush=True)
    for mode in ("batch", "layer"):
        if flt and flt not in "test_norm_act":
            continue
        try:
            test_norm_act(mode)
            print("test_norm_act", mode, "OK", flush=True)
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            print("test_norm_act", mode, "FAILED", repr(e)[:300], flush=True)


# ------------------------------------------------------------------------------------------ 3-channel ends as 1x1 GEMMs
@pytest.mark.parametrize("case", [dict(n=2, h=16, w=8, k=3, stride=1, cout=128), dict(n=3, h=32, w=16, k=5, stride=2, cout=64),
                                  dict(n=2, h=7, w=5, k=3, stride=2, cout=32)])
def test_patch_form_of_3_channel_input_conv(case):
    """cin = 3 (models.py:396, wgan_gp.py:419): im2col -> 1x1 conv on the SAME HWIO filter, and its filter gradient,
    against the float64 oracle convolution."""
    import ctypes as C
    _lib, SplitTensor, ptr, split_ref = _imports()
    n, h, w, k, stride, cout = (case[x] for x in ("n", "h", "w", "k", "stride", "cout"))
    g = torch.Generator().manual_seed(5)
    x = torch.randn((n, h, w, 3), generator=g)
    wt = torch.randn((k, k, 3, cout), generator=g) / np.sqrt(k * k * 3)
    bias = torch.randn((cout,), generator=g) * 0.1
    oh, ow = -(-h // stride), -(-w // stride)
    kp = (k * k * 3 + 7) // 8 * 8
    xs = padded_split(x)                                   # [n,h,w,8]
    patches = SplitTensor(n, oh, ow, kp, zero=False)
    ctx().im2col_small(xs.ref(), 3, k, k, stride, 0, patches.ref(), stream())
    # the filter read as a 1x1 conv with k*k*3 input channels
    w1 = wt.reshape(1, 1, k * k * 3, cout).cuda().contiguous()
    f, _ = pack_weights(w1, cin_pad=kp)
    out = SplitTensor(n, oh, ow, cout, zero=True)
    ep = _lib.ConvEpilogue()
    bd = bias.cuda()
    ep.bias = bd.data_ptr()
    ep.act = 1
    ep.upsample = 1
    ep.out = C.pointer(out.struct())
    ctx().conv2d_fwd(patches.ref(), ptr(f[0]), ptr(f[1]), 1, 1, 1, cout, C.byref(ep), stream())
    dy = torch.randn((n, oh, ow, cout), generator=g)
    dys = padded_split(dy)
    dw = torch.zeros((k, k, 3, cout), device="cuda")
    ctx().conv2d_bwd_filter(patches.ref(), dys.ref(), 1, 1, 1, k * k * 3, cout, ptr(dw), stream())
    torch.cuda.synchronize()
    wv = split_ref(wt).double().requires_grad_(True)
    y = T.conv2d_same(split_ref(x).double(), wv, bias.double(), stride)
    assert rel_err(out.float(), torch.relu(y).detach()) < 2e-5
    gw, = torch.autograd.grad(y, wv, split_ref(dy).double())
    assert rel_err(dw, gw) < 2e-5
    assert float(patches.float()[..., k * k * 3:].abs().max()) == 0.0 if kp > k * k * 3 else True


def test_patch_form_of_3_channel_output_conv():
    """cout = 3 (models.py:573): 1x1 conv to 27 tap-channels + col2im gather; data / filter gradients through the
    transposed patches of dy -- against the float64 oracle convolution and its autograd."""
    import ctypes as C
    _lib, SplitTensor, ptr, split_ref = _imports()
    n, h, w, cin, k, cout = 2, 16, 8, 256, 3, 3
    g = torch.Generator().manual_seed(6)
    x = torch.randn((n, h, w, cin), generator=g)
    wt = torch.randn((k, k, cin, cout), generator=g) / np.sqrt(k * k * cin)
    bias = torch.randn((cout,), generator=g) * 0.1
    dy = torch.randn((n, h, w, cout), generator=g)
    taps, J = k * k, k * k * cout
    xs = SplitTensor.from_float(x.cuda())
    wd = wt.cuda().contiguous()
    # ---- forward
    wf = torch.zeros((cin, J), device="cuda")
    ctx().permute_taps(ptr(wd), ptr(wf), taps, cin, cout, 0, stream())
    f, _ = pack_weights(wf.reshape(1, 1, cin, J).contiguous())
    ybuf = torch.zeros((n, h, w, 32), device="cuda")
    ep = _lib.ConvEpilogue()
    ep.act = 0
    ep.upsample = 1
    ep.out_f32 = ybuf.data_ptr()
    ep.out_f32_pix_stride = 32
    ctx().conv2d_fwd(xs.ref(), ptr(f[0]), ptr(f[1]), 1, 1, 1, J, C.byref(ep), stream())
    out32 = torch.zeros((n, h, w, cout), device="cuda")
    out8 = SplitTensor(n, h, w, 8, zero=False)
    bd = bias.cuda()
    ctx().col2im_small(ptr(ybuf), 32, n, h, w, k, k, cout, ptr(bd), ptr(out32), cout, out8.ref(), stream())
    # ---- gradients
    dys = padded_split(dy)
    dp = SplitTensor(n, h, w, 32, zero=False)
    ctx().im2col_small(dys.ref(), cout, k, k, 1, 1, dp.ref(), stream())
    wdg = torch.zeros((J, cin), device="cuda")
    ctx().permute_taps(ptr(wd), ptr(wdg), taps, cin, cout, 1, stream())
    fd, _ = pack_weights(wdg.reshape(1, 1, J, cin).contiguous(), cin_pad=32)
    dx = SplitTensor(n, h, w, cin, zero=True)
    ep2 = _lib.ConvEpilogue()
    ep2.act = 0
    ep2.upsample = 1
    ep2.out = C.pointer(dx.struct())
    ctx().conv2d_fwd(dp.ref(), ptr(fd[0]), ptr(fd[1]), 1, 1, 1, cin, C.byref(ep2), stream())
    dwv = torch.zeros((cin, J), device="cuda")
    ctx().conv2d_bwd_filter(xs.ref(), dp.ref(), 1, 1, 1, cin, J, ptr(dwv), stream())
    dw = torch.full((k, k, cin, cout), 0.5, device="cuda")            # mode 2 accumulates
    ctx().permute_taps(ptr(dwv), ptr(dw), taps, cin, cout, 2, stream())
    torch.cuda.synchronize()
    xv = split_ref(x).double().requires_grad_(True)
    wv = split_ref(wt).double().requires_grad_(True)
    y = T.conv2d_same(xv, wv, bias.double(), 1)
    gx, gw = torch.autograd.grad(y, [xv, wv], split_ref(dy).double())
    assert rel_err(out32, y.detach()) < 2e-5
    assert rel_err(out8.float()[..., :3], y.detach()) < 2e-5 and float(out8.float()[..., 3:].abs().max()) == 0.0
    assert rel_err(dx.float(), gx) < 2e-5
    assert rel_err(dw - 0.5, gw) < 2e-5
