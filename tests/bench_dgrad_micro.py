"""Dev tool (not a test): times single data-gradient launches with the epilogue features switched on one by one.
   python tests/bench_dgrad_micro.py            all shapes / variants
   MICRO_SHAPE=0 MICRO_VARIANT=full MICRO_ITERS=1 python tests/bench_dgrad_micro.py      (one launch, for ncu)"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dpig_b200  # noqa: E402
from dpig_b200 import _lib  # noqa: E402
from dpig_b200.tensor import SplitTensor, ptr  # noqa: E402

# n, in_h, in_w, cin (dx channels), cout (dy channels), k, stride
SHAPES = [(448, 48, 48, 128, 256, 3, 2), (64, 64, 32, 128, 256, 3, 2), (64, 64, 32, 512, 128, 1, 1),
          (448, 48, 48, 128, 128, 3, 1), (64, 128, 64, 256, 256, 3, 1)]
VARIANTS = ["out", "masked", "masked+colsum", "out+masked", "out+masked+colsum", "full"]   # + "wgrad": the filter gradient


def run(ctx, shape, variant, iters=10):
    n, h, w, cin, cout, k, stride = shape
    s = torch.cuda.current_stream().cuda_stream
    oh, ow = -(-h // stride), -(-w // stride)
    dy = SplitTensor(n, oh, ow, cout, zero=True)
    dy.buf.normal_(0, 1)
    wb = torch.randn((2, k * k, cin, cout), device="cuda").to(torch.bfloat16)
    out = SplitTensor(n, h, w, cin)
    out2 = SplitTensor(n, h, w, cin)
    add = SplitTensor(n, h, w, cin, zero=True)
    mask = torch.full((n * h * w, cin // 32), 0x55555555, dtype=torch.int32, device="cuda")
    colsum = torch.zeros(cin, device="cuda")
    ep = _lib.ConvEpilogue()
    ep.act = 0
    ep.upsample = 1
    if variant.startswith("out") or variant == "full":
        ep.out = C.pointer(out.struct())
    if "masked" in variant or variant == "full":
        ep.out_masked = C.pointer(out2.struct())
        ep.mask_in = mask.data_ptr()
    if "colsum" in variant or variant == "full":
        ep.colsum_masked = colsum.data_ptr()
    if variant == "full":
        ep.addend = C.pointer(add.struct())
    call = lambda: ctx.conv2d_bwd_data(dy.ref(), ptr(wb[0]), ptr(wb[1]), k, k, stride, h, w, cin, C.byref(ep), s)
    if variant == "wgrad":
        x = SplitTensor(n, h, w, cin, zero=True)
        x.buf.normal_(0, 1)
        dw = torch.zeros((k, k, cin, cout), device="cuda")
        call = lambda: ctx.conv2d_bwd_filter(x.ref(), dy.ref(), k, k, stride, cin, cout, ptr(dw), s)
    for _ in range(2 if iters > 1 else 0):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * n * oh * ow * cout * k * k * cin
    return ms, fl / ms / 1e9


if __name__ == "__main__":
    shapes = SHAPES
    if os.environ.get("MICRO_SHAPE"):
        shapes = [SHAPES[int(i)] for i in os.environ["MICRO_SHAPE"].split(",")]
    variants = os.environ.get("MICRO_VARIANT", ",".join(VARIANTS)).split(",")
    iters = int(os.environ.get("MICRO_ITERS", "10"))
    ctx = dpig_b200.Context(0)
    for sh in shapes:
        for v in variants:
            ms, tf = run(ctx, sh, v, iters)
            print("dgrad dx %4dx%3dx%3dx%4d <- dy c=%4d k%ds%d  %-18s %8.3f ms %8.1f TFLOP/s" %
                  (sh[0], sh[1], sh[2], sh[3], sh[4], sh[5], sh[6], v, ms, tf), flush=True)
