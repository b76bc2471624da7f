"""Dev tool (not a test): times single conv launches to separate mainloop from epilogue cost.
   python tests/bench_conv_micro.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dpig_b200  # noqa: E402
from dpig_b200 import _lib  # noqa: E402
from dpig_b200.tensor import SplitTensor, ptr  # noqa: E402


def run(ctx, n, h, w, cin, cout, k=3, mode="full", fast=False, iters=20):
    s = torch.cuda.current_stream().cuda_stream
    x = SplitTensor(n, h, w, cin, zero=True)
    x.buf.normal_(0, 1)
    wf = torch.randn((2, k * k, cout, cin), device="cuda").to(torch.bfloat16)
    bias = torch.zeros(cout, device="cuda")
    out = SplitTensor(n, h, w, cout)
    res = SplitTensor(n, h, w, cout, zero=True)
    mask = torch.zeros((n * h * w, cout // 32), dtype=torch.int32, device="cuda")
    ep = _lib.ConvEpilogue()
    ep.bias = bias.data_ptr()
    ep.act = 1
    ep.upsample = 1
    if mode in ("full", "nores"):
        ep.out = C.pointer(out.struct())
        ep.mask_out = mask.data_ptr()
    if mode == "full":
        ep.addend = C.pointer(res.struct())
    ctx.set_fast_mode(1 if fast else 0)
    for _ in range(3):
        ctx.conv2d_fwd(x.ref(), ptr(wf[0]), ptr(wf[1]), k, k, 1, cout, C.byref(ep), s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ctx.conv2d_fwd(x.ref(), ptr(wf[0]), ptr(wf[1]), k, k, 1, cout, C.byref(ep), s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 2.0 * n * h * w * cout * k * k * cin
    ctx.set_fast_mode(0)
    return ms, fl / ms / 1e9


if __name__ == "__main__":
    shapes = [(64, 128, 64, 128, 128), (64, 128, 64, 256, 256), (64, 64, 32, 512, 512), (448, 48, 48, 128, 128),
              (64, 32, 16, 640, 640)]
    variants = [("single", 0, {}), ("single-nomerge", 0, {"DPIG_CONV_MERGE": "0"}), ("pair", 2, {}),
                ("pair-nomerge", 2, {"DPIG_CONV_MERGE": "0"}), ("single-2stage", 0, {"DPIG_CONV_STAGES": "2"})]
    if len(sys.argv) > 1:
        variants = [v for v in variants if v[0] in sys.argv[1:]]
    if os.environ.get("MICRO_SHAPES"):  # e.g. MICRO_SHAPES=0,1 MICRO_MODES=none (profiling one launch under ncu)
        shapes = [shapes[int(i)] for i in os.environ["MICRO_SHAPES"].split(",")]
    if os.environ.get("MICRO_SHAPE"):   # explicit "n,h,w,cin,cout[,k]"
        shapes = [tuple(int(v) for v in os.environ["MICRO_SHAPE"].split(","))]
    modes = os.environ.get("MICRO_MODES", "full,nores,none").split(",")
    for name, tiling, env in variants:
        for k in ("DPIG_CONV_STAGES", "DPIG_CONV_MERGE"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ctx = dpig_b200.Context(0)
        ctx.set_pair_mode(tiling)
        for shp in shapes:
            n, h, w, cin, cout = shp[:5]
            kk = shp[5] if len(shp) > 5 else 3
            for mode in modes:
                ms, tf = run(ctx, n, h, w, cin, cout, k=kk, mode=mode, fast=False, iters=int(os.environ.get("MICRO_ITERS", "20")))
                print("%-14s %4dx%3dx%3d %4d->%4d  epilogue=%-5s %7.3f ms  %7.1f TFLOP/s (algorithmic)" % (
                    name, n, h, w, cin, cout, mode, ms, tf), flush=True)
