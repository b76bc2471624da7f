"""Bring-up aid (not a test): compares the engine's intermediate activations / activation gradients of the
G step with the float64 oracle's, tensor by tensor.   python tests/debug_grads_gpu.py [small|full] [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import nets  # noqa: E402
import test_stage1_gpu as t  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30)), float(b.abs().max())


def main():
    small = (sys.argv[1] if len(sys.argv) > 1 else "small") == "small"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    eng, cfg, p, ob = t._setup(small, batch)
    eng.g_grads()
    torch.cuda.synchronize()
    taps = {}
    out = nets.stage1_forward(p, cfg, ob, "dcgan", taps=taps)
    out["g_loss"].backward()
    rn = cfg.repeat_num
    rows = []

    def add(name, act, grad=None, masked=False):
        tp = taps[name]
        if act is not None:
            rows.append((name, "act",) + rel(act.float() if hasattr(act, "buf") else act, tp.detach()))
        if grad is not None:
            g = grad.float() if hasattr(grad, "buf") else grad
            rows.append((name, "grad",) + rel(g, tp.grad))

    add("xs", eng.xs, eng.g_xs)
    add("x_bg", eng.x_bg, eng.bg_pyr.g_in)
    add("rois", eng.rois, eng.roi_pyr.g_in)
    for tag, pyr in (("roi", eng.roi_pyr), ("bg", eng.bg_pyr), ("genc", eng.genc)):
        for i in range(rn):
            add("%s_a%d" % (tag, i), pyr.a[i], None)
            add("%s_y%d" % (tag, i), pyr.y[i], pyr.g_y[i])
            if i < rn - 1:
                add("%s_x%d" % (tag, i + 1), pyr.x_in[i + 1], None)
    add("g0", eng.g0, None)
    for i in range(rn):
        add("cat%d" % i, eng.cat[i], eng.g_cat[i])
        add("dec_a%d" % i, eng.dec_a[i], None)
        add("dec_y%d" % i, eng.dec_y[i], eng.dec_gy[i])
    add("G", eng.G, eng.g_G)
    for r in rows:
        print("%-10s %-5s relerr %.3e   (ref max %.3e)" % r)
    # masked gradients: compare gb = g_y * (conv2_pre > 0) using the oracle's own mask
    for tag, pyr in (("roi", eng.roi_pyr), ("bg", eng.bg_pyr), ("genc", eng.genc)):
        for i in range(rn):
            y = taps["%s_y%d" % (tag, i)]
            res = taps["%s_x%d" % (tag, i)] if i > 0 else {"roi": taps["rois"], "bg": taps["x_bg"], "genc": taps["g0"]}[tag]
            b = (y - res).detach()
            ref = y.grad * (b > 0)
            print("%-10s gb    relerr %.3e   (ref max %.3e)  zeros-in-b: %d of %d" % (
                ("%s_%d" % (tag, i),) + rel(pyr.gb[i].float(), ref) + (int((b == 0).sum()), b.numel())))


if __name__ == "__main__":
    main()
