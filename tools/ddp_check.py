"""Run under torchrun with 2+ ranks on GPUs: the data-parallel step with the overlapped ID_AE all-reduce (default),
with one all-reduce after the backward pass (DPIG_OVERLAP=0) and with the NCCL exchanges captured into the step graphs
(DPIG_GRAPHS=2) must leave identical weights on every rank and agree with each other up to the fp32-atomic noise.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_check.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dpig_b200  # noqa: E402
from dpig_b200 import ddp, engine, synth  # noqa: E402


def run(dist, overlap, graphs):
    os.environ["DPIG_OVERLAP"] = overlap
    os.environ["DPIG_GRAPHS"] = graphs
    kw = dict(img_h=32, img_w=16, hidden=64, roi_size=12, d_dim=64)
    cfg = engine.NetConfig(**kw)
    ctx = dpig_b200.Context(dist.local_rank)
    eng = engine.Stage1Engine(ctx, cfg, 2, mode="dcgan", dist=dist, device="cuda:%d" % dist.local_rank)
    eng.load_params(engine.init_params(cfg, seed=99))
    eng.g_lr = eng.d_lr = 2e-4
    init = eng.get_params()
    for i in range(5):
        eng.set_batch(synth.make_batch(2, 32, 16, seed=700 + 10 * i + dist.rank))
        eng.g_step()
        eng.set_batch(synth.make_batch(2, 32, 16, seed=705 + 10 * i + dist.rank))
        eng.d_step()
    torch.cuda.synchronize()
    return eng, init


def main():
    dist = ddp.Dist()
    torch.cuda.set_device(dist.local_rank)
    results = {}
    for name, overlap, graphs in (("overlap", "1", "1"), ("plain", "0", "1"), ("graphs", "1", "2")):
        eng, init = run(dist, overlap, graphs)
        p = eng.get_params()
        flat = torch.cat([torch.as_tensor(v).reshape(-1) for v in p.values()]).cuda()
        ref = flat.clone()
        dist.broadcast(ref, 0)
        same = bool(torch.equal(flat, ref))        # every rank holds rank 0's weights bit for bit
        results[name] = (p, init, same, bool(eng._graphs), eng.overlap_comm)
    ok = all(r[2] for r in results.values())
    base_p, init = results["plain"][0], results["plain"][1]
    worst = 0.0
    for name in ("overlap", "graphs"):
        for k in ("ID_AE/G/Conv_3/weights", "Encoder/G_encoder/Conv_5/weights", "Discriminator.3.Filters"):
            d = np.median(np.abs(results[name][0][k].astype(np.float64) - base_p[k]))
            moved = np.median(np.abs(base_p[k].astype(np.float64) - init[k]))
            worst = max(worst, d / max(moved, 1e-12))
    ok = ok and worst < 0.5 and results["graphs"][3] and results["overlap"][4] and not results["plain"][4]
    if dist.rank == 0:
        print("ddp_check ranks_identical=%s worst_median_diff/moved=%.3f graphs_captured=%s -> %s" % (
            [r[2] for r in results.values()], worst, results["graphs"][3], "OK" if ok else "FAIL"), flush=True)
    dist.barrier()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
