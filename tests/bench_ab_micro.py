"""Dev tool (not a test): same-process A/B of a tuning switch (dpig_ctx_set_option) on the layer shapes of a Stage-I step.
The two settings alternate (A B A B ...) per shape, so the power-capped clock drift of a long run cancels out.
   python tests/bench_ab_micro.py epi_specialise 0 1 [fwd|dgrad|wgrad ...]"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_conv_micro as CM  # noqa: E402
import bench_dgrad_micro as DM  # noqa: E402
import dpig_b200  # noqa: E402

# n, h, w, cin, cout, k
FWD = [(64, 128, 64, 128, 128, 3), (448, 48, 48, 128, 128, 3), (64, 128, 64, 256, 256, 3), (64, 32, 16, 384, 384, 3),
       (64, 8, 4, 640, 640, 3), (64, 128, 64, 32, 128, 1), (64, 64, 32, 512, 128, 1), (64, 128, 64, 192, 128, 1)]
# n, in_h, in_w, cin (dx), cout (dy), k, stride
DGRAD = [(448, 48, 48, 128, 256, 3, 2), (64, 64, 32, 128, 256, 3, 2), (64, 64, 32, 512, 128, 1, 1),
         (448, 48, 48, 128, 128, 3, 1), (64, 128, 64, 256, 256, 3, 1), (64, 32, 16, 384, 384, 3, 1)]
WGRAD = [(64, 128, 64, 128, 128, 3, 1), (448, 48, 48, 128, 128, 3, 1), (64, 128, 64, 256, 256, 3, 1),
         (64, 32, 16, 384, 384, 3, 1), (448, 12, 12, 384, 384, 3, 1), (64, 8, 4, 640, 640, 3, 1), (64, 16, 8, 512, 640, 3, 2),
         (64, 16, 8, 1024, 1024, 3, 1), (448, 3, 3, 640, 640, 3, 1)]


def main():
    opt, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    kinds = sys.argv[4:] or ["fwd", "dgrad", "wgrad"]
    rounds = int(os.environ.get("AB_ROUNDS", "3"))
    ctx = dpig_b200.Context(0)
    jobs = []
    if "fwd" in kinds:
        for shp in FWD:
            for mode in ("full", "nores"):
                jobs.append(("fwd %4dx%3dx%3d %4d->%4d k%d %-5s" % (shp + (mode,)),
                             lambda shp=shp, mode=mode: CM.run(ctx, *shp[:5], k=shp[5], mode=mode, iters=20)))
    if "dgrad" in kinds:
        for shp in DGRAD:
            for var in ("out", "out+masked+colsum", "full"):
                jobs.append(("dgrad %4dx%3dx%3d dx%4d<-dy%4d k%ds%d %-17s" % (shp + (var,)),
                             lambda shp=shp, var=var: DM.run(ctx, shp, var, iters=10)))
    if "wgrad" in kinds:
        for shp in WGRAD:
            jobs.append(("wgrad %4dx%3dx%3d %4d->%4d k%ds%d" % shp, lambda shp=shp: DM.run(ctx, shp, "wgrad", iters=10)))
    tot = {a: 0.0, b: 0.0}
    for name, fn in jobs:
        t = {a: [], b: []}
        for _ in range(rounds):
            for val in (a, b):
                ctx.set_option(opt, val)
                t[val].append(fn()[0])
        ma, mb = statistics.median(t[a]), statistics.median(t[b])
        tot[a] += ma
        tot[b] += mb
        print("%-62s %s=%d %8.4f ms   %s=%d %8.4f ms   %+6.1f %%" % (name, opt, a, ma, opt, b, mb, 100.0 * (mb - ma) / ma), flush=True)
    print("sum: %s=%d %.3f ms, %s=%d %.3f ms (%+.1f %%)" % (opt, a, tot[a], opt, b, tot[b], 100.0 * (tot[b] - tot[a]) / tot[a]))


if __name__ == "__main__":
    main()
