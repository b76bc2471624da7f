#!/usr/bin/env python
"""bench.py -- images/sec of the hot path on N B200s (BASELINE.json's metric), one JSON line on rank 0.

    python bench.py --gpus N --steps K --warmup W [--workload market|df256|stage2|sample]
                                                  (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...          (CPU arm: the oracle port on the host cores)

Workloads (BASELINE.json configs; `market` is the headline the metric is quoted on):
    market  configs[1]  Stage-I --model=1 (reference trainer.py:336-347): one g_optim + one d_optim update per step,
                        128x64, batch 64 per GPU, dcgan loss (as shipped) or --mode wgan-gp
    df256   configs[3]  Stage-I DeepFashion --model=101 (trainer_256.py:31-134), 256x256, batch 32 per GPU
    stage2  configs[2]  Stage-II --model=3 (trainer.py:812-867): per factor (Fg, Bg) one g_optim + 5 x (d_optim + clip),
                        every critic call on a fresh batch through the frozen encoder; batch 32 per GPU (256 on 8 GPUs)
    sample  configs[4]  sampling --model=13 (tester.py:573-613): tester.generate at batch 512, one GPU

`value` = whole-job images/s with each step's inputs already resident in HBM (device-to-device batch swap inside the
timed region); `e2e` = the same steps through the reference-facing call surface -- trainer.train() / tester.generate() --
with pinned HOST batches: H2D of every batch and D2H of the step's result inside the timed region.  `roofline` = the
dominant kernel (conv_umma_kernel, forward + data-gradient convolutions) timed with CUDA events per launch.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "market": dict(metric="images/sec (G+D step) Market-1501 128x64", batch=64, hw=(128, 64), model=1,
                   text="Stage-I Fg/Bg/Pose reconstruction (--model=1, %(mode)s loss), Market-1501 128x64, batch=%(b)d "
                        "per GPU: 1 g_optim + disc_ITERS d_optim per step (1 for dcgan, 5 for wgan-gp), separate batches "
                        "(trainer.py:336-347)"),
    "df256": dict(metric="images/sec (G+D step) DeepFashion 256x256", batch=32, hw=(256, 256), model=101,
                  text="Stage-I DeepFashion 256x256 (--model=101, trainer_256.py path, dcgan loss), batch=%(b)d per GPU: "
                       "1 g_optim + 1 d_optim per step, separate batches"),
    "stage2": dict(metric="images/sec (Stage-II step) Market-1501 128x64", batch=32, hw=(128, 64), model=3,
                   text="Stage-II appearance-sampling GAN (--model=3, wgan + clip, trainer.py:812-867), batch=%(b)d per "
                        "GPU: per factor 1 g_optim + 5 d_optim, 10 fresh batches through the frozen encoder per step; "
                        "images/s counts ONE batch per step"),
    "sample": dict(metric="images/sec (sampling forward) Market-1501 128x64", batch=512, hw=(128, 64), model=13,
                   text="Sampling (--model=13 tester.generate: encoder, Fg/Bg GaussianFCRes, pose auto-encoder, inflation, "
                        "U-Net, denorm, critic score, SSIM), batch=%(b)d, forward-only engine"),
}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    bf16_burst=d.get("bf16_tflops"),
                    source="MEASURED_PEAKS.json (bf16 sustained; kernel timed inside a long step)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1400.0, bf16_burst=None, source="fallback of B200_PROFILING.md (sustained)")


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant launches, from the committed
    `ncu --set full` captures (tools/ncu_capture.sh -> profiles/*_ncu_full_summary*.json); a bench run cannot sit under
    ncu itself.  Newest capture first."""
    out = []
    for name in ("r02_ncu_full_summary.json", "r01_ncu_full_summary_v3.json"):
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        try:
            with open(path) as fh:
                caps = json.load(fh)
        except Exception:
            continue
        # algorithmic bytes of the captured micro-shapes (tests/bench_conv_micro.py / bench_dgrad_micro.py), fp32-equivalent
        # split-bf16 activations of 64x128x64 images: tensors read + written once, 4 B / element
        shapes = {"conv_fwd_256": ("forward conv 256->256 3x3 + residual, 64x128x64 (pair kernel)", 3 * 64 * 128 * 64 * 256 * 4),
                  "conv_fwd_128": ("forward conv 128->128 3x3 + residual, 64x128x64 (wide-B kernel)", 3 * 64 * 128 * 64 * 128 * 4),
                  "wgrad_256": ("filter gradient 256->256 3x3, 64x128x64 (pair kernel)", 2 * 64 * 128 * 64 * 256 * 4)}
        for key, rows in caps.items():
            base = key.replace(".ncu-rep", "")
            if base not in shapes or not rows:
                continue
            r = max(rows, key=lambda x: x.get("time_ns", x.get("time_us", 0)))
            out.append({"launch": shapes[base][0], "dram_bytes": r["dram_read_bytes"] + r["dram_write_bytes"],
                        "algorithmic_bytes": shapes[base][1],
                        "tensor_pipe_pct": r.get("tensor_pipe_active_pct_elapsed", r.get("tensor_pipe_pct")),
                        "source": "profiles/" + name})
        if out:
            order = [v[0] for v in shapes.values()]
            out.sort(key=lambda e: order.index(e["launch"]))       # the 256 -> 256 forward launch first
            break
    return out or None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms by one background process that runs for the whole
    benchmark; window(t0, t1) summarises the samples that fall DURING a timed region (host timestamps)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.samples = []   # (host time, fields)

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t_end = time.time() + 5.0
            while not self.samples and time.time() < t_end:   # first sample before anything is timed
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def window(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [ln for (t, ln) in self.samples if t0 <= t <= t1 + 0.15]
        if not rows:   # region shorter than the sampling period: take the closest sample
            rows = [min(self.samples, key=lambda s: abs(s[0] - 0.5 * (t0 + t1)))[1]] if self.samples else []
        for ln in rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ CPU arm (oracle port)
def _host_threads():
    """Every host core this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is ONE
    process (rank 0), so it takes the whole box regardless."""
    import torch
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


class CpuStep:
    """One step of a workload on the fp32 PyTorch-CPU oracle port (oracle/nets.py: the restatement of the reference graph;
    TF 1.4 / Python 2.7 cannot run here) at a given batch."""

    def __init__(self, workload, batch, mode="dcgan"):
        import numpy as np
        import torch
        from dpig_b200 import synth
        from oracle import nets
        from oracle import tf_ops as T
        self.torch, self.nets, self.T, self.workload, self.batch = torch, nets, T, workload, batch
        df = workload == "df256"
        self.cfg = nets.NetConfig.deepfashion() if df else nets.NetConfig()
        cfg = self.cfg
        params = nets.init_params(cfg, seed=1234)

        def make(seed):
            b = synth.make_batch(batch, cfg.img_h, cfg.img_w, seed=seed)
            ob = dict(x=torch.tensor(b["x"]), mask=torch.tensor(b["mask"]), pose_rcv=torch.tensor(b["pose_rcv"]),
                      pose=T.pose_rasterize(torch.tensor(b["pose_rcv"]), cfg.img_h, cfg.img_w),
                      part_bbox=torch.tensor(b["part_bbox"][:, :7]), part_vis=torch.tensor(b["part_vis"][:, :7]))
            return ob
        self.batches = [make(123 + i) for i in range(2)]
        if workload in ("market", "df256"):
            self.tr = nets.Stage1Trainer(params, cfg, mode=mode, dtype=torch.float32)
        elif workload == "stage2":
            self.p1 = nets.to_torch(params, torch.float32)
            self.p2 = nets.to_torch(nets.init_stage2_params(seed=4321), torch.float32, requires_grad=True)
            self.ms = {k: torch.ones_like(v) for k, v in self.p2.items()}     # TF RMSProp slot starts at ones
        else:
            params = dict(params)
            params.update(nets.init_stage2_params(seed=4321))
            params.update(nets.init_pose_params(seed=777))
            self.p = nets.to_torch(params, torch.float32)
            rng = np.random.default_rng(5)
            self.z = (torch.tensor(rng.normal(0, 0.2, size=(batch, 224)).astype(np.float32)),
                      torch.tensor(rng.normal(0, 0.2, size=(batch, 128)).astype(np.float32)))

    def _stage2_call(self, factor, which, batch):
        torch, nets, T = self.torch, self.nets, self.T
        dim = 224 if factor == "fg" else 128
        with torch.no_grad():
            if which == "d":
                emb = nets.encoder(self.p1, self.cfg, batch)
                real = emb[:, :224] if factor == "fg" else emb[:, 224:]
            else:
                real = torch.zeros((self.batch, dim))
        z = torch.randn((self.batch, dim)) * 0.2
        out = nets.stage2_losses(self.p2, factor, real, z, "wgan")
        scope = ("Gaussian_FC_%s/" % ("Fg" if factor == "fg" else "Bg")) if which == "g" else ("%s_FCDis_" % ("Fg" if factor == "fg" else "Bg"))
        names = [k for k in self.p2 if k.startswith(scope)]
        grads = torch.autograd.grad(out["g_loss"] if which == "g" else out["d_loss"], [self.p2[k] for k in names])
        with torch.no_grad():
            for k, g in zip(names, grads):
                T.rmsprop_step(self.p2[k], g, self.ms[k], 2e-5, clip=0.01 if which == "d" else None)

    def step(self, i=0):
        if self.workload in ("market", "df256"):
            self.tr.g_step(self.batches[0])
            self.tr.d_step(self.batches[1])
        elif self.workload == "stage2":
            for factor in ("fg", "bg"):
                self._stage2_call(factor, "g", None)
                for j in range(5):
                    self._stage2_call(factor, "d", self.batches[j % 2])
        else:
            with self.torch.no_grad():
                self.nets.sample_factor_forward(self.p, self.cfg, self.batches[0], self.z[0], self.z[1], True, True, True)


def _time_cpu(workload, batch, warmup, steps, mode="dcgan"):
    st = CpuStep(workload, batch, mode)
    for i in range(warmup):
        st.step(i)
    times = []
    for i in range(steps):
        t0 = time.perf_counter()
        st.step(i)
        times.append(time.perf_counter() - t0)
    return times


def cpu_baseline(workload, mode="dcgan", budget_s=20.0):
    """`cpu_baseline` leg of the GPU arm (rank 0, N=1): a BOUNDED sample of the workload on the host cores."""
    threads = _host_threads()
    b = 2 if workload != "sample" else 4
    times = []
    st = CpuStep(workload, b, mode)
    st.step()                                   # warm-up (allocator, oneDNN primitive caches)
    t_end = time.perf_counter() + budget_s
    while len(times) < 8 and (not times or time.perf_counter() < t_end):
        t0 = time.perf_counter()
        st.step()
        times.append(time.perf_counter() - t0)
    t = statistics.median(times)
    return {"value": b / t, "unit": "images/s", "cores": threads, "kind": "port",
            "sample": "fp32 PyTorch-CPU oracle port (not TF1) of the %s step at batch=%d, median of %d steps (%.2f s each), "
                      "%d host threads" % (workload, b, len(times), t, threads)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (TF 1.4.1 / Python 2.7 are not
    installable here and nothing of the reference compiles: it is pure TF-1 graph code), all host threads, the same
    workload / metric; each step is a bounded sample: the workload's step at a reduced batch chosen so that W + K steps
    end within ~2.5 minutes.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    t_wall = time.perf_counter()
    threads = _host_threads()
    wl = WORKLOADS[args.workload]
    K, W = max(1, args.steps), max(0, args.warmup)
    # probe: seconds per image of this workload on this box (one warm + one timed step at a tiny batch)
    pb = 2 if args.workload != "sample" else 4
    probe = _time_cpu(args.workload, pb, 1, 1, args.mode)[0] / pb
    want = 16 if args.workload in ("market", "stage2") else (8 if args.workload == "df256" else 32)   # 16 = the shipped batch
    budget = 150.0
    b = int(max(1, min(want, budget / max(probe * (K + W), 1e-9))))
    if args.workload == "stage2":
        b = max(2, b - b % 2)
    times = _time_cpu(args.workload, b, W, K, args.mode)
    total = sum(times)
    value = b * K / total
    extra = {}
    if args.workload == "market":           # B=1 (config 1) and B=16 (the shipped batch) beside it
        t1 = statistics.median(_time_cpu("market", 1, 1, 3, args.mode))
        extra["batch1"] = {"images_per_s": 1.0 / t1, "s_per_step": t1, "steps": 3}
        if b != 16 and probe * 16 * 3 < 60.0:
            t16 = statistics.median(_time_cpu("market", 16, 1, 2, args.mode))
            extra["batch16"] = {"images_per_s": 16.0 / t16, "s_per_step": t16, "steps": 2}
        elif b == 16:
            extra["batch16"] = {"images_per_s": value, "s_per_step": total / K, "steps": K}
    sample = ("fp32 PyTorch-CPU oracle port (oneDNN; not TF1) of the %s step; every step is a bounded sample of the workload "
              "at batch=%d (GPU arm: batch %d per GPU); %d warm-up + %d timed steps, %.2f s per step, %d host threads"
              % (args.workload, b, args.batch or wl["batch"], W, K, total / K, threads))
    cb = {"value": value, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample}
    cb.update(extra)
    line = {
        "impl": "reference", "metric": wl["metric"], "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": 1000.0 * total / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": wl["text"] % dict(b=args.batch or wl["batch"], mode=args.mode), "cpu_sample_batch": b,
                   "host_threads": threads},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_wall,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
class PoolLoader:
    """`_load_batch_pair_pose` stand-in over a pool of pinned host batches: every call yields the next one (q2)."""

    def __init__(self, pool):
        self.pool, self.i = pool, 0

    def next_batch(self):
        b = self.pool[self.i % len(self.pool)]
        self.i += 1
        return b


def _flags(args, wl, batch, tmp, extra=()):
    H, W = wl["hw"]
    return ["--model=%d" % wl["model"], "--batch_size=%d" % batch, "--img_H=%d" % H, "--img_W=%d" % W,
            "--model_dir=%s" % tmp, "--log_dir=%s" % tmp, "--synthetic_data=true", "--gan_mode=%s" % args.mode,
            "--log_step=1000000000", "--lr_update_step=1000000000", "--g_lr=2e-5", "--d_lr=2e-5",
            "--random_seed=1234"] + list(extra)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (default: the workload's)")
    ap.add_argument("--workload", default="market", choices=sorted(WORKLOADS),
                    help="market = BASELINE.json configs[1] (the headline); the others are reported on request")
    ap.add_argument("--impl", default="dpig", choices=["dpig", "reference"])
    ap.add_argument("--mode", default="dcgan")
    ap.add_argument("--fast", action="store_true", help="single bf16 pass (NOT parity mode; labelled)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--detail", default="", help="write the per-launch conv timings of one iteration here")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    from dpig_b200 import config as cfgmod
    from dpig_b200 import ddp, synth

    wl = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = ddp.Dist() if world > 1 else None
    rank = dist.rank if dist else 0
    local = dist.local_rank if dist else 0
    torch.cuda.set_device(local)
    W = max(3, args.warmup)
    K = max(1, args.steps)
    B = args.batch or wl["batch"]
    H, Wd = wl["hw"]
    tmp = tempfile.mkdtemp(prefix="dpig_bench_")

    # pinned host batches; a step consumes several (one per optimiser call, reference q2)
    # critic updates per generator update: 1 for dcgan / lsgan, CRITIC_ITERS = 5 for wgan / wgan-gp (trainer.py:339-345)
    disc_iters = 1 if args.mode in ("dcgan", "lsgan") else 5
    per_step = {"market": 1 + disc_iters, "df256": 1 + disc_iters, "stage2": 10, "sample": 1}[args.workload]
    npool = 4 if per_step <= 2 else max(10, per_step)
    pool = []
    for i in range(npool):
        b = synth.make_batch(B, H, Wd, seed=1000 + 17 * rank + i)
        pool.append({k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in b.items()})
    dev_pool = [{k: v.cuda(non_blocking=True) for k, v in b.items()} for b in pool]
    batch_bytes = sum(int(v.numel() * v.element_size()) for v in pool[0].values())
    h2d = per_step * batch_bytes

    # ---- the reference-facing object of the workload (trainer / tester class of main.py's --model table)
    if args.workload in ("market", "df256"):
        from dpig_b200 import trainer, trainer_256
        conf, _ = cfgmod.get_config(_flags(args, wl, B, tmp))
        cls = trainer.DPIG_Encoder_GAN_BodyROI_FgBg if args.workload == "market" else trainer_256.DPIG_Encoder_GAN_BodyROI_256
        obj = cls(conf, loader=PoolLoader(pool), dist=dist)
        obj.init_net()
        eng, ctx = obj.net, obj.ctx
        d2h = 12

        def dev_step(i, timings=None):
            eng.set_batch(dev_pool[(per_step * i) % npool])
            eng.g_step(timings)
            for j in range(disc_iters):
                eng.set_batch(dev_pool[(per_step * i + 1 + j) % npool])
                eng.d_step(timings)

        def e2e_run(k0, k):
            sink = []
            obj.start_step, obj.max_step = 1 + k0, 1 + k0 + k          # step > 0: the G update is not skipped (q1)
            obj.train(on_step=lambda step, t: sink.append(t.net.losses()))   # D2H read of (g_gan, d_loss, L1) per step
            return sink
    elif args.workload == "stage2":
        from dpig_b200 import trainer_sub
        conf, _ = cfgmod.get_config(_flags(args, wl, B, tmp))
        obj = trainer_sub.DPIG_Encoder_subSampleAppNetFgBg_GAN_BodyROI(conf, loader=PoolLoader(pool), dist=dist)
        obj.init_net()
        eng, ctx = obj.net, obj.ctx
        d2h = 16
        dev_loader = PoolLoader(dev_pool)
        # `value` / `e2e` are measured on the reference's work: every critic call runs the WHOLE frozen encoder, as its
        # sess.run does (fg_embs / bg_embs are slices of the concatenated embedding, trainer.py:741-742).  The product default
        # runs only the pyramid the trained factor reads (exact, Stage2Engine.prune); that rate is measured after the timed
        # regions and reported beside the value as config.pruned_encoder.
        obj.s2.prune = False

        def dev_step(i, timings=None):
            obj.s2.train_iteration(1 + i, dev_loader.next_batch)

        def e2e_run(k0, k):
            sink = []
            obj.start_step, obj.max_step = 1 + k0, 1 + k0 + k
            obj.train(on_step=lambda step, t: sink.append([float(v) for f in ("fg", "bg") for v in t.s2.f[f].loss.cpu()]))
            return sink
    else:
        if world > 1:
            raise SystemExit("--workload sample is a one-GPU configuration (BASELINE.json configs[4])")
        from dpig_b200 import tester
        conf, _ = cfgmod.get_config(_flags(args, wl, B, tmp, ["--is_train=False", "--sample_fg=True", "--sample_bg=True",
                                                              "--sample_pose=True"]))
        obj = tester.DPIG_FourNetsFgBg_testOnlySampleFactor(conf, loader=PoolLoader(pool))
        obj.init_net()
        eng, ctx = obj.s1, obj.ctx
        d2h = B * H * Wd * 3 * 4 * 2 + B * 8          # G, pose image (float32 NHWC), score, ssim
        # BASELINE.json's configs[4] counts the appearance encoder in the sampling pass: `value` / `e2e` run it.  With all
        # three factors sampled (run_market_test.sh:64-80) the encoder is not an ancestor of the fetched G, tf.Session.run
        # never executes it and neither does the product default -- that rate follows as an extra (pruned_encoder).
        obj.encode_unused = True

        def dev_step(i, timings=None):
            eng.set_batch(dev_pool[i % npool])
            obj.generate_on_device()

        def e2e_run(k0, k):
            out = None
            for i in range(k):
                b = pool[(k0 + i) % npool]
                out = obj.generate(b["x"], None, b["pose_rcv"], b["part_bbox"], b["part_vis"], mask=b["mask"])
            return out
    if args.fast:
        ctx.set_fast_mode(1)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    e2e_run(0, W)                      # warm-up through the public surface (captures the step graphs on one GPU)
    for i in range(W):
        dev_step(i)
    barrier()

    sampler = ClockSampler(local)
    sampler.start()

    def timed(fn):
        barrier()
        t_host0 = time.time()
        launches0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        barrier()
        clocks = sampler.window(t_host0, time.time())
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if dist:
            dist.all_reduce_max(ms)
        return float(ms[0]), ctx.launch_count() - launches0, clocks

    ms_dev, launches, clocks = timed(lambda: [dev_step(i) for i in range(K)])
    ms_e2e, _, clocks_e2e = timed(lambda: e2e_run(W, K))
    pruned = None
    if args.workload == "stage2":
        obj.s2.prune = True
        for i in range(W):
            dev_step(i)                 # its own step graphs (captured on the third call of a kind)
        ms_p, launches_p, _ = timed(lambda: [dev_step(i) for i in range(K)])
        pruned = {"images_per_s": B * world * K / (ms_p * 1e-3), "ms_per_step": ms_p / K, "gpu_launches": launches_p,
                  "what": "product default (DPIG_STAGE2_PRUNE=1): a critic call runs only the encoder pyramid its factor "
                          "reads -- same updates, the other pyramid's output feeds nothing in that call"}
        obj.s2.prune = False
    elif args.workload == "sample":
        obj.encode_unused = False
        for i in range(W):
            dev_step(i)
        ms_p, launches_p, _ = timed(lambda: [dev_step(i) for i in range(K)])
        pruned = {"images_per_s": B * world * K / (ms_p * 1e-3), "ms_per_step": ms_p / K, "gpu_launches": launches_p,
                  "what": "product default = what tf.Session.run executes for these flags: every factor is sampled, the "
                          "encoder is not an ancestor of G and is not run"}
        obj.encode_unused = True

    # ---- per-launch roofline figures: one more step OUTSIDE the timed regions, replayed launch by launch with CUDA events
    # around every C-ABI call (the timed steps are whole-step CUDA graph replays on one GPU)
    tl = []
    if args.workload in ("market", "df256"):
        dev_step(0, tl)
    else:
        s = torch.cuda.current_stream().cuda_stream
        eng.set_batch(dev_pool[0])
        eng.p_fwd_enc.run(s, tl)
        if args.workload == "sample":
            eng.p_fwd_unet.run(s, tl)
            eng.p_d_fake_fwd.run(s, tl)
    barrier()
    sampler.stop()

    per = {}
    detail = []
    for name, flops, a, b, tag in tl:
        d = per.setdefault(name, [0, 0.0, 0.0])
        ms_call = a.elapsed_time(b)
        d[0] += 1
        d[1] += flops
        d[2] += ms_call
        if flops > 0:
            detail.append((name, tag, flops, ms_call))
    if args.detail and rank == 0:
        with open(args.detail, "w") as fh:
            for name, tag, flops, ms_call in detail:
                fh.write("%-18s %-58s %9.3f GFLOP %8.3f ms %8.1f TFLOP/s\n" % (name, tag, flops / 1e9, ms_call,
                                                                              flops / ms_call / 1e9 if ms_call > 0 else 0))
    iter_ms = ms_dev / K
    peaks = _peaks()
    zero = [0, 0.0, 0.0]
    conv_n = per.get("conv2d_fwd", zero)[0] + per.get("conv2d_bwd_data", zero)[0]
    conv_fl = per.get("conv2d_fwd", zero)[1] + per.get("conv2d_bwd_data", zero)[1]
    conv_ms = per.get("conv2d_fwd", zero)[2] + per.get("conv2d_bwd_data", zero)[2]
    wg = per.get("conv2d_bwd_filter", zero)
    passes = 1 if args.fast else 3
    ach = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    # for stage2 the event-timed program (one encoder forward) runs 10x per step
    reps = 10 if args.workload == "stage2" else 1
    traffic = _ncu_traffic()
    roofline = {
        "kernel": "conv_umma_kernel (tcgen05 implicit-GEMM conv: forward + data-gradient launches)",
        "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
        "frac": ach / peaks["bf16_tflops"], "peak_source": peaks["source"] + " -- of measured",
        # one number, per launch like `achieved` would be for that launch: DRAM bytes of the step's largest launch class
        # (the 256 -> 256 3x3 forward conv on 64x128x64 maps) from the committed `ncu --set full` capture; all three
        # captured launches with their algorithmic bytes beside them in traffic_detail
        "traffic": (traffic[0]["dram_bytes"] if traffic else None), "traffic_detail": traffic,
        "note": "achieved = algorithmic conv FLOPs (2*pixels*Cout*k*k*Cin) / CUDA-event time of the launches; the "
                "parity mode issues %d bf16 MMA passes per product, so executed tensor FLOPs = %dx algorithmic and the "
                "attainable frac is <= 1/%d by construction: BASELINE.json's '>= 40 %% of the conv roofline' cannot be "
                "met by a 3-pass fp32-equivalent kernel (cap 33.3 %%); frac_executed is the tensor pipe's own fraction"
                % (passes, passes, passes),
        "mma_passes": passes, "frac_executed": passes * ach / peaks["bf16_tflops"],
        "frac_executed_of_burst": (passes * ach / peaks["bf16_burst"]) if peaks.get("bf16_burst") else None,
        "timing": "per-launch CUDA events on the launching stream in one extra step replayed launch by launch right after "
                  "the timed regions (the timed steps themselves are whole-step CUDA graph replays on one GPU)",
        "launches_per_iteration": conv_n * reps, "share_of_iteration": reps * conv_ms / iter_ms if iter_ms else None,
        "wgrad_kernel": {"achieved": (wg[1] / (wg[2] * 1e-3) / 1e12) if wg[2] > 0 else 0.0, "unit": "TFLOP/s",
                         "launches_per_iteration": wg[0], "share_of_iteration": wg[2] / iter_ms if iter_ms else None},
        "kernel_time_ms_per_iteration": {k: round(v[2] * reps, 3) for k, v in sorted(per.items(), key=lambda kv: -kv[1][2])[:12]},
    }

    if rank == 0:
        n_img = B * world
        value = n_img * K / (ms_dev * 1e-3)
        e2e_v = n_img * K / (ms_e2e * 1e-3)
        graphs = bool(getattr(eng, "use_graphs", False)) if args.workload != "sample" else False
        line = {
            "metric": wl["metric"], "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16 hi/lo split x3 MMA, fp32 accumulate (fp32-equivalent forward; gradients carry ~1e-5 relative)"
                     if not args.fast else "bf16 (fast mode, NOT parity)",
            "data": "synthetic",
            "config": {"workload": wl["text"] % dict(b=B, mode=args.mode), "global_batch": n_img,
                       "parallelism": "dp%d" % world,
                       "l2": "working set per step (>8 GB of activations) exceeds the 126 MB L2; no flush needed",
                       "value_leg": "inputs resident in HBM; every optimiser call first swaps in its own batch with a "
                                    "device-to-device copy",
                       "e2e_leg": {"market": "trainer.DPIG_Encoder_GAN_BodyROI_FgBg.train()",
                                   "df256": "trainer_256.DPIG_Encoder_GAN_BodyROI_256.train()",
                                   "stage2": "trainer_sub.DPIG_Encoder_subSampleAppNetFgBg_GAN_BodyROI.train()",
                                   "sample": "tester.DPIG_FourNetsFgBg_testOnlySampleFactor.generate()"}[args.workload] +
                                  " with pinned host batches; losses / images read back every step"},
            "clocks": clocks,
            **({"pruned_encoder": pruned} if pruned else {}),
            "e2e": {"value": e2e_v, "unit": "images/s", "ms_per_step": ms_e2e / K, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "clocks": clocks_e2e},
            "gpu_launches": launches,
            "launch_mode": ("CUDA graph replay of each optimiser call (forward + backward + update%s), %d kernels per step"
                            % (" + NCCL exchanges" if world > 1 else "", launches // K)) if graphs else "eager launch lists",
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload, args.mode)
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()


if __name__ == "__main__":
    main()
