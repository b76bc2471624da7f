"""CPU tests (no GPU): golden-vector regression of the oracle, the C-ABI surface of libdpig.so, host-side
logic shared by engine / oracle, and the data-parallel plumbing on gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def test_oracle_reproduces_golden_small():
    import make_golden
    got = make_golden.compute(make_golden.SMALL, 2)
    ref = np.load(os.path.join(ROOT, "tests", "golden", "stage1_small_b2.npz"))
    for k in ("emb", "z", "G", "D_real", "D_fake", "L1", "g_loss", "d_loss"):
        assert np.abs(got[k] - ref[k]).max() < 1e-6, k
    assert np.array_equal(got["pose"], ref["pose"])


def test_library_builds_and_exports_every_declared_symbol():
    """include/dpig.h is the contract: every `int dpig_*(` / `const char* dpig_*(` declared there must be an
    exported symbol of libdpig.so and be bound in _lib.py.  No compute calls (no GPU here)."""
    import __graft_entry__
    lib_path = __graft_entry__.build()
    hdr = open(os.path.join(ROOT, "include", "dpig.h")).read()
    declared = set(re.findall(r"\b(dpig_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"dpig_tensor", "dpig_conv_epilogue"}
    lib = ctypes.CDLL(lib_path)
    for name in sorted(declared):
        assert hasattr(lib, name), "libdpig.so does not export %s" % name
    from dpig_b200 import _lib
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    # product path must fail loudly without a device (no CPU fallback)
    if not torch.cuda.is_available():
        with pytest.raises(_lib.DpigError):
            _lib.Context(0)


def test_struct_layouts_match_header(tmp_path):
    """ctypes mirrors of dpig_tensor / dpig_conv_epilogue must have the C compiler's layout of include/dpig.h."""
    from dpig_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "%s"\n'
                   'int main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu\\n", sizeof(dpig_tensor), offsetof(dpig_tensor,pix_stride),'
                   ' sizeof(dpig_conv_epilogue), offsetof(dpig_conv_epilogue,mask_neg), offsetof(dpig_conv_epilogue,out_f32),'
                   ' offsetof(dpig_conv_epilogue,upsample)); return 0;}\n' % os.path.join(ROOT, "include", "dpig.h"))
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", str(src), "-o", str(exe)])
    c = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    py = [ctypes.sizeof(_lib.Tensor), _lib.Tensor.pix_stride.offset, ctypes.sizeof(_lib.ConvEpilogue),
          _lib.ConvEpilogue.mask_neg.offset, _lib.ConvEpilogue.out_f32.offset, _lib.ConvEpilogue.upsample.offset]
    assert c == py, (c, py)


def test_engine_and_oracle_agree_on_parameter_names_and_shapes():
    from dpig_b200 import engine
    from oracle import nets
    for kw in (dict(), dict(img_h=32, img_w=16, hidden=64, roi_size=12)):
        a = engine.init_params(engine.NetConfig(**kw))
        b = nets.init_params(nets.NetConfig(**kw))
        assert list(a.keys()) == list(b.keys())
        assert all(a[k].shape == b[k].shape for k in a)
    # DeepFashion graph (--model=101): no background branch, 7 / 5 levels, hard-wired 16384-wide D rows
    for kw in (dict(), dict(img_h=128, img_w=128, hidden=64, roi_size=32)):
        ec, oc = engine.NetConfig.deepfashion(**kw), nets.NetConfig.deepfashion(**kw)
        a, b = engine.init_params(ec), nets.init_params(oc)
        assert list(a.keys()) == list(b.keys())
        assert all(a[k].shape == b[k].shape for k in a)
        assert (ec.d_row, ec.d_rows, ec.emb_dim) == (oc.d_row, ec.d_rows, oc.emb_dim)
    ec = engine.NetConfig.deepfashion()
    assert (ec.enc_repeat, ec.unet_repeat, ec.emb_dim, ec.d_row, ec.d_rows) == (7, 5, 224, 16384, 8)
    a = engine.init_params(ec)
    assert a["Encoder/G_encoder/Conv_22/weights"].shape == (3, 3, 896, 896)      # last ROI level (1x1 pixel maps)
    assert a["Encoder/G_encoder/fully_connected/weights"].shape == (896, 32)
    assert a["ID_AE/G/Conv/weights"].shape == (3, 3, 224 + 18, 128)
    assert a["ID_AE/G/fully_connected/weights"].shape == (16 * 16 * 640, 64)
    assert "Encoder/G_encoder/fully_connected_1/weights" not in a
    p = engine.init_params(engine.NetConfig())
    w = p["ID_AE/G/Conv_5/weights"]
    lim = np.sqrt(6.0 / (9 * w.shape[2] + 9 * w.shape[3]))       # slim xavier_uniform
    assert np.abs(w).max() <= lim and np.abs(w).max() > 0.95 * lim
    assert np.abs(p["Discriminator.2.Filters"]).max() <= 0.02 * np.sqrt(3.0) + 1e-7


def test_launch_programs_record_without_a_device():
    """Host logic of the engine: buffer geometry and the recorded C-ABI call lists of every graph variant are built
    on the CPU (no kernel runs; the context handle is absent), so shape / wiring errors surface without a GPU."""
    from dpig_b200 import _lib, engine

    class NoDeviceCtx:
        lib, handle = _lib.load(), None

    cases = [(engine.NetConfig(img_h=32, img_w=16, hidden=64, roi_size=12), "dcgan"),
             (engine.NetConfig(img_h=32, img_w=16, hidden=64, roi_size=12), "wgan-gp"),
             (engine.NetConfig.deepfashion(img_h=128, img_w=128, hidden=64, roi_size=32), "dcgan")]
    for cfg, mode in cases:
        eng = engine.Stage1Engine(NoDeviceCtx(), cfg, 2, mode=mode, device="cpu")
        names = [c[0] for c in eng.p_bwd_gen.calls]
        n_conv = sum(1 for v in eng.conv.values() if not v.wname.startswith("Discriminator"))
        # every Encoder / G convolution gets exactly one filter-gradient launch (the stem's pose rows use _rows)
        assert names.count("conv2d_bwd_filter") + names.count("conv2d_bwd_filter_rows") == n_conv
        if cfg.d_joint:
            assert eng.d_pair.n == 4 and eng.d_real.logits.numel() == 2 * cfg.d_rows == 4
            assert eng.d_fake.logits.data_ptr() == eng.d_pair.logits.data_ptr() + 4 * 4
            assert [c[0] for c in eng.p_d_pair_fwd.calls].count("unpack_f32") == 2 * cfg.d_rows
        else:
            assert [c[0] for c in eng.p_d_fake_fwd.calls].count("unpack_f32") == 1
    import pytest
    with pytest.raises(_lib.DpigError):
        engine.Stage1Engine(NoDeviceCtx(), cases[2][0], 2, mode="wgan-gp", device="cpu")


def test_step_size_and_overlap_slice_host_logic(monkeypatch):
    """Host logic of the data-parallel / graph-replay step: the device-side step size equals TF-Adam's lr_t computed the
    way dpig_adam_step does (float32 lr and betas, float64 arithmetic), and the all-reduce that overlaps the encoder's
    backward covers exactly the ID_AE parameters (one contiguous, 64-padded block of the generator arena) and is issued
    right before the appearance encoder's backward."""
    import math
    from dpig_b200 import _lib, engine

    class NoDeviceCtx:
        lib, handle = _lib.load(), None

    class FakeDist:
        world_size, local_rank, rank = 2, 0, 0
        calls = []

        def all_reduce_sum(self, t):
            self.calls.append(t.numel())

    cfg = engine.NetConfig(img_h=32, img_w=16, hidden=64, roi_size=12)
    eng = engine.Stage1Engine(NoDeviceCtx(), cfg, 2, mode="dcgan", device="cpu")
    assert eng.use_graphs and not eng.overlap_comm                      # one GPU: graphs on, nothing to overlap
    eng.g_lr, eng.t["g"] = 2e-5, 3
    lr32, b2 = float(np.float32(2e-5)), float(np.float32(0.999))
    assert eng._step_size("g") == lr32 * math.sqrt(1.0 - b2 ** 3) / (1.0 - 0.5 ** 3)
    eng_w = engine.Stage1Engine(NoDeviceCtx(), cfg, 2, mode="wgan", device="cpu")
    eng_w.d_lr = 5e-5
    assert eng_w._step_size("d") == float(np.float32(5e-5))             # RMSProp: the plain lr
    assert not any(c[0] is None and i > 0 and eng.p_bwd_gen.calls[i + 1][0] == "embedding_assemble"
                   for i, c in enumerate(eng.p_bwd_gen.calls[:-1]))     # no hook without torch.distributed

    monkeypatch.setenv("DPIG_GRAPHS", "1")
    d = engine.Stage1Engine(NoDeviceCtx(), cfg, 2, mode="dcgan", dist=FakeDist(), device="cpu")
    assert d.overlap_comm and d.use_graphs                              # N > 1: the NCCL exchanges are captured too
    monkeypatch.setenv("DPIG_GRAPHS", "0")
    assert not engine.Stage1Engine(NoDeviceCtx(), cfg, 2, mode="dcgan", dist=FakeDist(), device="cpu").use_graphs
    monkeypatch.setenv("DPIG_GRAPHS", "1")
    # the three slices that are all-reduced while the backward pass is still running are disjoint, 64-aligned blocks
    rngs = [d._slice_range(w) for w in ("idae", "roi", "bg")]
    assert all(r is not None and r[0] % 64 == 0 and r[1] % 64 == 0 for r in rngs), rngs
    srt = sorted(rngs)
    assert all(a[1] <= b[0] for a, b in zip(srt, srt[1:])), rngs
    for name, (off, n, _) in d.gp.specs.items():
        inside = [w for w, r in zip(("idae", "roi", "bg"), rngs) if r[0] <= off and off + n <= r[1]]
        want = ["idae"] if name.startswith("ID_AE/") else (
            ["roi"] if any(name.startswith(x + "/") for x in d.n_roi + [d.n_roi_fc]) else (
                ["bg"] if any(name.startswith(x + "/") for x in d.n_bg + [d.n_bg_fc]) else []))
        assert inside == want, (name, inside, want)
    lo, hi = d._idae_range()
    assert lo % 64 == 0 and hi % 64 == 0 and 0 < lo < hi <= d.gp.total
    for name, (off, n, _) in d.gp.specs.items():
        assert (lo <= off and off + n <= hi) == name.startswith("ID_AE/"), name
    names = [c[0] for c in d.p_bwd_gen.calls]
    k = names.index("embedding_assemble")
    assert names[k - 1] is None                                         # the python hook that starts the early all-reduce
    # every ID_AE filter gradient is launched before the hook
    idae_layers = sum(1 for v in d.conv.values() if v.wname.startswith("ID_AE/"))
    before = names[:k].count("conv2d_bwd_filter") + names[:k].count("conv2d_bwd_filter_rows")
    assert before == idae_layers, (before, idae_layers)
    monkeypatch.setenv("DPIG_OVERLAP", "0")
    assert not engine.Stage1Engine(NoDeviceCtx(), cfg, 2, mode="dcgan", dist=FakeDist(), device="cpu").overlap_comm


def test_prepare_dirs_and_save_config_follow_the_reference(tmp_path):
    """utils.prepare_dirs_and_logger / save_config (utils.py:110-154): model directory naming and params.json."""
    import json
    from dpig_b200 import config as C
    log = str(tmp_path / "logs")
    cfg, _ = C.get_config(["--dataset=Market_train_data", "--log_dir=%s" % log, "--data_dir=%s" % (tmp_path / "data")])
    C.prepare_dirs(cfg)
    assert re.fullmatch(r"Market_train_data_\d{4}_\d{6}", cfg.model_name) and cfg.model_dir == os.path.join(log, cfg.model_name)
    assert cfg.data_path == os.path.join(str(tmp_path / "data"), "Market_train_data") and os.path.isdir(cfg.model_dir)
    cfg, _ = C.get_config(["--dataset=Market_train_data", "--log_dir=%s" % log, "--load_path=run7"])
    assert C.prepare_dirs(cfg).model_dir == os.path.join(log, "Market_train_data_run7")
    cfg, _ = C.get_config(["--dataset=Market_train_data", "--log_dir=%s" % log, "--load_path=Market_train_data_x"])
    assert C.prepare_dirs(cfg).model_dir == os.path.join(log, "Market_train_data_x")
    cfg, _ = C.get_config(["--dataset=Market_train_data", "--log_dir=%s" % log, "--load_path=%s/abc" % log])
    assert C.prepare_dirs(cfg).model_dir == "%s/abc" % log
    cfg, _ = C.get_config(["--dataset=D", "--log_dir=%s" % log, "--model_dir=%s" % (tmp_path / "explicit"), "--model=13"])
    assert C.prepare_dirs(cfg).model_dir == str(tmp_path / "explicit")           # run_market_*.sh pass --model_dir
    params = json.load(open(C.save_config(cfg)))
    assert params["model"] == 13 and params["data_format"] == "NHWC" and params["model_dir"] == cfg.model_dir


def test_synthetic_batch_shapes_and_box_rule():
    from dpig_b200 import synth
    b = synth.make_batch(3, 128, 64, seed=7)
    assert b["x"].shape == (3, 128, 64, 3) and b["mask"].shape == (3, 128, 64, 1)
    assert b["part_bbox"].shape == (3, 37, 4) and b["part_vis"].shape == (3, 37)
    assert 0.2 < b["mask"].mean() < 0.5
    bb = b["part_bbox"]
    assert (bb[..., 0] <= bb[..., 2]).all() and (bb[..., 2] <= 127).all() and (bb[..., 3] <= 63).all()
    # invisible part -> [0,0,1,1] (convert_market.py:699-702)
    rcv = b["pose_rcv"].copy()
    rcv[:, :, 2] = 0
    bb0, v0 = synth.part_boxes(rcv, 128, 64)
    assert (bb0 == np.array([0, 0, 1, 1])).all() and (v0 == 0).all()


def test_sync_batchnorm_from_shard_sums_equals_global_batch():
    """The sync-BN hook all-reduces raw (sum x, sum x^2): check that statistic algebra against the oracle's
    global-batch BatchNorm (tflib/ops/batchnorm.py:29-30)."""
    from oracle import tf_ops as T
    g = torch.Generator().manual_seed(0)
    x = torch.randn((4, 8, 4, 16), generator=g, dtype=torch.float64)
    sc, of = torch.rand(16, generator=g, dtype=torch.float64) + 0.5, torch.randn(16, generator=g, dtype=torch.float64)
    ref = T.batchnorm_train(x, sc, of)
    shards = [x[:2], x[2:]]
    s1 = sum(s.sum(dim=(0, 1, 2)) for s in shards)
    s2 = sum((s * s).sum(dim=(0, 1, 2)) for s in shards)
    cnt = x.numel() / 16
    mean, var = s1 / cnt, s2 / cnt - (s1 / cnt) ** 2
    got = torch.cat([(s - mean) * torch.rsqrt(var + 1e-5) * sc + of for s in shards])
    assert (got - ref).abs().max() < 1e-10


_WORKER = r'''
import os, sys, torch
sys.path.insert(0, %r)
from dpig_b200 import ddp
d = ddp.Dist(backend="gloo")
t = torch.full((5,), float(d.rank + 1), dtype=torch.float64)
d.all_reduce_sum(t)
assert torch.equal(t, torch.full((5,), 3.0, dtype=torch.float64)), t
m = torch.tensor([float(d.rank)])
d.all_reduce_max(m)
assert float(m) == 1.0
# data-parallel gradient == global-batch gradient: mean loss over the global batch
w = torch.arange(4, dtype=torch.float64)
x = torch.arange(24, dtype=torch.float64).reshape(6, 4) / 10.0
sh = ddp.shard({"x": x.numpy()}, d.rank, d.world_size)["x"]
g_local = torch.tensor(sh).mean(dim=0)            # grad of mean(x @ w) over the local shard
d.all_reduce_sum(g_local)
g_local /= d.world_size
assert torch.allclose(g_local, x.mean(dim=0)), (g_local, x.mean(dim=0))
d.barrier()
print("rank", d.rank, "ok")
'''


def test_ddp_plumbing_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2", CUDA_VISIBLE_DEVICES="")
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs), outs


def test_bench_reference_arm_uses_all_host_threads_under_torchrun_env():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm, also under torchrun): torchrun exports
    OMP_NUM_THREADS=1 to its workers, which made the round-1 arm single-threaded at N > 1.  The arm takes every core of the
    affinity mask regardless, runs exactly the requested steps, and ranks other than 0 exit without work."""
    import json
    import subprocess
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
           "--workload", "sample"]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["steps"] == 1 and line["warmup"] == 0 and line["n_gpus"] == 2
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) == line["config"]["host_threads"]
    assert line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0 and line["metric"].startswith("images/sec")
    other = subprocess.run(cmd, env=dict(env, RANK="1"), capture_output=True, text=True, timeout=120)
    assert other.returncode == 0 and other.stdout.strip() == ""
