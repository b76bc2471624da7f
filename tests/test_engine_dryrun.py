"""CPU dry run of the engine's host logic: every launch program of Stage1Engine is RECORDED (buffers allocated on the
CPU, no kernel is called -- the library has no CPU path) and each recorded C-ABI call is checked against the ctypes
signature of include/dpig.h: arity and argument types.  Catches host-side breakage (a renamed buffer, a missing
argument, a program that references a gradient buffer in forward-only mode) without a GPU; the numerics are the
`-m gpu` tests' business."""
import ctypes as C
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from dpig_b200 import _lib, engine  # noqa: E402


class DryContext:
    """Stands in for _lib.Context where no device exists: the library is loaded (symbols resolve), nothing is launched."""

    def __init__(self):
        self.lib = _lib.load()
        self.handle = None
        self.replayed_launches = 0

    def launch_count(self):
        return 0

    def last_error(self):
        return ""


def _check_program(prog):
    n = 0
    for name, fn, args, flops, tag in prog.calls:
        if name is None:
            continue
        want = fn.argtypes[1:-1]          # minus ctx, minus the trailing stream
        assert len(args) == len(want), ("dpig_" + name, len(args), len(want))
        for i, (a, t) in enumerate(zip(args, want)):
            try:
                t.from_param(a)
            except (TypeError, C.ArgumentError) as e:      # pragma: no cover - the message is the point
                raise AssertionError("dpig_%s argument %d: %r does not convert to %s (%s)" % (name, i, a, t, e))
        n += 1
    return n


SMALL = dict(img_h=32, img_w=16, hidden=64, roi_size=12, d_dim=64)


@pytest.mark.parametrize("mode", ["dcgan", "wgan-gp", "wgan"])
def test_training_programs_record(mode):
    eng = engine.Stage1Engine(DryContext(), engine.NetConfig(**SMALL), 2, mode=mode, device="cpu")
    progs = [eng.p_fwd_gen, eng.p_fwd_enc, eng.p_fwd_unet, eng.p_bwd_gen, eng.p_d_fake_fwd, eng.p_d_real_fwd,
             eng.p_d_fake_bwd_data, eng.p_d_fake_bwd_par, eng.p_d_real_bwd_par]
    if mode == "wgan-gp":
        progs.append(eng.p_gp)
    assert sum(_check_program(p) for p in progs) > 150
    # conv + norm + LeakyReLU of the critic (wgan_gp.py:417-431): the conv epilogue emits the statistics, so a block is
    # two launches (conv, normalise-and-activate); no separate statistics pass
    names = [c[0] for c in eng.p_d_fake_fwd.calls if c[0]]
    assert "norm_stats" not in names
    assert names.count("conv2d_fwd") == 4 and names.count("norm_act_fwd") == 3


def test_deepfashion_joint_programs_record():
    cfg = engine.NetConfig.deepfashion(img_h=64, img_w=64, hidden=64, roi_size=16)
    eng = engine.Stage1Engine(DryContext(), cfg, 2, mode="dcgan", device="cpu")
    for p in (eng.p_fwd_gen, eng.p_bwd_gen, eng.p_d_pair_fwd, eng.p_d_pair_bwd_data, eng.p_d_pair_bwd_par):
        assert _check_program(p) > 0


def test_forward_only_engine_has_no_gradient_state():
    """inference=True (tester.py sampling at batch 512): forward programs only, no gradient / optimiser arenas."""
    cfg = engine.NetConfig(**SMALL)
    eng = engine.Stage1Engine(DryContext(), cfg, 4, mode="dcgan", device="cpu", inference=True)
    for p in (eng.p_fwd_gen, eng.p_fwd_enc, eng.p_fwd_unet, eng.p_d_fake_fwd, eng.p_d_real_fwd):
        assert _check_program(p) > 0
    assert not hasattr(eng, "p_bwd_gen") and not hasattr(eng, "g_cat") and not hasattr(eng.d_fake, "g_x")
    assert eng.gp.grad.numel() == 1 and eng.gp.m.numel() == 1
    with pytest.raises(_lib.DpigError):
        eng.g_step()
    full = engine.Stage1Engine(DryContext(), cfg, 4, mode="dcgan", device="cpu")

    def nbytes(e):
        seen, total = set(), 0
        stack = [e]
        while stack:
            o = stack.pop()
            if id(o) in seen:
                continue
            seen.add(id(o))
            if hasattr(o, "untyped_storage"):
                st = o.untyped_storage()
                if st.data_ptr() not in seen:
                    seen.add(st.data_ptr())
                    total += st.nbytes()
            elif isinstance(o, (list, tuple)):
                stack.extend(o)
            elif isinstance(o, dict):
                stack.extend(o.values())
            elif hasattr(o, "__dict__") and type(o).__module__.startswith("dpig_b200"):
                stack.extend(vars(o).values())
        return total

    assert nbytes(eng) < 0.6 * nbytes(full)


def test_df_sampler_engine_ceil_halving():
    """The BodyROI encoder of the DeepFashion sampler stages (--model=103 / 104): 48x48 crops through repeat_num+1 levels
    with TensorFlow's SAME / stride-2 sizes ceil(s / 2): 48 -> 24 -> 12 -> 6 -> 3 -> 2 -> 1 (models.py:275-325)."""
    assert [engine.halved(48, k) for k in range(7)] == [48, 24, 12, 6, 3, 2, 1]
    cfg = engine.NetConfig.deepfashion(img_h=256, img_w=256, hidden=64, roi_size=48, use_vis=False)
    eng = engine.Stage1Engine(DryContext(), cfg, 1, mode="dcgan", device="cpu", inference=True)
    assert [d[:2] for d in eng.roi_pyr.dims] == [(48, 48), (24, 24), (12, 12), (6, 6), (3, 3), (2, 2), (1, 1)]
    assert eng.roi_flat == 1 * 1 * 64 * 7          # same FC input as --model=101's 64x64 crops: the checkpoints interchange
    assert _check_program(eng.p_fwd_gen) > 0


def test_side_stream_schedule_of_the_programs():
    """Two-stream launch schedule (engine.Program): the background branch of the encoder forward is tagged for the side
    stream and joined before the embedding is assembled; in the backward programs every filter gradient is tagged for the
    side stream, and a python hook (an all-reduce of a finished gradient slice) is a join point."""
    eng = engine.Stage1Engine(DryContext(), engine.NetConfig(**SMALL), 2, mode="dcgan", device="cpu")
    p = eng.p_fwd_enc
    names = [c[0] for c in p.calls]
    join = next(i for i, c in enumerate(p.calls) if c[0] is None and c[1] is None)
    assert names[join + 1] == "embedding_assemble" and all(i < join for i in p.side)
    n_bg_convs = len(eng.n_bg)
    assert sum(1 for i in p.side if names[i] == "conv2d_fwd") == n_bg_convs and len(p.side) == n_bg_convs + 2
    # the main-stream calls between the first side call and the join are the ROI branch: they never touch a bg buffer
    b = eng.p_bwd_gen
    wg = [i for i, c in enumerate(b.calls) if c[0] == "conv2d_bwd_filter"]
    assert set(wg) - b.side == {i for i in wg if "patch form" in b.calls[i][4]}      # the two patch-form wgrads stay in order
    assert all(b.calls[i][0] == "conv2d_bwd_filter" for i in b.side)
    fwd_only = engine.Stage1Engine(DryContext(), engine.NetConfig(**SMALL), 2, mode="dcgan", device="cpu", inference=True)
    assert fwd_only.p_fwd_enc.side == p.side


def test_single_pyramid_encoder_programs():
    """Stage2Engine.prune / tester._appearance_branch launch the encoder with one pyramid only: the ROI program has no
    Bg-pyramid launch and no mask_split, the Bg program no crop_and_resize; both end in the embedding assembly, and together
    they hold the full program's launches (plus the shared stem twice)."""
    eng = engine.Stage1Engine(DryContext(), engine.NetConfig(**SMALL), 2, mode="dcgan", device="cpu", inference=True)
    full = [c[0] for c in eng.p_fwd_enc.calls if c[0]]
    roi = [c[0] for c in eng.p_fwd_enc_only["roi"].calls if c[0]]
    bg = [c[0] for c in eng.p_fwd_enc_only["bg"].calls if c[0]]
    for prog in (eng.p_fwd_enc, eng.p_fwd_enc_only["roi"], eng.p_fwd_enc_only["bg"]):
        assert _check_program(prog) > 0
    assert "crop_and_resize_fwd" in roi and "mask_split" not in roi
    assert "mask_split" in bg and "crop_and_resize_fwd" not in bg
    assert roi[-1] == bg[-1] == full[-1] == "embedding_assemble"
    stem = 5                      # pack, im2col, three stem convs
    assert len(roi) + len(bg) == len(full) + stem + 1
    # the tags name the layers: the ROI program touches no Bg-pyramid weights and vice versa
    tags = lambda prog: {c[4].split()[0] for c in prog.calls if c[0] == "conv2d_fwd"}      # noqa: E731
    w_roi, w_bg = {n + "/weights" for n in eng.n_roi}, {n + "/weights" for n in eng.n_bg}
    assert w_roi <= tags(eng.p_fwd_enc_only["roi"]) and not (w_bg & tags(eng.p_fwd_enc_only["roi"]))
    assert w_bg <= tags(eng.p_fwd_enc_only["bg"]) and not (w_roi & tags(eng.p_fwd_enc_only["bg"]))
    # no side-stream launches when a pyramid runs alone
    assert not eng.p_fwd_enc_only["bg"].side and not eng.p_fwd_enc_only["roi"].side
    # the DeepFashion encoder has one pyramid: nothing to prune
    df = engine.Stage1Engine(DryContext(), engine.NetConfig.deepfashion(img_h=64, img_w=64, hidden=64, roi_size=16), 2,
                             mode="dcgan", device="cpu", inference=True)
    assert df.p_fwd_enc_only == {}


def test_sampler_held_factors():
    """Which factors make the tester run the encoder (tester.py:536-552: a held factor reads the first sample's encoder
    embedding; a sampled one does not)."""
    from dpig_b200 import tester
    t = tester.DPIG_FourNetsFgBg_testOnlySampleFactor.__new__(tester.DPIG_FourNetsFgBg_testOnlySampleFactor)
    for fg, bg, want in ((True, True, []), (False, True, ["fg"]), (True, False, ["bg"]), (False, False, ["fg", "bg"])):
        t.sample_fg, t.sample_bg = fg, bg
        assert t._held_factors() == want
    t11 = tester.DPIG_FourNetsFgBg_testOnly.__new__(tester.DPIG_FourNetsFgBg_testOnly)
    t11.sample_app = True
    assert t11._held_factors() == []
    t11.sample_app = False
    assert t11._held_factors() == ["fg", "bg"]
