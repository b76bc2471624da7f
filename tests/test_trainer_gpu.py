"""GPU tests of the reference's call surface: `main.py --model=N` -> trainer class -> init_net() / train() / test()
(reference main.py:12-90, trainer.py:326-366, trainer_256.py:95-134) driving the engine for a few steps."""
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pytestmark = pytest.mark.gpu


def _run(tmp_path, model, extra):
    import dpig_b200  # noqa: F401
    from dpig_b200 import config as C
    from dpig_b200 import main as M
    argv = ["--model=%d" % model, "--is_train=True", "--batch_size=2", "--max_step=3", "--log_step=2", "--gpu=-1",
            "--model_dir=%s" % tmp_path, "--conv_hidden_num=64", "--an_unknown_flag=1"] + extra
    cfg, unparsed = C.get_config(argv)
    assert unparsed == ["--an_unknown_flag=1"]          # config.py:96 parse_known_args
    return M.main(cfg)


@pytest.mark.parametrize("model,extra,cls", [
    (1, ["--img_H=32", "--img_W=16"], "DPIG_Encoder_GAN_BodyROI_FgBg"),
    (101, ["--img_H=128", "--img_W=128", "--dataset=DF_train_data"], "DPIG_Encoder_GAN_BodyROI_256"),
])
def test_main_trains_and_generates(tmp_path, model, extra, cls):
    tr = _run(tmp_path, model, extra)
    assert type(tr).__name__ == cls
    recs = [json.loads(ln) for ln in open(os.path.join(str(tmp_path), "summary.jsonl"))]
    assert [r["step"] for r in recs] == [0, 1] and all(np.isfinite(r["loss/g_loss"]) for r in recs)
    assert set(recs[0]) >= {"loss/L1Loss", "loss/g_loss_only", "loss/g_loss", "loss/d_loss", "misc/g_lr", "misc/d_lr"}
    b = tr.loader.next_batch()
    g = tr.generate(b["x"], b["x"], b["pose_rcv"], b["part_bbox"], b["part_vis"], mask=b["mask"])
    assert g.shape == b["x"].shape and g.dtype == np.uint8
    # saver.save -> TensorFlow V2 checkpoint; a second trainer resumes from it (--ckpt_path) incl. the Adam slots
    from dpig_b200 import tf_checkpoint
    prefix = tr.save(2)
    z = tf_checkpoint.CheckpointReader(prefix)
    assert z.has_tensor("Encoder/G_encoder/Conv/weights") and z.has_tensor("Discriminator.Output.W")
    assert z.has_tensor("ID_AE/G/Conv_3/weights/Adam_1") and z.has_tensor("beta1_power_1") and int(z.get_tensor("step")) == 2
    if model == 101:
        assert z.get_variable_to_shape_map()["Discriminator.Output.W"] == [16384, 1]
        assert not z.has_tensor("Encoder/G_encoder/fully_connected_1/weights")
    import copy
    from dpig_b200 import main as M
    cfg2 = copy.copy(tr.config)
    cfg2.ckpt_path, cfg2.max_step, cfg2.model_dir = str(tmp_path), 0, str(tmp_path / "resumed")
    tr2 = M.main(cfg2)
    a, b2 = tr.net.get_state(), tr2.net.get_state()
    assert set(a) == set(b2)
    for k in a:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b2[k])), k
    assert tr2.net.t == tr.net.t


def test_main_rejects_unknown_models(tmp_path):
    """Every id of the reference's --model table is built (main.py:22-72); anything else is refused like there."""
    with pytest.raises(Exception, match="unknown --model"):
        _run(tmp_path, 999, [])


def test_main_model104_deepfashion_pose_sampler(tmp_path):
    """--model=104 (trainer_256.py:511-700): the pose-sampler stage at the DeepFashion normalisation; its preview runs the
    sampler stages' Stage-I graph (BodyROI encoder on 48x48 crops, no visibility gating)."""
    tr = _run(tmp_path, 104, ["--img_H=128", "--img_W=128", "--dataset=DF_train_data"])
    assert type(tr).__name__ == "DPIG_subnetSamplePoseRCV_GAN_BodyROI_256"
    b = tr.loader.next_batch()
    g = tr.generate(b["x"], b["x"], b["pose_rcv"], b["part_bbox"], part_vis=b["part_vis"], mask=b["mask"])
    assert g.shape == (2, 128, 128, 3) and g.dtype == np.uint8
    assert (tr.net.cfg.roi_size, tr.net.cfg.use_vis) == (48, False)
