"""Entry point mirroring the reference's main.py:12-90: `--model` id -> trainer class, `--is_train` -> train()/test().
Every id of the reference's table is implemented: 1 (Stage-I Market-1501, the BASELINE hot path), 2 / 3 / 4 (trainer_sub.py),
11 / 12 / 13 and 1001 / 1002 (tester.py), 101 - 104 (trainer_256.py: the DeepFashion 256x256 stages).  Under torchrun
(WORLD_SIZE > 1) the trainers that take data-parallel hooks get a ddp.Dist."""
import os

from .config import get_config, prepare_dirs, save_config

_MODEL_CLASSES = {
    1: "DPIG_Encoder_GAN_BodyROI_FgBg", 2: "DPIG_PoseRCV_AE_BodyROI", 3: "DPIG_Encoder_subSampleAppNetFgBg_GAN_BodyROI",
    4: "DPIG_subnetSamplePoseRCV_GAN_BodyROI", 11: "DPIG_FourNetsFgBg_testOnly", 12: "DPIG_FourNetsFgBg_testOnlyCondition",
    13: "DPIG_FourNetsFgBg_testOnlySampleFactor", 101: "DPIG_Encoder_GAN_BodyROI_256", 102: "DPIG_PoseRCV_AE_BodyROI_256",
    103: "DPIG_Encoder_subSampleAppNet_GAN_BodyROI_256", 104: "DPIG_subnetSamplePoseRCV_GAN_BodyROI_256",
    1001: "DPIG_ThreeNetsApp_testOnlyCondition_256", 1002: "DPIG_ThreeNetsApp_testOnlySampleFactor_256",
}


def main(config):
    from . import tester as TE
    from . import trainer as T
    from . import trainer_256 as T256
    from . import trainer_sub as TS
    prepare_dirs(config)            # main.py:13 prepare_dirs_and_logger
    if config.gpu > -1:
        os.environ["CUDA_DEVICE_ORDER"] = "PCI_BUS_ID"
        os.environ["CUDA_VISIBLE_DEVICES"] = str(config.gpu)
    config.data_format = "NHWC"
    name = _MODEL_CLASSES.get(config.model)
    if name is None:
        raise Exception("unknown --model=%r" % config.model)
    cls = getattr(T256, name, None) or getattr(T, name, None) or getattr(TS, name, None) or getattr(TE, name, None)   # main.py:4-6 import order (q9)
    if cls is None:
        raise NotImplementedError("--model=%d (%s) is outside this round's hot path (SURVEY.md §8f)" % (config.model, name))
    # one process per GPU under torchrun (WORLD_SIZE > 1): data-parallel hooks for the trainers that take them
    dist = None
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import inspect
        if "dist" in inspect.signature(cls.__init__).parameters:
            from . import ddp
            dist = ddp.Dist()
    trainer = cls(config, dist=dist) if dist is not None else cls(config)
    if getattr(config, "model_dir", None) and os.path.isdir(config.model_dir) and (dist is None or dist.rank == 0):
        save_config(config)         # again: now with synthetic_data_effective (which loader the run really uses)
    trainer.init_net()
    if config.is_train:
        trainer.train()
    else:
        trainer.test()
    return trainer


if __name__ == "__main__":
    cfg, _ = get_config()
    prepare_dirs(cfg)
    save_config(cfg)                # main.py:88-90
    main(cfg)
