"""Command-line surface of the reference (config.py:15-102): same flag names, types and defaults, unknown
flags tolerated (`parse_known_args`, config.py:96) so that run_market_*.sh command lines parse unchanged.
Flags the reference parses but never reads (optimizer, gamma, lambda_k, L1Loss_weight, interpolate_*, ...)
are accepted and likewise ignored."""
import argparse
import json
import os
from datetime import datetime


def str2bool(v):
    return str(v).lower() in ("true", "1")


def build_parser():
    p = argparse.ArgumentParser()
    net = p.add_argument_group("Network")
    net.add_argument("--img_H", type=int, default=128)
    net.add_argument("--img_W", type=int, default=64)
    net.add_argument("--conv_hidden_num", type=int, default=128, choices=[64, 128])
    net.add_argument("--z_num", type=int, default=64)
    data = p.add_argument_group("Data")
    data.add_argument("--dataset", type=str, default="CelebA")
    data.add_argument("--split", type=str, default="train")
    data.add_argument("--batch_size", type=int, default=16)
    data.add_argument("--grayscale", type=str2bool, default=False)
    data.add_argument("--num_worker", type=int, default=4)
    for name in ("ckpt_path", "pretrained_path", "pretrained_appSample_path", "pretrained_poseAE_path",
                 "pretrained_poseSample_path", "FeaLossModel_path", "z_emb_dir"):
        data.add_argument("--" + name, type=str, default=None)
    tr = p.add_argument_group("Training")
    tr.add_argument("--is_train", type=str2bool, default=True)
    tr.add_argument("--test_one_by_one", type=str2bool, default=False)
    tr.add_argument("--optimizer", type=str, default="adam")
    tr.add_argument("--start_step", type=int, default=0)
    tr.add_argument("--max_step", type=int, default=500000)
    tr.add_argument("--lr_update_step", type=int, default=100000)
    tr.add_argument("--L1Loss_weight", type=float, default=20)
    tr.add_argument("--d_lr", type=float, default=0.00008)
    tr.add_argument("--g_lr", type=float, default=0.00008)
    tr.add_argument("--beta1", type=float, default=0.5)
    tr.add_argument("--beta2", type=float, default=0.999)
    tr.add_argument("--gamma", type=float, default=0.5)
    tr.add_argument("--lambda_k", type=float, default=0.001)
    tr.add_argument("--use_gpu", type=str2bool, default=True)
    tr.add_argument("--gpu", type=int, default=-1)
    tr.add_argument("--model", type=int, default=0)
    tr.add_argument("--D_arch", type=str, default="DCGAN")
    for name in ("sample_app", "sample_fg", "sample_bg", "sample_pose", "one_app_per_batch", "interpolate_fg",
                 "interpolate_fg_up", "interpolate_fg_down", "interpolate_bg", "interpolate_pose", "inverse_fg",
                 "inverse_bg", "inverse_pose"):
        tr.add_argument("--" + name, type=str2bool, default=False)
    misc = p.add_argument_group("Misc")
    misc.add_argument("--load_path", type=str, default="")
    misc.add_argument("--log_step", type=int, default=200)
    misc.add_argument("--save_model_secs", type=int, default=1000)
    misc.add_argument("--num_log_samples", type=int, default=3)
    misc.add_argument("--log_level", type=str, default="INFO", choices=["INFO", "DEBUG", "WARN"])
    misc.add_argument("--log_dir", type=str, default="logs")
    misc.add_argument("--model_dir", type=str, default=None)
    misc.add_argument("--data_dir", type=str, default="data")
    misc.add_argument("--test_data_path", type=str, default=None)
    misc.add_argument("--sample_per_image", type=int, default=64)
    misc.add_argument("--random_seed", type=int, default=123)
    # additions of this implementation (not in the reference)
    misc.add_argument("--gan_mode", type=str, default="dcgan", choices=["dcgan", "wgan", "wgan-gp", "lsgan"],
                      help="loss / critic-norm mode of wgan_gp.WGAN_GP (the shipped Stage-I trainer hard-codes dcgan)")
    misc.add_argument("--synthetic_data", type=str2bool, default=None,
                      help="unset: read the TFRecord pair files under <data_dir>/<dataset> when they exist, else synthetic batches "
                           "with a warning; true / false force either (datasets.py, trainer.make_loader)")
    return p


def get_config(argv=None):
    config, unparsed = build_parser().parse_known_args(argv)
    config.data_format = "NHWC"   # main.py:18 overrides config.py:97-101
    return config, unparsed


def prepare_dirs(config):
    """utils.prepare_dirs_and_logger (utils.py:110-141) minus the logging handler: `--load_path` inside `--log_dir` IS the
    model directory, otherwise it names it (prefixed with the dataset unless it already starts with it); without it the
    name is <dataset>_<MMDD_HHMMSS>.  An explicit `--model_dir` (what run_market_*.sh pass) wins.  Also sets
    `config.data_path = <data_dir>/<dataset>` (utils.py:136) and creates the log / model directories."""
    model_name = None
    if config.load_path:
        if config.load_path.startswith(config.log_dir):
            config.model_dir = config.load_path
        elif config.load_path.startswith(config.dataset):
            model_name = config.load_path
        else:
            model_name = "%s_%s" % (config.dataset, config.load_path)
    else:
        model_name = "%s_%s" % (config.dataset, datetime.now().strftime("%m%d_%H%M%S"))
    config.model_name = model_name
    if getattr(config, "model_dir", None) is None:
        config.model_dir = os.path.join(config.log_dir, model_name)
    config.data_path = os.path.join(config.data_dir, config.dataset)
    for path in (config.log_dir, config.model_dir):
        os.makedirs(path, exist_ok=True)
    return config


def save_config(config):
    """utils.save_config (utils.py:146-154): the parsed flags as <model_dir>/params.json."""
    param_path = os.path.join(config.model_dir, "params.json")
    with open(param_path, "w") as fp:
        json.dump(config.__dict__, fp, indent=4, sort_keys=True)
    return param_path
