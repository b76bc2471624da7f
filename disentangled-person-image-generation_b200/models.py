"""Mirror of the reference's models.py call surface for the live Stage-I/II graph functions -- same names,
argument order and return tuples:

    GeneratorCNN_ID_Encoder_BodyROIVis_FgBgFeaTwoBranch   models.py:390-471
    GeneratorCNN_ID_UAEAfterResidual                       models.py:518-576
    GaussianFCRes                                          models.py:474-486

In the reference these build TF sub-graphs; here they run the corresponding forward launch programs of a cached
engine.Stage1Engine / stage2 tape on CUDA arrays and return results as fp32 CUDA tensors.  `variables` is the
dict of parameters (TF variable name -> fp32 tensor view) of the scope, the analogue of
tf.contrib.framework.get_variables(vs).  `reuse=True` reuses the cached engine (shared weights), like a TF
variable scope.  Only data_format='NHWC' and activation_fn=relu (what every trainer passes, trainer.py:581-595)
are supported for the conv nets.
"""
import numpy as np
import torch

from . import _lib, engine, stage2, synth
from . import tflib as lib
from .tensor import ptr

_engines = {}


def relu(x):
    return torch.relu(x)


def LeakyReLU(x, alpha=0.3):
    """models.LeakyReLU (models.py:137) -- alpha 0.3; the trainers resolve wgan_gp.LeakyReLU (0.2) instead."""
    return torch.maximum(alpha * x, x)


def _engine_for(batch, img_h, img_w, hidden, z_num, roi_size, repeat_num):
    key = (batch, img_h, img_w, hidden, z_num, roi_size, repeat_num)
    if key not in _engines:
        cfg = engine.NetConfig(img_h=img_h, img_w=img_w, hidden=hidden, z_num=z_num, roi_size=roi_size,
                               repeat_num=repeat_num)
        eng = engine.Stage1Engine(lib.context(), cfg, batch, mode="dcgan")
        eng.load_params(engine.init_params(cfg))
        _engines[key] = eng
    return _engines[key]


def _variables(eng, prefix):
    return {k: eng.gp.view(k) for k in eng.gp.specs if k.startswith(prefix)}


def GeneratorCNN_ID_Encoder_BodyROIVis_FgBgFeaTwoBranch(x, fg_mask, ROI_bboxs, ROI_vis, bbox_num, z_num, repeat_num,
                                                        hidden_num, data_format, activation_fn=relu,
                                                        keep_part_prob=1.0, roi_size=48, reuse=False):
    """x [B,H,W,3], fg_mask [B,H,W,1], ROI_bboxs int [B,bbox_num,4] (y1,x1,y2,x2 px), ROI_vis [B,bbox_num].
    Returns (fea_all [B,bbox_num*z_num + 4*z_num], fea_list, conv_fea_list, variables)."""
    if data_format != "NHWC" or keep_part_prob != 1.0 or bbox_num != 7 or z_num != 32:
        raise Exception("only the configuration the reference trainers pass is implemented")
    B, H, W, _ = x.shape
    eng = _engine_for(B, H, W, hidden_num, 64, roi_size, repeat_num)
    eng.set_batch(dict(x=x, pose_rcv=torch.zeros((B, 18, 3)), mask=fg_mask, part_bbox=ROI_bboxs, part_vis=ROI_vis))
    eng.run_encoder()
    fea_all = eng.emb.clone()
    fea_list = list(torch.split(fea_all[:, :bbox_num * z_num], z_num, dim=1)) + [fea_all[:, bbox_num * z_num:]]
    conv_fea_list = [eng.rois.float()[i * B:(i + 1) * B] for i in range(bbox_num)] + [eng.x_bg.float()]
    return fea_all, fea_list, conv_fea_list, _variables(eng, "Encoder/G_encoder")


def GeneratorCNN_ID_UAEAfterResidual(x, pose, input_channel, z_num, repeat_num, hidden_num, data_format,
                                     activation_fn=relu, min_fea_map_H=8, noise_dim=0, reuse=False, pose_rcv=None):
    """x: the spatially tiled embedding [B,H,W,E] (trainer.py:588-590; only x[:,0,0,:] is read -- it is constant
    over space by construction) ; pose: [B,H,W,18] maps, or pass the raw keypoints as pose_rcv [B,18,3] to have the
    maps rasterised on the GPU.  Returns (out [B,H,W,input_channel], z [B,z_num], variables)."""
    if data_format != "NHWC" or noise_dim != 0 or input_channel != 3:
        raise Exception("only the configuration the reference trainers pass is implemented")
    B, H, W, E = x.shape
    eng = _engine_for(B, H, W, hidden_num, z_num, 48 if H >= 128 else 12, repeat_num)
    eng.emb.copy_(x[:, 0, 0, :].to(eng.device, torch.float32))
    s = torch.cuda.current_stream().cuda_stream
    if pose_rcv is not None:
        eng.pose_rcv.copy_(torch.as_tensor(pose_rcv, dtype=torch.float32).to(eng.device))
    prog = engine.Program(eng.ctx)
    eng._prog_unet_forward(prog)
    if pose_rcv is None:
        # maps given: drop the two calls that rasterise from keypoints (the maps, and their 3x3 patches for the stem's
        # patch-form contraction), inject the maps into the stem-input slice and unfold the patches from there
        prog.calls = [c for c in prog.calls if c[0] not in ("pose_rasterize", "pose_patch")]
        sl = eng.gin.slice(0, eng.cfg.keypoints)
        sl.set_from_float(torch.as_tensor(pose, dtype=torch.float32).to(eng.device))
        eng.ctx.im2col_small(eng.gin.ref(), eng.cfg.keypoints, 3, 3, 1, 0, eng.pose_patch.ref(), s)
    prog.run(s)
    return eng.G.clone(), eng.z.clone(), _variables(eng, "ID_AE/G")


def GaussianFCRes(z_shape, out_channel, repeat_num, hidden_num, data_format, mean=0.0, stddev=0.2,
                  activation_fn=relu, reuse=False, z=None, scope="G_FC"):
    """noise [B, z_shape[-1]] -> residual MLP -> [B, out_channel] (models.py:474-486).  activation_fn must be relu
    or a LeakyReLU(0.2) (what trainer.py:753-757 passes)."""
    B, zin = z_shape[0], z_shape[-1]
    key = ("fc", scope, B, zin, out_channel, hidden_num, repeat_num)
    act = _lib.ACT_RELU if activation_fn in (relu, torch.relu) else _lib.ACT_LRELU
    if key not in _engines:
        dev = torch.device("cuda", torch.cuda.current_device())
        grp = engine.ParamGroup(stage2.fc_res_specs(scope, zin, hidden_num, out_channel, repeat_num), dev)
        rng = np.random.default_rng(0)
        for name, (off, n, shape) in grp.specs.items():
            if name.endswith("weights"):
                lim = np.sqrt(6.0 / (shape[0] + shape[1]))
                grp.view(name).copy_(torch.as_tensor(rng.uniform(-lim, lim, size=shape).astype(np.float32)))
        tape = stage2.FCTape(lib.context(), grp, B, dev)
        zin_node = stage2._Node(B, zin, dev)
        names = list(grp.specs)
        lay = [(names[2 * i], names[2 * i + 1]) for i in range(len(names) // 2)]
        h = tape.linear(zin_node, *lay[0], act=act)
        for r in range(repeat_num):
            a = tape.linear(h, *lay[1 + 2 * r], act=act)
            b = tape.linear(a, *lay[2 + 2 * r], act=act)
            h = tape.add(h, b)
        out = tape.linear(h, *lay[-1])
        _engines[key] = (grp, tape, zin_node, out, tape.forward_program())
    grp, tape, zin_node, out, prog = _engines[key]
    if z is None:
        zin_node.data.normal_(mean, stddev)
    else:
        zin_node.data.copy_(torch.as_tensor(z, dtype=torch.float32))
    prog.run(torch.cuda.current_stream().cuda_stream)
    return out.data.clone(), {k: grp.view(k) for k in grp.specs}
