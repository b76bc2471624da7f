"""Trainer call surface of the reference for the DeepFashion 256x256 Stage-I model (--model=101):

    class DPIG_Encoder_GAN_BodyROI_256    reference trainer_256.py:10-134

Same step loop as the Market-1501 trainer (trainer_256.py:95-134 repeats trainer.py:326-366); the graph differs
(trainer_256.py:31-68): `models.GeneratorCNN_ID_Encoder_BodyROIVis` (no Fg/Bg branch) with repeat_num+1 levels on
64x64 ROI crops, the U-Net with repeat_num-1 levels, and ONE DCGANDiscriminator call on concat([x, G]) whose
16384-wide reshape turns every 256x256 image into 8 logits -- engine.NetConfig.deepfashion().
MODE is hard-wired to 'dcgan' (trainer_256.py:28).
"""
from . import engine
from .trainer import DPIG_Encoder_GAN_BodyROI_FgBg
from .trainer_sub import DPIG_PoseRCV_AE_BodyROI


class DPIG_Encoder_GAN_BodyROI_256(DPIG_Encoder_GAN_BodyROI_FgBg):
    def __init__(self, config, loader=None, dist=None):
        super().__init__(config, loader=loader, dist=dist)
        self.gan_mode = "dcgan"

    def _net_config(self):
        return engine.NetConfig.deepfashion(img_h=self.img_H, img_w=self.img_W, hidden=self.conv_hidden_num,
                                            z_num=self.z_num)


class DPIG_PoseRCV_AE_BodyROI_256(DPIG_PoseRCV_AE_BodyROI):
    """--model=102 (trainer_256.py:404-509): the pose auto-encoder stage on DeepFashion keypoints.  The graph and the
    step are those of --model=2 (trainer.py:626-708; the two build_model / train bodies differ only in commented-out
    preview code): PoseEncoderFCRes -> PoseDecoderFCRes on (r / img_H, c / img_W, v) normalised to [-1, 1] with the
    256 x 256 image size of the flags, loss mean((pose - G_pose)^2) * 20, Adam(beta1 = .5) on the PoseAE variables."""
