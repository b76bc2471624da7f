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


class DPIG_Encoder_GAN_BodyROI_256(DPIG_Encoder_GAN_BodyROI_FgBg):
    def __init__(self, config, loader=None, dist=None):
        super().__init__(config, loader=loader, dist=dist)
        self.gan_mode = "dcgan"

    def _net_config(self):
        return engine.NetConfig.deepfashion(img_h=self.img_H, img_w=self.img_W, hidden=self.conv_hidden_num,
                                            z_num=self.z_num)
