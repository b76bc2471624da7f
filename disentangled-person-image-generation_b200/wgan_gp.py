"""Mirror of the reference's wgan_gp.py call surface that trainers use: LeakyReLU (wgan_gp.py:23-24), the norm
switch Batchnorm (wgan_gp.py:34-40), class WGAN_GP (wgan_gp.py:95-117) with DCGANDiscriminator (:407-440) and
FCDiscriminator (:399-405).  Eager, op by op over tflib (forward only) -- for reference-style scripting and
tests; training uses engine.Stage1Engine / stage2.Stage2Engine."""
import torch

from . import tflib as lib
from .tflib.ops import batchnorm as _bn
from .tflib.ops import conv2d as _conv
from .tflib.ops import layernorm as _ln
from .tflib.ops import linear as _lin


def LeakyReLU(x, alpha=0.2):
    return torch.maximum(alpha * x, x)


def LeakyReLULayer(name, n_in, n_out, inputs):
    return LeakyReLU(_lin.Linear(name + ".Linear", n_in, n_out, inputs, initialization="he"))


def Batchnorm(name, axes, inputs, MODE):
    if ("Discriminator" in name) and (MODE == "wgan-gp"):
        if axes != [0, 2, 3]:
            raise Exception("Layernorm over non-standard axes is unsupported")
        return _ln.Layernorm(name, [1, 2, 3], inputs)
    return _bn.Batchnorm(name, axes, inputs, fused=True)


class WGAN_GP(object):
    def __init__(self, DATA_DIR="", MODE="wgan-gp", DIM=64, BATCH_SIZE=64, ITERS=200000, LAMBDA=10,
                 G_OUTPUT_DIM=128 * 64 * 3, IMG_H=128, IMG_W=64):
        self.DATA_DIR, self.MODE, self.DIM, self.BATCH_SIZE, self.ITERS = DATA_DIR, MODE, DIM, BATCH_SIZE, ITERS
        self.LAMBDA, self.G_OUTPUT_DIM, self.IMG_H, self.IMG_W = LAMBDA, G_OUTPUT_DIM, IMG_H, IMG_W
        self.CRITIC_ITERS = 5
        self.N_GPUS = 1
        self.DEVICES = ["/gpu:{}".format(i) for i in range(self.N_GPUS)]

    def FCDiscriminator(self, inputs, input_dim, FC_DIM=512, n_layers=3, reuse=False, name=""):
        output = LeakyReLULayer(name + "Discriminator.Input", input_dim, FC_DIM, inputs)
        for i in range(n_layers):
            output = LeakyReLULayer(name + "Discriminator.{}".format(i), FC_DIM, FC_DIM, output)
        output = _lin.Linear(name + "Discriminator.Out", FC_DIM, 1, output)
        return output.reshape(-1)

    def DCGANDiscriminator(self, inputs, input_dim=3, dim=64, bn=True, nonlinearity=LeakyReLU, name=""):
        """inputs: NCHW fp32 CUDA tensor (the reference transposes before the call, trainer.py:601-602)."""
        _conv.set_weights_stdev(0.02)
        _lin.set_weights_stdev(0.02)
        output = _conv.Conv2D(name + "Discriminator.1", input_dim, dim, 5, inputs, stride=2)
        output = nonlinearity(output)
        for i, (ci, co) in enumerate(((dim, 2 * dim), (2 * dim, 4 * dim), (4 * dim, 8 * dim)), start=2):
            output = _conv.Conv2D(name + "Discriminator.%d" % i, ci, co, 5, output, stride=2)
            if bn:
                output = Batchnorm(name + "Discriminator.BN%d" % i, [0, 2, 3], output, self.MODE)
            output = nonlinearity(output)
        output = output.reshape(-1, 8 * 4 * 8 * dim)      # NCHW flatten incl. the 256x256 quirk
        output = _lin.Linear(name + "Discriminator.Output", 8 * 4 * 8 * dim, 1, output)
        _conv.unset_weights_stdev()
        _lin.unset_weights_stdev()
        return output.reshape(-1)
