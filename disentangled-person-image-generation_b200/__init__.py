"""dpig_b200 -- B200-native hot path of charliememory/Disentangled-Person-Image-Generation.

Layout (only what the hot path needs):
  csrc/        hand-written sm_100a CUDA kernels + the C ABI (include/dpig.h) -> libdpig.so
  _lib.py      ctypes binding (the Python<->CUDA boundary)
  tensor.py    device-memory containers (torch owns HBM; kernels see raw pointers)
  engine.py    static launch programs for the Stage-I graph (forward, backward, optimiser)
  models.py / wgan_gp.py / tflib/  host-side mirror of the reference's operator interface
  trainer.py / tester.py / main.py / config.py  the reference's call surface
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
from ._lib import Context, DpigError  # noqa: F401
