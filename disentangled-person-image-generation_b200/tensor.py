"""Device-memory containers for the DPIG kernels.

torch is used only to own HBM (allocation, H2D/D2H copies, streams); the kernels see raw pointers.
`SplitTensor` is the Python handle of `struct dpig_tensor`: an NHWC activation stored as two bf16
planes (hi, lo) with value = hi + lo.
"""
import ctypes as C

import torch

from . import _lib


def ptr(t):
    """Raw device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


class SplitTensor:
    """NHWC split-bf16 tensor, possibly a channel slice of a wider (concat) buffer."""

    def __init__(self, n, h, w, c, device="cuda", c_alloc=None, _buf=None, _c0=0, zero=False):
        self.n, self.h, self.w, self.c = int(n), int(h), int(w), int(c)
        self.c_alloc = int(c_alloc or c)
        if _buf is None:
            alloc = torch.zeros if zero else torch.empty
            _buf = alloc((2, self.n, self.h, self.w, self.c_alloc), dtype=torch.bfloat16, device=device)
        self.buf = _buf
        self.c0 = int(_c0)
        self._struct = None

    # ---- views
    def slice(self, c0, c):
        assert 0 <= c0 and c0 + c <= self.c
        return SplitTensor(self.n, self.h, self.w, c, c_alloc=self.c_alloc, _buf=self.buf, _c0=self.c0 + c0)

    def batch_slice(self, n0, n):
        assert self.c0 == 0 or True
        v = SplitTensor(n, self.h, self.w, self.c, c_alloc=self.c_alloc, _buf=self.buf[:, n0:n0 + n], _c0=self.c0)
        return v

    @property
    def hi(self):
        return self.buf[0][..., self.c0:self.c0 + self.c]

    @property
    def lo(self):
        return self.buf[1][..., self.c0:self.c0 + self.c]

    def struct(self):
        if self._struct is None:
            base_hi = self.buf[0].data_ptr() + 2 * self.c0
            base_lo = self.buf[1].data_ptr() + 2 * self.c0
            self._struct = _lib.Tensor(base_hi, base_lo, self.n, self.h, self.w, self.c, self.c_alloc)
        return self._struct

    def ref(self):
        return C.byref(self.struct())

    # ---- host-side conversions (tests, I/O); the hot path never calls these
    def float(self):
        return self.hi.float() + self.lo.float()

    def set_from_float(self, x):
        x = x.to(self.buf.device, torch.float32)
        hi = x.to(torch.bfloat16)
        lo = (x - hi.float()).to(torch.bfloat16)
        self.hi.copy_(hi)
        self.lo.copy_(lo)
        return self

    @staticmethod
    def from_float(x, c_alloc=None):
        n, h, w, c = x.shape
        t = SplitTensor(n, h, w, c, device=x.device, c_alloc=c_alloc, zero=True)
        return t.set_from_float(x)


def split_ref(x):
    """fp32 value a split tensor would hold for x (hi + lo), for oracle-side comparisons."""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.float() + lo.float()
