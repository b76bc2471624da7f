"""ctypes binding of libdpig.so (C ABI declared in include/dpig.h).

This is the whole Python<->CUDA boundary: plain pointers and sizes, no torch types cross it.
There is deliberately NO fallback: if the library or an sm_100 device is missing, every entry
point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdpig.so")

OK = 0
ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2
NORM_LAYER, NORM_BATCH, NORM_INSTANCE = 0, 1, 2
GAN_DCGAN, GAN_WGAN, GAN_WGAN_GP, GAN_LSGAN = 0, 1, 2, 3
GAN_MODES = {"dcgan": GAN_DCGAN, "wgan": GAN_WGAN, "wgan-gp": GAN_WGAN_GP, "lsgan": GAN_LSGAN}


class DpigError(RuntimeError):
    pass


class Tensor(C.Structure):
    """struct dpig_tensor"""
    _fields_ = [("hi", C.c_void_p), ("lo", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32),
                ("w", C.c_int32), ("c", C.c_int32), ("pix_stride", C.c_int64)]


class ConvEpilogue(C.Structure):
    """struct dpig_conv_epilogue"""
    _fields_ = [("bias", C.c_void_p), ("act", C.c_int32), ("alpha", C.c_float),
                ("addend", C.POINTER(Tensor)), ("mask_in", C.c_void_p), ("mask_neg", C.c_float),
                ("mask_out", C.c_void_p), ("out", C.POINTER(Tensor)), ("out_masked", C.POINTER(Tensor)),
                ("out_f32", C.c_void_p), ("out_f32_pix_stride", C.c_int64), ("upsample", C.c_int32),
                ("class_bias", C.c_void_p), ("colsum_masked", C.c_void_p), ("stat_sums", C.c_void_p),
                ("stat_mode", C.c_int32)]


_P = C.c_void_p
_T = C.POINTER(Tensor)
_I = C.c_int32
_L = C.c_int64
_F = C.c_float
_D = C.c_double

# name -> argtypes after the leading ctx pointer (all return int)
_SIGNATURES = {
    "dpig_ctx_set_fast_mode": [C.c_int],
    "dpig_ctx_set_pair_mode": [C.c_int],
    "dpig_ctx_set_option": [C.c_char_p, C.c_int],
    "dpig_weight_pack": [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "dpig_conv2d_fwd": [_T, _P, _P, _I, _I, _I, _I, C.POINTER(ConvEpilogue), _P],
    "dpig_conv2d_bwd_data": [_T, _P, _P, _I, _I, _I, _I, _I, _I, C.POINTER(ConvEpilogue), _P],
    "dpig_conv2d_bwd_filter": [_T, _T, _I, _I, _I, _I, _I, _P, _P],
    "dpig_conv2d_bwd_filter_rows": [_T, _T, _I, _I, _I, _I, _I, _P, _I, _P],
    "dpig_weight_pack_rows": [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "dpig_stem_class_bias": [_P, _I, _I, _I, _I, _P, _P],
    "dpig_stem_tap_sums": [_T, _P, _P, _P],
    "dpig_conv2d_small_fwd": [_P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _I, _F, _T, _P, _P, _P],
    "dpig_conv2d_small_bwd_data": [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _P, _P],
    "dpig_conv2d_small_bwd_filter": [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _P, _P],
    "dpig_im2col_small": [_T, _I, _I, _I, _I, _I, _T, _P],
    "dpig_col2im_small": [_P, _L, _I, _I, _I, _I, _I, _I, _P, _P, _L, _T, _P],
    "dpig_permute_taps": [_P, _P, _I, _I, _I, _I, _P],
    "dpig_bias_grad": [_T, _P, _P],
    "dpig_bias_grad_f32": [_P, _L, _I, _P, _P],
    "dpig_ew_combine": [_T, _T, _T, _T, _P, _L, _P, _F, _I, _P],
    "dpig_ew_combine_colsum": [_T, _T, _T, _T, _P, _L, _P, _F, _I, _P, _P],
    "dpig_pack_f32": [_P, _L, _I, _T, _P],
    "dpig_unpack_f32": [_T, _P, _L, _P],
    "dpig_mask_split": [_T, _P, _T, _T, _P],
    "dpig_broadcast_embedding": [_P, _I, _T, _P],
    "dpig_spatial_sum": [_T, _P, _P],
    "dpig_embedding_assemble": [_P, _P, _P, _I, _I, _I, _I, _P, _I, _P],
    "dpig_crop_and_resize_fwd": [_T, _P, _P, _P, _I, _T, _P],
    "dpig_crop_and_resize_bwd": [_T, _P, _P, _P, _I, _P, _I, _I, _I, _I, _P],
    "dpig_linear_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "dpig_linear_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "dpig_add_f32": [_P, _P, _P, _L, _F, _F, _P],
    "dpig_act_bwd_f32": [_P, _P, _L, _F, _P],
    "dpig_norm_stats": [_P, _I, _I, _I, _I, _I, _P, _P],
    "dpig_norm_act_fwd": [_P, _I, _I, _I, _I, _I, _F, _P, _D, _P, _P, _I, _F, _P, _T, _P, _P],
    "dpig_norm_act_bwd_reduce": [_T, _P, _P, _P, _F, _I, _P, _P, _P, _P, _P],
    "dpig_norm_act_bwd_apply": [_T, _P, _P, _P, _F, _I, _P, _P, _D, _T, _P],
    "dpig_layernorm_jvp_fwd": [_P, _P, _I, _I, _I, _I, _P, _P, _P, _F, _P, _T, _P],
    "dpig_layernorm_jvp_bwd": [_T, _P, _F, _P, _P, _P, _P, _P, _P, _P, _T, _P, _P],
    "dpig_loss_l1": [_P, _P, _L, _F, _P, _P, _P],
    "dpig_loss_gan": [_I, _P, _P, _I, _P, _P, _P, _P, _P],
    "dpig_pose_ae_loss": [_P, _P, _P, _I, _I, _F, _P, _P, _P, _P, _P],
    "dpig_gp_interpolate": [_P, _P, _P, _I, _L, _P, _P],
    "dpig_gp_penalty": [_P, _I, _L, _F, _P, _P, _P, _P],
    "dpig_adam_step": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _I, _F, _P],
    "dpig_rmsprop_step": [_P, _P, _P, _L, _F, _F, _F, _F, _F, _P],
    "dpig_adam_step_dev": [_P, _P, _P, _P, _L, _P, _F, _F, _F, _F, _P],
    "dpig_rmsprop_step_dev": [_P, _P, _P, _L, _P, _F, _F, _F, _F, _P],
    "dpig_clip": [_P, _L, _F, _F, _P],
    "dpig_denorm_u8": [_P, _L, _P, _P],
    "dpig_ssim_gray_u8": [_P, _P, _I, _I, _I, _P, _P],
    "dpig_pose_rasterize": [_P, _I, _I, _I, _I, _I, _T, _P, _P],
    "dpig_pose_patch": [_P, _I, _I, _I, _I, _I, _I, _I, _T, _P],
}

EXPORTS = sorted(list(_SIGNATURES) + ["dpig_ctx_create", "dpig_ctx_destroy", "dpig_last_error",
                                      "dpig_launch_count", "dpig_version", "dpig_crc32c"])

_lib = None


def load():
    """dlopen libdpig.so (building it is __graft_entry__.build()'s job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DpigError("libdpig.so is missing at %s -- run `python __graft_entry__.py` (build) first; "
                        "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.dpig_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.dpig_ctx_create.restype = C.c_int
    lib.dpig_ctx_destroy.argtypes = [C.c_void_p]
    lib.dpig_ctx_destroy.restype = None
    lib.dpig_last_error.argtypes = [C.c_void_p]
    lib.dpig_last_error.restype = C.c_char_p
    lib.dpig_launch_count.argtypes = [C.c_void_p]
    lib.dpig_launch_count.restype = C.c_ulonglong
    lib.dpig_version.argtypes = []
    lib.dpig_version.restype = C.c_char_p
    for name, args in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = [C.c_void_p] + args
        fn.restype = C.c_int
    _lib = lib
    return lib


class Context:
    """One dpig_ctx (device + host thread). Methods mirror the C entry points minus the `dpig_` prefix
    and raise DpigError with dpig_last_error() on a non-zero return code."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.dpig_ctx_create(int(device), C.byref(h))
        if rc != OK:
            raise DpigError("dpig_ctx_create(device=%d) failed with %d: no sm_100 (B200) device available; "
                            "this library has no CPU or other-GPU fallback" % (device, rc))
        self.handle = h
        self.device = device
        self.replayed_launches = 0

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.dpig_ctx_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def last_error(self):
        return self.lib.dpig_last_error(self.handle).decode()

    def set_fast_mode(self, fast):
        """fast=True: single bf16 pass on the hi planes (NOT the parity mode)."""
        self.call("ctx_set_fast_mode", int(bool(fast)))

    def set_pair_mode(self, mode):
        """2-CTA cluster conv kernel: 0 never, 1 where it measured faster (default), 2 wherever the shape allows."""
        self.call("ctx_set_pair_mode", int(mode))

    def set_option(self, name, value):
        """Tuning switch by name (see dpig_ctx_set_option); results are identical under every setting."""
        self.call("ctx_set_option", name.encode(), int(value))

    def launch_count(self):
        """Kernels launched through this context, including the kernels of replayed CUDA graphs (the library counts a
        launch when it is issued or captured; every replay adds the captured count, see engine.StepGraph)."""
        return int(self.lib.dpig_launch_count(self.handle)) + self.replayed_launches

    def call(self, name, *args):
        rc = getattr(self.lib, "dpig_" + name)(self.handle, *args)
        if rc != OK:
            raise DpigError("dpig_%s failed (%d): %s" % (name, rc, self.last_error()))

    def __getattr__(self, name):
        if name.startswith("_") or ("dpig_" + name) not in _SIGNATURES:
            raise AttributeError(name)
        return lambda *a: self.call(name, *a)
