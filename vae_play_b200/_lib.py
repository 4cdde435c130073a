"""ctypes binding of libvaeplay_b200.so (include/vaeplay_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc; if that is impossible the
import raises.  Every compute entry point takes raw device pointers and the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvaeplay_b200.so")

F32, BF16 = 0, 1
ACT = {None: 0, "none": 0, "relu": 1, "lrelu": 2, "tanh": 3, "sigmoid": 4}
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC = 0, 1, 2


class VpConvGeom(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n", "hi", "wi", "ci", "ho", "wo", "co", "kh", "kw", "stride", "pad", "transposed")]


class VaePlayError(RuntimeError):
    pass


_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_u64 = C.c_uint64
_f = C.c_float

# name -> argtypes, exactly as declared in include/vaeplay_b200.h
SIGNATURES = {
    "vp_pack_weight": [_p, _p, _i, _i, _i, _i, _i64, _i64, _i64, _p],
    "vp_unpack_wgrad": [_p, _p, _i, _i, _i, _i64, _i64, _i64, _p],
    "vp_conv_fwd": [C.POINTER(VpConvGeom), _p, _p, _p, _p, _i, _i, _i, _f, _i, _p],
    "vp_conv_dgrad": [C.POINTER(VpConvGeom), _p, _p, _p, _i, _i, _i, _p],
    "vp_conv_wgrad": [C.POINTER(VpConvGeom), _p, _p, _p, _i, _i, _p],
    "vp_norm_stats": [_p, _p, _i, _i64, _i64, _i, _p],
    "vp_norm_finalize": [_p, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _i64, _i64, _i, _p],
    "vp_norm_apply_act": [_p, _p, _p, _p, _i, _i64, _i64, _i, _i, _f, _p],
    "vp_norm_bwd_reduce": [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i64, _i64, _i, _i, _f, _p],
    "vp_norm_bwd_apply": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i64, _i64, _i, _i, _f, _p],
    "vp_bn_rows_fwd": [_p, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p, _i, _i64, _i, _i, _f, _p],
    "vp_bn_rows_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i64, _i, _i, _f, _p],
    "vp_colsum": [_p, _p, _p, _i, _i64, _i, _p],
    "vp_reparam_kl_fwd": [_p, _p, _i64, _p, _u64, _u64, _p, _i, _p, _i, _p, _p, _i64, _i, _p],
    "vp_reparam_kl_bwd": [_p, _p, _i64, _p, _p, _i, _p, _p, _p, _i, _i64, _i64, _i, _p],
    "vp_philox_normal": [_p, _i64, _u64, _u64, _p, _i, _p],
    "vp_philox_advance": [_p, _u64, _p],
    "vp_recon_loss_fwd": [_p, _p, _i64, _i, _p, _p, _p, _p],
    "vp_recon_loss_bwd": [_p, _p, _i64, _i, _p, _p, _p],
    "vp_bce_dice_fwd": [_p, _p, _i64, _i64, _f, _p, _p, _p, _p],
    "vp_bce_dice_bwd": [_p, _p, _i64, _i64, _f, _p, _p, _p, _p],
    "vp_nchw_to_nhwc": [_p, _p, _i, _i, _i, _i, _i, _p],
    "vp_nhwc_to_nchw": [_p, _p, _i, _i, _i, _i, _i, _p],
    "vp_cast": [_p, _i, _p, _i, _i64, _p],
    "vp_axpy": [_f, _p, _p, _i64, _p],
    "vp_sum_into": [_p, _i64, _f, _p, _p],
    "vp_fill_from": [_p, _f, _p, _i64, _p],
    "vp_debug_umma_probe": [_p, _p, _p, _i, _i, _i, _p],
    "vp_set_workspace": [_p, C.c_size_t],
    "vp_conv_fwd_cl": [C.POINTER(VpConvGeom), _p, _p, _p, _p, _i, _i, _f, _p],
    "vp_conv_fwd_cl_stats": [C.POINTER(VpConvGeom), _p, _p, _p, _p, _i, C.POINTER(C.c_int), _p],
    "vp_thin_conv_fwd_stats": [C.POINTER(VpConvGeom), _p, _p, _p, _p, _i, C.POINTER(C.c_int), _p],
    "vp_norm_finalize_parts": [_p, _i, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _i64, _i, _p],
    "vp_conv_dgrad_cl": [C.POINTER(VpConvGeom), _p, _p, _p, _i, _p],
    "vp_conv_wgrad_cl": [C.POINTER(VpConvGeom), _p, _p, _p, _i, _p],
    "vp_transpose_bt": [_p, _p, _i, _i, _i, _i, _p],
    "vp_thin_conv_fwd": [C.POINTER(VpConvGeom), _p, _p, _p, _p, _i, _i, _f, _p],
    "vp_thin_conv_dgrad": [C.POINTER(VpConvGeom), _p, _p, _p, _i, _p],
    "vp_thin_conv_wgrad": [C.POINTER(VpConvGeom), _p, _p, _p, _i, _p],
    "vp_rmsprop_step": [_p, _p, _p, _p, _i, _f, _f, _f, _f, _p],
    "vp_rmsprop_step_shadow": [_p, _p, _p, _p, _p, _i, _f, _f, _f, _f, _i, _p],
}
PLAIN = {"vp_last_error": (C.c_char_p, []), "vp_abi_version": (_i, []), "vp_device_arch": (_i, []),
         "vp_launch_count": (_u64, [])}

_lib = None


def load(build_if_missing: bool = True):
    """Load (building first if needed) the CUDA library.  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise VaePlayError(f"{LIB_PATH} is missing; run `python -m vae_play_b200.build`")
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _i
    for name, (res, args) in PLAIN.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    if lib.vp_abi_version() != 1:
        raise VaePlayError("libvaeplay_b200.so ABI version mismatch; rebuild with `python -m vae_play_b200.build --force`")
    _lib = lib
    return lib


_workspace = {}


def ensure_workspace(device, nbytes: int = 32 << 20):
    """Register a per-device scratch buffer for split-K partial sums (vp_set_workspace)."""
    import torch
    key = (device.type, device.index)
    if key not in _workspace:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _workspace[key] = buf
        call("vp_set_workspace", C.c_void_p(buf.data_ptr()), nbytes)
    return _workspace[key]


def check(rc: int, what: str):
    if rc != 0:
        msg = load().vp_last_error()
        raise VaePlayError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def call(name: str, *args):
    check(getattr(load(), name)(*args), name)


def launch_count() -> int:
    return int(load().vp_launch_count())
