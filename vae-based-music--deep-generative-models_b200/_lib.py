"""ctypes binding of csrc/libvqvae_b200.so (C ABI: include/vqb.h).

There is NO CPU fallback: if the shared library has not been built, or the device is not sm_100, every call
raises.  (`tests/fake_backend.py` can inject a CPU stand-in through `set_backend` so that the host-side logic —
tape, variable ordering, sharding arithmetic — is unit-testable without a GPU; nothing in the package does.)
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# VQB_LIB: another build of the same library (A/B timing of two kernel variants on one GPU box); default = the in-tree build
LIB_PATH = os.environ.get("VQB_LIB") or os.path.join(_HERE, "csrc", "libvqvae_b200.so")

PREC_FP32, PREC_TF32, PREC_BF16, PREC_BF16X2, PREC_BF16X3, PREC_FP16X2 = 0, 1, 2, 3, 4, 5
PRECISIONS = {"fp32": PREC_FP32, "tf32": PREC_TF32, "bf16": PREC_BF16, "bf16x2": PREC_BF16X2, "bf16x3": PREC_BF16X3,
              "fp16x2": PREC_FP16X2}


class VQBError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("B", "L", "C_in", "C_out", "k", "stride", "dilation", "relu_in", "precision")]


class ResblockDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "L", "C", "F", "dilation", "precision")]


class ResstackDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("L", C.c_int32), ("C", C.c_int32), ("n_blocks", C.c_int32),
                ("dilations", C.c_int32 * 4), ("precision", C.c_int32)]


class VQDesc(C.Structure):
    _fields_ = [("N", C.c_int64), ("D", C.c_int32), ("K", C.c_int32), ("beta", C.c_float),
                ("precision", C.c_int32)]


class TailDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "L", "C_in", "C_mid")]


TAIL_GBUF = 272  # VQB_TAIL_GBUF

_P = C.c_void_p
_CD, _RD, _VD, _TD = C.POINTER(ConvDesc), C.POINTER(ResblockDesc), C.POINTER(VQDesc), C.POINTER(TailDesc)
_SD = C.POINTER(ResstackDesc)

# name -> (restype, argtypes); mirrors include/vqb.h one to one
SIGNATURES = {
    "vqb_version": (C.c_int, []),
    "vqb_last_error": (C.c_char_p, []),
    "vqb_kernel_launch_count": (C.c_int64, []),
    "vqb_device_check": (C.c_int, [C.c_int]),
    "vqb_conv1d_supports": (C.c_int, [_CD, C.c_int]),
    "vqb_conv1d_transpose_supports": (C.c_int, [_CD, C.c_int]),
    "vqb_reduce_begin": (C.c_int, []),
    "vqb_reduce_flush": (C.c_int, [_P]),
    "vqb_conv1d_fwd": (C.c_int, [_CD, _P, _P, _P, _P, _P, _P]),
    "vqb_conv1d_dgrad": (C.c_int, [_CD, _P, _P, _P, _P, _P, _P]),
    "vqb_conv1d_wgrad_workspace_bytes": (C.c_size_t, [_CD]),
    "vqb_conv1d_wgrad": (C.c_int, [_CD, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "vqb_conv1d_transpose_fwd": (C.c_int, [_CD, _P, _P, _P, _P, _P]),
    "vqb_conv1d_transpose_dgrad": (C.c_int, [_CD, _P, _P, _P, _P]),
    "vqb_conv1d_transpose_wgrad_workspace_bytes": (C.c_size_t, [_CD]),
    "vqb_conv1d_transpose_wgrad": (C.c_int, [_CD, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "vqb_resblock_supports": (C.c_int, [_RD]),
    "vqb_resblock_fwd": (C.c_int, [_RD, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vqb_resblock_bwd_data": (C.c_int, [_RD, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vqb_resblock_fwd_masks": (C.c_int, [_RD, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vqb_resblock_bwd_data_masks": (C.c_int, [_RD, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vqb_resblock_wgrad_workspace_bytes": (C.c_size_t, [_RD]),
    "vqb_resblock_wgrad": (C.c_int, [_RD, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "vqb_resblock_wgrad_batch_workspace_bytes": (C.c_size_t, [_RD, C.c_int32]),
    "vqb_resblock_wgrad_batch": (C.c_int, [_RD, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "vqb_resstack_supports": (C.c_int, [_SD]),
    "vqb_resstack_workspace_bytes": (C.c_size_t, [_SD]),
    "vqb_resstack_fwd": (C.c_int, [_SD, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "vqb_resstack_fwd_private_ws": (C.c_int, [_SD, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "vqb_resstack_bwd_data": (C.c_int, [_SD, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "vqb_resstack_bwd_data_packed": (C.c_int, [_SD, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "vqb_dec_tail_supports": (C.c_int, [_TD]),
    "vqb_dec_tail_fwd": (C.c_int, [_TD, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vqb_dec_tail_bwd_workspace_bytes": (C.c_size_t, [_TD]),
    "vqb_dec_tail_bwd": (C.c_int, [_TD, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "vqb_vq_fwd_workspace_bytes": (C.c_size_t, [_VD]),
    "vqb_vq_fwd": (C.c_int, [_VD, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "vqb_vq_fwd_bf16": (C.c_int, [_VD, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "vqb_vq_bwd": (C.c_int, [_VD, _P, _P, _P, C.c_float, _P, _P]),
    "vqb_vq_ema_update": (C.c_int, [C.c_int32, C.c_int32, C.c_double, C.c_float, _P, _P, _P, _P, _P, _P, _P, _P]),
    "vqb_gather_rows": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_int32, C.c_int64, C.c_int64, _P, _P]),
    "vqb_restart_ids": (C.c_int, [C.c_int64, C.c_int32, C.c_uint64, _P, _P, _P]),
    "vqb_gather_codes": (C.c_int, [_P, C.c_int32, C.c_int32, _P, C.c_int64, _P, _P]),
    "vqb_stft_frames": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "vqb_spec_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "vqb_spec_mag": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, C.c_size_t, _P]),
    "vqb_spec_diff": (C.c_int, [_P, _P, C.c_int64, C.c_int64, _P, _P, C.c_size_t, _P]),
    "vqb_spec_loss": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    "vqb_spec_grad": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P, _P]),
    "vqb_stft_frames_bwd": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "vqb_reduce_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "vqb_mse": (C.c_int, [_P, _P, C.c_int64, C.c_float, _P, _P, _P, _P, C.c_size_t, _P]),
    "vqb_adam_step": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float,
                                C.c_float, _P, _P]),
    "vqb_adam_step_dev": (C.c_int, [_P, _P, _P, _P, C.c_int64, _P, C.c_float, C.c_float, C.c_float, C.c_float, _P, _P]),
    "vqb_increment": (C.c_int, [_P, _P]),
    "vqb_lincomb": (C.c_int, [C.c_int32, _P, _P, _P, _P, _P]),
}

_BACKEND = None  # ctypes.CDLL, or an injected test double
_DEVICE = None
_LAUNCHES = 0  # number of vqb_* compute calls issued (bench.py reports kernels via the C side counter too)


def load_library(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise VQBError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    return lib


def set_backend(backend, device):
    """Test seam only: install a stand-in for the shared library and the torch device buffers live on."""
    global _BACKEND, _DEVICE
    _BACKEND, _DEVICE = backend, (torch.device(device) if device is not None else None)


def lib():
    global _BACKEND, _DEVICE
    if _BACKEND is None:
        if not torch.cuda.is_available():
            raise VQBError("no CUDA device visible: libvqvae_b200 runs on B200 (sm_100a) only; there is no CPU fallback")
        backend = load_library()
        dev = torch.cuda.current_device()
        rc = backend.vqb_device_check(dev)
        if rc != 0:
            raise VQBError(backend.vqb_last_error().decode())
        _BACKEND, _DEVICE = backend, torch.device("cuda", dev)
    return _BACKEND


def device() -> torch.device:
    lib()
    return _DEVICE


def is_native() -> bool:
    return isinstance(_BACKEND, C.CDLL)


def stream() -> int:
    if _DEVICE is not None and _DEVICE.type == "cuda":
        return torch.cuda.current_stream(_DEVICE).cuda_stream
    return 0


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().vqb_last_error()
        raise VQBError(f"{what} failed with code {rc}: {msg.decode() if isinstance(msg, bytes) else msg}")


def ptr(t):
    return None if t is None else t.data_ptr()


def call(name: str, *args):
    global _LAUNCHES
    _LAUNCHES += 1
    check(getattr(lib(), name)(*args), name)


def launches() -> int:
    return _LAUNCHES
