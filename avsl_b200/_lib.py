"""ctypes binding of libavfe.so (include/avfe.h).  No CPU fallback: if the library is missing
or CUDA is unavailable every entry point raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libavfe.so"
_lib = None

AVFE_F32, AVFE_F16, AVFE_BF16 = 0, 1, 2
FUSE_CONCAT, FUSE_SUM, FUSE_WSUM = 0, 1, 2

# name -> (restype, argtypes); mirrors include/avfe.h one to one
_SIGNATURES = {
    "avfe_version": (c_int, []),
    "avfe_strerror": (c_char_p, [c_int]),
    "avfe_launch_count": (c_uint64, []),
    "avfe_pad_or_trim_f32": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "avfe_pad_or_trim_ragged_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "avfe_peak_normalize_f32": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "avfe_logmel_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int]),
    "avfe_logmel_f32": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p,
                                c_void_p, c_size_t, c_void_p]),
    "avfe_logmel_pack_bytes": (c_size_t, []),
    "avfe_logmel_prepare": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "avfe_logmel_prepared_f32": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_size_t, c_void_p]),
    "avfe_logmel_ragged_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_size_t, c_void_p]),
    "avfe_spec_mask_f32": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_int, c_float, c_void_p]),
    "avfe_spec_time_warp_f32": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "avfe_add_noise_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "avfe_add_noise": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p,
                               c_void_p, c_void_p, c_size_t, c_void_p]),
    "avfe_logfbank_num_frames": (c_int64, [c_int64]),
    "avfe_logfbank_workspace_bytes": (c_size_t, []),
    "avfe_logfbank_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int, c_int,
                                  c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "avfe_logfbank_prepare": (c_int, [c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "avfe_logfbank_prepared_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int, c_int,
                                  c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "avfe_bgr2gray_u8": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "avfe_warp_affine_u8": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p,
                                    c_void_p]),
    "avfe_lip_workspace_bytes": (c_size_t, [c_int64]),
    "avfe_lip_roi_batch": (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_int64,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                   c_int, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_size_t, c_void_p]),
    "avfe_lip_roi_collate": (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_int64,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                     c_int, c_float, c_float, c_void_p, c_int64, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_size_t, c_void_p]),
    "avfe_landmarks_interpolate": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p,
                                           c_void_p]),
    "avfe_similarity_fit": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "avfe_cut_patch_u8": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p,
                                  c_void_p, c_void_p]),
    "avfe_video_feats_u8": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_float, c_float,
                                    c_void_p, c_void_p]),
    "avfe_video_feats_workspace_bytes": (c_size_t, []),
    "avfe_video_feats": (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_int, c_double, c_double, c_void_p,
                                 c_void_p, c_size_t, c_void_p]),
    "avfe_fuse": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_int, c_int64,
                          c_int64, c_int64, c_void_p, c_void_p]),
    "avfe_fuse_layernorm": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_int, c_int64,
                                    c_int64, c_int64, c_void_p, c_void_p, c_float, c_void_p, c_void_p]),
    "avfe_fuse_layernorm_tma_ok": (c_int, [c_int, c_int64, c_int64, c_int64]),
    "avfe_fuse_layernorm_pitched": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_int, c_int64,
                                            c_int64, c_int64, c_int64, c_void_p, c_void_p, c_float, c_void_p, c_void_p]),
    "avfe_proj_fold_bytes": (c_size_t, [c_int64, c_int64]),
    "avfe_proj_fold": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "avfe_fuse_ln_proj_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "avfe_fuse_ln_proj": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                  c_int64, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "avfe_fuse_backward": (c_int, [c_void_p, c_void_p, c_int, c_float, c_float, c_int, c_int64, c_int64, c_int64,
                                   c_void_p, c_void_p, c_void_p]),
    "avfe_fuse_layernorm_backward_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int]),
    "avfe_fuse_layernorm_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_int, c_int64,
                                             c_int64, c_int64, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class AvfeError(RuntimeError):
    """A libavfe entry point returned a non-zero avfe_status."""

    def __init__(self, fn: str, status: int, message: str):
        super().__init__(f"{fn} failed: {message} (status {status})")
        self.status = status


def lib_path() -> Path:
    return _LIB_PATH


def load() -> ctypes.CDLL:
    """Load libavfe.so (once).  Raises if it has not been built — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise RuntimeError(
            f"{_LIB_PATH} is missing: build it with `python -m avsl_b200.build` "
            "(nvcc, sm_100a).  avsl_b200 has no CPU fallback.")
    lib = ctypes.CDLL(os.fspath(_LIB_PATH))
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(fn: str, status: int) -> None:
    if status != 0:
        msg = load().avfe_strerror(status).decode()
        raise AvfeError(fn, status, msg)


def call(fn: str, *args) -> None:
    check(fn, getattr(load(), fn)(*args))


def launch_count() -> int:
    return int(load().avfe_launch_count())


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("avsl_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def ptr(t) -> c_void_p:
    """Device pointer of a torch tensor (or NULL for None)."""
    return c_void_p(0 if t is None else t.data_ptr())


def stream_ptr() -> c_void_p:
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
