"""ctypes binding of libtauv_b200.so — the only way any wrapper in this package reaches the GPU.

There is deliberately no fallback: if the shared object is missing or the tensors are not CUDA
fp32 on an sm_100 device, the call raises.  (The CPU oracle under ``oracle/`` is test
infrastructure and is never imported from here.)
"""
from __future__ import annotations

import ctypes
import os
import re
import threading
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_void_p
from pathlib import Path

import torch

from . import _build

_LIB = None
_LOCK = threading.Lock()

# Error codes of include/tauv_b200.h
E_NULL, E_SHAPE, E_K_RANGE, E_KERNEL, E_WORKSPACE, E_UNSUPPORTED, E_ARCH, E_ALIGN = range(-1, -9, -1)

_F = POINTER(c_float)
_D = POINTER(c_double)
_I64 = POINTER(c_int64)
_I32 = POINTER(c_int32)
_U8 = POINTER(c_uint8)

# name -> (restype, argtypes); mirrors include/tauv_b200.h one-to-one.
_SIGNATURES = {
    "tauv_version": (c_int, []),
    "tauv_last_error": (c_char_p, []),
    "tauv_check_device": (c_int, []),
    "tauv_heatmap_nms": (c_int, [_F, _F, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "tauv_heatmap_topk_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "tauv_heatmap_topk": (c_int, [_F, c_int, c_int, c_int, c_int, c_int, c_int, _I64, _I64, _F, c_void_p, c_size_t,
                                  c_void_p]),
    "tauv_centernet_boxes": (c_int, [_I64, _F, c_int, c_int, c_int, c_int, _F, _I64, _F, _I64, _F, _I64, c_int, c_int,
                                     c_int, c_int, c_int, c_int, c_float, _D, _F, _F, _I32, c_void_p]),
    "tauv_centernet_decode": (c_int, [_F, c_int, c_int, c_int, c_int, c_int, _F, _I64, _F, _I64, _F, _I64, c_int,
                                      c_int, c_int, c_int, c_float, _I64, _I64, _F, _D, _F, _F, _I32, c_void_p,
                                      c_size_t, c_void_p]),
    "tauv_centernet_block_maxima": (c_int, [_F, c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "tauv_heatmap_topk_stage1": (c_int, [_F, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "tauv_centernet_decode_stage2": (c_int, [c_int, c_int, c_int, c_int, c_int, _F, _I64, _F, _I64, _F, _I64, c_int,
                                             c_int, c_int, c_int, c_float, _I64, _I64, _F, _D, _F, _F, _I32, c_void_p,
                                             c_size_t, c_void_p]),
    "tauv_gather_at": (c_int, [_F, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, _I64, _I64, c_int, c_int, _F,
                               c_void_p]),
    "tauv_scatter_add_at": (c_int, [_F, _I64, c_int, c_int, c_int, _F, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "tauv_centernet_keypoint_assoc": (c_int, [_I64, _D, _I32, c_int, c_int, _I64, _I64, _F, c_int, _F, _I64, _I32, c_int,
                                              c_int, c_int, c_int, c_double, _U8, _F, _F, _F, c_void_p]),
    "tauv_angle_decode": (c_int, [_F, _F, c_int64, c_double, _F, c_void_p]),
    "tauv_depth_decode": (c_int, [_F, c_int64, _F, c_void_p]),
    "tauv_gaussian_encode": (c_int, [_U8, _I64, _F, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_double,
                                     _F, c_void_p]),
    "tauv_centernet_focal_loss_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "tauv_centernet_focal_loss": (c_int, [_F, _U8, _I64, _F, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                          c_double, c_double, c_double, _D, _I64, c_void_p, c_size_t, c_void_p]),
    "tauv_centernet_focal_loss_reduce": (c_int, [_D, _I64, c_int, _F, _I64, c_void_p]),
    "tauv_centernet_focal_loss_backward": (c_int, [_F, _U8, _I64, _F, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                                   c_int, c_double, c_double, c_double, _I64, _F, _F, c_void_p]),
    "tauv_keypoint_encode": (c_int, [_U8, _I64, _F, _I64, _F, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_double, c_double, _F, _F, _F, c_void_p]),
    "tauv_out_index_offset": (c_int, [_F, c_int64, c_int, c_int, c_int, c_int, c_int, _I64, _F, c_void_p]),
    "tauv_gaussian_splat": (c_int, [c_int, c_int, c_int, c_int, c_double, _F, c_void_p]),
    "tauv_yolact_anchors": (c_int, [POINTER(c_int), POINTER(c_int), c_int, c_int, _F, _F, c_void_p]),
    "tauv_yolact_box_decode": (c_int, [_F, _F, c_int, c_int, c_int, c_float, c_float, _F, c_void_p]),
    "tauv_yolact_box_encode": (c_int, [_F, _F, c_int, c_int, c_int, c_float, c_float, _F, c_void_p]),
    "tauv_iou_matrix": (c_int, [_F, _F, c_int, c_int, c_int, c_int, _F, c_void_p]),
    "tauv_yolact_scores": (c_int, [_F, c_int, c_int, c_int, _F, _I32, c_void_p]),
    "tauv_yolact_nms_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "tauv_yolact_fast_nms": (c_int, [_F, _F, c_int, c_int, c_int, c_int, c_int, c_float, c_float, _I64, _I32,
                                     c_void_p, c_size_t, c_void_p]),
    "tauv_yolact_detect": (c_int, [_F, _F, _F, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_float, c_float,
                                   _I64, _I32, _F, _F, _I32, c_void_p, c_size_t, c_void_p]),
    "tauv_yolact_assemble_mask": (c_int, [_F, _F, _F, c_int, c_int, c_int, c_int, _F, _F, c_void_p]),
    "tauv_yolact_assemble_mask_batched": (c_int, [_F, _F, _I64, _I32, _F, c_int, c_int, c_int, c_int, c_int, c_int,
                                                  _F, c_void_p]),
    "tauv_yolact_mask_depth_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "tauv_yolact_mask_depth": (c_int, [_F, _F, _F, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, _D, _I64,
                                       c_void_p, c_size_t, c_void_p]),
    "tauv_yolact_mask_depth_batched": (c_int, [_F, _F, _I64, _I32, _F, c_int, c_int, c_int, c_int, c_int, c_int,
                                               c_void_p, c_int, c_int, _D, _I64, c_void_p, c_size_t, c_void_p]),
    "tauv_yolact_mask_binary_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "tauv_yolact_mask_binary": (c_int, [_F, _F, _F, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _U8, c_void_p,
                                        c_size_t, c_void_p]),
    "tauv_yolact_mask_binary_batched": (c_int, [_F, _F, _I64, _I32, _F, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                                c_int, c_int, _U8, c_void_p, c_size_t, c_void_p]),
    "tauv_box_to_mask": (c_int, [_F, c_int, c_int, _F, c_void_p]),
    "tauv_yolact_match_anchors": (c_int, [_F, _F, _U8, c_int, c_int, c_int, c_float, c_float, c_float, c_float, _I64,
                                          _F, _U8, _U8, _F, c_void_p]),
    "tauv_yolact_class_box_loss_workspace_bytes": (c_size_t, [c_int, c_int]),
    "tauv_yolact_class_box_loss": (c_int, [_F, _F, _F, _U8, _U8, _I64, _I64, c_int, c_int, c_int, c_int, c_int, _U8, _I32,
                                           _D, _I64, c_void_p, c_size_t, c_void_p]),
    "tauv_yolact_class_box_loss_backward": (c_int, [_F, _F, _F, _U8, _U8, _I64, _I64, c_int, c_int, c_int, c_int, c_int,
                                                    _I64, _F, _F, _F, _F, c_void_p]),
    "tauv_keypoint_affinity_loss_partials": (c_size_t, [c_int, c_int, c_int, c_int]),
    "tauv_keypoint_affinity_loss": (c_int, [_F, _U8, _I64, _F, _I64, _F, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                            c_int, c_int, c_double, _D, c_void_p]),
    "tauv_keypoint_affinity_loss_backward": (c_int, [_F, _U8, _I64, _F, _I64, _F, c_int, c_int, c_int, c_int, c_int, c_int,
                                                     c_int, c_int, c_int, c_double, _F, _F, c_void_p]),
    "tauv_yolact_pack_heads": (c_int, [POINTER(c_void_p), _I32, c_int, c_int, c_int, c_int, _F, c_void_p]),
    "tauv_yolact_pack_heads_backward": (c_int, [_F, _F, _I32, c_int, c_int, c_int, c_int, POINTER(c_void_p), c_void_p]),
    "tauv_yolact_loss_reduce": (c_int, [_D, _I64, c_int, c_int, _D, c_int, _F, _I64, c_void_p]),
    "tauv_yolact_mask_loss_partials": (c_int, []),
    "tauv_yolact_mask_loss_records_bytes": (c_size_t, [c_int, c_int]),
    "tauv_yolact_mask_loss": (c_int, [_F, _F, _I32, _I64, _I64, _F, c_void_p, c_int, _U8, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_int, c_int, _D, c_void_p, _D, c_void_p]),
    "tauv_yolact_mask_loss_backward": (c_int, [_F, _F, _I32, _I64, _I64, _F, c_void_p, c_int, _U8, c_int, c_int, c_int, c_int, c_int,
                                               c_int, c_int, c_int, _D, c_void_p, _I64, _F, _F, _F, c_void_p]),
}


def header_symbols() -> list[str]:
    """Every function the public header declares (used by the symbol-export test)."""
    text = (Path(__file__).resolve().parent.parent / "include" / "tauv_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tauv_[a-z0-9_]+)\s*\(", text)))


def lib_path() -> Path:
    return _build.LIB_PATH


def load(build_if_missing: bool = True) -> ctypes.CDLL:
    """Load (building in-tree first if needed) the shared object and set every prototype."""
    global _LIB
    if _LIB is not None:
        return _LIB
    with _LOCK:
        if _LIB is not None:
            return _LIB
        path = _build.LIB_PATH
        if not _build.is_fresh() and os.environ.get("TAUV_SKIP_BUILD_CHECK") != "1":
            # a missing library, or one built from other sources / flags than those in the tree (the stamp next to it
            # records what it was built from): stale kernels must never pass for the current ones
            if not build_if_missing:
                raise RuntimeError(f"{path} is missing or stale and building was disabled; there is no fallback path")
            _build.build()
        lib = ctypes.CDLL(str(path))
        missing = []
        for name, (res, args) in _SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                missing.append(name)
                continue
            fn.restype = res
            fn.argtypes = args
        if missing:  # a stale / partial .so must never be papered over
            raise RuntimeError(f"{path} does not export {missing}; rebuild with tauv_vision_b200.build(force=True)")
        _LIB = lib
    return _LIB


class TauvError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libtauv_b200 error {code}: {message}")
        self.code = code
        self.message = message


def check(code: int) -> None:
    if code == 0:
        return
    msg = load().tauv_last_error().decode("utf-8", "replace")
    if code == E_KERNEL:
        raise AssertionError(msg)  # reference: `assert kernel_size >= 1 and kernel_size % 2 == 1`
    if code == E_K_RANGE:
        raise RuntimeError(msg)  # reference: torch.topk "selected index k out of range"
    raise TauvError(code, msg)


# ---- tensor helpers ---------------------------------------------------------------------------

def require_cuda(*tensors: torch.Tensor) -> torch.device:
    """All tensors must live on one CUDA device; no CPU path exists."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "tauv_vision_b200 runs on CUDA (sm_100a) tensors only; got a tensor on "
                f"'{t.device}'. There is no CPU fallback.")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} vs {t.device}")
    if dev is None:
        raise RuntimeError("no tensor arguments")
    return dev


def f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32 + contiguous (no copy when already so)."""
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t if t.is_contiguous() else t.contiguous()


def ptr(t, ctype):
    if t is None:
        return ctypes.cast(None, ctype)
    return ctypes.cast(t.data_ptr(), ctype)


def fptr(t):
    return ptr(t, _F)


def dptr(t):
    return ptr(t, _D)


def i64ptr(t):
    return ptr(t, _I64)


def i32ptr(t):
    return ptr(t, _I32)


def u8ptr(t):
    return ptr(t, _U8)


def strides_arg(t: torch.Tensor, n: int):
    s = list(t.stride())[:n]
    while len(s) < n:
        s.append(0)
    return (c_int64 * n)(*s)


def stream_ptr(dev: torch.device) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(dev).cuda_stream)


_WS_CACHE: dict = {}


def workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    """Per-(thread, device, stream) scratch buffer, grown on demand; owned by torch's allocator."""
    key = (threading.get_ident(), dev.index, torch.cuda.current_stream(dev).cuda_stream)
    buf = _WS_CACHE.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=dev)
        _WS_CACHE[key] = buf
    return buf
