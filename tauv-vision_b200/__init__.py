"""tauv_vision_b200 — B200-native (sm_100a) detection-head hot path of TAUV-Vision.

Sub-packages mirror the reference's module paths so that call sites only change their import
root (or call :func:`patch_reference` and change nothing):

    tauv_vision.centernet.model.decode  ->  tauv_vision_b200.centernet.model.decode
    tauv_vision.centernet.model.loss    ->  tauv_vision_b200.centernet.model.loss   (target encode half)
    tauv_vision.yolact.model.boxes      ->  tauv_vision_b200.yolact.model.boxes
    tauv_vision.yolact.model.nms        ->  tauv_vision_b200.yolact.model.nms
    tauv_vision.yolact.model.masks      ->  tauv_vision_b200.yolact.model.masks
    tauv_vision.yolact.model.anchors    ->  tauv_vision_b200.yolact.model.anchors
    tauv_vision.yolact.model.loss       ->  tauv_vision_b200.yolact.model.loss      (anchor matching half)

Everything computes in hand-written CUDA behind the C ABI of ``include/tauv_b200.h``; there is no
CPU path and no fallback.
"""
from . import _build, _lib  # noqa: F401
from ._lib import TauvError, load as load_library  # noqa: F401

__version__ = "0.1.0"


def build(force: bool = False):
    """Compile libtauv_b200.so in-tree (cross-compiles for sm_100a without a GPU)."""
    return _build.build(force=force)


def patch_reference(*args, **kwargs):
    from .patch import patch_reference as _p
    return _p(*args, **kwargs)
