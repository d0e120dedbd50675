"""``patch_reference()`` — point an installed TAUV-Vision at the B200 kernels without editing it.

The reference has no plugin or operator registry: callers bind module-level functions by name at import
(``from tauv_vision.centernet.model.decode import decode_keypoints`` — centernet/node/centernet_node.py:21;
``from tauv_vision.yolact.model.nms import nms`` — yolact/node/yolact_node.py:24-26;
``from tauv_vision.yolact.model.boxes import box_encode, iou_matrix, box_to_mask`` — yolact/model/loss.py:4).
So a drop-in has to do two things: replace the attribute on the defining module, and replace every copy of
the old function object that an already-imported caller holds in its globals.  This module does both and
returns a handle that undoes it.
"""
from __future__ import annotations

import importlib
import sys
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Tuple

# reference module -> (our module, names that are replaced)
PATCH_TABLE: Dict[str, Tuple[str, Tuple[str, ...]]] = {
    "tauv_vision.centernet.model.decode": (
        "tauv_vision_b200.centernet.model.decode",
        ("heatmap_nms", "heatmap_detect", "decode", "decode_keypoints", "angle_get_bins", "angle_decode",
         "depth_decode", "Detection", "KeypointDetection")),
    "tauv_vision.centernet.model.loss": (
        "tauv_vision_b200.centernet.model.loss",
        ("generate_heatmap", "generate_keypoint_heatmap", "out_index_for_position", "gaussian_splat")),
    "tauv_vision.yolact.model.boxes": (
        "tauv_vision_b200.yolact.model.boxes",
        ("box_xy_swap", "box_to_corners", "corners_to_box", "box_encode", "box_decode", "iou_matrix",
         "box_to_mask")),
    "tauv_vision.yolact.model.nms": ("tauv_vision_b200.yolact.model.nms", ("nms",)),
    "tauv_vision.yolact.model.masks": ("tauv_vision_b200.yolact.model.masks", ("assemble_mask",)),
    "tauv_vision.yolact.model.anchors": ("tauv_vision_b200.yolact.model.anchors", ("get_anchor",)),
    "tauv_vision.yolact.model.loss": ("tauv_vision_b200.yolact.model.loss", ("loss",)),   # yolact/scripts/train.py:14,246
}


_ABSENT = object()  # the reference module had no such attribute (gaussian_splat is missing from the snapshot)


@dataclass
class PatchHandle:
    """What was replaced where; ``undo()`` restores every binding."""
    replaced: List[Tuple[object, str, object]] = field(default_factory=list)  # (namespace owner, name, old value)
    skipped: List[str] = field(default_factory=list)                          # reference modules not importable

    def undo(self) -> None:
        for owner, name, old in reversed(self.replaced):
            if old is _ABSENT:
                delattr(owner, name)
            else:
                setattr(owner, name, old)
        self.replaced.clear()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.undo()


def _rebind_callers(old, new, handle: PatchHandle, skip: Iterable[object]) -> None:
    """Replace ``old`` by identity in the globals of every imported module that copied it."""
    skip_ids = {id(m) for m in skip}
    for mod in list(sys.modules.values()):
        if mod is None or id(mod) in skip_ids:
            continue
        d = getattr(mod, "__dict__", None)
        if not isinstance(d, dict):
            continue
        for name, val in list(d.items()):
            if val is old:
                handle.replaced.append((mod, name, old))
                setattr(mod, name, new)


def patch_reference(modules: Optional[Iterable[str]] = None, strict: bool = False,
                    rebind_callers: bool = True) -> PatchHandle:
    """Replace the reference's hot-path functions with the CUDA-backed ones.

    modules        : subset of ``PATCH_TABLE`` keys (default: all).
    strict         : raise if a reference module cannot be imported (default: record it in ``handle.skipped`` —
                     ``tauv_vision.centernet.model.*`` needs matplotlib and spatialmath at import time).
    rebind_callers : also replace copies held by already-imported modules (``from x import f`` bindings).

    Names beyond the reference's (``decode_packed``, ``nms_batched``, ``detect``, ``assemble_mask_batched`` ...)
    are not injected: a patched reference exposes exactly its own API.
    """
    handle = PatchHandle()
    for ref_name in (modules if modules is not None else PATCH_TABLE):
        ours_name, names = PATCH_TABLE[ref_name]
        try:
            ref_mod = importlib.import_module(ref_name)
        except Exception as e:  # noqa: BLE001 - optional dependencies of the reference
            if strict:
                raise
            handle.skipped.append(f"{ref_name}: {type(e).__name__}: {e}")
            continue
        ours = importlib.import_module(ours_name)
        for name in names:
            new = getattr(ours, name)
            old = getattr(ref_mod, name, _ABSENT)
            handle.replaced.append((ref_mod, name, old))
            setattr(ref_mod, name, new)
            if rebind_callers and old is not _ABSENT and old is not new:
                _rebind_callers(old, new, handle, skip=(ref_mod, ours))
    return handle
