"""CPU tests of the host-side logic: frame sharding + the world_size-2 gloo gather, and patch_reference()."""
import importlib
import os
import socket
import sys
import textwrap
import types

import numpy as np
from pathlib import Path

import pytest
import torch

from tauv_vision_b200 import shard
from tauv_vision_b200.patch import PATCH_TABLE, patch_reference


def test_frame_range_partitions_the_batch():
    for n in (0, 1, 7, 64, 256, 257):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard.frame_range(r, world, n) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))     # contiguous, in order
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert shard.frame_range(3, 8, 256) == (96, 128)                                   # config #4: 32 per GPU
    with pytest.raises(ValueError):
        shard.frame_range(2, 2, 4)


def test_concat_host_keeps_frame_order_and_none_fields():
    a = {"score": np.arange(6, dtype=np.float32).reshape(2, 3), "depth": None}
    b = {"score": np.arange(6, 9, dtype=np.float32).reshape(1, 3), "depth": None}
    out = shard.concat_host([a, None, b])
    assert out["depth"] is None and out["score"].shape == (3, 3)
    assert out["score"][:, 0].tolist() == [0.0, 3.0, 6.0]
    with pytest.raises(ValueError):
        shard.concat_host([a, {"score": b["score"], "depth": np.zeros((1, 3), np.float32)}])


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank: int, world: int, port: int, n_frames: int, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # the "decode" of this rank's frames: values that encode the global frame number
        frames = torch.arange(n_frames, dtype=torch.float32)[:, None].repeat(1, 4)
        mine = shard.shard_frames(frames, rank, world)
        lo, hi = shard.frame_range(rank, world, n_frames)
        local = {"score": mine.numpy() * 10.0, "count": np.arange(lo, hi, dtype=np.int32), "depth": None}
        got = shard.gather_host(local, dst=0)
        # the packed single-collective form (tensors in, one gather, one host copy on dst)
        packed = shard.gather_frames({"score": mine * 10.0, "count": torch.arange(lo, hi, dtype=torch.int32),
                                      "yx": mine.double()[:, :2].reshape(-1, 1, 2)}, n_frames, dst=0)
        if rank == 0:
            assert packed["score"][:, 0].tolist() == [10.0 * i for i in range(n_frames)]
            assert packed["count"].tolist() == list(range(n_frames)) and packed["count"].dtype == np.int32
            assert packed["yx"].shape == (n_frames, 1, 2) and packed["yx"].dtype == np.float64
        else:
            assert packed is None
        # timing reduction used by bench.py: max over ranks
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            q.put((got["score"][:, 0].tolist(), got["count"].tolist(), got["depth"], float(t)))
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world2_gloo_gather_is_in_frame_order():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, world, n = _free_port(), 2, 7
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    score0, count, depth, tmax = q.get(timeout=100)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert score0 == [10.0 * i for i in range(n)]
    assert count == list(range(n)) and depth is None and tmax == 2.0


def test_gather_host_without_process_group_is_identity():
    local = {"score": np.zeros((2, 3), np.float32)}
    assert shard.gather_host(local) is local
    one = shard.gather_frames({"score": torch.ones((2, 3))}, 2)
    assert one["score"].shape == (2, 3) and one["score"].dtype == np.float32


# ---- patch_reference -------------------------------------------------------------------------------------------

@pytest.fixture()
def fake_reference(tmp_path, monkeypatch):
    """A stand-in `tauv_vision` tree with the reference's module paths and function names (bodies irrelevant) plus
    a caller module that binds them by name at import, the way the nodes and scripts do."""
    root = tmp_path / "fake_ref"
    for ref_name, (_, names) in PATCH_TABLE.items():
        parts = ref_name.split(".")
        d = root
        for p in parts[:-1]:
            d = d / p
            d.mkdir(parents=True, exist_ok=True)
            (d / "__init__.py").touch()
        body = "\n".join(f"def {n}(*a, **k):\n    return 'reference {n}'\n" for n in names if n != "gaussian_splat")
        (d / (parts[-1] + ".py")).write_text(body)
    (root / "caller_node.py").write_text(textwrap.dedent("""
        from tauv_vision.yolact.model.nms import nms
        from tauv_vision.yolact.model.boxes import box_decode as bd
        from tauv_vision.centernet.model.decode import decode_keypoints
        def run():
            return nms, bd, decode_keypoints
    """))
    monkeypatch.syspath_prepend(str(root))
    before = set(sys.modules)
    yield root
    for name in set(sys.modules) - before:
        if name.startswith("tauv_vision.") or name in ("tauv_vision", "caller_node"):
            sys.modules.pop(name, None)


def test_patch_reference_replaces_modules_and_caller_bindings(fake_reference):
    caller = importlib.import_module("caller_node")
    ref_nms = importlib.import_module("tauv_vision.yolact.model.nms")
    ref_loss = importlib.import_module("tauv_vision.centernet.model.loss")
    assert caller.nms() == "reference nms" and not hasattr(ref_loss, "gaussian_splat")
    from tauv_vision_b200.centernet.model import decode as D
    from tauv_vision_b200.yolact.model import boxes as BX, nms as NM

    handle = patch_reference(strict=True)
    assert not handle.skipped
    assert ref_nms.nms is NM.nms
    assert caller.run() == (NM.nms, BX.box_decode, D.decode_keypoints)      # `from x import f` copies re-bound
    assert callable(ref_loss.gaussian_splat)                                 # the function the snapshot lacks
    for ref_name, (ours_name, names) in PATCH_TABLE.items():
        ref_mod, ours = importlib.import_module(ref_name), importlib.import_module(ours_name)
        assert all(getattr(ref_mod, n) is getattr(ours, n) for n in names)

    handle.undo()
    assert caller.nms() == "reference nms" and ref_nms.nms() == "reference nms"
    assert not hasattr(ref_loss, "gaussian_splat")


def test_patch_reference_without_the_reference_reports_skips():
    # (tauv_vision is not installed in the test environment)
    if "tauv_vision" in sys.modules or importlib.util.find_spec("tauv_vision") is not None:
        pytest.skip("a tauv_vision package is importable here")
    handle = patch_reference()
    assert len(handle.skipped) == len(PATCH_TABLE) and not handle.replaced
    with pytest.raises(ModuleNotFoundError):
        patch_reference(strict=True)


def test_patched_signatures_match_reference_positional_order():
    """Positional order of every public function = the reference's."""
    import inspect
    from tauv_vision_b200.centernet.model import decode as D, loss as L
    from tauv_vision_b200.yolact.model import boxes as BX, masks as MK, nms as NM, anchors as AN, loss as YL
    want = {
        YL.loss: ["prediction", "truth", "config"],
        D.heatmap_nms: ["heatmap", "kernel_size"],
        D.heatmap_detect: ["heatmap", "n_detections"],
        D.decode: ["prediction", "model_config", "n_detections", "score_threshold"],
        D.decode_keypoints: ["prediction", "model_config", "object_config", "M_projection", "n_detections",
                             "keypoint_n_detections", "score_threshold", "keypoint_score_threshold",
                             "keypoint_angle_threshold"],
        D.angle_decode: ["predicted_bin", "predicted_offset", "theta_range", "bin_overlap"],
        D.depth_decode: ["prediction"],
        L.generate_heatmap: ["truth", "model_config", "train_config", "object_config"],
        L.generate_keypoint_heatmap: ["truth", "model_config", "train_config", "object_config"],
        L.out_index_for_position: ["position", "model_config"],
        BX.box_encode: ["box", "anchor", "config"],
        BX.box_decode: ["box_encoding", "anchor", "config"],
        BX.iou_matrix: ["box_a", "box_b"],
        BX.box_to_mask: ["box", "img_size"],
        NM.nms: ["classification", "box", "top_k", "iou_threshold", "confidence_threshold"],
        MK.assemble_mask: ["mask_prototype", "mask_coeff", "box"],
        AN.get_anchor: ["fpn_i", "fpn_size", "config"],
    }
    for fn, names in want.items():
        got = list(inspect.signature(fn).parameters)[:len(names)]
        assert got == names, (fn.__name__, got)
        # anything beyond the reference's parameters must be optional
        extra = list(inspect.signature(fn).parameters.values())[len(names):]
        assert all(p.default is not inspect.Parameter.empty for p in extra), fn.__name__


REFERENCE_SRC = Path("/root/reference/src")


@pytest.mark.skipif(not REFERENCE_SRC.exists(), reason="the reference tree only exists in the build container")
def test_patch_reference_on_the_real_tree_keeps_cpu_callers_working(monkeypatch):
    """patch_reference() against the unmodified reference: the layout helpers it re-binds are called on CPU tensors by
    the reference's own data path (datasets/segmentation_dataset/segmentation_dataset.py:119, utils/plot.py:49,65,
    yolact/scripts/train.py:143-145) and must keep working there; the kernels-backed functions refuse CPU tensors."""
    import importlib
    import sys
    import torch
    monkeypatch.syspath_prepend(str(REFERENCE_SRC))
    for name in [m for m in sys.modules if m == "tauv_vision" or m.startswith("tauv_vision.")]:
        monkeypatch.delitem(sys.modules, name)
    ref_boxes = importlib.import_module("tauv_vision.yolact.model.boxes")
    original = ref_boxes.box_to_corners
    box = torch.tensor([[[0.5, 0.4, 0.2, 0.1], [0.25, 0.75, 0.5, 0.5]]])
    want_corners, want_swap = ref_boxes.box_to_corners(box), ref_boxes.box_xy_swap(box)
    with patch_reference(modules=["tauv_vision.yolact.model.boxes", "tauv_vision.yolact.model.nms"]) as handle:
        assert not handle.skipped
        assert ref_boxes.box_to_corners is not original
        assert torch.equal(ref_boxes.box_to_corners(box), want_corners)      # CPU in, CPU out, same numbers
        assert torch.equal(ref_boxes.box_xy_swap(box), want_swap)
        assert torch.equal(ref_boxes.corners_to_box(ref_boxes.box_to_corners(box)), ref_boxes.corners_to_box(want_corners))
        with pytest.raises(RuntimeError, match="CUDA"):
            ref_boxes.iou_matrix(box, box)                                    # a kernel: no CPU path
    assert ref_boxes.box_to_corners is original
