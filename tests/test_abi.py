"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/tauv_b200.h declares, and
rejects bad arguments with the documented codes — all without touching a GPU."""
import ctypes
import subprocess

import pytest

import tauv_vision_b200 as tv
from tauv_vision_b200 import _build, _lib


@pytest.fixture(scope="module")
def lib():
    return tv.load_library()


def test_builds_in_tree_for_sm100a():
    path = tv.build()
    assert path == _build.LIB_PATH and path.exists()
    assert path.is_relative_to(_build.PKG_DIR), "the .so must live in-tree so that it travels to the GPU box"
    out = subprocess.run(["cuobjdump", "-lelf", str(path)], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_exports_every_header_symbol(lib):
    names = _lib.header_symbols()
    assert len(names) >= 27
    assert set(names) == set(_lib._SIGNATURES), "ctypes prototypes and header out of sync"
    for n in names:
        assert getattr(lib, n) is not None


def test_exports_nothing_the_header_does_not_declare():
    """The default build has no debug hooks and no undeclared entry points (they exist only with -DTAUV_DEBUG)."""
    out = subprocess.run(["nm", "-D", "--defined-only", str(_build.LIB_PATH)], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line and line.split()[-1].startswith("tauv_")}
    assert exported, "nm found no tauv_ symbols"
    assert exported == set(_lib.header_symbols()), sorted(exported ^ set(_lib.header_symbols()))
    strings = subprocess.run(["strings", str(_build.LIB_PATH)], capture_output=True, text=True).stdout
    for knob in ("TAUV_NO_FUSE", "TAUV_MASK_SIMT", "TAUV_SCORES_OLD", "TAUV_MASK_NO_TMA"):
        assert knob not in strings, f"the default library still reads {knob} from the environment"


def test_version_and_error_text(lib):
    assert lib.tauv_version() == 100
    assert isinstance(lib.tauv_last_error(), bytes)


def test_argument_errors_need_no_gpu(lib):
    null_f = ctypes.cast(None, _lib._F)
    assert lib.tauv_heatmap_nms(null_f, null_f, 1, 1, 4, 4, 3, 0, None) == _lib.E_NULL
    buf = (ctypes.c_float * 16)()
    p = ctypes.cast(buf, _lib._F)
    assert lib.tauv_heatmap_nms(p, p, 1, 1, 4, 4, 2, 0, None) == _lib.E_KERNEL
    assert b"kernel_size" in lib.tauv_last_error()
    assert lib.tauv_heatmap_nms(p, p, 0, 1, 4, 4, 3, 0, None) == _lib.E_SHAPE
    i64 = ctypes.cast((ctypes.c_int64 * 64)(), _lib._I64)
    assert lib.tauv_heatmap_topk(p, 1, 1, 4, 4, 17, 0, i64, i64, p, None, 0, None) == _lib.E_K_RANGE
    assert b"out of range" in lib.tauv_last_error()
    assert lib.tauv_heatmap_topk(p, 1, 1, 4, 4, 4, 7, i64, i64, p, None, 0, None) == _lib.E_SHAPE  # bad mode
    assert lib.tauv_iou_matrix(p, p, 2, 3, 1, 1, p, None) == _lib.E_SHAPE  # batch dims do not broadcast
    with pytest.raises(AssertionError):
        _lib.check(_lib.E_KERNEL)
    with pytest.raises(RuntimeError):
        _lib.check(_lib.E_K_RANGE)


def test_workspace_queries(lib):
    n = lib.tauv_heatmap_topk_workspace_bytes(64, 80, 128, 128, 100)
    assert n >= 64 * 80 * 100 * 8 and n % 256 == 0
    assert lib.tauv_heatmap_topk_workspace_bytes(0, 80, 128, 128, 100) == 0
    assert lib.tauv_yolact_nms_workspace_bytes(64, 19248, 81, 200) >= 64 * 19248 * 4


def test_no_cpu_fallback():
    import torch
    from tauv_vision_b200.centernet.model import decode as D
    from tauv_vision_b200.yolact.model import boxes as Bx
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        D.heatmap_nms(torch.zeros((1, 1, 4, 4)), 3)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        Bx.iou_matrix(torch.zeros((1, 1, 4)), torch.zeros((1, 1, 4)))
    # the three pure layout helpers have no kernel behind them and work wherever the tensor lives: the reference's data
    # loader, collate and plots call them on CPU tensors (segmentation_dataset.py:119, yolact/scripts/train.py:143-145)
    b = torch.tensor([[[0.5, 0.4, 0.2, 0.1]]])
    assert torch.equal(Bx.box_xy_swap(b), torch.tensor([[[0.4, 0.5, 0.1, 0.2]]]))
    assert torch.allclose(Bx.corners_to_box(Bx.box_to_corners(b)), b)


def test_product_never_imports_the_oracle():
    for p in _build.PKG_DIR.rglob("*.py"):
        text = p.read_text()
        assert "import oracle" not in text and "from oracle" not in text, f"{p} imports the oracle"
