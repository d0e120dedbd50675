"""Shared test helpers: golden loading and the tie-aware comparators (hard parts)."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"

# Tolerances of the north star: indices / labels / keep sets exact; fp32 values 1e-5 relative.
RTOL = 1e-5
ATOL_DENORMAL = 1e-30


def golden(name: str) -> dict:
    with np.load(GOLDEN / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


def t(a, dtype=None) -> torch.Tensor:
    x = torch.from_numpy(np.ascontiguousarray(a))
    return x.to(dtype) if dtype is not None else x


def assert_close(actual, expected, rtol=RTOL, atol=ATOL_DENORMAL, what=""):
    a = np.asarray(actual.detach().cpu() if isinstance(actual, torch.Tensor) else actual, dtype=np.float64)
    e = np.asarray(expected.detach().cpu() if isinstance(expected, torch.Tensor) else expected, dtype=np.float64)
    assert a.shape == e.shape, f"{what}: shape {a.shape} vs {e.shape}"
    bad = ~(np.abs(a - e) <= atol + rtol * np.abs(e))
    bad &= ~(np.isnan(a) & np.isnan(e))
    if bad.any():
        i = np.argwhere(bad)[0]
        raise AssertionError(f"{what}: {bad.sum()} / {bad.size} mismatches, first at {tuple(i)}: "
                             f"{a[tuple(i)]!r} vs {e[tuple(i)]!r}")


def assert_equal(actual, expected, what=""):
    a = np.asarray(actual.detach().cpu() if isinstance(actual, torch.Tensor) else actual)
    e = np.asarray(expected.detach().cpu() if isinstance(expected, torch.Tensor) else expected)
    assert a.shape == e.shape, f"{what}: shape {a.shape} vs {e.shape}"
    if not np.array_equal(a, e):
        i = np.argwhere(a != e)[0]
        raise AssertionError(f"{what}: {(a != e).sum()} / {a.size} mismatches, first at {tuple(i)}: "
                             f"{a[tuple(i)]!r} vs {e[tuple(i)]!r}")


def flat_index(index, label, H, W):
    index = np.asarray(index)
    return np.asarray(label) * (H * W) + index[..., 0] * W + index[..., 1]


def assert_topk_tie_aware(flat_a, score_a, flat_e, score_e, rtol=RTOL, what="", allow_swaps=0):
    """Ranked lists must agree exactly, except that inside a run of (nearly) equal expected scores the
    entries may come in any order, and a run cut by the k boundary may hold any of its members.

    ``allow_swaps`` > 0 additionally tolerates that many adjacent near-tie transpositions caused by the
    last-ulp difference between our expf and ATen's (counted and returned)."""
    flat_a, score_a = np.asarray(flat_a), np.asarray(score_a, dtype=np.float64)
    flat_e, score_e = np.asarray(flat_e), np.asarray(score_e, dtype=np.float64)
    assert flat_a.shape == flat_e.shape, f"{what}: shape"
    k = flat_e.shape[-1]
    flat_a, score_a = flat_a.reshape(-1, k), score_a.reshape(-1, k)
    flat_e, score_e = flat_e.reshape(-1, k), score_e.reshape(-1, k)
    swaps = 0
    for b in range(flat_e.shape[0]):
        # scores themselves agree rank by rank
        assert_close(score_a[b], score_e[b], rtol=rtol, atol=1e-12, what=f"{what} frame {b} scores")
        i = 0
        while i < k:
            j = i + 1
            while j < k and abs(score_e[b, j] - score_e[b, i]) <= rtol * abs(score_e[b, i]) + 1e-12:
                j += 1
            ea, aa = set(flat_e[b, i:j].tolist()), set(flat_a[b, i:j].tolist())
            if ea != aa:
                if j == k:  # run truncated by k: any members of the (unknown) full run are acceptable
                    pass
                else:
                    swaps += len(ea ^ aa)
                    if swaps > allow_swaps:
                        raise AssertionError(f"{what} frame {b}: ranks [{i},{j}) differ: got {sorted(aa)} "
                                             f"expected {sorted(ea)} (scores {score_e[b, i:j]})")
            i = j
    return swaps
