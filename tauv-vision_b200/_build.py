"""In-tree build of libtauv_b200.so (sm_100a only) with plain nvcc.

The shared object lands next to this file (``tauv-vision_b200/lib/libtauv_b200.so``) so that it
travels with the source tree; nothing is installed into site-packages and nothing is JIT-cached.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_DIR = PKG_DIR / "lib"
OBJ_DIR = LIB_DIR / "obj"
LIB_PATH = LIB_DIR / "libtauv_b200.so"
INCLUDE = PKG_DIR.parent / "include"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr", "--expt-extended-lambda",
    "-Xcompiler", "-fPIC",
    # Nothing that changes fp32 results: no --use_fast_math, and FMA contraction stays off
    # because every comparison-feeding expression must round exactly like the reference's
    # separate ATen ops.
    "--fmad=false",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libtauv_b200 has no CPU or prebuilt fallback")


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*")) + list(INCLUDE.glob("*.h"))):
        if p.is_file():
            h.update(p.name.encode())
            h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(os.environ.get("TAUV_EXTRA_NVCC", "").encode())  # e.g. -DTAUV_DEBUG (experiment hooks, tools/*_trace.py)
    return h.hexdigest()


def is_fresh() -> bool:
    stamp = LIB_DIR / "build.stamp"
    return LIB_PATH.exists() and stamp.exists() and stamp.read_text().strip() == _fingerprint()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every csrc/*.cu for sm_100a and link libtauv_b200.so.  Cross-compiles without a GPU."""
    if not force and is_fresh():
        return LIB_PATH
    nvcc = _nvcc()
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    srcs = _sources()
    if not srcs:
        raise RuntimeError(f"no CUDA sources under {CSRC}")

    def compile_one(src: Path) -> Path:
        obj = OBJ_DIR / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("TAUV_EXTRA_NVCC", "").split(), "-I", str(INCLUDE), "-c", str(src),
               "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH),
            *map(str, objs), "-cudart", "static"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    (LIB_DIR / "build.stamp").write_text(_fingerprint())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
