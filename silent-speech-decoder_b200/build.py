"""In-tree build of libssd_ctc.so (nvcc, sm_100a only).

``python -m silent_speech_decoder_b200.build`` or ``__graft_entry__.build()``.
nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels
with the repo snapshot to the GPU box.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import List

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
LIB_PATH = PKG_DIR / "libssd_ctc.so"
OBJ_DIR = PKG_DIR / "build"
SOURCES = ["greedy.cu", "beam.cu", "ctc_loss.cu", "ctc_loss_fast.cu", "edit.cu", "aux.cu", "capi.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libssd_ctc.so cannot be built")


def _stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h")) + [Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build_native(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a and link libssd_ctc.so. Returns its path."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = _nvcc()
    OBJ_DIR.mkdir(exist_ok=True)
    logs: List[str] = []

    def compile_one(src: str) -> Path:
        obj = OBJ_DIR / (Path(src).stem + ".o")
        extra = os.environ.get("SSD_NVCC_FLAGS", "").split()  # e.g. -DSSD_LOSS_TIMELINE for profiles/timeline_fast.py
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", str(INCLUDE), "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs.append(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}")
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    logs.append(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}")
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    (OBJ_DIR / "build.log").write_text("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    return LIB_PATH


if __name__ == "__main__":
    p = build_native(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
