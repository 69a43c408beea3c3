"""In-tree build of libvattn_b200.so (nvcc, sm_100a only).

`python -m vattention_b200.build` or `build_library()`; __graft_entry__.build()
calls this.  Objects go to build/obj, the library next to this file so it
travels with the repo snapshot to the GPU box.  No torch headers are involved:
the library is a plain C-ABI shared object (include/vattn_b200.h).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
OBJ_DIR = REPO / "build" / "obj"
LIB_PATH = PKG_DIR / "libvattn_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
          f"-I{REPO / 'include'}", f"-I{CSRC}"]


def _sources() -> list[Path]:
    return sorted(list(CSRC.glob("*.cpp")) + list(CSRC.glob("*.cu")))


def _headers_mtime() -> float:
    hs = list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list((REPO / "include").glob("*.h"))
    return max((h.stat().st_mtime for h in hs), default=0.0)


def _compile(src: Path, hdr_mtime: float, force: bool, verbose: bool) -> Path:
    obj = OBJ_DIR / (src.name + ".o")
    if (not force and obj.exists() and obj.stat().st_mtime > src.stat().st_mtime
            and obj.stat().st_mtime > hdr_mtime):
        return obj
    cmd = [NVCC, *COMMON, *ARCH_FLAGS, "-c", str(src), "-o", str(obj)]
    if src.suffix == ".cu":
        cmd += ["-Xptxas", "-v"] if verbose else []
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed on {src.name}")
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr, flush=True)
    return obj


def build_library(force: bool = False, verbose: bool = False, watchdog: bool = False) -> Path:
    """watchdog=True builds libvattn_b200_dbg.so (-DVATTN_WATCHDOG: mbarrier waits that never complete
    print the barrier and trap instead of hanging); load it with VATTN_B200_LIB=<path>."""
    global OBJ_DIR, LIB_PATH
    if watchdog:
        OBJ_DIR = REPO / "build" / "obj_dbg"
        LIB_PATH = PKG_DIR / "libvattn_b200_dbg.so"
        if "-DVATTN_WATCHDOG" not in COMMON:
            COMMON.append("-DVATTN_WATCHDOG")
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    srcs = _sources()
    hdr_mtime = _headers_mtime()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, hdr_mtime, force, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < newest:
        cmd = [NVCC, "-shared", *ARCH_FLAGS, "-o", str(LIB_PATH), *map(str, objs),
               "-ldl", "-lpthread"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB_PATH


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, watchdog="--watchdog" in sys.argv)
    print(p)
