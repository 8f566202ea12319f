"""Builds libdcvit.so (the sm_100a kernel library) in-tree with nvcc.

nvcc cross-compiles for sm_100a without a GPU, so this runs on the CPU build box;
the resulting .so is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = CSRC / "build"
LIB_PATH = PKG_DIR / "libdcvit.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    *os.environ.get("DCV_NVCC_EXTRA", "").split(),  # e.g. -DDCV_ATTN_TIMELINE for tools/attn_timeline.py
]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _headers() -> list[Path]:
    return sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [PKG_DIR.parent / "include" / "dcvit.h"]


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def _compile(src: Path, obj: Path, extra=()) -> str:
    cmd = [NVCC, *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    return r.stderr


# validation variants: extra -D flags, separate object directory and library name (libdcvit_<variant>.so)
VARIANTS = {"erf": ["-DDCV_GELU_ERF"], "timeline": ["-DDCV_ATTN_TIMELINE"]}


def build(force: bool = False, verbose: bool = False, variant: str = "") -> Path:
    build_dir = BUILD_DIR / variant if variant else BUILD_DIR
    lib_path = PKG_DIR / f"libdcvit_{variant}.so" if variant else LIB_PATH
    extra = VARIANTS[variant] if variant else []
    build_dir.mkdir(parents=True, exist_ok=True)
    hdrs = _headers()
    jobs = []
    objs = []
    for src in _sources():
        obj = build_dir / (src.stem + ".o")
        objs.append(obj)
        if force or _stale(obj, [src, *hdrs]):
            jobs.append((src, obj))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            logs = list(ex.map(lambda j: _compile(*j, extra), jobs))
        (build_dir / "ptxas.log").write_text("\n".join(logs))
        if verbose:
            sys.stderr.write("\n".join(logs))
    if force or jobs or _stale(lib_path, objs):
        cmd = [NVCC, "-shared", "-o", str(lib_path), *map(str, objs), "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib_path


if __name__ == "__main__":
    var = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else ""
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=var)
    print(p)
