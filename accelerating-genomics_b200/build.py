"""Build recipe for libagx.so (CUDA, sm_100a only), the C drivers and the peaks microbenchmark.

Everything is compiled in-tree with explicit nvcc / gcc commands so the artefacts travel with the
gpurun snapshot; nothing is JIT-compiled at run time.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libagx.so"
BIN = ROOT / "drivers" / "bin"

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-O3,-Wall", "--expt-relaxed-constexpr"]

LIB_SOURCES = ["api.cu", "sw_kernels.cu", "sw_long.cu", "sw_parse.cu", "pairhmm_kernels.cu", "pairhmm_parse.cu"]


def _run(cmd, **kw):
    print("+", " ".join(str(c) for c in cmd), flush=True)
    subprocess.run([str(c) for c in cmd], check=True, **kw)


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build_lib(force: bool = False, verbose_ptxas: bool = False) -> Path:
    srcs = [CSRC / s for s in LIB_SOURCES]
    deps = srcs + [CSRC / "common.cuh", ROOT / "include" / "agx.h"]
    objs = []
    for s in srcs:
        o = s.with_suffix(".o")
        if force or _stale(o, deps):
            cmd = [NVCC, *ARCH, *NVCC_FLAGS, "-c", s, "-o", o]
            if verbose_ptxas:
                cmd += ["-Xptxas", "-v"]
            _run(cmd)
        objs.append(o)
    if force or _stale(LIB, objs):
        _run([NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lpthread"])
    return LIB


def build_drivers(force: bool = False) -> None:
    BIN.mkdir(parents=True, exist_ok=True)
    gcc = shutil.which("gcc") or "gcc"
    for name in ("smithWaterman", "smithWatermanGpu", "smithWatermanAlign", "pairHMM"):
        src = ROOT / "drivers" / f"{name}.c"
        if not src.exists():
            continue
        out = BIN / name
        if force or _stale(out, [src, ROOT / "include" / "agx.h", LIB]):
            _run([gcc, "-O2", "-Wall", "-std=c11", "-I", ROOT / "include", src, "-o", out,
                  f"-L{PKG}", "-lagx", f"-Wl,-rpath,{PKG}", "-Wl,-rpath,$ORIGIN/../../accelerating-genomics_b200",
                  "-lm"])
    peaks = CSRC / "peaks.cu"
    if peaks.exists():
        out = BIN / "agx_peaks"
        if force or _stale(out, [peaks]):
            _run([NVCC, *ARCH, "-O3", "-std=c++17", "-lineinfo", peaks, "-o", out, "-ldl"])


def build_oracle() -> None:
    """The CPU oracle is TEST infrastructure; building it here is not using it."""
    _run(["make", "-C", ROOT / "oracle", "liboracle.so"])
    if Path("/root/reference").exists():
        _run(["make", "-C", ROOT / "oracle", "ref"])


def build_all(force: bool = False) -> None:
    build_lib(force)
    build_drivers(force)
    build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
