"""Build helper: compiles the CUDA sources under fastdem_b200/csrc for sm_100a into
fastdem_b200/libfastdem_b200.so (in-tree, so the .so travels to the GPU box with the
snapshot).  nvcc cross-compiles without a GPU.

-fmad=false / -ffp-contract=off are REQUIRED, not tuning: cell indices must be
bit-identical to the CPU oracle, which means no FMA contraction on either side."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
BUILD = REPO / "build"
LIB = PKG_DIR / "libfastdem_b200.so"

SOURCES = ["kernels.cu", "kernels_tile.cu", "kernels_raycast.cu", "sort.cu", "capi.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
]


if os.environ.get("FDEM_PROBES") == "1":   # tuning probes in K3t (tools/phase_probe.py); never in the product build
    NVCC_FLAGS.append("-DFDEM_PROBES")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libfastdem_b200.so)")


def _deps() -> list[Path]:
    return sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.inc")) + sorted(CSRC.glob("*.cuh")) + [
        REPO / "include" / "fastdem_b200.h", Path(__file__)]


def _stale(target: Path, sources: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(s.stat().st_mtime > t for s in sources)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    nvcc = _nvcc()
    deps = _deps()
    objs, procs = [], []
    for src in SOURCES:
        s = CSRC / src
        o = BUILD / (s.stem + ".o")
        objs.append(o)
        if force or _stale(o, [s] + deps):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(s), "-o", str(o)]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out.decode(errors='replace')}")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-Xcompiler", "-fPIC",
               "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout.decode(errors='replace')}")
    return LIB


def build_oracle(force: bool = False) -> Path:
    """Test infrastructure: the CPU oracle (oracle/).  Building the checker is not using it."""
    odir = REPO / "oracle"
    lib = odir / "libfdem_oracle.so"
    if force or _stale(lib, [odir / "fdem_oracle.hpp", odir / "fdem_oracle_capi.cpp", odir / "fdem_oracle_io.cpp", odir / "Makefile"]):
        r = subprocess.run(["make", "-C", str(odir)] + (["-B"] if force else []),
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode != 0:
            raise RuntimeError(f"oracle build failed:\n{r.stdout.decode(errors='replace')}")
    return lib


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
    print(build_oracle(force="--force" in sys.argv))
