"""Build recipe for libsmcrt_gpu.so (sm_100a only, in-tree so the .so travels with the snapshot)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIB_DIR = ROOT / "lib"
LIB_PATH = LIB_DIR / "libsmcrt_gpu.so"

SOURCES = [CSRC / "engine.cu", CSRC / "host" / "host.cpp"]
HEADERS = [CSRC / "kernels.cuh", CSRC / "step_body.inc", CSRC / "step_macros.inc", CSRC / "step_macros_undef.inc", CSRC / "device_scene.cuh", CSRC / "host_math.hpp", CSRC / "host" / "toml_lite.hpp",
           ROOT.parent / "include" / "smcrt.h", ROOT.parent / "include" / "smcrt_host.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # FP32 division / sqrt as MUFU + 1 Newton step (<= 2 ulp) instead of the IEEE sequences: the SDF and ray-distance maths
    # tolerate it (parity tests hold the 1e-6 bar) and the kernel is issue bound (+23 % packets/s).  FP64 code is unaffected.
    "-prec-div=false", "-prec-sqrt=false",
    "-Xcompiler", "-fPIC",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libsmcrt_gpu.so cannot be built (this package has no CPU fallback)")


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, out: Path | None = None) -> Path:
    """Compile the CUDA engine + host layer into rsmcrt_b200/lib/libsmcrt_gpu.so."""
    if not force and out is None and not needs_build():
        return LIB_PATH
    LIB_DIR.mkdir(parents=True, exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"]
    cmd += os.environ.get("SMCRT_NVCC_EXTRA", "").split()  # experiments only; the shipped flags are NVCC_FLAGS
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", str(out or LIB_PATH), *map(str, SOURCES), "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
