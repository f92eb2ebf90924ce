"""Build recipe for libsmcrt_gpu.so (sm_100a only, in-tree so the .so travels with the snapshot).

Six objects -- engine.cu, host.cpp and trace_inst.cu once per <PATHLEN, HASDET> pair (the 52 trace-kernel instantiations) --
compiled in parallel and linked into ONE shared object.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIB_DIR = ROOT / "lib"
OBJ_DIR = ROOT / "build"
LIB_PATH = LIB_DIR / "libsmcrt_gpu.so"

SOURCES = [CSRC / "engine.cu", CSRC / "trace_inst.cu", CSRC / "host" / "host.cpp"]
HEADERS = [CSRC / "kernels.cuh", CSRC / "step_body.inc", CSRC / "step_macros.inc", CSRC / "step_macros_undef.inc", CSRC / "device_scene.cuh",
           CSRC / "host_math.hpp", CSRC / "host" / "toml_lite.hpp", ROOT.parent / "include" / "smcrt.h", ROOT.parent / "include" / "smcrt_host.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # FP32 division / sqrt as MUFU + 1 Newton step (<= 2 ulp) instead of the IEEE sequences: the SDF and ray-distance maths
    # tolerate it (parity tests hold the 1e-6 bar) and the kernel is issue bound (+23 % packets/s).  FP64 code is unaffected.
    "-prec-div=false", "-prec-sqrt=false",
    "-Xcompiler", "-fPIC",
]

# (object name, source, extra defines)
UNITS = [("engine", CSRC / "engine.cu", [])] + \
        [(f"trace_pl{pl}_hd{hd}", CSRC / "trace_inst.cu", [f"-DSMCRT_INST_PL={pl}", f"-DSMCRT_INST_HD={hd}"]) for pl in (0, 1) for hd in (0, 1)] + \
        [("host", CSRC / "host" / "host.cpp", [])]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libsmcrt_gpu.so cannot be built (this package has no CPU fallback)")


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    return any(p.exists() and p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, out: Path | None = None) -> Path:
    """Compile the CUDA engine + host layer into rsmcrt_b200/lib/libsmcrt_gpu.so."""
    if not force and out is None and not needs_build():
        return LIB_PATH
    LIB_DIR.mkdir(parents=True, exist_ok=True)
    obj_dir = OBJ_DIR if out is None else Path(str(out) + ".obj")
    obj_dir.mkdir(parents=True, exist_ok=True)
    base = [_nvcc(), *NVCC_FLAGS, "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"]
    base += os.environ.get("SMCRT_NVCC_EXTRA", "").split()  # experiments only; the shipped flags are NVCC_FLAGS
    if verbose:
        base += ["-Xptxas", "-v"]

    def compile_unit(unit):
        name, src, defs = unit
        if not src.exists():
            return name, None, ""
        obj = obj_dir / (name + ".o")
        cmd = [*base, *defs, "-c", "-o", str(obj), str(src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return name, obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 1)) as pool:
        done = list(pool.map(compile_unit, UNITS))
    objs = [str(o) for _, o, _ in done if o is not None]
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out or LIB_PATH), *objs, "-ldl"]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(link) + "\n" + res.stdout + res.stderr)
    if verbose:
        for name, _, log in done:
            print(f"==== {name}\n{log}")
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
