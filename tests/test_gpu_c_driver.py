"""The Fortran shim's side of the C ABI, compiled.

fortran/gpu_bridge.f90 cannot be compiled here (no Fortran compiler in the image); tests/c_driver/shim_replay.c replays its exact
call sequence from C with the arrays laid out as the shim passes them (column-major (16,n) / (8,n) / (20,n) blocks, 0-based node
indices, accumulate = 1), built with gcc against include/smcrt.h and linked to libsmcrt_gpu.so.  Its tallies must equal those of
the ctypes path the rest of the suite uses -- bit for bit for the integer-valued ones.  The scene bytes come from the oracle's
own builder (oracle/scenes.py), so this is also the independent builder driving the engine end to end.
"""
import struct
import subprocess

import numpy as np
import pytest

from conftest import RES, ROOT
from rsmcrt_b200 import api as A

LIB_DIR = ROOT / "rsmcrt_b200" / "lib"


def build_driver(tmp_path):
    exe = tmp_path / "shim_replay"
    r = subprocess.run(["gcc", "-O2", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(ROOT / "tests" / "c_driver" / "shim_replay.c"),
                        "-L", str(LIB_DIR), "-lsmcrt_gpu", f"-Wl,-rpath,{LIB_DIR}", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def dump_scene(deck, path):
    s, (n3, m3), (sk, ss, sp), (dk, dp, dn, _) = deck.scene, deck.grid, deck.source, deck.detectors
    with open(path, "wb") as f:
        f.write(struct.pack("<8i3d", len(s.kind), s.n_top, len(dk), sk, ss, *n3, *m3))
        for a in (s.kind, s.first_child, s.n_child, s.top_node):
            f.write(np.ascontiguousarray(a, "<i4").tobytes())
        for a in (s.xform, s.params, s.mus, s.mua, s.hgg, s.n, sp):
            f.write(np.ascontiguousarray(a, "<f8").tobytes())      # (n,16) C order == xform(16,n) Fortran order
        f.write(np.ascontiguousarray(dk, "<i4").tobytes())
        f.write(np.ascontiguousarray(dn, "<i4").tobytes())
        f.write(np.ascontiguousarray(dp, "<f8").tobytes())


def test_c_driver_compiles_against_the_header(tmp_path):
    """(no GPU needed) gcc -Wall -Werror accepts include/smcrt.h from C and the driver links against the shared object."""
    build_driver(tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("deck,n,mode", [("validation1.toml", 300_000, 1), ("test_dects.toml", 30_000, 1), ("skin_b200.toml", 30_000, 3),
                                         ("omg.toml", 20_000, 1)])
def test_shim_call_sequence_equals_the_ctypes_path(tmp_path, smcrt, deck, n, mode):
    from oracle import scenes
    exe = build_driver(tmp_path)
    d = scenes.load(RES / deck)
    dump_scene(d, tmp_path / "scene.bin")
    r = subprocess.run([str(exe), str(tmp_path / "scene.bin"), str(tmp_path / "out.bin"), str(n), str(d.iseed), str(mode), "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    raw = (tmp_path / "out.bin").read_bytes()
    nv, nb = struct.unpack_from("<2q", raw, 0)
    off = 16
    absorb = np.frombuffer(raw, "<f4", nv, off); off += 4 * nv
    jmean = None
    if mode & 2:
        jmean = np.frombuffer(raw, "<f4", nv, off); off += 4 * nv
    bins = np.frombuffer(raw, "<f8", nb, off); off += 8 * nb
    nscatt, launched, lost = struct.unpack_from("<3d", raw, off)
    # the ctypes path: the product's own TOML -> scene layer + Engine
    e = smcrt.Engine(1)
    e.apply(smcrt.Config.load(RES / deck))
    e.run(n, d.iseed, tally_mode=mode)
    ref = e.fetch(jmean=bool(mode & 2), absorb=True)
    e.close()
    assert launched == n and lost == ref["counters"]["lost"] and nscatt == ref["counters"]["nscatt"]
    assert np.array_equal(absorb, ref["absorb"].reshape(-1, order="F"))          # unit deposits: exact
    if len(ref["det_bins"]):
        assert np.array_equal(bins[:len(ref["det_bins"])], ref["det_bins"])      # Q40.24 fixed point: exact
    if jmean is not None:
        assert np.allclose(jmean, ref["jmean"].reshape(-1, order="F"), rtol=2e-4, atol=1e-7)   # float atomics: order only
