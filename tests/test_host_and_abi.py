"""CPU tests of the host mirror (TOML schema/defaults, geom_name dispatch, writers) and of the C-ABI surface.
No compute entry point is called: there is no GPU here and the engine has no CPU fallback."""
import ctypes as C
import re
import struct

import numpy as np
import pytest

from conftest import RES, ROOT, GOLDEN
from rsmcrt_b200 import api as A, _lib


# ------------------------------------------------------------------ C ABI
def declared_symbols():
    names = []
    for h in ("smcrt.h", "smcrt_host.h"):
        text = (ROOT / "include" / h).read_text()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b(smcrt_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(smcrt):
    lib = smcrt.load()
    declared = declared_symbols()
    assert len(declared) >= 45
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/*.h but not exported by libsmcrt_gpu.so"
    # and the ctypes table covers exactly the declared set
    assert sorted(_lib.PROTOTYPES) == declared


def test_fails_loudly_without_a_gpu(smcrt):
    import subprocess, sys
    # in a child process so a CUDA-less driver state cannot leak into this one
    code = ("import rsmcrt_b200 as R\n"
            "try:\n    R.Engine(1)\n    print('CREATED')\nexcept R.SmcrtError as e:\n    print('ERR', e)\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT).stdout
    import torch
    if torch.cuda.is_available():
        assert "CREATED" in out
    else:
        assert "ERR" in out and "no CPU fallback" in out


def test_philox_host_matches_known_answer(smcrt):
    # counter = (event, id_lo, id_hi, 0), key = (seed_lo, seed_hi): all-zero -> Random123 KAT
    assert [hex(x) for x in smcrt.philox(0, 0, 0)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]


# ------------------------------------------------------------------ parse_params: keys and defaults
MINIMAL = """
[source]
name = "pencil"
position = [0.0, 0.0, 0.0]
direction = "z"
[grid]
[geometry]
"""


def test_defaults(smcrt):
    cfg = smcrt.Config.loads(MINIMAL)
    assert cfg.grid == ((200, 200, 200), (1.0, 1.0, 1.0))        # parse.f90:92-110
    assert cfg.nphotons == 1000000 and cfg.iseed == 123456789     # parse_source.f90:60, parse.f90:170
    assert cfg.geom_name == "sphere" and cfg.source_name == "pencil"
    k, sub, p = cfg.source
    assert k == A.SRC_PENCIL and list(p[3:6]) == [0, 0, 1]
    assert p[15] == 0.5 and p[16] == 1.0 and p[17] == 0.5 and p[18] == 0.5 and p[19] == 0.6 and p[20] == 0.04
    s = cfg.scene  # geom "sphere": sphere r=1 at the origin (mus=1, mua=0, g=0, n=1) + bounding box 2^3
    assert list(s.kind) == [A.SPHERE, A.BOX] and s.params[0, 0] == 1.0 and list(s.params[1, :3]) == [1, 1, 1]
    assert list(s.mus) == [1.0, 0.0] and list(s.n) == [1.0, 1.0]
    assert not cfg.render_source


@pytest.mark.parametrize("d,vec", [("x", (1, 0, 0)), ("-x", (-1, 0, 0)), ("y", (0, 1, 0)), ("-y", (0, -1, 0)), ("-z", (0, 0, -1))])
def test_cardinal_directions(smcrt, d, vec):
    cfg = smcrt.Config.loads(MINIMAL.replace('"z"', f'"{d}"'))
    assert tuple(cfg.source[2][3:6]) == vec


@pytest.mark.parametrize("bad,msg", [
    (MINIMAL.replace('direction = "z"', 'direction = "up"'), "cardinal"),
    (MINIMAL.replace('direction = "z"\n', ""), "direction"),
    (MINIMAL.replace("[grid]", ""), "grid"),
    (MINIMAL.replace("[geometry]", '[geometry]\ngeom_name="nonesuch"'), "no such routine"),
    (MINIMAL.replace("[geometry]", '[geometry]\ngeom_name="box"\nnumOptProp=2'), "numOptProp"),
    (MINIMAL.replace('name = "pencil"', 'name = "laser"'), "No such source"),
    (MINIMAL.replace("[geometry]", '[geometry]\nmua=[1.0, 2.0]'), "mua"),
    (MINIMAL + '[[detectors]]\ntype="circle"\nposition=[0.0,0.0,0.0]\n', "ID"),
    (MINIMAL + '[[detectors]]\ntype="annulus"\nID="a"\nposition=[0.0,0.0,0.0]\nradius1=0.3\nradius2=0.2\n', "Radii"),
    (MINIMAL + '[[detectors]]\ntype="sphere"\nID="a"\nposition=[0.0,0.0,0.0]\n', "Invalid detector type"),
    (MINIMAL.replace('name = "pencil"', 'name = "focus"'), "rotation"),
    ("[source\nname=1", "toml"),
])
def test_parse_errors(smcrt, bad, msg):
    with pytest.raises(smcrt.SmcrtError) as e:
        smcrt.Config.loads(bad)
    assert msg.lower() in str(e.value).lower()


def test_shipped_configs_parse_and_build(smcrt):
    expect = {"sphere.toml": ("sphere_scene", 41), "validation1.toml": ("box", 2), "validation2.toml": ("box", 2),
              "validation3.toml": ("box", 2), "scat_test.toml": ("scat_test", 2), "scat_test2.toml": ("scat_test2", 1),
              "test_dects.toml": ("scat_test", 2), "validateFibreDect.toml": ("box", 2), "omg.toml": ("omg", 2),
              "aptran.toml": ("aptran", 3), "egg_test.toml": ("egg", 4)}
    for name, (geom, ntop) in expect.items():
        cfg = smcrt.Config.load(RES / name)
        assert cfg.geom_name == geom and cfg.scene.n_top == ntop, name
    # jacques / skin / lens do not exist in the reference dispatcher: builder-defined here
    assert smcrt.Config.load(RES / "jacques.toml").scene.n_top == 2
    lens = smcrt.Config.load(RES / "lens.toml").scene
    assert lens.kind[lens.top_node[0]] == A.MODEL_INTERSECTION and lens.n[0] == 1.5
    skin = smcrt.Config.load(RES / "skin_b200.toml").scene
    assert skin.n_top == 6 and skin.params[:5, 2].sum() == pytest.approx(0.05)  # half thicknesses of the five layers
    with pytest.raises(smcrt.SmcrtError) as e:  # the shipped skin.toml omits point1..3 (parse_source.f90:206-214)
        smcrt.Config.load(RES / "skin.toml")
    assert "point1" in str(e.value)
    with pytest.raises(smcrt.SmcrtError) as e:
        smcrt.Config.load(RES / "vessels.toml")  # needs res/{edges,nodes,radii}.dat
    assert "edges.dat" in str(e.value)


def test_validation1_scene_and_detectors(smcrt):
    cfg = smcrt.Config.load(RES / "validation1.toml")
    s = cfg.scene
    assert list(s.params[0, :3]) == [50.0, 50.0, 0.01] and list(s.params[1, :3]) == [50.0, 50.0, 0.015]  # half lengths
    assert (s.mus[0], s.mua[0], s.hgg[0], s.n[0]) == (90.0, 10.0, 0.75, 1.0)
    assert (s.mus[1], s.mua[1], s.n[1]) == (0.0, 0.0, 1.0)
    kind, p, nb, ids = cfg.detectors
    assert list(kind) == [1, 1] and list(nb) == [100, 100] and ids == ["this is a test", "1"]
    assert list(p[0, :7]) == [0, 0, -0.01, 0, 0, -1, 20] and list(p[1, :7]) == [0, 0, 0.01, 0, 0, 1, 20]
    assert cfg.render_source


def test_omg_scene_is_a_smooth_union_model(smcrt, oracle):
    cfg = smcrt.Config.load(RES / "omg.toml")
    s = cfg.scene
    assert s.kind[s.top_node[0]] == A.MODEL_SMOOTHUNION and s.n_child[s.top_node[0]] == 10
    kids = s.kind[s.first_child[s.top_node[0]]: s.first_child[s.top_node[0]] + 10]
    assert list(kids) == [A.TORUS] + [A.CYLINDER] * 9
    assert s.params[s.top_node[0], 0] == pytest.approx(0.09)
    assert (s.mus[0], s.mua[0], s.n[0]) == (10.0, 0.16, 2.65)
    # the letters are inside the medium, the corners of the box are not
    osc = oracle.OracleScene.from_config(cfg)
    assert osc.sdf(1, [[0.0, 0.0, -0.125]])[0] < 0 and osc.sdf(1, [[0.2, 0.0, -0.7]])[0] < 0 and osc.sdf(1, [[0.9, 0.9, 0.9]])[0] > 0


def test_sphere_scene_fixture(smcrt):
    """setup_sphere_scene draws from the UNSEEDED compiler RNG in the reference; here a fixed stream, pinned."""
    import json
    cfg = smcrt.Config.load(RES / "sphere.toml")
    s = cfg.scene
    table = np.column_stack([s.params[:40, 0], -s.xform[:40, 3], -s.xform[:40, 7], -s.xform[:40, 11]])
    gold = np.array(json.loads((GOLDEN / "sphere_scene_40.json").read_text())["radius_x_y_z"])
    assert table == pytest.approx(gold, abs=1e-15)
    assert ((table[:, 0] >= 0.001) & (table[:, 0] < 0.25)).all()
    assert (np.abs(table[:, 1:]) <= 1 - table[:, :1]).all()  # every sphere inside the 2^3 box
    assert list(s.n[:40]) == [1.37] * 40 and s.n[40] == 1.0 and s.mus[40] == 1e-17


def test_vessels_reads_reference_dat_formats(smcrt, tmp_path):
    (tmp_path / "nodes.dat").write_text("0 0 0\n100 0 0\n100 50 20\n")
    (tmp_path / "edges.dat").write_text("1 2\n2 3\n")
    (tmp_path / "radii.dat").write_text("5\n4\n3\n")
    cfg = smcrt.Config.load(RES / "vessels.toml", res_dir=tmp_path)
    s = cfg.scene
    assert list(s.kind) == [A.CAPSULE, A.CAPSULE, A.BOX]
    # nodes rescaled by res=0.001 and recentred on each axis (setupGeometry.f90:629-639)
    assert list(s.params[0, :6]) == pytest.approx([-0.05, -0.025, -0.01, 0.05, -0.025, -0.01])
    assert s.params[0, 6] == pytest.approx(0.005) and s.params[1, 6] == pytest.approx(0.004)
    assert (s.mus[0], s.mua[0]) == (94.0, 231.0) and (s.mus[2], s.mua[2]) == (357.0, 0.458)


# ------------------------------------------------------------------ writers (src/writer.f90)
def test_nrrd_writer_bytes(smcrt, tmp_path):
    a = np.arange(2 * 3 * 4, dtype=np.float32).reshape((2, 3, 4), order="F")
    path = tmp_path / "a.nrrd"
    smcrt.write_nrrd(path, a, meta='units = "cm"\n')
    raw = path.read_bytes()
    hdr = b"NRRD0004\ntype: float\ndimension: 3\nsizes: 4 3 2\nspace dimension: 3\nencoding: raw\nendian: little\nunits = \"cm\"\n\n\n"
    assert raw.startswith(hdr)                       # sizes reversed like write_hdr (writer.f90:316-318)
    data = np.frombuffer(raw[len(hdr):], np.float32)
    assert (data == np.arange(24)).all()             # x fastest, little endian
    # tools/read_nrrd_class.py splits the header at the first blank line and reads raw float32 after it
    head, _, body = raw.partition(b"\n\n")
    assert b"sizes: 4 3 2" in head


def test_normalise_fluence(smcrt):
    a = np.ones((4, 5, 6), np.float32)
    out = smcrt.normalise_fluence(a, (4, 5, 6), (1.0, 2.0, 3.0), 1000)
    assert out == pytest.approx(4 * 5 * 6 / 1000.0)  # writer.f90:25-52


def test_detector_file_format(smcrt, tmp_path):
    cfg = smcrt.Config.load(RES / "test_dects.toml")
    bins = np.arange(11 + 11 + 121, dtype=float)
    cfg.write_detectors(bins, tmp_path)
    d1 = np.fromfile(tmp_path / "detector_1.dat", np.float64)     # circle: type, len(ID), ID chars, nphotons, radius, pos, dir, pairs
    assert list(d1[:4]) == [1.0, 1.0, float(ord("1")), 100000.0] and d1[4] == 0.5
    assert list(d1[5:11]) == [-1, 0, 0, -1, 0, 0]
    pairs = d1[11:].reshape(-1, 2)
    assert len(pairs) == 11 and pairs[:, 1] == pytest.approx(np.arange(11))
    assert pairs[:, 0] == pytest.approx((np.arange(1, 12) - 0.5) * 0.05)
    d2 = np.fromfile(tmp_path / "detector_2.dat", np.float64)     # annulus: type 3, ..., r1, r2, pos, dir, pairs offset by r1
    assert d2[0] == 3.0 and list(d2[4:6]) == [0.5, 1.0]
    pairs = d2[12:].reshape(-1, 2)
    assert pairs[:, 0] == pytest.approx((np.arange(1, 12) - 0.5) * 0.05 + 0.5) and pairs[:, 1] == pytest.approx(np.arange(11, 22))
    assert (tmp_path / "detector_3.dat").stat().st_size == 0      # camera: "not yet implmented" (writer.f90:127-128)


def test_metadata_is_a_toml_dump_of_the_dict(smcrt):
    meta = smcrt.Config.load(RES / "validation1.toml").metadata
    import tomllib
    d = tomllib.loads(meta)
    assert d["mua%   1"] == 10.0 and d["mus%   1"] == 90.0 and d["BoxDimensions%   3"] == 0.02
    assert d["focus_type"] == "gaussian" and d["units"] == "cm"


def test_inverse_evaluate(smcrt):
    """src/kernelsMod.f90:1753-1787: -mean |total/N - target| over detectors with a target (-1 = none)."""
    lib = smcrt.load()
    totals = np.array([500.0, 250.0, 1000.0])
    targets = np.array([0.4, -1.0, 0.9])
    err = C.c_double(0)
    assert lib.smcrt_inverse_evaluate(3, totals.ctypes.data_as(C.POINTER(C.c_double)), targets.ctypes.data_as(C.POINTER(C.c_double)), 1000,
                                      C.byref(err)) == 0
    assert err.value == pytest.approx(-(abs(0.5 - 0.4) + abs(1.0 - 0.9)) / 2)
    none = np.array([-1.0, -1.0, -1.0])
    assert lib.smcrt_inverse_evaluate(3, totals.ctypes.data_as(C.POINTER(C.c_double)), none.ctypes.data_as(C.POINTER(C.c_double)), 1000,
                                      C.byref(err)) != 0


def test_escape_cell_centre(smcrt):
    """src/kernelsMod.f90:576-591: voxel centre of the symmetry grid, two row-vector rotations, shift."""
    lib = smcrt.load()
    P = C.POINTER(C.c_double)
    out = np.zeros(3)
    pos = np.array([0.1, 0.2, 0.3])
    assert lib.smcrt_escape_cell_centre(1, 2, 4, 4, 4, 4, 1.0, 2.0, 4.0, None, None, pos.ctypes.data_as(P), out.ctypes.data_as(P)) == 0
    assert np.allclose(out, [-0.75 + 0.1, -0.5 + 0.2, 3.0 + 0.3])
    # a rotation about z by 90 degrees stored like the reference's rotate_z (row-vector convention, translation in row 4)
    rz = np.zeros((4, 4)); rz[0, 1] = 1.0; rz[1, 0] = -1.0; rz[2, 2] = 1.0; rz[3, 3] = 1.0   # v.M: (x, y) -> (-y, x)
    M = np.asfortranarray(rz)
    assert lib.smcrt_escape_cell_centre(1, 2, 4, 4, 4, 4, 1.0, 2.0, 4.0, M.ctypes.data_as(P), None, None, out.ctypes.data_as(P)) == 0
    v = np.array([-0.75, -0.5, 3.0, 1.0]) @ rz
    assert np.allclose(out, v[:3])


def test_checkpoint_file_format_round_trip(smcrt, tmp_path):
    """writer.f90:426-457: `tomlfile=<name>` / `photons_run=<n>` as two formatted lines, then the raw float32 jmean appended as a
    stream; kernelsMod.f90:52-66 reads it back by scanning for '=' and taking the stream position after line 2."""
    rng = np.random.default_rng(5)
    jm = rng.random((4, 3, 5), dtype=np.float32)
    p = tmp_path / "check.ckpt"
    smcrt.checkpoint_write(p, "res/validation1.toml", 1234567, jm)
    raw = p.read_bytes()
    head = b"tomlfile=res/validation1.toml\nphotons_run=1234567\n"
    assert raw.startswith(head) and len(raw) == len(head) + 4 * jm.size
    assert np.array_equal(np.frombuffer(raw[len(head):], np.float32), jm.reshape(-1, order="F"))   # x fastest
    name, run, back = smcrt.checkpoint_read(p, jm.size)
    assert name == "res/validation1.toml" and run == 1234567 and np.array_equal(back, jm.reshape(-1, order="F"))
    with pytest.raises(smcrt.SmcrtError):
        smcrt.checkpoint_read(p, jm.size + 1)                      # grid larger than the file
    with pytest.raises(smcrt.SmcrtError):
        smcrt.checkpoint_read(tmp_path / "missing.ckpt")


# ------------------------------------------------------------------ the product's TOML -> scene builder against a second implementation
ALL_DECKS = sorted(p.name for p in RES.glob("*.toml") if p.name not in ("default.toml", "skin.toml"))


@pytest.mark.parametrize("deck", ALL_DECKS)
def test_host_builder_matches_the_oracles_own_builder(smcrt, deck, tmp_path):
    """rsmcrt_b200/csrc/host/host.cpp (C++, the product) and oracle/scenes.py (Python, test infrastructure) restate the reference's
    parse + setupGeometry layer independently of each other (parse_*.f90, src/setupGeometry.f90).  Every shipped deck must come
    out of both as the same flattened bytes: node table, transforms, parameters, optics, grid, source slots, detector table in
    dects(:) order.  A wrong builder is then not common-mode between the engine and the oracle that checks it."""
    from oracle import scenes
    res_dir = None
    if deck == "vessels.toml":
        import sys
        sys.path.insert(0, str(RES.parent / "tools"))
        import make_vessels
        make_vessels.make(tmp_path, 60, 3)
        res_dir = tmp_path
    try:
        cfg = smcrt.Config.load(RES / deck, res_dir=res_dir)
    except smcrt.SmcrtError:
        with pytest.raises(ValueError):  # a deck the reference's parser stops on (exp.toml has no source position): both refuse it
            scenes.load(RES / deck, res_dir)
        return
    d = scenes.load(RES / deck, res_dir)
    a, b = cfg.scene, d.scene
    assert a.n_top == b.n_top and len(a.kind) == len(b.kind)
    for name in ("kind", "first_child", "n_child", "top_node"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    for name in ("xform", "params", "mus", "mua", "hgg", "n"):
        assert np.allclose(getattr(a, name), getattr(b, name), rtol=1e-14, atol=1e-15), name
    assert cfg.grid == d.grid
    ka, sa, pa = cfg.source
    kb, sb, pb = d.source
    assert (ka, sa) == (kb, sb) and np.allclose(pa, pb, rtol=1e-15, atol=0)
    da, db = cfg.detectors, d.detectors
    assert np.array_equal(da[0], db[0]) and np.array_equal(da[2], db[2]) and list(da[3]) == list(db[3])
    assert np.allclose(da[1], db[1], rtol=1e-15, atol=0)
    assert (cfg.nphotons, cfg.iseed, cfg.geom_name, cfg.source_name) == (d.nphotons, d.iseed, d.geom_name, d.source_name)


def test_normalise_fluence_takes_64_bit_packet_counts():
    """normalise_fluence (src/writer.f90:25-52): x nx*ny*nz / nphotons.  The reference's nphotons is a default integer; the engine's
    counts are 64-bit (it traces 4e9 packets per second), and a 3e9-packet job must not normalise by a wrapped negative number
    (ADVICE r1)."""
    a = np.full((4, 5, 6), 3.0e9, np.float32)
    out = A.normalise_fluence(a, (4, 5, 6), (1.0, 2.0, 3.0), 3_000_000_000)
    assert np.allclose(out, 4 * 5 * 6, rtol=1e-6)
    with pytest.raises(A.SmcrtError):
        A.normalise_fluence(a, (4, 5, 6), (1.0, 2.0, 3.0), 0)
