"""Pins the CPU oracle against the reference's OWN known-answer tests, transcribed from
/root/reference/test/** (file:line cited per test).  CPU only.  If these pass, the oracle's deterministic
components are the reference's; the GPU parity tests (tests/test_gpu_*.py) then compare the engine with it."""
import numpy as np
import pytest

from rsmcrt_b200 import api as A

WP = dict(rel=1e-12, abs=1e-12)


def one(oracle, kind, params, xform=None, optics=(0, 0, 0, 0)):
    scene = A.Scene.from_primitives([(kind, xform, params)], [optics])
    return oracle.OracleScene(scene)


def model(oracle, op, children, k=0.0):
    """children: list of (kind, xform, params)."""
    n = 1 + len(children)
    kind = np.array([op] + [c[0] for c in children], np.int32)
    xf = np.tile(np.eye(4).reshape(-1), (n, 1))
    par = np.zeros((n, 8))
    par[0, 0] = k
    for i, c in enumerate(children):
        if c[1] is not None:
            xf[i + 1] = np.asarray(c[1]).reshape(-1)
        par[i + 1, :len(c[2])] = c[2]
    first = np.zeros(n, np.int32); first[0] = 1
    nch = np.zeros(n, np.int32); nch[0] = len(children)
    s = A.Scene(kind, first, nch, xf, par, np.array([0], np.int32), np.zeros(1), np.zeros(1), np.zeros(1), np.zeros(1))
    return oracle.OracleScene(s)


def ev(osc, *p):
    return float(osc.sdf(1, [p])[0])


# ------------------------------------------------------------------ test/SDF/test_SDF.f90:678-1036 (primitives)
def test_sphere(oracle):  # :678-721
    s = one(oracle, A.SPHERE, [1.0])
    assert ev(s, 0, 0, 0) == pytest.approx(-1.0, **WP)
    for p in [(0, 1, 0), (0, 0, 1), (1, 0, 0), (0, -1, 0), (0, 0, -1), (-1, 0, 0)]:
        assert ev(s, *p) == pytest.approx(0.0, **WP)
    r = np.sqrt(np.float32(1.0) / 3.0)  # sqrt(1./3._wp): default-real 1. promoted
    assert ev(s, r, r, r) == pytest.approx(0.0, abs=1e-7)


def test_box(oracle):  # :723-766 (box(lengths) stores half lengths, sdfs.f90:455)
    s = one(oracle, A.BOX, [1.0, 1.0, 1.0])
    assert ev(s, 0, 0, 0) == pytest.approx(-1.0, **WP)
    for p in [(0, 1, 0), (0, 0, 1), (1, 0, 0), (0, -1, 0), (0, 0, -1), (-1, 0, 0), (1, 1, 1)]:
        assert ev(s, *p) == pytest.approx(0.0, **WP)


def test_cylinder(oracle):  # :768-813
    s = one(oracle, A.CYLINDER, [0, 0, -1, 0, 0, 1, 1.0])
    assert ev(s, 0, 0, 0) == pytest.approx(-1.0, **WP)
    for p in [(0, 1, 0), (0, 0, 1), (1, 0, 0), (0, -1, 0), (0, 0, -1), (-1, 0, 0)]:
        assert ev(s, *p) == pytest.approx(0.0, **WP)
    r = np.sqrt(0.5)
    assert ev(s, r, r, 0) == pytest.approx(0.0, abs=1e-7)


def test_torus(oracle):  # :815-834
    s = one(oracle, A.TORUS, [0.5, 1.0])
    assert ev(s, 0, 0, 0) == pytest.approx(-0.5, **WP)
    assert ev(s, 1.5, 0, 0) == pytest.approx(0.0, **WP)


def test_segment(oracle):  # :836-873 (radius fixed at 0.1)
    s = one(oracle, A.SEGMENT, [-1, 0, 0, 1, 0, 0])
    for p in [(0, 0, 0), (-1, 0, 0), (1, 0, 0)]:
        assert ev(s, *p) == pytest.approx(-0.1, **WP)
    for p in [(1, 1.1, 0), (0, 1.1, 0), (0, 0, 1.1)]:
        assert ev(s, *p) == pytest.approx(1.0, rel=1e-12)


def test_triprism(oracle):  # :875-899
    s = one(oracle, A.TRIPRISM, [1.0, 5.0])
    assert ev(s, 0, 0, 5) == pytest.approx(0.0, **WP)
    assert ev(s, 0, 1, 0) == pytest.approx(0.0, **WP)


def test_capsule(oracle):  # :901-931
    s = one(oracle, A.CAPSULE, [-1, 0, 0, 1, 0, 0, 1.0])
    assert ev(s, 0, 0, 0) == pytest.approx(-1.0, **WP)
    assert ev(s, 0, 1, 0) == pytest.approx(0.0, **WP)
    assert ev(s, 2, 0, 0) == pytest.approx(0.0, **WP)


def test_plane(oracle):  # :933-967
    s = one(oracle, A.PLANE, [0, 0, 1.0])
    for p in [(0, 0, 0), (0, 1, 0), (2, 0, 0)]:
        assert ev(s, *p) == pytest.approx(0.0, **WP)
    assert ev(s, 0, 0, -1) == pytest.approx(-1.0, **WP)
    assert ev(s, 0, 0, 1) == pytest.approx(1.0, **WP)


def test_cone(oracle):  # :969-995
    s = one(oracle, A.CONE, [0, 0, 0, 0, 0, 1, 5.0, 0.0])
    assert ev(s, 0, 0, 1) == pytest.approx(0.0, **WP)
    assert ev(s, 1, 1, 0) == pytest.approx(0.0, **WP)


def test_egg(oracle):  # :997-1036
    r1, r2, h = 2.5, 0.75, 1.5
    s = one(oracle, A.EGG, [r1, r2, h])
    assert ev(s, 0, 0, 0) == pytest.approx(-r1, **WP)
    assert ev(s, r1, 0, 0) == pytest.approx(0.0, **WP)
    assert ev(s, 0, r1 + 2 * r2, 0) == pytest.approx(0.0, abs=1e-5)
    assert ev(s, r1, r1, 0) == pytest.approx(0.630294, abs=1e-5)


# ------------------------------------------------------------------ CSG / modifiers / normals
def test_intersection(oracle):  # :189-233
    m = model(oracle, A.MODEL_INTERSECTION, [(A.SPHERE, None, [0.25]), (A.BOX, None, [0.5, 0.5, 0.5])], 1.0)
    assert ev(m, 0, 0, 0) == pytest.approx(-0.25, **WP)
    assert ev(m, 0.25, 0, 0) == pytest.approx(0.0, **WP)
    assert ev(m, np.float32(0.4), 0, 0) > 0


def test_subtraction(oracle):  # :235-265
    m = model(oracle, A.MODEL_SUBTRACTION, [(A.SPHERE, None, [0.25]), (A.BOX, None, [0.5, 0.5, 0.5])], 1.0)
    assert ev(m, 0, 0, 0) == pytest.approx(0.25, **WP)
    assert ev(m, 0.25, 0, 0) == pytest.approx(0.0, **WP)


def test_bend(oracle):  # :267-301
    b = model(oracle, A.MOD_BEND, [(A.BOX, None, [0.5, 0.5, 0.5])], 10.0)
    box = one(oracle, A.BOX, [0.5, 0.5, 0.5])
    f = np.float32
    assert ev(b, 0, 0, 0) < 0
    assert ev(b, f(0.6), 0, 0) > 0
    assert ev(b, f(0.4), f(-0.4), f(-0.4)) > 0
    assert ev(box, f(0.4), f(-0.4), f(-0.4)) < 0


def test_calc_normal(oracle):  # :128-170
    s = one(oracle, A.SPHERE, [1.0])
    for p in [(1, 0, 0), (0, 1, 0), (0, 0, 1)]:
        n = s.normal(1, [p])[0]
        assert n == pytest.approx(np.array(p, float), abs=1e-9)


def test_albedo_of_nonabsorbing_medium(oracle):  # :172-187, test/optical_props/test_opticalprops.f90:45-79
    assert oracle.mono(0, 0, 0, 0)["albedo"] == 1.0
    m = oracle.mono(1.0, 2.0, 0.5, 1.3)
    assert m["kappa"] == 3.0 and m["albedo"] == pytest.approx(1 / 3) and m["g2"] == 0.25 and m["n"] == 1.3


def test_model_left_fold_smooth_union(oracle):  # sdf_base.f90:146-161 with sdfModifiers.f90:443-459
    a, b, c = (A.SPHERE, None, [0.3]), (A.SPHERE, A_translate_inv([0.4, 0, 0]), [0.3]), (A.SPHERE, A_translate_inv([0, 0.4, 0]), [0.2])
    m = model(oracle, A.MODEL_SMOOTHUNION, [a, b, c], 0.09)
    p = np.array([0.2, 0.1, 0.05])
    d = [np.linalg.norm(p) - 0.3, np.linalg.norm(p - [0.4, 0, 0]) - 0.3, np.linalg.norm(p - [0, 0.4, 0]) - 0.2]

    def su(d1, d2, k=0.09):
        h = max(k - abs(d1 - d2), 0) / k
        return min(d1, d2) - h * h * h * k / 6

    assert ev(m, *p) == pytest.approx(su(su(d[0], d[1]), d[2]), abs=1e-14)


def A_translate_inv(c):
    m = np.eye(4)
    m[3, :3] = -np.asarray(c, float)
    return m.reshape(-1, order="F")


# ------------------------------------------------------------------ transforms: test_SDF.f90:303-676, test/matrix, test/vector
def test_rotation_align(oracle):  # :303-341
    a, b = np.array([0, 0, 1.0]), np.array([1.0, 0, 0])
    m = oracle.mat("orc_rotation_align", a, b)
    assert oracle.vec_dot_mat(a, m) == pytest.approx(b, abs=1e-14)
    a, b = np.array([1.0, 2, 1]), np.array([1.0, 4, 5])
    a, b = a / np.linalg.norm(a), b / np.linalg.norm(b)
    assert oracle.vec_dot_mat(a, oracle.mat("orc_rotation_align", a, b)) == pytest.approx(b, abs=1e-14)


def test_rotmat_equals_rotate_axes(oracle):  # :343-369
    assert (oracle.mat("orc_rotmat", [0, 0, 1.0], 45.0) == oracle.mat("orc_rotate_z", 45.0)).all()
    assert (oracle.mat("orc_rotmat", [0, 1.0, 0], 45.0) == oracle.mat("orc_rotate_y", 45.0)).all()
    assert (oracle.mat("orc_rotmat", [1.0, 0, 0], 45.0) == oracle.mat("orc_rotate_x", 45.0)).all()


@pytest.mark.parametrize("angle", [45.0, 90.0, 0.0, 64.45])
def test_rotate_xyz_elements(oracle, angle):  # :371-620 (element (i,j) -> m[i-1, j-1])
    a = np.deg2rad(angle)
    c, s = np.cos(a), np.sin(a)
    x = oracle.mat("orc_rotate_x", angle)
    assert x[:3, :3] == pytest.approx(np.array([[1, 0, 0], [0, c, s], [0, -s, c]]), abs=1e-15)
    y = oracle.mat("orc_rotate_y", angle)
    assert y[:3, :3] == pytest.approx(np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]]), abs=1e-15)
    z = oracle.mat("orc_rotate_z", angle)
    assert z[:3, :3] == pytest.approx(np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]]), abs=1e-15)


def test_identity_translate_skew(oracle):  # :622-676
    assert (oracle.mat("orc_identity") == np.eye(4)).all()
    t = oracle.mat("orc_translate", [1.0, 2.0, 3.0])
    assert (t[3, :3] == [1, 2, 3]).all() and (t[:3, :3] == np.eye(3)).all()
    k = oracle.mat("orc_skew", [1.0, 2.0, 3.0])
    assert k[:3, :3] == pytest.approx(np.array([[0, 3, -2], [-3, 0, 1], [2, -1, 0]]))  # out(:,1)=[0,-az,ay] etc.


def test_vec_dot_identity(oracle):  # test/vector/test_vec3.f90:494-514
    v = np.array([1.0, 2.0, 3.0])
    assert (oracle.vec_dot_mat(v, np.eye(4)) == v).all()
    assert oracle.vec_dot_mat(v, oracle.mat("orc_translate", [0.5, -1, 2])) == pytest.approx(v + [0.5, -1, 2])


def test_matrix_invert(oracle):  # test/matrix/test_matrix.f90:205-229
    a = np.zeros((4, 4))
    a[:, 0] = [4, 0, 2, 1]; a[:, 1] = [0, 0, 2, 0]; a[:, 2] = [0, 1, 2, 0]; a[:, 3] = [1, 0, 0, 1]
    b = np.zeros((4, 4))
    t = 1 / 3
    b[:, 0] = [t, -t, 0, -t]; b[:, 1] = [0, -1, 1, 0]; b[:, 2] = [0, 0.5, 0, 0]; b[:, 3] = [-t, t, 0, 1 + t]
    c = oracle.mat("orc_invert", a)
    assert c == pytest.approx(b, abs=1e-14)
    assert oracle.mat("orc_matmul", a, c) == pytest.approx(np.eye(4), abs=1e-14)
    rng = np.random.default_rng(0)
    for _ in range(20):
        m = rng.normal(size=(4, 4))
        assert oracle.mat("orc_matmul", m, oracle.mat("orc_invert", m)) == pytest.approx(np.eye(4), abs=1e-9)


# ------------------------------------------------------------------ test/fresnel/test_fresnel.f90:34-189
def test_simple_refract_normal_incidence(oracle):  # :118-152
    d, R, fl = oracle.fresnel([[0, 0, -1.0]], [[0, 0, 1.0]], 1.0, 1.33, 0.3)
    assert fl[0] == 0 and np.pi - np.arccos(d[0, 2]) == pytest.approx(0.0, abs=1e-10)


def test_simple_reflect_tir(oracle):  # :154-189 (50 degrees, 1.33 -> 1.0 is beyond the critical angle)
    th = np.deg2rad(50.0)
    I = np.array([np.sin(th), 0, np.cos(th)])
    d, R, fl = oracle.fresnel([I], [[0, 0, 1.0]], 1.33, 1.0, 0.999)
    assert fl[0] == 1 and R[0] == 1.0
    assert d[0] == pytest.approx([I[0], I[1], -I[2]], abs=1e-15)


def test_complex_refract_frequency(oracle):  # :34-75: 1e6 draws, 45 deg, 1 -> 1.33
    th = np.deg2rad(180 + 45.0)
    I = np.array([abs(np.sin(th)), 0.0, np.cos(th)])
    I /= np.linalg.norm(I)
    n = 1_000_000
    xi = np.random.default_rng(123456789).random(n)
    d, R, fl = oracle.fresnel(np.tile(I, (n, 1)), np.tile([0, 0, 1.0], (n, 1)), 1.0, 1.33, xi)
    snell = np.arcsin(1.0 / 1.33 * np.sin(np.deg2rad(45.0)))
    ok = np.abs((np.pi - np.arccos(d[:, 2])) - snell) < 1e-10
    assert ok.mean() == pytest.approx(1 - R[0], abs=5e-4)
    assert (ok == (fl == 0)).all()


def test_complex_reflect_frequency(oracle):  # :77-116: 45 deg, 1.33 -> 1.0
    th = np.deg2rad(45.0)
    I = np.array([np.sin(th), 0.0, np.cos(th)])
    I /= np.linalg.norm(I)
    n = 1_000_000
    xi = np.random.default_rng(5).random(n)
    d, R, fl = oracle.fresnel(np.tile(I, (n, 1)), np.tile([0, 0, 1.0], (n, 1)), 1.33, 1.0, xi)
    hit = (d[:, 2] == -I[2]) & (fl == 1)
    assert hit.mean() == pytest.approx(R[0], abs=5e-4)
    # unpolarised Fresnel coefficient, independent formula
    ci, st = np.cos(th), 1.33 * np.sin(th)
    ct = np.sqrt(1 - st * st)
    rs, rp = (1.33 * ci - ct) / (1.33 * ci + ct), (1.33 * ct - ci) / (1.33 * ct + ci)
    assert R[0] == pytest.approx(0.5 * (rs * rs + rp * rp), rel=1e-12)


# ------------------------------------------------------------------ test/detector/test_detector.f90:27-182, test/geometry/test_geometry.f90:182-253
def det_scene(oracle, kind, p, nbins):
    scene = A.Scene.from_primitives([(A.BOX, None, [50, 50, 50])], [(0, 0, 0, 1)])
    q = np.zeros((1, 20)); q[0, :len(p)] = p
    return oracle.OracleScene(scene, ((10, 10, 10), (50.0, 50.0, 50.0)), None, ([kind], q, [nbins]))


def test_hit_circle(oracle):  # test_detector.f90:27-69
    s = det_scene(oracle, A.DET_CIRCLE, [0.5, 0, 0, 1, 0, 0, 0.5], 100)
    hit, b = s.detector(1, [[0, 0, 0]], [[1.0, 0, 0]], 1.0)
    assert hit[0] == 1 and b[0] == 1


def test_hit_camera(oracle):  # :72-126
    s = det_scene(oracle, A.DET_CAMERA, [-1, -1, -1, 0, 2, 0, 0, 0, 2, 100.0], 100)
    assert s.detector(1, [[10.0, 0, 0]], [[-1.0, 0, 0]], 1.0)[0][0] == 1
    assert s.detector(1, [[10.0, 0, 0]], [[1.0, 0, 0]], 1.0)[0][0] == 0


def test_hit_annulus(oracle):  # :128-182
    s = det_scene(oracle, A.DET_ANNULUS, [0.5, 0, 0, 1, 0, 0, 0.5, 1.0], 100)
    assert s.detector(1, [[0, 0.75, 0]], [[1.0, 0, 0]], 1.0)[0][0] == 1
    assert s.detector(1, [[0, 0, 0]], [[1.0, 0, 0]], 1.0)[0][0] == 0


def test_plane_and_circle_intersection(oracle):  # test_geometry.f90:182-253
    s = det_scene(oracle, A.DET_CIRCLE, [0, 0, 0, 0, 0, 1, 10.0], 100)
    hit, b = s.detector(1, [[0, 0, -10.0]], [[0, 0, 1.0]], 10.0)  # hit point = origin -> radius 0 -> bin 1
    assert hit[0] == 1 and b[0] == 1
    assert s.detector(1, [[0, 0, -20.0]], [[0, 0, -1.0]], 100.0)[0][0] == 0
    assert s.detector(1, [[0, 0, -10.0]], [[0, 0, 1.0]], 9.0)[0][0] == 0  # t > pointSep (detectors.f90:160)


def test_detector_defaults(oracle, smcrt):  # test/parse/test_parse.f90:613-700: stored nbins = user nbins + 1
    cfg = smcrt.Config.loads('''
[source]
name="point"
position=[0.0,0.0,0.0]
[grid]
[geometry]
geom_name="scat_test"
[[detectors]]
type="annulus"
ID="a"
position=[0.0,0.0,0.0]
[[detectors]]
type="circle"
ID="c"
position=[0.0,0.0,0.0]
nbins=10
''')
    kind, p, nb, ids = cfg.detectors
    assert list(kind) == [A.DET_CIRCLE, A.DET_ANNULUS] and ids == ["c", "a"]  # circles first (parse_detectors.f90:119-137)
    assert list(nb) == [10, 100]
    assert p[0, 3:6] == pytest.approx([0, 0, -1]) and p[0, 6] == 1.0
    assert p[1, 6] == 0.1 and p[1, 7] == 0.2
    osc = oracle.OracleScene.from_config(cfg)
    assert osc.det_total == 11 + 101


# ------------------------------------------------------------------ test/photon/test_photon.f90:64-255
def src_scene(oracle, kind, sub, p, grid):
    scene = A.Scene.from_primitives([(A.BOX, None, [10, 10, 10])], [(0, 0, 0, 1)])
    return oracle.OracleScene(scene, grid, (kind, sub, p))


def test_uniform_source(oracle):  # :64-121
    p = np.zeros(24)
    p[3:6] = [1, 0, 0]
    p[6:9] = [-7.5, -1, -1]; p[9:12] = [0, 2, 0]; p[12:15] = [0, 0, 2]
    s = src_scene(oracle, A.SRC_UNIFORM, 0, p, ((200, 200, 200), (7.5, 7.5, 7.5)))
    xi = np.random.default_rng(1).random((10000, 4))
    pos, d, cell, ok = s.emit(xi)
    assert (pos[:, 0] == -7.5 + 7.9e-7).all()
    assert (np.abs(pos[:, 1:]) <= 1.0).all()
    assert (d == [1, 0, 0]).all()


def test_point_and_pencil_source(oracle):  # :123-166, :213-255
    p = np.zeros(24)
    p[0:3] = [0.0, 0.5, -0.25]; p[3:6] = [1, 0, 0]
    xi = np.random.default_rng(2).random((1000, 4))
    for kind in (A.SRC_POINT, A.SRC_PENCIL):
        pos, d, cell, ok = src_scene(oracle, kind, 0, p, ((200, 200, 200), (1.0, 1.0, 1.0))).emit(xi)
        assert (pos == [0.0, 0.5, -0.25]).all()
        assert np.linalg.norm(d, axis=1) == pytest.approx(1.0, abs=1e-12)
    # isotropy of the point source: <cos> = 0, <cos^2> = 1/3
    assert abs(d[:, 2].mean()) < 0.06


def test_circular_source(oracle):  # :168-211
    p = np.zeros(24)
    p[0:3] = [0, 0, 1.0]; p[3:6] = [0, 0, -1.0]; p[15] = 2.5
    s = src_scene(oracle, A.SRC_CIRCULAR, 0, p, ((200, 200, 200), (1.1, 1.1, 1.1)))
    pos, d, cell, ok = s.emit(np.random.default_rng(12345678).random((10000, 4)))
    r = np.hypot(pos[:, 0], pos[:, 1])
    assert (r <= 2.5 + 1e-12).all() and (pos[:, 2] == 1.0).all()
    assert 0.6 < (r ** 2).mean() / (2.5 ** 2 / 2) < 1.4  # uniform disc: <r^2> = R^2/2


# ------------------------------------------------------------------ RNG: Philox4x32-10 known answers (Random123 kat_vectors)
def test_philox_known_answers(oracle):
    assert [hex(x) for x in oracle.philox_raw([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in oracle.philox_raw([0xffffffff] * 4, [0xffffffff] * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in oracle.philox_raw([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_uniform_conversions(oracle):  # test/random/test_random.f90: ranges
    u = oracle.uniforms([0, 0xffffffff, 0x80000000, 0])
    assert u[0] == 0.0 and u[1] == 1 - 2.0 ** -24 and u[2] == 0.5
    assert u[4] == 2.0 ** -32  # tau draw is in (0, 1]
    assert oracle.uniforms([0, 0, 0, 0xffffffff])[4] == 1.0
