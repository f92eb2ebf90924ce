"""oracle/scenes.py -- TEST INFRASTRUCTURE: the oracle's OWN path from an input deck to the flattened scene.

An independent Python restatement (tomllib + numpy) of the reference's set-up layer, written from the reference sources and
NOT from rsmcrt_b200/csrc/host/host.cpp, so that
  * the oracle (and bench.py --impl reference) never loads libsmcrt_gpu.so, and
  * tests/test_host_and_abi.py can check the product's TOML -> scene builder against a second implementation value by value
    (a wrong builder is no longer common-mode between engine and oracle).

    parse_source      src/parse/parse_source.f90:58-255          parse_grid        src/parse/parse.f90:92-110
    parse_geometry    src/parse/parse_geometry.f90:45-282        parse_detectors   src/parse/parse_detectors.f90:17-349
    scene builders    src/setupGeometry.f90:10-652; `jacques`, `skin`, `lens` are builder-defined

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import math
import tomllib
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

# node kinds / sources / detectors: the numbering of include/smcrt.h (the flattened bytes are the interface)
SPHERE, BOX, TORUS, CYLINDER, TRIPRISM, SEGMENT, CAPSULE, CONE, EGG, PLANE = range(1, 11)
MODEL_UNION, MODEL_SMOOTHUNION, MODEL_SUBTRACTION, MODEL_INTERSECTION = 20, 21, 22, 23
MOD_REVOLUTION = 30
SRC = {"point": 1, "pencil": 2, "uniform": 3, "circular": 4, "focus": 5, "annulus": 6, "dslit": 7, "aperture": 8}
FOCUS_SUB = {"square": 1, "circle": 2, "gaussian": 3}
ANNULUS_SUB = {"tophat": 1, "besselAnnulus": 2, "gaussian": 3}
DET_CIRCLE, DET_ANNULUS, DET_FIBRE, DET_CAMERA = 1, 2, 3, 4
NODE_PARAMS, SOURCE_PARAMS, DET_PARAMS = 8, 24, 20


# ---------------------------------------------------------------------------------------------- 4x4 kit (Fortran M(i,j))
def translate(o):
    """src/sdfs/sdfHelpers.f90:168-182: the offset sits in ROW 4."""
    m = np.eye(4)
    m[3, :3] = o
    return m


def rotate_y(deg):
    """src/sdfs/sdfHelpers.f90:43-62 (columns given as r(:,j))."""
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    m = np.eye(4)
    m[:, 0] = [c, 0, s, 0]
    m[:, 2] = [-s, 0, c, 0]
    return m


def invert(m):
    return np.linalg.inv(m)


def flat(m):
    """M(i,j) -> 16 doubles in Fortran (column-major) storage order."""
    return np.asarray(m, float).reshape(4, 4).reshape(-1, order="F")


# ---------------------------------------------------------------------------------------------- scene tree
@dataclass
class Node:
    kind: int
    params: list
    xform: np.ndarray = field(default_factory=lambda: np.eye(4))
    kids: list = field(default_factory=list)


@dataclass
class FlatScene:
    kind: np.ndarray
    first_child: np.ndarray
    n_child: np.ndarray
    xform: np.ndarray
    params: np.ndarray
    top_node: np.ndarray
    mus: np.ndarray
    mua: np.ndarray
    hgg: np.ndarray
    n: np.ndarray

    @property
    def n_top(self):
        return len(self.top_node)


def flatten(tops):
    """tops: list of (Node, mus, mua, hgg, n).  Children of a node occupy consecutive slots (first_child, n_child)."""
    kind, first, nch, xf, par, top = [], [], [], [], [], []

    def slot():
        kind.append(0); first.append(0); nch.append(0); xf.append(np.zeros(16)); par.append(np.zeros(NODE_PARAMS))
        return len(kind) - 1

    def fill(i, nd):
        kind[i] = nd.kind
        xf[i] = flat(nd.xform)
        p = np.zeros(NODE_PARAMS)
        p[:len(nd.params)] = nd.params
        par[i] = p
        nch[i] = len(nd.kids)
        if nd.kids:
            ids = [slot() for _ in nd.kids]
            first[i] = ids[0]
            for j, k in zip(ids, nd.kids):
                fill(j, k)

    opt = []
    for nd, *o in tops:
        i = slot()
        top.append(i)
        fill(i, nd)
        opt.append(o)
    o = np.asarray(opt, float).reshape(len(tops), 4)
    return FlatScene(np.array(kind, np.int32), np.array(first, np.int32), np.array(nch, np.int32), np.array(xf), np.array(par),
                     np.array(top, np.int32), o[:, 0].copy(), o[:, 1].copy(), o[:, 2].copy(), o[:, 3].copy())


def sphere(r, at=None):
    return Node(SPHERE, [r], np.eye(4) if at is None else invert(translate(at)))


def box(lengths, at=None):
    """box(lengths) stores HALF lengths (src/sdfs/sdfs.f90:455)."""
    return Node(BOX, [0.5 * v for v in lengths], np.eye(4) if at is None else invert(translate(at)))


def cylinder(a, b, r, xform=None):
    return Node(CYLINDER, [*a, *b, r], np.eye(4) if xform is None else xform)


class SplitMix64:
    """The seeded stand-in for the reference's UNSEEDED ranu() in setup_sphere_scene."""

    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def uni(self):
        M = 0xFFFFFFFFFFFFFFFF
        self.s = (self.s + 0x9E3779B97F4A7C15) & M
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        z ^= z >> 31
        return (z >> 11) / 9007199254740992.0

    def ranu(self, a, b):
        return a + self.uni() * (b - a)


# ---------------------------------------------------------------------------------------------- the deck
@dataclass
class Deck:
    """What `setup()` leaves behind, as far as the packet path reads it."""
    scene: FlatScene
    grid: tuple        # ((nxg, nyg, nzg), (xmax, ymax, zmax))
    source: tuple      # (kind, subtype, params[24])
    detectors: tuple   # (kind[n], params[n, 20], nbins[n], ids)
    nphotons: int
    iseed: int
    geom_name: str
    source_name: str


def _vec3(t, key, default=None, required=False):
    v = t.get(key)
    if v is None or not isinstance(v, list):
        if required:
            raise ValueError(f"'{key}' is required")
        return None if default is None else list(default)
    if len(v) != 3:
        raise ValueError(f"'{key}': expected vector of size 3")
    return [float(x) for x in v]


def _unit(v):
    l = math.sqrt(sum(x * x for x in v))
    return [x / l for x in v]


def parse_source(t):
    name = t.get("name", "point")
    if name not in SRC:
        raise ValueError("No such source!")
    p = np.zeros(SOURCE_PARAMS)
    if name != "uniform":
        p[0:3] = _vec3(t, "position", required=True)
    if name in ("focus", "annulus", "dslit", "aperture"):
        rot = _vec3(t, "rotation", required=True)
        p[21:24] = _unit(rot)
    d = t.get("direction")
    if isinstance(d, list):
        p[3:6] = _vec3(t, "direction")
    elif isinstance(d, str):
        axis = {"x": 0, "y": 1, "z": 2}[d.lstrip("-")]
        p[3 + axis] = -1.0 if d.startswith("-") else 1.0
    elif name not in ("point", "annulus", "focus"):
        raise ValueError("Need to specify direction for source type!")
    corners = [[-1, -1, 1], [2, 0, 0], [0, 2, 0]]        # parse_source.f90:52-56
    for k, key in enumerate(("point1", "point2", "point3")):
        v = t.get(key)
        if isinstance(v, list):
            corners[k] = [float(x) for x in v[:3]]
        elif name == "uniform":
            raise ValueError(f"Uniform source requires {key} variable")
    p[6:9], p[9:12], p[12:15] = corners
    p[15] = t.get("wavelength", 500.0) if name in ("dslit", "aperture") else t.get("radius", 0.5)   # slot 15: see include/smcrt.h
    p[16] = t.get("focalLength", 1.0)
    p[17] = t.get("beam_size", 0.5)
    p[18] = t.get("rlo", 0.5)
    p[19] = t.get("rhi", 0.6)
    p[20] = t.get("sigma", 0.04)
    sub = 0
    if name == "focus":
        sub = FOCUS_SUB[t.get("focus_type", "gaussian")]
    elif name == "annulus":
        sub = ANNULUS_SUB[t.get("annulus_type", "gaussian")]
    return (SRC[name], sub, p), int(t.get("nphotons", 1000000)), name


def parse_detectors(arr):
    groups = {"circle": [], "annulus": [], "fibre": [], "camera": []}
    for t in arr or []:
        ty = t.get("type")
        if "ID" not in t:
            raise ValueError("Need to specify a detector ID")
        if ty not in groups:
            raise ValueError("Invalid detector type. Valid types are [circle, annulus, camera]")
        p = np.zeros(DET_PARAMS)
        d = _vec3(t, "direction", default=[0, 0, -1])
        if ty == "circle":
            p[0:3] = _vec3(t, "position", required=True)
            p[3:6] = _unit(d)
            p[6] = t.get("radius", 1.0)
            groups[ty].append((DET_CIRCLE, p, int(t.get("nbins", 100)), t["ID"]))
        elif ty == "annulus":
            p[0:3] = _vec3(t, "position", required=True)
            p[3:6] = d                                        # NOT normalised (parse_detectors.f90:318)
            p[6], p[7] = t.get("radius1", 0.1), t.get("radius2", 0.2)
            if p[7] <= p[6]:
                raise ValueError("Radii are invalid")
            groups[ty].append((DET_ANNULUS, p, int(t.get("nbins", 100)), t["ID"]))
        elif ty == "fibre":
            p[0:3] = _vec3(t, "position", required=True)
            p[3:6] = _unit(d)
            f1, f2 = t.get("focalLength1", 1.0), t.get("focalLength2", 1.0)
            a1, a2 = t.get("f1Aperture", 1.0), t.get("f2Aperture", 1.0)
            p[6:10] = f1, f2, a1, a2
            p[10] = t.get("frontOffset", 0.0)
            p[11] = t.get("backOffset", f2)
            p[12] = t.get("frontToPinSep", f1)
            p[13] = t.get("pinToBackSep", f2)
            p[14] = t.get("pinAperture", max(a1, a2))
            p[15] = t.get("acceptanceAngle", 90.0)            # the shipped decks write `acceptAngle`: ignored by the reference too
            p[16] = t.get("coreDiameter", 0.01)
            groups[ty].append((DET_FIBRE, p, int(t.get("nbins", 1)), t["ID"]))
        else:
            p[0:3] = _vec3(t, "p1", default=[-1, -1, -1])
            p[3:6] = _vec3(t, "p2", default=[2, 0, 0])
            p[6:9] = _vec3(t, "p3", default=[0, 2, 0])
            p[9] = t.get("maxval", 100.0)
            groups[ty].append((DET_CAMERA, p, int(t.get("nbins", 100)), t["ID"]))
    # dects(:) order: circles, annuli, fibres, cameras -- not file order (parse_detectors.f90:119-137)
    rows = groups["circle"] + groups["annulus"] + groups["fibre"] + groups["camera"]
    n = len(rows)
    return (np.array([r[0] for r in rows], np.int32), np.array([r[1] for r in rows]).reshape(n, DET_PARAMS),
            np.array([r[2] for r in rows], np.int32), [str(r[3]) for r in rows])


def build_scene(g, res_dir=None):
    """[geometry] table -> top-level SDF list (src/setupGeometry.f90; array order = layer index)."""
    name = g.get("geom_name", "sphere")
    nopt = int(g.get("numOptProp", 1))
    arr = lambda key, d: [float(x) for x in g[key]] if isinstance(g.get(key), list) else [d] * nopt
    mua, mus, hgg, nref = arr("mua", 0.0), arr("mus", 1.0), arr("hgg", 0.0), arr("n", 1.0)
    pos = _vec3(g, "position", default=[0, 0, 0])
    bbox = _vec3(g, "boundingBox", default=[2, 2, 2])
    vac = (0.0, 0.0, 0.0, 1.0)
    if name == "sphere":                                       # :10-71
        return [(sphere(g.get("sphereRadius", 1.0), pos), mus[0], mua[0], hgg[0], nref[0]), (box(bbox), *vac)]
    if name in ("box", "test_box"):                            # :73-147
        return [(box(_vec3(g, "BoxDimensions", default=[1, 1, 1]), pos), mus[0], mua[0], hgg[0], nref[0]), (box(bbox), *vac)]
    if name == "egg":                                          # :149-248: yolk, albumen, shell, bounding box
        d = 3.0 * math.sqrt(2.0 - math.sqrt(2.0))
        r1, r2, h = g.get("BottomSphereRadius", 3.0), g.get("TopSphereRadius", d), g.get("SphereSep", d)
        k = 1.0 - g.get("ShellThickness", 0.05)
        rev = lambda a, b, c: Node(MOD_REVOLUTION, [0.0, *pos], kids=[Node(EGG, [a, b, c])])
        return [(sphere(g.get("YolkRadius", 1.5), pos), mus[2], mua[2], hgg[2], nref[2]),
                (rev(r1 * k, r2 * k, h * k), mus[1], mua[1], hgg[1], nref[1]),
                (rev(r1, r2, h), mus[0], mua[0], hgg[0], nref[0]), (box(bbox), *vac)]
    if name == "sphere_scene":                                 # :250-294
        rng = SplitMix64(0x5343454E45343021)
        tops = []
        for _ in range(int(g.get("num_spheres", 10))):
            r = rng.ranu(0.001, 0.25)
            c = [rng.ranu(-1.0 + r, 1.0 - r) for _ in range(3)]
            tops.append((sphere(r, c), 0.0, 0.0, 0.9, 1.37))
        return tops + [(box([2, 2, 2]), 1e-17, 1e-17, 0.0, 1.0)]
    if name == "aptran":                                       # :335-363
        return [(sphere(0.5, [0, 0, 0]), 0.0, 1e-17, 0.0, 1.33), (box([2, 2, 2]), 0.0, 1e-17, 0.0, 1.0),
                (box([2.01, 2.01, 2.01]), 0.0, 1e7, 0.0, 1.0)]
    if name == "exp":                                          # :365-407
        a, b, hg = [-8, 0, 0], [8, 0, 0], g.get("hgga", 0.7)
        return [(cylinder(a, b, 1.55), g.get("musc", 0.0), g.get("muac", 0.01), hg, 1.3),
                (cylinder(a, b, 1.75), g.get("musb", 0.0), g.get("muab", 0.01), hg, 1.5), (box([20, 20, 20]), *vac)]
    if name == "scat_test":                                    # :409-435
        return [(sphere(1.0), g.get("tau", 10.0), 0.0, 0.0, 1.0), (box([2, 2, 2]), *vac)]
    if name == "scat_test2":                                   # :437-464
        return [(box([200, 200, 200]), g.get("tau", 10.0), 1e-17, hgg[0], 1.0)]
    if name == "omg":                                          # :466-549
        kids = [Node(TORUS, [0.2, 0.05], invert(translate([0, 0, -0.7])))]
        segs = [((-.25, 0, -.25), (-.25, 0, .25)), ((-.25, 0, -.25), (.25, 0, .0)), ((.25, 0, .0), (-.25, 0, .25)),
                ((-.25, 0, .25), (.25, 0, .25)), ((-.25, 0, .5), (.25, 0, .5)), ((-.25, 0, .5), (-.25, 0, .75)),
                ((.25, 0, .5), (.25, 0, .75)), ((.25, 0, .75), (0, 0, .75)), ((0, 0, .625), (0, 0, .75))]
        for i, (a, b) in enumerate(segs):
            kids.append(cylinder(a, b, 0.05, invert(rotate_y(90.0)) if i == 0 else None))
        return [(Node(MODEL_SMOOTHUNION, [0.09], kids=kids), 10.0, 0.16, 0.0, 2.65), (box([2, 2, 2]), *vac)]
    if name == "vessels":                                      # :552-652
        d = Path(res_dir or "res")
        edges = np.loadtxt(d / "edges.dat", ndmin=2)
        nodes = np.loadtxt(d / "nodes.dat", ndmin=2)
        radii = np.loadtxt(d / "radii.dat", ndmin=1)
        res = 0.001
        mx = np.abs(nodes).max(axis=0)
        nodes = ((nodes / mx) - 0.5) * mx * res
        tops = []
        for e in edges:
            i1, i2 = int(e[0]) - 1, int(e[1]) - 1
            tops.append((Node(CAPSULE, [*nodes[i1], *nodes[i2], radii[i1] * res]), 94.0, 231.0, 0.9, 1.37))
        return tops + [(box([0.32, 0.18, 0.26]), 357.0, 0.458, 0.9, 1.37)]
    if name == "jacques":                                      # builder-defined
        return [(box([2, 2, 2]), 100.0, 1.0, 0.9, 1.38), (box([2.02, 2.02, 2.02]), *vac)]
    if name == "skin":                                         # builder-defined five-layer stack
        layers = [(0.002, 1000.0, 0.10, 0.86, 1.50), (0.008, 450.0, 1.50, 0.80, 1.34), (0.020, 300.0, 0.70, 0.90, 1.40),
                  (0.050, 200.0, 0.50, 0.95, 1.39), (0.020, 150.0, 0.20, 0.75, 1.44)]
        tops, top = [], 0.05
        for t, s, a, gg, n in layers:
            tops.append((box([0.1, 0.1, t], [0.0, 0.0, top - 0.5 * t]), s, a, gg, n))
            top -= t
        return tops + [(box([0.102, 0.102, 0.102]), *vac)]
    if name == "lens":                                         # builder-defined bi-convex lens
        kids = [sphere(1.0, [0, 0, 0.8]), sphere(1.0, [0, 0, -0.8])]
        return [(Node(MODEL_INTERSECTION, [0.0], kids=kids), 0.0, 0.0, 0.0, 1.5), (box([2, 2, 2]), *vac)]
    raise ValueError("no such routine")                        # src/setup.f90:58-59


def load(path, res_dir=None) -> Deck:
    return loads(Path(path).read_text(), res_dir)


def loads(text, res_dir=None) -> Deck:
    root = tomllib.loads(text)
    if "source" not in root:
        raise ValueError("Simulation needs Source table")
    source, nphotons, sname = parse_source(root["source"])
    g = root.get("grid")
    if g is None:
        raise ValueError("Need grid table in input param file")
    grid = ((int(g.get("nxg", 200)), int(g.get("nyg", 200)), int(g.get("nzg", 200))),
            (float(g.get("xmax", 1.0)), float(g.get("ymax", 1.0)), float(g.get("zmax", 1.0))))
    geo = root.get("geometry")
    if geo is None:
        raise ValueError("Need geometry table in input param file")
    scene = flatten(build_scene(geo, res_dir))
    dets = parse_detectors(root.get("detectors"))
    sim = root.get("simulation", {})
    return Deck(scene, grid, source, dets, nphotons, int(sim.get("iseed", 123456789)), geo.get("geom_name", "sphere"), sname)
