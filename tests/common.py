"""Shared scene builders for the tests (same flattened bytes go to the engine and to the oracle)."""
import numpy as np

from rsmcrt_b200 import api as A


def fmat_translate_inv(c):
    """invert(translate(c)) in Fortran storage: p' = p - c  (translation in row 4)."""
    m = np.eye(4)
    m[3, :3] = -np.asarray(c, float)
    return m.reshape(-1, order="F")


def fmat_from_rows(m):
    """m[i-1, j-1] = Fortran M(i,j) -> flat column-major."""
    return np.asarray(m, float).reshape(4, 4).reshape(-1, order="F")


def zoo_scene(oracle):
    """One top-level SDF of every primitive kind (some with transforms) + a smooth-union model + modifiers."""
    rot = oracle.mat("orc_invert", oracle.mat("orc_rotmat", [1.0, 2.0, 0.5], 37.0))
    rt = oracle.mat("orc_matmul", rot, oracle.mat("orc_invert", oracle.mat("orc_translate", [0.1, -0.2, 0.05])))
    F = lambda m: np.asarray(m).reshape(-1, order="F")
    prims = [
        (A.SPHERE, fmat_translate_inv([0.2, 0.1, -0.3]), [0.45]),
        (A.BOX, F(rt), [0.3, 0.2, 0.4]),
        (A.TORUS, F(rot), [0.5, 0.12]),
        (A.CYLINDER, None, [-0.3, 0.1, 0.0, 0.4, -0.2, 0.3, 0.15]),
        (A.TRIPRISM, fmat_translate_inv([0.0, 0.1, 0.0]), [0.4, 0.3]),
        (A.SEGMENT, None, [-0.2, -0.2, 0.1, 0.5, 0.2, 0.0]),
        (A.CAPSULE, F(rt), [-0.2, 0.0, 0.1, 0.3, 0.2, -0.1, 0.08]),
        (A.CONE, None, [0.0, -0.3, 0.0, 0.1, 0.4, 0.05, 0.3, 0.05]),
        (A.EGG, None, [0.4, 0.2, 0.3]),
        (A.PLANE, F(rot), [0.0, 0.6, 0.8]),
    ]
    nn = len(prims)
    kind = [p[0] for p in prims]
    xf = [np.eye(4).reshape(-1) if p[1] is None else np.asarray(p[1]) for p in prims]
    par = [list(p[2]) + [0.0] * (8 - len(p[2])) for p in prims]
    first = [0] * nn
    nch = [0] * nn
    top = list(range(nn))

    def add(k, m, p, fc=0, nc=0):
        kind.append(k); xf.append(np.eye(4).reshape(-1) if m is None else np.asarray(m)); par.append(list(p) + [0.0] * (8 - len(p)))
        first.append(fc); nch.append(nc)
        return len(kind) - 1

    # smooth-union model of torus + 2 cylinders (omg-like)
    m = add(A.MODEL_SMOOTHUNION, None, [0.09]); top.append(m)
    c0 = add(A.TORUS, fmat_translate_inv([0, 0, -0.2]), [0.2, 0.05])
    add(A.CYLINDER, None, [-.25, 0, -.25, .25, 0, .0, .05])
    add(A.CYLINDER, F(rot), [.25, 0, .0, -.25, 0, .25, .05])
    first[m], nch[m] = c0, 3
    # subtraction(sphere, box), intersection, union
    for op in (A.MODEL_SUBTRACTION, A.MODEL_INTERSECTION, A.MODEL_UNION):
        m = add(op, None, [0.0]); top.append(m)
        c0 = add(A.SPHERE, fmat_translate_inv([0.1, 0.0, 0.0]), [0.35])
        add(A.BOX, None, [0.3, 0.3, 0.3])
        first[m], nch[m] = c0, 2
    # revolution(egg), extrude(segment), onion(sphere), twist(box), bend(box), elongate(torus)
    for mk, mp, ck, cm, cp in (
        (A.MOD_REVOLUTION, [0.0, 0.05, -0.1, 0.02], A.EGG, None, [0.4, 0.25, 0.3]),
        (A.MOD_EXTRUDE, [0.2], A.SEGMENT, None, [-0.3, -0.1, 0.0, 0.4, 0.3, 0.0]),
        (A.MOD_ONION, [0.05], A.SPHERE, fmat_translate_inv([0.0, 0.2, 0.0]), [0.4]),
        (A.MOD_TWIST, [1.5], A.BOX, None, [0.3, 0.15, 0.5]),
        (A.MOD_BEND, [0.8], A.BOX, None, [0.5, 0.1, 0.2]),
        (A.MOD_ELONGATE, [0.1, 0.2, 0.05], A.TORUS, None, [0.3, 0.08]),
    ):
        m = add(mk, None, mp); top.append(m)
        c0 = add(ck, cm, cp)
        first[m], nch[m] = c0, 1
    # nested: onion(extrude-less) -> union(model) inside a revolution to exercise stack depth
    m = add(A.MOD_ONION, None, [0.02]); top.append(m)
    u = add(A.MODEL_UNION, None, [0.0]); first[m], nch[m] = u, 1
    c0 = add(A.SPHERE, fmat_translate_inv([0.2, 0.0, 0.0]), [0.25])
    add(A.CAPSULE, None, [-0.3, 0.0, 0.0, 0.1, 0.2, 0.1, 0.1])
    first[u], nch[u] = c0, 2
    nt = len(top)
    return A.Scene(np.array(kind, np.int32), np.array(first, np.int32), np.array(nch, np.int32), np.array(xf), np.array(par),
                   np.array(top, np.int32), np.full(nt, 1.0), np.full(nt, 0.1), np.full(nt, 0.5), np.full(nt, 1.3))


def random_dirs(rng, n):
    v = rng.normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def zscore(a, b, var_a, var_b):
    return (a - b) / np.sqrt(np.maximum(var_a + var_b, 1e-300))
