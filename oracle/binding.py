"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE (see oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"

dp = C.POINTER(C.c_double)
fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int32)
up = C.POINTER(C.c_uint32)


class OrcCounters(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("nscatt", "sdf_evals", "bounces", "launched", "emit_retries", "lost", "sweeps", "det_hits")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(force=False):
    src = HERE / "oracle.cpp"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        r = subprocess.run(["make", "-C", str(HERE), "-B" if force else "-s"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB.exists():
        build()
    L = C.CDLL(str(LIB))
    L.orc_scene_create.restype = C.c_void_p
    L.orc_scene_create.argtypes = [C.c_int, ip, ip, ip, dp, dp, C.c_int, ip, dp, dp, dp, dp]
    L.orc_scene_free.argtypes = [C.c_void_p]
    L.orc_set_grid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
    L.orc_set_source.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
    L.orc_set_optprops.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
    L.orc_set_flags.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_set_detectors.restype = C.c_int64
    L.orc_set_detectors.argtypes = [C.c_void_p, C.c_int, ip, dp, ip]
    L.orc_sdf_eval.argtypes = [C.c_void_p, C.c_int, C.c_int64, dp, dp]
    L.orc_sdf_normal.argtypes = [C.c_void_p, C.c_int, C.c_int64, dp, dp]
    L.orc_locate_layer.restype = C.c_int
    L.orc_locate_layer.argtypes = [C.c_void_p, dp, C.c_int]
    L.orc_fresnel.argtypes = [C.c_int64, dp, dp, dp, dp, dp, dp, dp, ip]
    L.orc_scatter.argtypes = [C.c_int64, dp, dp, dp, dp]
    L.orc_emit.argtypes = [C.c_void_p, C.c_int64, dp, dp, dp, ip, ip]
    L.orc_detector.argtypes = [C.c_void_p, C.c_int, C.c_int64, dp, dp, dp, ip, ip]
    L.orc_get_voxel.argtypes = [C.c_void_p, dp, ip]
    L.orc_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, up]
    L.orc_philox_raw.argtypes = [up, up, up]
    L.orc_uniforms.argtypes = [up, dp]
    for name in ("orc_rotate_x", "orc_rotate_y", "orc_rotate_z"):
        getattr(L, name).argtypes = [C.c_double, dp]
    L.orc_rotmat.argtypes = [dp, C.c_double, dp]
    L.orc_rotation_align.argtypes = [dp, dp, dp]
    L.orc_translate.argtypes = [dp, dp]
    L.orc_identity.argtypes = [dp]
    L.orc_skew.argtypes = [dp, dp]
    L.orc_invert.argtypes = [dp, dp]
    L.orc_matmul.argtypes = [dp, dp, dp]
    L.orc_vec_dot_mat.argtypes = [dp, dp, dp]
    L.orc_mono.argtypes = [C.c_double] * 4 + [dp]
    L.orc_test_kernel.restype = C.c_double
    L.orc_test_kernel.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_int, C.c_int, dp]
    L.orc_max_threads.restype = C.c_int
    L.orc_run.restype = C.c_double
    L.orc_run.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                          C.c_int, fp, fp, fp, dp, C.POINTER(OrcCounters), ip, ip, dp, ip]
    _lib = L
    return L


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _P(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def mat(fn, *args):
    """4x4 helpers; returns a (4,4) numpy array indexed [i-1, j-1] like the Fortran (i,j)."""
    L = load()
    out = np.zeros(16)
    cargs = []
    for a in args:
        if np.isscalar(a):
            cargs.append(C.c_double(a))
        else:
            cargs.append(_P(_d(a).reshape(-1, order="F") if np.asarray(a).ndim == 2 else _d(a), C.c_double))
    getattr(L, fn)(*cargs, _P(out, C.c_double))
    return out.reshape(4, 4, order="F")


def vec_dot_mat(v, m):
    out = np.zeros(3)
    load().orc_vec_dot_mat(_P(_d(v), C.c_double), _P(_d(np.asarray(m).reshape(-1, order="F")), C.c_double), _P(out, C.c_double))
    return out


def mono(mus, mua, hgg, n):
    out = np.zeros(7)
    load().orc_mono(mus, mua, hgg, n, _P(out, C.c_double))
    return dict(zip(("mus", "mua", "hgg", "g2", "n", "kappa", "albedo"), out))


def philox(seed, pid, event):
    out = np.zeros(4, np.uint32)
    load().orc_philox(seed, pid, event, _P(out, C.c_uint32))
    return out


def philox_raw(ctr, key):
    out = np.zeros(4, np.uint32)
    c, k = np.asarray(ctr, np.uint32), np.asarray(key, np.uint32)
    load().orc_philox_raw(_P(c, C.c_uint32), _P(k, C.c_uint32), _P(out, C.c_uint32))
    return out


def uniforms(words):
    out = np.zeros(5)
    w = np.asarray(words, np.uint32)
    load().orc_uniforms(_P(w, C.c_uint32), _P(out, C.c_double))
    return out


def fresnel(dir, nrm, n1, n2, xi):
    dir, nrm = _d(dir).reshape(-1, 3), _d(nrm).reshape(-1, 3)
    n = len(dir)
    n1, n2, xi = (_d(np.broadcast_to(v, n)) for v in (n1, n2, xi))
    out, R, fl = np.zeros((n, 3)), np.zeros(n), np.zeros(n, np.int32)
    load().orc_fresnel(n, _P(dir, C.c_double), _P(nrm, C.c_double), _P(n1, C.c_double), _P(n2, C.c_double), _P(xi, C.c_double),
                       _P(out, C.c_double), _P(R, C.c_double), _P(fl, C.c_int32))
    return out, R, fl


def scatter(dir, hgg, xi):
    dir, xi = _d(dir).reshape(-1, 3), _d(xi).reshape(-1, 2)
    n = len(dir)
    hgg = _d(np.broadcast_to(hgg, n))
    out = np.zeros((n, 3))
    load().orc_scatter(n, _P(dir, C.c_double), _P(hgg, C.c_double), _P(xi, C.c_double), _P(out, C.c_double))
    return out


class OracleScene:
    """The oracle's view of one simulation set-up (same flattened bytes as the engine receives)."""

    def __init__(self, scene, grid=None, source=None, detectors=None):
        L = load()
        self.L = L
        a = [_i(scene.kind), _i(scene.first_child), _i(scene.n_child), _d(scene.xform), _d(scene.params), _i(scene.top_node),
             _d(scene.mus), _d(scene.mua), _d(scene.hgg), _d(scene.n)]
        self.n_top = len(a[5])
        self.h = C.c_void_p(L.orc_scene_create(len(a[0]), _P(a[0], C.c_int32), _P(a[1], C.c_int32), _P(a[2], C.c_int32),
                                               _P(a[3], C.c_double), _P(a[4], C.c_double), self.n_top, _P(a[5], C.c_int32),
                                               _P(a[6], C.c_double), _P(a[7], C.c_double), _P(a[8], C.c_double), _P(a[9], C.c_double)))
        self.grid_shape = (200, 200, 200)
        self.det_total = 0
        if grid is not None:
            self.set_grid(*grid[0], *grid[1])
        if source is not None:
            self.set_source(*source)
        if detectors is not None:
            self.set_detectors(*detectors[:3])

    @classmethod
    def from_deck(cls, deck):
        """deck: oracle.scenes.Deck -- the oracle's own TOML -> scene path (no product code involved)."""
        o = cls(deck.scene, deck.grid, deck.source, deck.detectors)
        o.deck = deck
        return o

    @classmethod
    def from_toml(cls, path, res_dir=None):
        from . import scenes
        return cls.from_deck(scenes.load(path, res_dir))

    @classmethod
    def from_config(cls, cfg):
        """cfg: anything with .toml_text (and .res_dir).  Only the deck's TEXT is taken from it: scene, source and detectors are
        built by oracle/scenes.py, independently of the product's host layer."""
        from . import scenes
        return cls.from_deck(scenes.loads(cfg.toml_text, getattr(cfg, "res_dir", None)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_scene_free(self.h)
            self.h = None

    def set_grid(self, nx, ny, nz, xm, ym, zm):
        self.L.orc_set_grid(self.h, nx, ny, nz, xm, ym, zm)
        self.grid_shape = (nx, ny, nz)

    def set_source(self, kind, subtype, p):
        p = _d(p)
        self.L.orc_set_source(self.h, kind, subtype, _P(p, C.c_double))

    def set_optprops(self, top_index, mus, mua, hgg, n):
        self.L.orc_set_optprops(self.h, top_index, mus, mua, hgg, n)

    def set_flags(self, bugcompat=True, launch_mask_le=False):
        self.L.orc_set_flags(self.h, int(bugcompat), int(launch_mask_le))

    def set_detectors(self, kind, p, nbins):
        k, p, nb = _i(kind), _d(p), _i(nbins)
        self.det_total = int(self.L.orc_set_detectors(self.h, len(k), _P(k, C.c_int32), _P(p, C.c_double), _P(nb, C.c_int32)))

    def sdf(self, top_index, pos):
        pos = _d(pos).reshape(-1, 3)
        n = len(pos)
        out = np.zeros(n if top_index > 0 else n * self.n_top)
        self.L.orc_sdf_eval(self.h, top_index, n, _P(pos, C.c_double), _P(out, C.c_double))
        return out if top_index > 0 else out.reshape(n, self.n_top)

    def normal(self, top_index, pos):
        pos = _d(pos).reshape(-1, 3)
        out = np.zeros((len(pos), 3))
        self.L.orc_sdf_normal(self.h, top_index, len(pos), _P(pos, C.c_double), _P(out, C.c_double))
        return out

    def locate_layer(self, pos, le=False):
        p = _d(pos)
        return int(self.L.orc_locate_layer(self.h, _P(p, C.c_double), int(le)))

    def emit(self, xi4):
        xi4 = _d(xi4).reshape(-1, 4)
        n = len(xi4)
        pos, dir, cell, ok = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3), np.int32), np.zeros(n, np.int32)
        self.L.orc_emit(self.h, n, _P(xi4, C.c_double), _P(pos, C.c_double), _P(dir, C.c_double), _P(cell, C.c_int32), _P(ok, C.c_int32))
        return pos, dir, cell, ok

    def detector(self, det_index, start, dir, seg_len):
        start, dir = _d(start).reshape(-1, 3), _d(dir).reshape(-1, 3)
        n = len(start)
        seg_len = _d(np.broadcast_to(seg_len, n))
        hit, b = np.zeros(n, np.int32), np.zeros(n, np.int32)
        self.L.orc_detector(self.h, det_index, n, _P(start, C.c_double), _P(dir, C.c_double), _P(seg_len, C.c_double),
                            _P(hit, C.c_int32), _P(b, C.c_int32))
        return hit, b

    def get_voxel(self, pos):
        p = _d(pos)
        c = np.zeros(3, np.int32)
        self.L.orc_get_voxel(self.h, _P(p, C.c_double), _P(c, C.c_int32))
        return c

    def test_kernel(self, nphotons, seed, end_early=True, nthreads=0):
        """test_kernel of the reference (kernelsMod.f90:2069-2182) -> (mean scatters, moments[2,4,3])."""
        m = np.zeros(24)
        nsc = self.L.orc_test_kernel(self.h, int(nphotons), int(seed), int(end_early), int(nthreads), _P(m, C.c_double))
        return nsc, m.reshape(2, 4, 3)

    def run(self, nphotons, seed, id_offset=0, tally_mode=1, survival_bias=False, threshold=-1.0, chance=-1.0, nthreads=0,
            rng_mode=0, per_packet=False, grids=True):
        """-> dict(seconds, jmean, absorb, emission, det_bins, counters[, fate, nscatt, pos, events])."""
        nv = int(np.prod(self.grid_shape))
        out = {}
        if grids:
            jm, ab, em = (np.zeros(nv, np.float32) for _ in range(3))
            pj, pa, pe = (_P(x, C.c_float) for x in (jm, ab, em))
        else:
            jm = ab = em = None
            pj = pa = pe = None
        bins = np.zeros(max(self.det_total, 1))
        cn = OrcCounters()
        if per_packet:
            fate, nsc, ev = (np.zeros(nphotons, np.int32) for _ in range(3))
            pos = np.zeros((nphotons, 3))
            pp = (_P(fate, C.c_int32), _P(nsc, C.c_int32), _P(pos, C.c_double), _P(ev, C.c_int32))
        else:
            pp = (None, None, None, None)
        secs = self.L.orc_run(self.h, int(nphotons), int(seed), int(id_offset), int(tally_mode), int(survival_bias), threshold,
                              chance, int(nthreads), int(rng_mode), pj, pa, pe, _P(bins, C.c_double), C.byref(cn), *pp)
        out["seconds"] = secs
        if grids:
            out["jmean"] = jm.reshape(self.grid_shape, order="F")
            out["absorb"] = ab.reshape(self.grid_shape, order="F")
            out["emission"] = em.reshape(self.grid_shape, order="F")
        out["det_bins"] = bins[:self.det_total]
        out["counters"] = cn.as_dict()
        if per_packet:
            out.update(fate=fate, nscatt=nsc, pos=pos, events=ev)
        return out


def max_threads():
    return int(load().orc_max_threads())
