"""Python face of the host mirror: thin objects over the C ABI (include/smcrt.h, include/smcrt_host.h).

Names follow the reference's driver layer (src/kernelsMod.f90): `Config` is what `setup()` leaves behind
(`state`, `dict`, `dects`, `array`), `Engine.run` is the photon loop of `run_MCRT`, `default_MCRT` the
program entry.  Everything that computes lives in libsmcrt_gpu.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import Counters, SmcrtError, check

TALLY_ABSORB, TALLY_PATHLENGTH, TALLY_EMISSION = 1, 2, 4
FATE_ABSORBED, FATE_ESCAPED, FATE_ROULETTE, FATE_LOST = 0, 1, 2, 3

# node kinds / sources / detectors (include/smcrt.h)
SPHERE, BOX, TORUS, CYLINDER, TRIPRISM, SEGMENT, CAPSULE, CONE, EGG, PLANE = range(1, 11)
MODEL_UNION, MODEL_SMOOTHUNION, MODEL_SUBTRACTION, MODEL_INTERSECTION = 20, 21, 22, 23
MOD_REVOLUTION, MOD_EXTRUDE, MOD_ONION, MOD_TWIST, MOD_BEND, MOD_ELONGATE = 30, 31, 32, 33, 34, 35
SRC_POINT, SRC_PENCIL, SRC_UNIFORM, SRC_CIRCULAR, SRC_FOCUS, SRC_ANNULUS, SRC_DSLIT, SRC_APERTURE = range(1, 9)
DET_CIRCLE, DET_ANNULUS, DET_FIBRE, DET_CAMERA = 1, 2, 3, 4


def _p(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


@dataclass
class Scene:
    """Flattened `array(:)` of type(sdf): the bytes that cross the C ABI."""
    kind: np.ndarray
    first_child: np.ndarray
    n_child: np.ndarray
    xform: np.ndarray   # (n_nodes, 16) Fortran column-major 4x4
    params: np.ndarray  # (n_nodes, 8)
    top_node: np.ndarray
    mus: np.ndarray
    mua: np.ndarray
    hgg: np.ndarray
    n: np.ndarray

    @property
    def n_top(self):
        return len(self.top_node)

    @staticmethod
    def from_primitives(prims, optics):
        """prims: list of (kind, xform16 or None, params list); optics: list of (mus, mua, hgg, n); all top-level."""
        nn = len(prims)
        xf = np.tile(np.eye(4).reshape(-1), (nn, 1))
        pr = np.zeros((nn, _lib.NODE_PARAMS))
        kind = np.zeros(nn, np.int32)
        for i, (k, m, p) in enumerate(prims):
            kind[i] = k
            if m is not None:
                xf[i] = np.asarray(m, float).reshape(-1)
            pr[i, :len(p)] = p
        o = np.asarray(optics, float).reshape(nn, 4)
        return Scene(kind, np.zeros(nn, np.int32), np.zeros(nn, np.int32), xf, pr, np.arange(nn, dtype=np.int32),
                     o[:, 0].copy(), o[:, 1].copy(), o[:, 2].copy(), o[:, 3].copy())


class Config:
    """`setup()`'s product: parsed TOML (`state`, `dict`, `dects`) + the scene built by the geom_name dispatch."""

    def __init__(self, handle, toml_text="", res_dir=None):
        self._h = handle
        self._L = _lib.load()
        self.toml_text = toml_text  # the deck as read (tests hand it to the oracle's own TOML -> scene path, oracle/scenes.py)
        self.res_dir = res_dir

    @classmethod
    def load(cls, toml_path, res_dir=None):
        L = _lib.load()
        h = C.c_void_p()
        check(L.smcrt_config_load(str(toml_path).encode(), None if res_dir is None else str(res_dir).encode(), C.byref(h)))
        with open(toml_path, "r") as f:
            return cls(h, f.read(), res_dir)

    @classmethod
    def loads(cls, text, res_dir=None):
        L = _lib.load()
        h = C.c_void_p()
        check(L.smcrt_config_loads(text.encode(), None if res_dir is None else str(res_dir).encode(), C.byref(h)))
        return cls(h, text, res_dir)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.smcrt_config_free(self._h)
            self._h = None

    @property
    def grid(self):
        n = np.zeros(3, np.int32)
        m = np.zeros(3)
        self._L.smcrt_config_grid(self._h, _p(n, C.c_int32), _p(m, C.c_double))
        return tuple(int(v) for v in n), tuple(float(v) for v in m)

    @property
    def nphotons(self):
        return int(self._L.smcrt_config_nphotons(self._h))

    @property
    def iseed(self):
        return int(self._L.smcrt_config_iseed(self._h))

    @property
    def geom_name(self):
        return self._L.smcrt_config_geom_name(self._h).decode()

    @property
    def source_name(self):
        return self._L.smcrt_config_source_name(self._h).decode()

    @property
    def render_source(self):
        return bool(self._L.smcrt_config_render_source(self._h))

    @property
    def metadata(self):
        return self._L.smcrt_config_metadata(self._h).decode()

    @property
    def source(self):
        k, s = C.c_int32(), C.c_int32()
        p = np.zeros(_lib.SOURCE_PARAMS)
        self._L.smcrt_config_source(self._h, C.byref(k), C.byref(s), _p(p, C.c_double))
        return int(k.value), int(s.value), p

    @property
    def detectors(self):
        n = self._L.smcrt_config_n_detectors(self._h)
        kind = np.zeros(max(n, 1), np.int32)
        nb = np.zeros(max(n, 1), np.int32)
        p = np.zeros((max(n, 1), _lib.DET_PARAMS))
        if n:
            self._L.smcrt_config_detectors(self._h, _p(kind, C.c_int32), _p(p, C.c_double), _p(nb, C.c_int32))
        ids = [self._L.smcrt_config_detector_id(self._h, i).decode() for i in range(n)]
        return kind[:n], p[:n], nb[:n], ids

    @property
    def scene(self) -> Scene:
        nn, nt = C.c_int32(), C.c_int32()
        self._L.smcrt_config_scene_sizes(self._h, C.byref(nn), C.byref(nt))
        nn, nt = nn.value, nt.value
        s = Scene(np.zeros(nn, np.int32), np.zeros(nn, np.int32), np.zeros(nn, np.int32), np.zeros((nn, 16)),
                  np.zeros((nn, _lib.NODE_PARAMS)), np.zeros(nt, np.int32), np.zeros(nt), np.zeros(nt), np.zeros(nt), np.zeros(nt))
        self._L.smcrt_config_scene(self._h, _p(s.kind, C.c_int32), _p(s.first_child, C.c_int32), _p(s.n_child, C.c_int32),
                                   _p(s.xform, C.c_double), _p(s.params, C.c_double), _p(s.top_node, C.c_int32),
                                   _p(s.mus, C.c_double), _p(s.mua, C.c_double), _p(s.hgg, C.c_double), _p(s.n, C.c_double))
        return s

    def write_detectors(self, det_bins, out_dir):
        b = _f64(det_bins)
        check(self._L.smcrt_write_detectors(self._h, _p(b, C.c_double), str(out_dir).encode()))


class Engine:
    """An smcrt_ctx: the sm_100a photon-transport engine on 1..N GPUs of this process."""

    def __init__(self, n_gpus=1, device_ids=None):
        self._L = _lib.load()
        self._h = C.c_void_p()
        ids = None if device_ids is None else _p(_i32(device_ids), C.c_int32)
        check(self._L.smcrt_create(C.byref(self._h), int(n_gpus), ids))
        self.n_voxels = 0
        self.n_top = 0

    def close(self):
        if getattr(self, "_h", None):
            self._L.smcrt_destroy(self._h)
            self._h = None

    __del__ = close

    # ---- set-up ------------------------------------------------------------------------------------
    def set_grid(self, nxg, nyg, nzg, xmax, ymax, zmax):
        check(self._L.smcrt_set_grid(self._h, nxg, nyg, nzg, xmax, ymax, zmax))
        self.grid_shape = (nxg, nyg, nzg)
        self.n_voxels = nxg * nyg * nzg

    def set_scene(self, s: Scene):
        a = [_i32(s.kind), _i32(s.first_child), _i32(s.n_child), _f64(s.xform), _f64(s.params), _i32(s.top_node),
             _f64(s.mus), _f64(s.mua), _f64(s.hgg), _f64(s.n)]
        check(self._L.smcrt_set_scene(self._h, len(a[0]), _p(a[0], C.c_int32), _p(a[1], C.c_int32), _p(a[2], C.c_int32),
                                      _p(a[3], C.c_double), _p(a[4], C.c_double), len(a[5]), _p(a[5], C.c_int32),
                                      _p(a[6], C.c_double), _p(a[7], C.c_double), _p(a[8], C.c_double), _p(a[9], C.c_double)))
        self.n_top = len(a[5])

    def set_optprops(self, top_index, mus, mua, hgg, n):
        check(self._L.smcrt_set_optprops(self._h, top_index, mus, mua, hgg, n))

    def set_source(self, kind, subtype, params):
        p = _f64(params)
        assert p.size == _lib.SOURCE_PARAMS
        check(self._L.smcrt_set_source(self._h, kind, subtype, _p(p, C.c_double)))

    def set_detectors(self, kind, params, nbins):
        k, p, nb = _i32(kind), _f64(params), _i32(nbins)
        n = len(k)
        self._n_det = n
        check(self._L.smcrt_set_detectors(self._h, n, _p(k, C.c_int32) if n else None, _p(p, C.c_double) if n else None,
                                          _p(nb, C.c_int32) if n else None))

    def set_tolerances(self, eps0=-1.0, eps_rel=-1.0, max_steps=-1):
        check(self._L.smcrt_set_tolerances(self._h, eps0, eps_rel, max_steps))

    def apply(self, cfg: Config):
        check(self._L.smcrt_config_apply(cfg._h, self._h))
        (nx, ny, nz), _ = cfg.grid
        self.grid_shape = (nx, ny, nz)
        self.n_voxels = nx * ny * nz
        self.n_top = cfg.scene.n_top
        self._n_det = int(self._L.smcrt_config_n_detectors(cfg._h))

    # ---- run ---------------------------------------------------------------------------------------
    def run(self, nphotons, seed, id_offset=0, tally_mode=TALLY_ABSORB, survival_bias=False, threshold=-1.0, chance=-1.0):
        check(self._L.smcrt_run(self._h, int(nphotons), int(seed), int(id_offset), int(tally_mode), int(survival_bias),
                                threshold, chance))

    def run_async(self, nphotons, seed, id_offset=0, tally_mode=TALLY_ABSORB, survival_bias=False, threshold=-1.0, chance=-1.0):
        check(self._L.smcrt_run_async(self._h, int(nphotons), int(seed), int(id_offset), int(tally_mode), int(survival_bias),
                                      threshold, chance))

    def wait(self):
        check(self._L.smcrt_wait(self._h))

    @property
    def last_run_ms(self):
        return float(self._L.smcrt_last_run_ms(self._h))

    @property
    def launch_count(self):
        return int(self._L.smcrt_launch_count(self._h))

    @property
    def det_bins_total(self):
        return int(self._L.smcrt_det_bins_total(self._h))

    def fetch(self, jmean=False, absorb=True, emission=False, detectors=True):
        """-> dict(jmean, absorb, emission: float32 (nxg,nyg,nzg) Fortran-ordered views; det_bins; counters)."""
        out = {}
        ptr = {}
        for name, want in (("jmean", jmean), ("absorb", absorb), ("emission", emission)):
            if want:
                out[name] = np.zeros(self.n_voxels, np.float32)
                ptr[name] = _p(out[name], C.c_float)
            else:
                ptr[name] = None
        nb = self.det_bins_total
        bins = np.zeros(max(nb, 1))
        cn = Counters()
        check(self._L.smcrt_fetch(self._h, ptr["jmean"], ptr["absorb"], ptr["emission"],
                                  _p(bins, C.c_double) if (detectors and nb) else None, C.byref(cn), 0))
        for name in list(out):
            out[name] = out[name].reshape(self.grid_shape, order="F")
        out["det_bins"] = bins[:nb]
        out["counters"] = cn.as_dict()
        return out

    def fetch_into(self, absorb=None, jmean=None, emission=None, det_bins=None, accumulate=False):
        """smcrt_fetch into caller-owned (ideally pinned, see pin_host) float32 / float64 buffers. -> counters dict."""
        cn = Counters()
        ptr = lambda a, t: None if a is None else _p(a, t)
        check(self._L.smcrt_fetch(self._h, ptr(jmean, C.c_float), ptr(absorb, C.c_float), ptr(emission, C.c_float),
                                  ptr(det_bins, C.c_double), C.byref(cn), int(accumulate)))
        return cn.as_dict()

    @staticmethod
    def pin_host(a: np.ndarray):
        check(_lib.load().smcrt_pin_host(a.ctypes.data, a.nbytes))

    @staticmethod
    def unpin_host(a: np.ndarray):
        check(_lib.load().smcrt_unpin_host(a.ctypes.data))

    def reset_tallies(self):
        check(self._L.smcrt_reset_tallies(self._h))

    # ---- multi-process reduce ----------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(_lib.load().smcrt_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, nranks, rank, uid: bytes):
        check(self._L.smcrt_comm_init(self._h, nranks, rank, C.create_string_buffer(uid, 128)))

    def comm_reduce(self, root=0):
        check(self._L.smcrt_comm_reduce(self._h, root))

    # ---- deterministic-component probes ---------------------------------------------------------
    def probe_sdf(self, top_index, pos, normals=False):
        pos = _f64(pos).reshape(-1, 3)
        n = len(pos)
        dist = np.zeros(n if top_index > 0 else n * self.n_top)
        nrm = np.zeros((n, 3))
        check(self._L.smcrt_probe_sdf(self._h, top_index, n, _p(pos, C.c_double), _p(dist, C.c_double),
                                      _p(nrm, C.c_double) if normals else None))
        if top_index == 0:
            dist = dist.reshape(n, self.n_top)
        return (dist, nrm) if normals else dist

    def probe_fresnel(self, dir, nrm, n1, n2, xi):
        dir, nrm = _f64(dir).reshape(-1, 3), _f64(nrm).reshape(-1, 3)
        n = len(dir)
        n1, n2, xi = (_f64(np.broadcast_to(v, n)) for v in (n1, n2, xi))
        out, R, fl = np.zeros((n, 3)), np.zeros(n), np.zeros(n, np.int32)
        check(self._L.smcrt_probe_fresnel(self._h, n, _p(dir, C.c_double), _p(nrm, C.c_double), _p(n1, C.c_double),
                                          _p(n2, C.c_double), _p(xi, C.c_double), _p(out, C.c_double), _p(R, C.c_double),
                                          _p(fl, C.c_int32)))
        return out, R, fl

    def probe_scatter(self, dir, hgg, xi):
        dir, xi = _f64(dir).reshape(-1, 3), _f64(xi).reshape(-1, 2)
        n = len(dir)
        hgg = _f64(np.broadcast_to(hgg, n))
        out = np.zeros((n, 3))
        check(self._L.smcrt_probe_scatter(self._h, n, _p(dir, C.c_double), _p(hgg, C.c_double), _p(xi, C.c_double), _p(out, C.c_double)))
        return out

    def probe_emit(self, xi4):
        xi4 = _f64(xi4).reshape(-1, 4)
        n = len(xi4)
        pos, dir, cell = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3), np.int32)
        check(self._L.smcrt_probe_emit(self._h, n, _p(xi4, C.c_double), _p(pos, C.c_double), _p(dir, C.c_double), _p(cell, C.c_int32)))
        return pos, dir, cell

    def probe_detector(self, det_index, start, dir, seg_len):
        start, dir = _f64(start).reshape(-1, 3), _f64(dir).reshape(-1, 3)
        n = len(start)
        seg_len = _f64(np.broadcast_to(seg_len, n))
        hit, b = np.zeros(n, np.int32), np.zeros(n, np.int32)
        check(self._L.smcrt_probe_detector(self._h, det_index, n, _p(start, C.c_double), _p(dir, C.c_double), _p(seg_len, C.c_double),
                                           _p(hit, C.c_int32), _p(b, C.c_int32)))
        return hit, b

    def probe_ray(self, top_index, pos, dirs):
        """-> (dist, bound, exact) of the directional step bound (smcrt_probe_ray)."""
        pos = np.ascontiguousarray(pos, np.float64).reshape(-1, 3)
        dirs = np.ascontiguousarray(dirs, np.float64).reshape(-1, 3)
        n = len(pos)
        d, b, ex = np.zeros(n), np.zeros(n), np.zeros(n, np.int32)
        check(self._L.smcrt_probe_ray(self._h, int(top_index), n, _p(pos, C.c_double), _p(dirs, C.c_double), _p(d, C.c_double),
                                      _p(b, C.c_double), _p(ex, C.c_int32)))
        return d, b, ex

    def run_sources(self, positions, nphotons_per_source, seed, id_offset=0, tally_mode=TALLY_ABSORB, survival_bias=False,
                    threshold=-1.0, chance=-1.0):
        """Batched isotropic point sources (escape-function drivers). -> (det_totals (n_src, n_det), layer (n_src,))."""
        pos = np.ascontiguousarray(positions, np.float64).reshape(-1, 3)
        tot = np.zeros((len(pos), max(getattr(self, "_n_det", 0), 1)))
        layer = np.zeros(len(pos), np.int32)
        check(self._L.smcrt_run_sources(self._h, len(pos), _p(pos, C.c_double), int(nphotons_per_source), int(seed), int(id_offset),
                                        int(tally_mode), int(survival_bias), threshold, chance, _p(tot, C.c_double), _p(layer, C.c_int32)))
        return tot, layer

    @property
    def last_fetch_bytes(self):
        return int(self._L.smcrt_last_fetch_bytes(self._h))

    def kernel_variant(self, tally_mode=TALLY_ABSORB):
        """-1 until the first large run has timed the candidates; then 0..5 (smcrt_kernel_variant)."""
        return int(self._L.smcrt_kernel_variant(self._h, int(tally_mode)))

    # ---- trackHistory ------------------------------------------------------------------------------
    def set_track_history(self, flags):
        f = _i32(flags)
        check(self._L.smcrt_set_track_history(self._h, len(f), _p(f, C.c_int32) if len(f) else None))

    def history_hits(self, max_hits=1 << 20):
        """-> (packet ids, 1-based detector indices, total hits seen) since the last reset, sorted by packet id."""
        ids, det, total = np.zeros(max_hits, np.uint64), np.zeros(max_hits, np.int32), C.c_int64(0)
        check(self._L.smcrt_history_hits(self._h, max_hits, _p(ids, C.c_uint64), _p(det, C.c_int32), C.byref(total)))
        n = min(max_hits, int(total.value))
        return ids[:n], det[:n], int(total.value)

    def history_replay(self, ids, seed, survival_bias=False, max_vertices=256):
        """-> (vertices (n, max_vertices, 4), n_vertices (n,), hit_vertex (n,))   (smcrt_history_replay)"""
        ids = np.ascontiguousarray(ids, np.uint64)
        n = len(ids)
        v, nv, hv = np.zeros((n, max_vertices, 4), np.float32), np.zeros(n, np.int32), np.zeros(n, np.int32)
        check(self._L.smcrt_history_replay(self._h, n, _p(ids, C.c_uint64), int(seed), int(survival_bias), int(max_vertices),
                                           _p(v, C.c_float), _p(nv, C.c_int32), _p(hv, C.c_int32)))
        return v, nv, hv

    def inverse_mcrt(self, top_index, find_mask, targets, max_steps, nphotons, seed, bounds=None, tally_mode=TALLY_ABSORB):
        """inverse_MCRT's search loop (smcrt_inverse_mcrt). -> (table (max_steps, 5): mus, mua, g, n, error; best row)."""
        t = _f64(targets)
        table = np.zeros((int(max_steps), 5))
        best = C.c_int(0)
        b = None if bounds is None else _f64(bounds)
        check(self._L.smcrt_inverse_mcrt(self._h, int(top_index), int(find_mask), None if b is None else _p(b, C.c_double), int(max_steps),
                                         int(nphotons), int(seed), int(tally_mode), _p(t, C.c_double), _p(table, C.c_double), C.byref(best)))
        return table, int(best.value)

    @property
    def segment_mode(self):
        """-Dpathlength deposits of this scene: 0 = deposit kernel, 1 = inline walks, -1 = not timed yet (smcrt_segment_mode)."""
        return int(self._L.smcrt_segment_mode(self._h))

    @property
    def segments_per_packet(self):
        return float(self._L.smcrt_segments_per_packet(self._h))

    def bench_red(self, pattern, span=333, n_ops=1 << 30):
        """red.global.add.f32 operations per second on this context's path-length grid (smcrt_bench_red)."""
        out = C.c_double(0.0)
        check(self._L.smcrt_bench_red(self._h, int(pattern), int(span), int(n_ops), C.byref(out)))
        return out.value

    def trace_packets(self, n, seed, id_offset=0, tally_mode=TALLY_ABSORB, survival_bias=False):
        fate, nsc, ev, sw = (np.zeros(n, np.int32) for _ in range(4))
        pos = np.zeros((n, 3))
        check(self._L.smcrt_trace_packets(self._h, n, int(seed), int(id_offset), int(tally_mode), int(survival_bias),
                                          _p(fate, C.c_int32), _p(nsc, C.c_int32), _p(pos, C.c_double), _p(ev, C.c_int32),
                                          _p(sw, C.c_int32)))
        return {"fate": fate, "nscatt": nsc, "pos": pos, "events": ev, "sweeps": sw}


def philox(seed, packet_id, event):
    out = np.zeros(4, np.uint32)
    _lib.load().smcrt_probe_philox(int(seed), int(packet_id), int(event), _p(out, C.c_uint32))
    return out


def normalise_fluence(array, grid_shape, half_extent, nphotons):
    a = np.ascontiguousarray(np.asarray(array, np.float32).reshape(-1, order="F"))
    check(_lib.load().smcrt_normalise_fluence(_p(a, C.c_float), *grid_shape, *half_extent, int(nphotons)))
    return a.reshape(grid_shape, order="F")


def write_nrrd(path, array, meta=""):
    a = np.asarray(array, np.float32)
    flat = np.ascontiguousarray(a.reshape(-1, order="F"))
    check(_lib.load().smcrt_write_nrrd_f32(str(path).encode(), _p(flat, C.c_float), a.shape[0], a.shape[1], a.shape[2],
                                           meta.encode() if meta else None))


def checkpoint_write(path, toml_filename, nphotons_run, jmean):
    """writer.f90:426-457: two header lines + the raw float32 jmean grid (x fastest)."""
    flat = np.ascontiguousarray(np.asarray(jmean, np.float32).reshape(-1, order="F"))
    check(_lib.load().smcrt_checkpoint_write(str(path).encode(), str(toml_filename).encode(), int(nphotons_run), _p(flat, C.c_float), flat.size))


def checkpoint_read(path, n_voxels=0):
    """-> (toml_filename, nphotons_run, jmean flat float32 or None)   (the load_checkpoint branch, kernelsMod.f90:52-72)"""
    name = C.create_string_buffer(1024)
    run = C.c_int64()
    jm = np.zeros(int(n_voxels), np.float32) if n_voxels else None
    check(_lib.load().smcrt_checkpoint_read(str(path).encode(), name, 1024, C.byref(run), None if jm is None else _p(jm, C.c_float),
                                            int(n_voxels)))
    return name.value.decode(), int(run.value), jm


def default_MCRT(input_file, res_dir=None, out_dir="data", n_gpus=1, tally_mode=-1, survival_bias=False, nphotons=-1):
    """`program mcpolar` -> default_MCRT (app/main.f90:24, src/kernelsMod.f90:29-83). Returns (photons/s, counters)."""
    pps = C.c_double()
    cn = Counters()
    check(_lib.load().smcrt_default_mcrt(str(input_file).encode(), None if res_dir is None else str(res_dir).encode(),
                                         str(out_dir).encode(), n_gpus, tally_mode, int(survival_bias), int(nphotons),
                                         C.byref(pps), C.byref(cn)))
    return pps.value, cn.as_dict()
