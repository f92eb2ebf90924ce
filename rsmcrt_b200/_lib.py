"""ctypes binding of libsmcrt_gpu.so (include/smcrt.h + include/smcrt_host.h).

The shared object is the product; this module only declares prototypes.  It fails loudly when the library
is missing: there is no CPU / PyTorch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from . import build as _build

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_uint32_p = C.POINTER(C.c_uint32)

NODE_PARAMS = 8
SOURCE_PARAMS = 24
DET_PARAMS = 20


class Counters(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("nscatt", "sdf_evals", "bounces", "launched", "emit_retries", "lost", "sweeps", "det_hits", "voxel_crossings", "deposit_atomics")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/smcrt.h and include/smcrt_host.h declare: name -> (restype, argtypes)
PROTOTYPES = {
    # smcrt.h
    "smcrt_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, c_int32_p]),
    "smcrt_destroy": (None, [C.c_void_p]),
    "smcrt_last_error": (C.c_char_p, []),
    "smcrt_version": (C.c_char_p, []),
    "smcrt_set_grid": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]),
    "smcrt_set_scene": (C.c_int, [C.c_void_p, C.c_int, c_int32_p, c_int32_p, c_int32_p, c_double_p, c_double_p, C.c_int,
                                  c_int32_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "smcrt_set_optprops": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]),
    "smcrt_set_source": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p]),
    "smcrt_set_detectors": (C.c_int, [C.c_void_p, C.c_int, c_int32_p, c_double_p, c_int32_p]),
    "smcrt_det_bins_total": (C.c_int64, [C.c_void_p]),
    "smcrt_set_tolerances": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int64]),
    "smcrt_run": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_double]),
    "smcrt_run_async": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_double]),
    "smcrt_wait": (C.c_int, [C.c_void_p]),
    "smcrt_last_run_ms": (C.c_double, [C.c_void_p]),
    "smcrt_set_track_history": (C.c_int, [C.c_void_p, C.c_int, c_int32_p]),
    "smcrt_history_hits": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_uint64), c_int32_p, C.POINTER(C.c_int64)]),
    "smcrt_history_replay": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_uint64), C.c_uint64, C.c_int, C.c_int, c_float_p, c_int32_p, c_int32_p]),
    "smcrt_inverse_mcrt": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, C.c_int, C.c_int64, C.c_uint64, C.c_int, c_double_p, c_double_p,
                                     C.POINTER(C.c_int)]),
    "smcrt_segment_mode": (C.c_int, [C.c_void_p]),
    "smcrt_segments_per_packet": (C.c_double, [C.c_void_p]),
    "smcrt_launch_count": (C.c_int64, [C.c_void_p]),
    "smcrt_fetch": (C.c_int, [C.c_void_p, c_float_p, c_float_p, c_float_p, c_double_p, C.POINTER(Counters), C.c_int]),
    "smcrt_reset_tallies": (C.c_int, [C.c_void_p]),
    "smcrt_pin_host": (C.c_int, [C.c_void_p, C.c_uint64]),
    "smcrt_unpin_host": (C.c_int, [C.c_void_p]),
    "smcrt_comm_unique_id": (C.c_int, [C.c_char_p]),
    "smcrt_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "smcrt_comm_reduce": (C.c_int, [C.c_void_p, C.c_int]),
    "smcrt_probe_sdf": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_double_p, c_double_p, c_double_p]),
    "smcrt_probe_ray": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p, c_int32_p]),
    "smcrt_probe_fresnel": (C.c_int, [C.c_void_p, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                      c_double_p, c_double_p, c_int32_p]),
    "smcrt_probe_scatter": (C.c_int, [C.c_void_p, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p]),
    "smcrt_probe_emit": (C.c_int, [C.c_void_p, C.c_int64, c_double_p, c_double_p, c_double_p, c_int32_p]),
    "smcrt_probe_detector": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_double_p, c_double_p, c_double_p, c_int32_p, c_int32_p]),
    "smcrt_trace_packets": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, C.c_int64, C.c_int, C.c_int, c_int32_p, c_int32_p,
                                      c_double_p, c_int32_p, c_int32_p]),
    "smcrt_last_fetch_bytes": (C.c_uint64, [C.c_void_p]),
    "smcrt_kernel_variant": (C.c_int, [C.c_void_p, C.c_int]),
    "smcrt_run_sources": (C.c_int, [C.c_void_p, C.c_int64, c_double_p, C.c_int64, C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_double,
                                    C.c_double, c_double_p, c_int32_p]),
    "smcrt_bench_red": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_double)]),
    "smcrt_probe_philox": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint32, c_uint32_p]),
    # smcrt_host.h
    "smcrt_config_load": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "smcrt_config_loads": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "smcrt_config_free": (None, [C.c_void_p]),
    "smcrt_config_grid": (C.c_int, [C.c_void_p, c_int32_p, c_double_p]),
    "smcrt_config_nphotons": (C.c_int64, [C.c_void_p]),
    "smcrt_config_iseed": (C.c_int64, [C.c_void_p]),
    "smcrt_config_geom_name": (C.c_char_p, [C.c_void_p]),
    "smcrt_config_source_name": (C.c_char_p, [C.c_void_p]),
    "smcrt_config_render_source": (C.c_int, [C.c_void_p]),
    "smcrt_config_source": (C.c_int, [C.c_void_p, c_int32_p, c_int32_p, c_double_p]),
    "smcrt_config_n_detectors": (C.c_int, [C.c_void_p]),
    "smcrt_config_detectors": (C.c_int, [C.c_void_p, c_int32_p, c_double_p, c_int32_p]),
    "smcrt_config_detector_id": (C.c_char_p, [C.c_void_p, C.c_int]),
    "smcrt_config_scene_sizes": (C.c_int, [C.c_void_p, c_int32_p, c_int32_p]),
    "smcrt_config_scene": (C.c_int, [C.c_void_p, c_int32_p, c_int32_p, c_int32_p, c_double_p, c_double_p, c_int32_p,
                                     c_double_p, c_double_p, c_double_p, c_double_p]),
    "smcrt_config_apply": (C.c_int, [C.c_void_p, C.c_void_p]),
    "smcrt_inverse_evaluate": (C.c_int, [C.c_int, c_double_p, c_double_p, C.c_int64, c_double_p]),
    "smcrt_escape_cell_centre": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                           c_double_p, c_double_p, c_double_p, c_double_p]),
    "smcrt_normalise_fluence": (C.c_int, [c_float_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int64]),
    "smcrt_write_nrrd_f32": (C.c_int, [C.c_char_p, c_float_p, C.c_int, C.c_int, C.c_int, C.c_char_p]),
    "smcrt_write_detectors": (C.c_int, [C.c_void_p, c_double_p, C.c_char_p]),
    "smcrt_checkpoint_write": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int64, c_float_p, C.c_int64]),
    "smcrt_checkpoint_read": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int64), c_float_p, C.c_int64]),
    "smcrt_config_metadata": (C.c_char_p, [C.c_void_p]),
    "smcrt_history_write": (C.c_int, [C.c_char_p, C.c_int64, C.c_int, c_float_p, c_int32_p]),
    "smcrt_config_detector_track": (C.c_int, [C.c_void_p, C.c_int]),
    "smcrt_config_history_filename": (C.c_char_p, [C.c_void_p]),
    "smcrt_default_mcrt": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int64, c_double_p,
                                     C.POINTER(Counters)]),
}

_lib = None


def lib_path() -> Path:
    return _build.LIB_PATH


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load libsmcrt_gpu.so, declaring every prototype.  Raises if the library cannot be built/loaded."""
    global _lib
    if _lib is not None:
        return _lib
    import os
    path = Path(os.environ["SMCRT_LIB"]) if os.environ.get("SMCRT_LIB") else _build.LIB_PATH  # override: kernel-variant experiments
    if not path.exists():
        if not build_if_missing:
            raise RuntimeError(f"{path} is missing; run `python -m rsmcrt_b200.build` (no CPU fallback exists)")
        _build.build()
    lib = C.CDLL(str(path), mode=C.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class SmcrtError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        raise SmcrtError(load().smcrt_last_error().decode(errors="replace"))
