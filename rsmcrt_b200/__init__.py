"""rsmcrt_b200 — B200 (sm_100a) photon-packet transport engine behind signedMCRT/RSMCRT's run_MCRT seam.

The product is rsmcrt_b200/lib/libsmcrt_gpu.so (CUDA engine + C++ host mirror, C ABI in include/*.h);
this package is the ctypes face used by tests/, bench.py and __graft_entry__.py.
"""
from .api import (Config, Engine, Scene, checkpoint_read, checkpoint_write, default_MCRT, normalise_fluence, philox, write_nrrd,  # noqa: F401
                  TALLY_ABSORB, TALLY_EMISSION, TALLY_PATHLENGTH)
from ._lib import SmcrtError, load, lib_path  # noqa: F401

__version__ = "0.1.0"
