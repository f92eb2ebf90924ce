"""Every kernel variant against the oracle, directly.

The engine picks one of six kernels per scene by timing: plain at 2/3/4 resident CTAs, compacted, queue-scheduled at
2/3.  The other GPU tests reach only the variant an untuned context starts with (and smcrt_trace_packets excludes the LEAN builds), so
here each scheduler is PINNED (SMCRT_VARIANT_FORCE, read once per process: hence one subprocess per variant) and the oracle-parity
subset -- validation1 tallies, scat_test, skin, sphere path-length, test_dects, validation1 path-length -- runs through plain
smcrt_run against oracle results computed once on the same Philox streams.  Variant 4 is `trace_queued<...,SIMPLE,LEAN>`, the kernel
that produces the bench headline.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import RES, ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle_results(tmp_path_factory, oracle):
    sys.path.insert(0, str(ROOT / "tests"))
    from variant_parity_worker import CASES
    out = {}
    for name, (deck, n, seed, mode) in CASES.items():
        o = oracle.OracleScene.from_toml(RES / deck).run(n, seed, tally_mode=mode)
        out[name + ".absorb"] = o["absorb"]
        if mode & 2:
            out[name + ".jmean"] = o["jmean"]
        if mode & 4:
            out[name + ".emission"] = o["emission"]
        out[name + ".det_bins"] = o["det_bins"]
        out[name + ".nscatt"] = np.float64(o["counters"]["nscatt"])
        out[name + ".lost"] = np.float64(o["counters"]["lost"])
    path = tmp_path_factory.mktemp("oracle") / "oracle.npz"
    np.savez(path, **out)
    return path


@pytest.mark.parametrize("variant", [0, 3, 4, 5])
def test_variant_matches_oracle(oracle_results, variant):
    env = dict(os.environ, SMCRT_VARIANT_FORCE=str(variant))
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "variant_parity_worker.py"), str(oracle_results)], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.stdout.strip(), r.stderr[-2000:]
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["fails"] == [] and r.returncode == 0, (info["fails"], r.stderr[-1000:])
