"""CPU check of the per-chunk assignment routine of dz_assign (lsap_small.cuh, compiled for the host by a test harness) against
scipy.optimize.linear_sum_assignment(maximize=True), with the tie patterns the pipeline produces: identical rows (inactive local
speakers share one embedding), constant rows (NaN scores replaced by the global minimum), fewer clusters than speakers."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lsap(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "lsap_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "host_shim", "lsap_host.cpp")], check=True)
    L = C.CDLL(so)
    L.lsap_host.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]

    def run(soft):
        soft = np.ascontiguousarray(soft, dtype=np.float64)
        Cn, S, K = soft.shape
        hard = np.empty((Cn, S), dtype=np.int8)
        L.lsap_host(soft.ctypes.data, Cn, S, K, hard.ctypes.data)
        return hard
    return run


def _scipy(soft):
    hard = -2 * np.ones(soft.shape[:2], dtype=np.int8)
    for c, cost in enumerate(soft):
        for s, k in zip(*linear_sum_assignment(cost, maximize=True)):
            hard[c, s] = k
    return hard


@pytest.mark.parametrize("K", [1, 2, 3, 4, 5, 7, 12, 31, 40, 127])
@pytest.mark.parametrize("S", [1, 2, 3, 4])
def test_matches_scipy_including_ties(lsap, S, K):
    r = np.random.default_rng(100 * S + K)
    n = 400
    soft = 2 - 2 * r.random((n, S, K))
    # tie patterns: identical rows, constant rows, quantised scores, a whole constant matrix
    for c in range(0, n, 4):
        if S > 1:
            soft[c, r.integers(1, S)] = soft[c, 0]
    for c in range(1, n, 4):
        rows = r.choice(S, size=int(r.integers(1, S + 1)), replace=False)
        soft[c, rows] = soft[c].min()
    soft[2::8] = np.round(soft[2::8] * 3) / 3
    soft[6::16] = 0.5
    got, ref = lsap(soft), _scipy(soft)
    bad = np.argwhere((got != ref).any(1)).ravel()
    assert bad.size == 0, (S, K, bad[:5], got[bad[:3]], ref[bad[:3]])
