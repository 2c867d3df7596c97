"""CPU check of the dz_dendrogram_cut algorithm: the kernel body (diarizen_b200/csrc/dendro_cut.cuh) compiled for the host
with one "thread" (tests/host_shim/dendro_cut_host.cpp, a test harness - not part of the product library) against
scipy.fcluster, against the pinned glue oracle, and - through the product's own host logic - against the goldens the
reference's clustering classes produced (tests/golden/glue_clustering.npz).  The CUDA instantiation of the same body is
checked on the GPU (tests/test_glue_golden_gpu.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from scipy.cluster.hierarchy import fcluster, linkage

from oracle import pipeline_oracle as po

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("shim") / "dendro_cut_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "host_shim", "dendro_cut_host.cpp")], check=True)
    L = C.CDLL(so)
    L.dendro_cut_host.restype = C.c_int
    L.dendro_cut_host.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]

    def cut(Z, threshold, mcs=1, lo=1, hi=None, num=None, force=-1):
        Z = np.ascontiguousarray(Z, dtype=np.float64)
        n = Z.shape[0] + 1
        labels = np.empty(n, dtype=np.int32)
        info = np.zeros(8, dtype=np.int32)
        assert L.dendro_cut_host(Z.ctypes.data, n, float(threshold), mcs, lo, hi if hi is not None else n, num or 0, force,
                                 labels.ctypes.data, info.ctypes.data) == 0
        return labels, {"num_large": int(info[0]), "iteration": int(info[1]), "found_only": bool(info[2]), "num_flat": int(info[3]),
                        "num_large_at_threshold": int(info[4]), "target": int(info[5])}
    return cut


def _points(n, k, noise, seed, dup=False):
    r = np.random.default_rng(seed)
    x = (r.standard_normal((k, 64))[r.integers(0, k, n)] + noise * r.standard_normal((n, 64))).astype(np.float32)
    if dup and n > 8:
        x[1::5] = x[0::5][: len(x[1::5])]
    return x / np.linalg.norm(x, axis=1, keepdims=True)


@pytest.mark.parametrize("n,k,noise,dup", [(2, 1, 0.1, False), (3, 2, 0.3, False), (40, 3, 0.4, False), (300, 6, 0.5, False),
                                           (300, 6, 0.5, True), (1200, 9, 0.7, False)])
def test_numbering_matches_scipy(shim, n, k, noise, dup):
    Z = linkage(_points(n, k, noise, n + k, dup), method="centroid", metric="euclidean")
    for t in [0.0, 0.05, 0.3, 0.6, 0.7, 0.9, 1.2, 2.5] + list(np.random.default_rng(0).choice(Z[:, 2], size=min(8, n - 1))):
        labels, info = shim(Z, t)
        ref = fcluster(Z, t, criterion="distance") - 1
        assert np.array_equal(labels, ref), (n, t)
        assert info["num_flat"] == ref.max() + 1 and info["iteration"] == -1
    _Z = Z.copy()
    _Z[:, 2] = np.arange(n - 1)
    for it in sorted(set([0, (n - 1) // 3, (n - 1) // 2, n - 2])):
        labels, _ = shim(Z, 0.7, force=it)
        assert np.array_equal(labels, fcluster(_Z, it, criterion="distance") - 1), (n, it)


def test_cluster_selection_matches_pinned_oracle(shim):
    """cut + absorb_small_clusters == oracle ahc_cluster (itself equal to the reference's cluster() on the goldens) over random
    settings that reach the re-cut branches (too few / too many / exact num_clusters) and the small-cluster absorption."""
    from diarizen_b200.clustering import absorb_small_clusters
    r = np.random.default_rng(1)
    hits = {"recut": 0, "found_only": 0, "small": 0}
    for trial in range(120):
        n = int(r.integers(5, 260))
        x = _points(n, int(r.integers(1, 9)), float(r.uniform(0.2, 0.9)), 100 + trial, dup=trial % 7 == 0)
        thr = float(r.uniform(0.3, 1.2))
        mcs = int(r.integers(1, 25))
        lo = int(r.integers(1, 6))
        hi = int(r.integers(lo, 12))
        num = int(r.integers(1, 8)) if trial % 3 == 0 else None
        num_c, lo_c, hi_c = po.set_num_clusters(n, num, lo, hi)
        ref = po.ahc_cluster(x.copy(), thr, mcs, lo_c, hi_c, num_c)
        Z = linkage(x, method="centroid", metric="euclidean")
        eff = min(mcs, max(1, round(0.1 * n)))
        labels, info = shim(Z, thr, eff, lo_c, hi_c, num_c)
        got = absorb_small_clusters(x, labels.astype(np.int64), eff)
        assert np.array_equal(got, ref), (trial, n, thr, mcs, lo_c, hi_c, num_c, info)
        hits["recut"] += info["iteration"] >= 0
        hits["found_only"] += info["found_only"]
        hits["small"] += not np.array_equal(got, labels)
    assert min(hits.values()) > 3, hits


def _params(z, key):
    return eval(str(z[key]), {"__builtins__": {}}, {"dict": dict})


@pytest.mark.parametrize("name", ["plain", "small", "recut_down", "recut_up", "recut_exact", "tiny", "many", "loose"])
def test_product_host_logic_on_reference_goldens(shim, name, monkeypatch):
    """diarizen_b200.clustering.AgglomerativeClustering.__call__ with its three device steps replaced by CPU stand-ins
    (scipy linkage + the host build of the cut kernel body + scipy's Hungarian) == the reference class's output."""
    from scipy.optimize import linear_sum_assignment
    from diarizen_b200 import clustering as cl

    class HostDendrogram:
        def __init__(self, unit, device=None):
            self.Z_ = linkage(unit.astype(np.float32), method="centroid", metric="euclidean")

        def cut(self, threshold, mcs=1, lo=1, hi=None, num=None, force_iteration=-1):
            return shim(self.Z_, threshold, mcs, lo, hi, num, force_iteration)

    def host_assign(soft, device=None):
        hard = -2 * np.ones(soft.shape[:2], dtype=np.int8)
        for c, cost in enumerate(soft):
            for s, k in zip(*linear_sum_assignment(cost, maximize=True)):
                hard[c, s] = k
        return hard

    monkeypatch.setattr(cl, "DeviceDendrogram", HostDendrogram)
    monkeypatch.setattr(cl, "device_assign", host_assign)
    z = np.load(os.path.join(G, "glue_clustering.npz"))
    prm = _params(z, f"{name}__params")
    a = cl.AgglomerativeClustering()
    a.threshold, a.min_cluster_size = prm["threshold"], prm["mcs"]
    hard, soft, cent = a(z[f"{name}__embeddings"], z[f"{name}__segmentations"].astype(np.float32), num_clusters=prm.get("num"),
                         min_clusters=prm["min"], max_clusters=prm["max"])
    assert np.array_equal(np.asarray(hard).astype(np.int16), z[f"{name}__hard"])
    assert np.allclose(soft, z[f"{name}__soft"], atol=1e-9, equal_nan=True)
    assert np.allclose(cent, z[f"{name}__centroids"], atol=1e-9)
