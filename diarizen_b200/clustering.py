"""Clustering stage: host-side mirror of `pyannote.audio.pipelines.clustering.AgglomerativeClustering`
(reference: pyannote-audio/pyannote/audio/pipelines/clustering.py:76-322, 325-513) with the O(N^2) / O(N^2 log N)
parts on the GPU: float64 distance matrix (dz_pdist), centroid-linkage merge loop (dz_linkage_centroid) and the
per-chunk constrained assignment (dz_assign).  Selection / bookkeeping logic (a few thousand scalars) stays in numpy.

Call convention and hyper-parameters are the reference's: attributes `threshold`, `method`, `min_cluster_size`,
`metric`; `__call__(embeddings (C,S,D), segmentations (C,T,S), num_clusters, min_clusters, max_clusters)`
-> `(hard_clusters (C,S) int8, soft_clusters (C,S,K), centroids (K,D))`.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib


# ---------------------------------------------------------------------------------------------------------
# flat clusters from a dendrogram: scipy.cluster.hierarchy.fcluster(Z, t, criterion="distance") restated
# (third-party; call sites clustering.py:418,457,477).  Numbering follows scipy's traversal so that cluster
# ids - which leak into the RTTM speaker labels - are identical.
# ---------------------------------------------------------------------------------------------------------
def fcluster_distance(Z: np.ndarray, t: float) -> np.ndarray:
    n = Z.shape[0] + 1
    left = Z[:, 0].astype(np.int64)
    right = Z[:, 1].astype(np.int64)
    # maximum merge height inside each subtree (handles the inversions of centroid linkage)
    md = np.empty(n - 1, dtype=np.float64)
    for i in range(n - 1):
        m = Z[i, 2]
        if left[i] >= n:
            m = max(m, md[left[i] - n])
        if right[i] >= n:
            m = max(m, md[right[i] - n])
        md[i] = m
    T = np.zeros(n, dtype=np.int32)
    visited = np.zeros(2 * n, dtype=bool)
    stack = [2 * n - 2]
    n_cluster = 0
    leader = -1
    while stack:
        root = stack[-1] - n
        lc, rc = left[root], right[root]
        if leader == -1 and md[root] <= t:
            leader = root
            n_cluster += 1
        if lc >= n and not visited[lc]:
            visited[lc] = True
            stack.append(lc)
            continue
        if rc >= n and not visited[rc]:
            visited[rc] = True
            stack.append(rc)
            continue
        if lc < n:
            if leader == -1:
                n_cluster += 1
            T[lc] = n_cluster
        if rc < n:
            if leader == -1:
                n_cluster += 1
            T[rc] = n_cluster
        if leader == root:
            leader = -1
        stack.pop()
    return T


def _cosine_cdist(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """scipy.spatial.distance.cdist(metric="cosine") in float64: 1 - u.v / (|u| |v|)."""
    A = A.astype(np.float64)
    B = B.astype(np.float64)
    na = np.sqrt(np.einsum("ij,ij->i", A, A))
    nb = np.sqrt(np.einsum("ij,ij->i", B, B))
    with np.errstate(divide="ignore", invalid="ignore"):
        return 1.0 - (A @ B.T) / (na[:, None] * nb[None, :])


def device_linkage_centroid(unit_embeddings: np.ndarray, device=None) -> np.ndarray:
    """scipy linkage(method="centroid", metric="euclidean") on the GPU -> Z (N-1, 4) float64."""
    L = _lib.lib()
    dev = torch.device(device if device is not None else "cuda")
    x = torch.as_tensor(np.ascontiguousarray(unit_embeddings, dtype=np.float32), device=dev)
    N, D = x.shape
    dist = torch.empty((N, N), dtype=torch.float64, device=dev)
    Z = torch.empty((N - 1, 4), dtype=torch.float64, device=dev)
    ws = torch.empty(int(L.dz_linkage_workspace_bytes(N)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(L.dz_pdist(C.c_void_p(x.data_ptr()), N, D, C.c_void_p(dist.data_ptr()), st))
        _lib.check(L.dz_linkage_centroid(C.c_void_p(dist.data_ptr()), N, C.c_void_p(Z.data_ptr()), C.c_void_p(ws.data_ptr()), st))
    return Z.cpu().numpy()


def device_assign(soft: np.ndarray, device=None) -> np.ndarray:
    """Per-chunk constrained assignment (clustering.py:159-173) on the GPU."""
    L = _lib.lib()
    dev = torch.device(device if device is not None else "cuda")
    Cn, S, K = soft.shape
    s = torch.as_tensor(np.ascontiguousarray(soft, dtype=np.float64), device=dev)
    hard = torch.empty((Cn, S), dtype=torch.int8, device=dev)
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(L.dz_assign(C.c_void_p(s.data_ptr()), Cn, S, K, C.c_void_p(hard.data_ptr()), st))
    return hard.cpu().numpy()


class AgglomerativeClustering:
    """reference: clustering.py:325-513 (only method="centroid" with metric="cosine" is what DiariZen configures:
    diarizen/pipelines/inference.py:64-71)."""

    def __init__(self, metric: str = "cosine", max_num_embeddings: float = np.inf, constrained_assignment: bool = True,
                 device=None):
        self.metric = metric
        self.max_num_embeddings = max_num_embeddings
        self.constrained_assignment = constrained_assignment
        self.threshold = 0.7
        self.method = "centroid"
        self.min_cluster_size = 30
        self.device = device

    # -- clustering.py:88-109
    @staticmethod
    def set_num_clusters(num_embeddings, num_clusters=None, min_clusters=None, max_clusters=None):
        min_clusters = num_clusters or min_clusters or 1
        min_clusters = max(1, min(num_embeddings, min_clusters))
        max_clusters = num_clusters or max_clusters or num_embeddings
        max_clusters = max(1, min(num_embeddings, max_clusters))
        if min_clusters > max_clusters:
            raise ValueError(
                f"min_clusters must be smaller than (or equal to) max_clusters "
                f"(here: min_clusters={min_clusters:g} and max_clusters={max_clusters:g}).")
        if min_clusters == max_clusters:
            num_clusters = min_clusters
        return num_clusters, min_clusters, max_clusters

    # -- clustering.py:111-157; `active_frames` / `single_frames` are (C,S) counts (device kernel dz_embedding_masks)
    def filter_embeddings(self, embeddings: np.ndarray, active_frames: np.ndarray, single_frames: np.ndarray,
                          num_frames: int, min_frames_ratio: float = 0.1):
        active = active_frames > 0
        valid = ~np.any(np.isnan(embeddings), axis=2)
        min_frames = round(min_frames_ratio * num_frames)
        ci, si = np.where(active * valid * (single_frames >= min_frames))
        if len(ci) < 2:
            ci, si = np.where(active * valid * (single_frames >= 0))
        return embeddings[ci, si], ci, si

    # -- clustering.py:363-513
    def cluster(self, embeddings: np.ndarray, min_clusters: int, max_clusters: int, num_clusters: Optional[int] = None):
        if self.method != "centroid" or self.metric != "cosine":
            raise ValueError("only method='centroid' with metric='cosine' is implemented (the DiariZen configuration)")
        n = embeddings.shape[0]
        mcs = min(self.min_cluster_size, max(1, round(0.1 * n)))
        if n == 1:
            return np.zeros((1,), dtype=np.uint8)
        emb = embeddings.copy()
        with np.errstate(divide="ignore", invalid="ignore"):
            emb /= np.linalg.norm(emb, axis=-1, keepdims=True)
        import os as _os, time as _time
        _t0 = _time.perf_counter()
        Z = device_linkage_centroid(emb, self.device)
        _t1 = _time.perf_counter()
        clusters = fcluster_distance(Z, self.threshold) - 1
        if _os.environ.get("DZ_TIMING") is not None:
            print(f"[dz timing] linkage N={n}: device pdist+linkage {(_t1 - _t0) * 1e3:.1f} ms, fcluster {(_time.perf_counter() - _t1) * 1e3:.1f} ms")
        uniq, counts = np.unique(clusters, return_counts=True)
        large = uniq[counts >= mcs]
        nlarge = len(large)
        if nlarge < min_clusters:
            num_clusters = min_clusters
        elif nlarge > max_clusters:
            num_clusters = max_clusters
        if num_clusters is not None and nlarge != num_clusters:
            _Z = np.copy(Z)
            _Z[:, 2] = np.arange(n - 1)
            best_it, best_n = n - 1, 1
            for it in np.argsort(np.abs(Z[:, 2] - self.threshold)):
                if _Z[it, 3] < mcs:
                    continue
                clusters = fcluster_distance(_Z, it) - 1
                uniq, counts = np.unique(clusters, return_counts=True)
                large = uniq[counts >= mcs]
                nlarge = len(large)
                if abs(nlarge - num_clusters) < abs(best_n - num_clusters):
                    best_it, best_n = it, nlarge
                if nlarge == num_clusters:
                    break
            if best_n != num_clusters:
                clusters = fcluster_distance(_Z, best_it) - 1
                uniq, counts = np.unique(clusters, return_counts=True)
                large = uniq[counts >= mcs]
                nlarge = len(large)
                print(f"Found only {nlarge} clusters. Using a smaller value than {mcs} for `min_cluster_size` might help.")
        if nlarge == 0:
            clusters[:] = 0
            return clusters
        small = uniq[counts < mcs]
        if len(small) == 0:
            return clusters
        lc = np.vstack([np.mean(emb[clusters == k], axis=0) for k in large])
        scn = np.vstack([np.mean(emb[clusters == k], axis=0) for k in small])
        d = _cosine_cdist(lc, scn)
        for sk, lk in enumerate(np.argmin(d, axis=0)):
            clusters[clusters == small[sk]] = large[lk]
        _, clusters = np.unique(clusters, return_inverse=True)
        return clusters

    # -- clustering.py:175-245
    def assign_embeddings(self, embeddings, ci, si, train_clusters):
        K = int(np.max(train_clusters)) + 1
        Cn, S, D = embeddings.shape
        train = embeddings[ci, si]
        centroids = np.vstack([np.mean(train[train_clusters == k], axis=0) for k in range(K)])
        soft = 2 - _cosine_cdist(embeddings.reshape(Cn * S, D), centroids).reshape(Cn, S, K)
        if self.constrained_assignment:
            sc = np.nan_to_num(soft, nan=np.nanmin(soft))
            hard = device_assign(sc, self.device)
        else:
            hard = np.argmax(soft, axis=2).astype(np.int8)
        return hard, soft, centroids

    # -- clustering.py:247-322
    def __call__(self, embeddings: np.ndarray, segmentations=None, num_clusters=None, min_clusters=None,
                 max_clusters=None, frame_stats: Optional[Tuple[np.ndarray, np.ndarray]] = None, **kwargs):
        if frame_stats is None:
            seg = np.asarray(getattr(segmentations, "data", segmentations))
            active_frames = np.sum(seg, axis=1)
            single = (np.sum(seg, axis=2, keepdims=True) == 1)
            single_frames = np.sum(seg * single, axis=1)
            T = seg.shape[1]
        else:
            active_frames, single_frames, T = frame_stats
        train, ci, si = self.filter_embeddings(embeddings, active_frames, single_frames, T)
        n = train.shape[0]
        num_clusters, min_c, max_c = self.set_num_clusters(n, num_clusters, min_clusters, max_clusters)
        if max_c < 2:
            Cn, S, _ = embeddings.shape
            return (np.zeros((Cn, S), dtype=np.int8), np.ones((Cn, S, 1)), np.mean(train, axis=0, keepdims=True))
        tc = self.cluster(train, min_c, max_c, num_clusters)
        return self.assign_embeddings(embeddings, ci, si, tc)
