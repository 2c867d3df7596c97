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


def _subtree_max(Z: np.ndarray) -> np.ndarray:
    n = Z.shape[0] + 1
    left = Z[:, 0].astype(np.int64)
    right = Z[:, 1].astype(np.int64)
    md = np.empty(n - 1, dtype=np.float64)
    for i in range(n - 1):
        m = Z[i, 2]
        if left[i] >= n:
            m = max(m, md[left[i] - n])
        if right[i] >= n:
            m = max(m, md[right[i] - n])
        md[i] = m
    return md


def fcluster_maxclust(Z: np.ndarray, max_nc: int) -> np.ndarray:
    """scipy.cluster.hierarchy.fcluster(Z, t, criterion="maxclust") restated (third-party; call site clustering.py:652 with
    ahc_criterion="maxclust").  scipy bisects over merge INDICES using the subtree-max height of merge i as the trial
    threshold - with centroid linkage those heights are not sorted, and the bisection is reproduced as is."""
    n = Z.shape[0] + 1
    left = Z[:, 0].astype(np.int64)
    right = Z[:, 1].astype(np.int64)
    md = _subtree_max(Z)

    def exceeds(thresh):
        nc = 0
        visited = np.zeros(2 * n, dtype=bool)
        stack = [2 * n - 2]
        while stack:
            root = stack[-1] - n
            lc, rc = left[root], right[root]
            if md[root] <= thresh:
                nc += 1
                if nc > max_nc:
                    return True
                stack.pop()
                visited[lc] = visited[rc] = True
                continue
            if not visited[lc]:
                visited[lc] = True
                if lc >= n:
                    stack.append(lc)
                    continue
                nc += 1
                if nc > max_nc:
                    return True
            if not visited[rc]:
                visited[rc] = True
                if rc >= n:
                    stack.append(rc)
                    continue
                nc += 1
                if nc > max_nc:
                    return True
            stack.pop()
        return False

    if max_nc >= n:                      # the installed scipy (1.18) answers singletons outright
        return np.arange(1, n + 1, dtype=np.int32)
    lo, hi = -1, n - 1                   # the bisection may end on merge 0 (n - 1 flat clusters)
    while hi - lo > 1:
        i = (lo + hi) >> 1
        if exceeds(md[i]):
            lo = i
        else:
            hi = i
    return fcluster_distance(Z, md[hi] if hi < n - 1 else md[n - 2])


def _cosine_cdist(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """scipy.spatial.distance.cdist(metric="cosine") in float64: 1 - u.v / (|u| |v|)."""
    A = A.astype(np.float64)
    B = B.astype(np.float64)
    na = np.sqrt(np.einsum("ij,ij->i", A, A))
    nb = np.sqrt(np.einsum("ij,ij->i", B, B))
    with np.errstate(divide="ignore", invalid="ignore"):
        return 1.0 - (A @ B.T) / (na[:, None] * nb[None, :])


class DeviceDendrogram:
    """Centroid-linkage dendrogram of unit-norm embeddings, built and kept on the GPU: float64 distance matrix (dz_pdist),
    merge loop (dz_linkage_centroid, bit-identical to scipy.cluster.hierarchy.linkage(method="centroid")), and the
    flat-cluster selection (dz_dendrogram_cut).  Only the N labels and eight counters come back to the host."""

    def __init__(self, unit_embeddings: np.ndarray, device=None, variant: int = 0):
        L = _lib.lib()
        self.device = torch.device(device if device is not None else "cuda")
        x = torch.as_tensor(np.ascontiguousarray(unit_embeddings, dtype=np.float32), device=self.device)
        self.n, D = x.shape
        N = self.n
        dist = torch.empty((N, N), dtype=torch.float64, device=self.device)
        self._Z = torch.empty((N - 1, 4), dtype=torch.float64, device=self.device)
        ws = torch.empty(int(L.dz_linkage_workspace_bytes(N)), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(L.dz_pdist(C.c_void_p(x.data_ptr()), N, D, C.c_void_p(dist.data_ptr()), st))
            _lib.check(L.dz_linkage_centroid_variant(C.c_void_p(dist.data_ptr()), N, C.c_void_p(self._Z.data_ptr()), C.c_void_p(ws.data_ptr()), st,
                                                     int(variant)))
        self.row_rescans = int(ws[24 * N:24 * N + 8].view(torch.int64).item()) if variant != 1 else None   # lazy-kernel counter

    def Z(self) -> np.ndarray:
        """scipy-layout linkage matrix (N-1, 4) float64 on the host."""
        return self._Z.cpu().numpy()

    def cut(self, threshold: float, min_cluster_size: int = 1, min_clusters: int = 1, max_clusters: Optional[int] = None,
            num_clusters: Optional[int] = None, force_iteration: int = -1):
        """-> (labels (N,) int32 in scipy's fcluster numbering minus one, info dict)."""
        L = _lib.lib()
        N = self.n
        labels = torch.empty(N, dtype=torch.int32, device=self.device)
        info = torch.zeros(8, dtype=torch.int32, device=self.device)
        ws = torch.empty(int(L.dz_dendrogram_cut_workspace_bytes(N)), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(L.dz_dendrogram_cut(C.c_void_p(self._Z.data_ptr()), N, float(threshold), int(min_cluster_size), int(min_clusters),
                                           int(max_clusters if max_clusters is not None else N), int(num_clusters or 0), int(force_iteration),
                                           C.c_void_p(labels.data_ptr()), C.c_void_p(info.data_ptr()), C.c_void_p(ws.data_ptr()), st))
        i = info.cpu().numpy()
        return labels.cpu().numpy(), {"num_large": int(i[0]), "iteration": int(i[1]), "found_only": bool(i[2]), "num_flat": int(i[3]),
                                      "num_large_at_threshold": int(i[4]), "target": int(i[5])}


def device_linkage_centroid(unit_embeddings: np.ndarray, device=None, variant: int = 0) -> np.ndarray:
    """scipy linkage(method="centroid", metric="euclidean") on the GPU -> Z (N-1, 4) float64."""
    return DeviceDendrogram(unit_embeddings, device, variant).Z()


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


def frame_statistics(segmentations) -> Tuple[np.ndarray, np.ndarray, int]:
    """(C,T,S) {0,1} -> per (chunk, speaker): active frames, frames where it is the only active speaker; T."""
    seg = np.asarray(getattr(segmentations, "data", segmentations))
    alone = seg * (seg.sum(axis=2, keepdims=True) == 1)
    return seg.sum(axis=1), alone.sum(axis=1), seg.shape[1]


def absorb_small_clusters(unit_embeddings: np.ndarray, labels: np.ndarray, min_size: int) -> np.ndarray:
    """Every flat cluster with fewer than `min_size` members joins the large cluster whose centroid is nearest in cosine
    distance; the survivors are renumbered 0..K-1 in increasing order of their old number (clustering.py:494-513).
    With no small cluster the labels are returned untouched (their scipy numbering then reaches the RTTM)."""
    counts = np.bincount(labels)
    ids = np.flatnonzero(counts)
    big, small = ids[counts[ids] >= min_size], ids[counts[ids] < min_size]
    if big.size == 0:
        return np.zeros_like(labels)
    if small.size == 0:
        return labels
    mean_of = lambda k: np.mean(unit_embeddings[labels == k], axis=0)
    nearest = np.argmin(_cosine_cdist(np.stack([mean_of(k) for k in big]), np.stack([mean_of(k) for k in small])), axis=0)
    remap = np.full(counts.size, -1, dtype=np.int64)
    remap[big] = np.arange(big.size)            # big is sorted: rank == position
    remap[small] = remap[big[nearest]]
    return remap[labels]


class AgglomerativeClustering:
    """reference: clustering.py:325-513 (only method="centroid" with metric="cosine" is what DiariZen configures:
    diarizen/pipelines/inference.py:64-71).  Same call convention and hyper-parameter attributes; the dendrogram, the
    choice of the cut and the assignment run on the GPU."""

    def __init__(self, metric: str = "cosine", max_num_embeddings: float = np.inf, constrained_assignment: bool = True,
                 device=None):
        self.metric = metric
        self.max_num_embeddings = max_num_embeddings
        self.constrained_assignment = constrained_assignment
        self.threshold = 0.7
        self.method = "centroid"
        self.min_cluster_size = 30
        self.device = device

    @staticmethod
    def set_num_clusters(num_embeddings, num_clusters=None, min_clusters=None, max_clusters=None):
        """Clamp the requested cluster-count range to [1, num_embeddings]; a fixed `num_clusters` collapses the range."""
        lo = min(num_embeddings, num_clusters or min_clusters or 1)
        hi = min(num_embeddings, num_clusters or max_clusters or num_embeddings)
        lo, hi = max(1, lo), max(1, hi)
        if lo > hi:
            raise ValueError(f"min_clusters must be smaller than (or equal to) max_clusters (here: min_clusters={lo:g} and max_clusters={hi:g}).")
        return (lo if lo == hi else num_clusters), lo, hi

    def filter_embeddings(self, embeddings: np.ndarray, active_frames: np.ndarray, single_frames: np.ndarray,
                          num_frames: int, min_frames_ratio: float = 0.1):
        """Embeddings used to build the clusters: active speakers with a finite embedding and at least
        round(ratio * T) single-speaker frames (all active ones when fewer than two qualify).  `active_frames` /
        `single_frames` are the (C,S) counters of dz_embedding_masks."""
        ok = (active_frames > 0) & ~np.isnan(embeddings).any(axis=2)
        keep = ok & (single_frames >= round(min_frames_ratio * num_frames))
        if np.count_nonzero(keep) < 2:
            keep = ok
        ci, si = np.nonzero(keep)
        return embeddings[ci, si], ci, si

    def cluster(self, embeddings: np.ndarray, min_clusters: int, max_clusters: int, num_clusters: Optional[int] = None):
        """(N,D) training embeddings -> (N,) cluster indices."""
        if self.method != "centroid" or self.metric != "cosine":
            raise ValueError("only method='centroid' with metric='cosine' is implemented (the DiariZen configuration)")
        n = embeddings.shape[0]
        if n == 1:
            return np.zeros((1,), dtype=np.uint8)
        min_size = min(self.min_cluster_size, max(1, round(0.1 * n)))
        with np.errstate(divide="ignore", invalid="ignore"):
            unit = embeddings / np.linalg.norm(embeddings, axis=-1, keepdims=True)
        labels, info = DeviceDendrogram(unit, self.device).cut(self.threshold, min_size, min_clusters, max_clusters, num_clusters)
        if info["found_only"]:
            print(f"Found only {info['num_large']} clusters. Using a smaller value than {min_size} for `min_cluster_size` might help.")
        self.last_cut = info
        return absorb_small_clusters(unit, labels.astype(np.int64), min_size)

    def assign_embeddings(self, embeddings, ci, si, train_clusters):
        """Centroids = plain means of the raw training embeddings per cluster; soft score = 2 - cosine distance of every
        (chunk, speaker) embedding to every centroid; hard = one distinct cluster per local speaker, maximal total score."""
        K = int(np.max(train_clusters)) + 1
        Cn, S, D = embeddings.shape
        train = embeddings[ci, si]
        centroids = np.stack([train[train_clusters == k].mean(axis=0) for k in range(K)])
        soft = 2 - _cosine_cdist(embeddings.reshape(Cn * S, D), centroids).reshape(Cn, S, K)
        if self.constrained_assignment:
            hard = device_assign(np.nan_to_num(soft, nan=np.nanmin(soft)), self.device)
        else:
            hard = np.argmax(soft, axis=2).astype(np.int8)
        return hard, soft, centroids

    def __call__(self, embeddings: np.ndarray, segmentations=None, num_clusters=None, min_clusters=None,
                 max_clusters=None, frame_stats: Optional[Tuple[np.ndarray, np.ndarray, int]] = None, **kwargs):
        active_frames, single_frames, T = frame_stats if frame_stats is not None else frame_statistics(segmentations)
        train, ci, si = self.filter_embeddings(embeddings, active_frames, single_frames, T)
        num_clusters, lo, hi = self.set_num_clusters(train.shape[0], num_clusters, min_clusters, max_clusters)
        if hi < 2:      # a single speaker is imposed: nothing to cluster
            Cn, S, _ = embeddings.shape
            return (np.zeros((Cn, S), dtype=np.int8), np.ones((Cn, S, 1)), np.mean(train, axis=0, keepdims=True))
        return self.assign_embeddings(embeddings, ci, si, self.cluster(train, lo, hi, num_clusters))


# ---------------------------------------------------------------------------------------------------------
# VBx: AHC initialisation + variational-Bayes GMM over PLDA-space x-vectors
# ---------------------------------------------------------------------------------------------------------
def _unit_rows(x: np.ndarray) -> np.ndarray:
    return x / np.linalg.norm(x, axis=1, ord=2)[:, None]


class PldaTransform:
    """The x-vector -> PLDA latent space map of diarizen/clustering/VBx.py:146-178, built once per model from
    <hub>/plda/xvec_transform.npz and <hub>/plda/plda.npz (the reference rebuilds it on every call)."""

    def __init__(self, xvec_transform, plda):
        from scipy.linalg import eigh   # one generalized 128x128 eigenproblem at model-load time
        self.mean1 = np.asarray(xvec_transform["mean1"], dtype=np.float64)
        self.mean2 = np.asarray(xvec_transform["mean2"], dtype=np.float64)
        self.lda = np.asarray(xvec_transform["lda"], dtype=np.float64)
        self.mu = np.asarray(plda["mu"], dtype=np.float64)
        tr = np.asarray(plda["tr"], dtype=np.float64)
        psi = np.asarray(plda["psi"], dtype=np.float64)
        within = np.linalg.inv(tr.T.dot(tr))
        between = np.linalg.inv((tr.T / psi).dot(tr))
        ev, vec = eigh(between, within)
        self.psi = ev[::-1].copy()
        self.basis = vec.T[::-1].copy()

    @classmethod
    def from_dir(cls, plda_dir):
        import os
        return cls(np.load(os.path.join(plda_dir, "xvec_transform.npz")), np.load(os.path.join(plda_dir, "plda.npz")))

    def __call__(self, x: np.ndarray, lda_dim: int) -> np.ndarray:
        d_in, d_out = self.lda.shape
        y = np.sqrt(d_in) * _unit_rows(np.asarray(x, dtype=np.float64) - self.mean1)
        y = np.sqrt(d_out) * _unit_rows(y.dot(self.lda) - self.mean2)
        return (y - self.mu).dot(self.basis.T)[:, :lda_dim]


def device_vb_gmm(X: np.ndarray, Phi: np.ndarray, gamma0: np.ndarray, Fa: float, Fb: float, max_iters: int,
                  epsilon: float = 1e-4, device=None):
    """diarizen/clustering/VBx.py:73-113 (loopProb = 0): both halves of each iteration run on the GPU in float64
    (dz_vbx_model / dz_vbx_resp); the host keeps only the S-vector prior and the scalar ELBO convergence test."""
    L = _lib.lib()
    dev = torch.device(device if device is not None else "cuda")
    N, D = X.shape
    S = gamma0.shape[1]
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        Xd = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float64)).to(dev)
        phi = torch.from_numpy(np.ascontiguousarray(Phi, dtype=np.float64)).to(dev)
        rho = (Xd * torch.sqrt(phi)).contiguous()
        G = (-0.5 * ((Xd * Xd).sum(dim=1) + D * float(np.log(2 * np.pi)))).contiguous()
        gamma = torch.from_numpy(np.ascontiguousarray(gamma0, dtype=np.float64)).to(dev)
        alpha = torch.empty((S, D), dtype=torch.float64, device=dev)
        invl = torch.empty((S, D), dtype=torch.float64, device=dev)
        acc = torch.zeros((S + 1,), dtype=torch.float64, device=dev)
        pi = np.ones(S) / S
        hist = []
        for it in range(max_iters):
            _lib.check(L.dz_vbx_model(gamma.data_ptr(), rho.data_ptr(), phi.data_ptr(), N, D, S, Fa / Fb,
                                      alpha.data_ptr(), invl.data_ptr(), stream))
            pid = torch.from_numpy(pi).to(dev)
            acc.zero_()
            _lib.check(L.dz_vbx_resp(rho.data_ptr(), G.data_ptr(), alpha.data_ptr(), invl.data_ptr(), phi.data_ptr(),
                                     pid.data_ptr(), N, D, S, Fa, gamma.data_ptr(), acc.data_ptr(),
                                     acc.data_ptr() + 8 * S, stream))
            model_term = torch.sum(torch.log(invl) - invl - alpha * alpha + 1)
            a = acc.cpu().numpy()
            pi = a[:S] / a[:S].sum()
            elbo = float(a[S]) + Fb * 0.5 * float(model_term)
            hist.append(elbo)
            if it > 0 and elbo - hist[-2] < epsilon:
                break
        return gamma.cpu().numpy(), pi, hist


class VBxClustering(AgglomerativeClustering):
    """reference: pyannote-audio/pyannote/audio/pipelines/clustering.py:601-700 (the HF checkpoints' default method)."""

    def __init__(self, metric: str = "cosine", max_num_embeddings: float = np.inf, constrained_assignment: bool = True,
                 plda_dir: str = "", lda_dim: int = 128, maxIters: int = 20, device=None):
        super().__init__(metric=metric, max_num_embeddings=max_num_embeddings,
                         constrained_assignment=constrained_assignment, device=device)
        self.ahc_criterion = "distance"
        self.ahc_threshold = 0.6
        self.plda_dir = plda_dir
        self.lda_dim = lda_dim
        self.maxIters = maxIters
        self.Fa = 0.07
        self.Fb = 0.8
        self._plda = None
        self._plda_key = None

    def plda(self) -> PldaTransform:
        if self._plda is None or self._plda_key != self.plda_dir:
            self._plda = PldaTransform.from_dir(self.plda_dir)
            self._plda_key = self.plda_dir
        return self._plda

    def __call__(self, embeddings: np.ndarray, segmentations=None, num_clusters=None, min_clusters=None,
                 max_clusters=None, frame_stats: Optional[Tuple[np.ndarray, np.ndarray]] = None, **kwargs):
        active_frames, single_frames, T = frame_stats if frame_stats is not None else frame_statistics(segmentations)
        train, _, _ = self.filter_embeddings(embeddings, active_frames, single_frames, T, min_frames_ratio=0.1)
        Cn, S, D = embeddings.shape
        if train.shape[0] < 2:
            return (np.zeros((Cn, S), dtype=np.int8), np.ones((Cn, S, 1)), np.mean(train, axis=0, keepdims=True))
        normed = train / np.linalg.norm(train, axis=1, keepdims=True)
        dendrogram = DeviceDendrogram(normed, self.device)
        if self.ahc_criterion == "distance":
            ahc, _ = dendrogram.cut(self.ahc_threshold)
        elif self.ahc_criterion == "maxclust":
            ahc = fcluster_maxclust(dendrogram.Z(), int(self.ahc_threshold)) - 1
        else:
            raise ValueError(f"unsupported ahc_criterion {self.ahc_criterion!r}")
        _, ahc = np.unique(ahc, return_inverse=True)
        plda = self.plda()
        fea = plda(train, self.lda_dim)
        # VBx.py:117-119: one-hot AHC labels, softmax-smoothed with factor 7
        q0 = np.zeros((len(ahc), int(ahc.max()) + 1))
        q0[np.arange(len(ahc)), ahc] = 7.0
        q0 = np.exp(q0 - q0.max(axis=1, keepdims=True))
        q0 /= q0.sum(axis=1, keepdims=True)
        q, sp, _ = device_vb_gmm(fea, plda.psi[:self.lda_dim], q0, self.Fa, self.Fb, self.maxIters, device=self.device)
        centroids = q[:, sp > 1e-7].T @ train.reshape(-1, D)
        soft = 2 - _cosine_cdist(embeddings.reshape(Cn * S, D), centroids).reshape(Cn, S, -1)
        if self.constrained_assignment:
            sc = np.nan_to_num(soft, nan=np.nanmin(soft))
            hard = device_assign(sc, self.device)
        else:
            hard = np.argmax(soft, axis=2)
        # clustering.py:696-697: np.unique over ALL entries - an unassigned (-2) slot shifts every label up by one
        _, hard = np.unique(hard, return_inverse=True)
        return hard.reshape(Cn, S), soft, centroids
