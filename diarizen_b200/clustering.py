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

    lo, hi = 0, n - 1
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
    dev = torch.device(device if device is not None else "cuda:0")
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
        if frame_stats is None:
            seg = np.asarray(getattr(segmentations, "data", segmentations))
            active_frames = np.sum(seg, axis=1)
            single = (np.sum(seg, axis=2, keepdims=True) == 1)
            single_frames = np.sum(seg * single, axis=1)
            T = seg.shape[1]
        else:
            active_frames, single_frames, T = frame_stats
        train, _, _ = self.filter_embeddings(embeddings, active_frames, single_frames, T, min_frames_ratio=0.1)
        Cn, S, D = embeddings.shape
        if train.shape[0] < 2:
            return (np.zeros((Cn, S), dtype=np.int8), np.ones((Cn, S, 1)), np.mean(train, axis=0, keepdims=True))
        normed = train / np.linalg.norm(train, axis=1, keepdims=True)
        Z = device_linkage_centroid(normed, self.device)
        if self.ahc_criterion == "distance":
            ahc = fcluster_distance(Z, self.ahc_threshold) - 1
        elif self.ahc_criterion == "maxclust":
            ahc = fcluster_maxclust(Z, int(self.ahc_threshold)) - 1
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
