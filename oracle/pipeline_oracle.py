"""TEST INFRASTRUCTURE ONLY - numpy/scipy restatement of the CPU glue of `DiariZenPipeline.__call__`.

Follows (pa = pyannote-audio/pyannote/audio):
  diarizen/pipelines/inference.py:121-192          orchestration
  pa/core/inference.py:237-409                      slide (windowing, zero-padded last chunk)
  pa/core/inference.py:543-666                      aggregate (overlap-add, mean / sum)
  pa/pipelines/utils/diarization.py:122-157         speaker_count
  pa/pipelines/speaker_diarization.py:228-375       get_embeddings (mask selection), :377-425 reconstruct
  pa/pipelines/utils/diarization.py:193-239         to_diarization
  pa/pipelines/clustering.py:47-322, :363-513       filter / AHC / assign
  pa/utils/signal.py:254-317                        Binarize
`pyannote.core` (SlidingWindow, Annotation, RTTM formatting) is not in /root/reference: its semantics are restated
from SURVEY.md Appendix B - parity unpinned by any reference test.  scipy (linkage, fcluster, cdist,
linear_sum_assignment, median_filter) is used as-is, exactly as the reference does.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

SR = 16000
FRAME_DURATION = 400 / SR   # receptive field size of the conv stack (model_wavlm_conformer.py:126-150)
FRAME_STEP = 320 / SR


@dataclass
class SlidingWindow:
    start: float
    duration: float
    step: float

    def closest_frame(self, t: float) -> int:
        return int(np.rint((t - self.start - 0.5 * self.duration) / self.step))

    def middle(self, i: int) -> float:
        # pyannote.core: window[i] = Segment(start + i * step, start + i * step + duration); Segment.middle = .5 * (start + end).
        # The order of operations matters for the third decimal printed in the RTTM (pinned by tests/golden/glue_*.npz).
        s = self.start + i * self.step
        return 0.5 * (s + (s + self.duration))


def slide_windows(wav: np.ndarray, window: int, step: int) -> np.ndarray:
    """core/inference.py:281-299 -> (C, window) float32, last chunk zero padded when needed."""
    n = wav.shape[0]
    chunks = []
    num = 0
    if n >= window:
        num = (n - window) // step + 1
        chunks = [wav[c * step:c * step + window] for c in range(num)]
    has_last = (n < window) or ((n - window) % step > 0)
    if has_last:
        last = wav[num * step:]
        chunks.append(np.pad(last, (0, window - last.shape[0])))
    return np.stack(chunks).astype(np.float32)


def aggregate(scores: np.ndarray, chunk_duration: float, chunk_step: float, skip_average: bool,
              missing: float = 0.0, epsilon: float = 1e-12) -> np.ndarray:
    """core/inference.py:543-666 with hamming=False, warm_up=(0,0).  scores (C, T, K), NaN = absent."""
    C, T, K = scores.shape
    frames = SlidingWindow(0.0, FRAME_DURATION, FRAME_STEP)   # start overridden by the chunk start (:577-581)
    mask = 1.0 - np.isnan(scores)
    sc = np.nan_to_num(scores, nan=0.0)
    num_frames = frames.closest_frame(0.0 + chunk_duration + (C - 1) * chunk_step + 0.5 * frames.duration) + 1
    agg = np.zeros((num_frames, K), dtype=np.float32)
    cnt = np.zeros((num_frames, K), dtype=np.float32)
    amask = np.zeros((num_frames, K), dtype=np.float32)
    for c in range(C):
        s0 = frames.closest_frame(c * chunk_step + 0.5 * frames.duration)
        agg[s0:s0 + T] += (sc[c] * mask[c])[: max(0, num_frames - s0)]
        cnt[s0:s0 + T] += mask[c][: max(0, num_frames - s0)]
        amask[s0:s0 + T] = np.maximum(amask[s0:s0 + T], mask[c][: max(0, num_frames - s0)])
    avg = agg if skip_average else agg / np.maximum(cnt, epsilon)
    avg[amask == 0.0] = missing
    return avg


def speaker_count(binarized: np.ndarray, chunk_duration: float, chunk_step: float) -> np.ndarray:
    """utils/diarization.py:122-157 -> (F, 1) uint8."""
    c = aggregate(np.sum(binarized, axis=-1, keepdims=True), chunk_duration, chunk_step, skip_average=False)
    return np.rint(c).astype(np.uint8)


def embedding_masks(binarized: np.ndarray, min_num_frames: int) -> np.ndarray:
    """speaker_diarization.py:271-320: per (chunk, speaker) the clean (non-overlap) mask when it keeps more than
    `min_num_frames` frames, else the full mask.  -> (C, S, T) float32."""
    seg = np.nan_to_num(binarized, nan=0.0).astype(np.float32)
    clean = seg * (np.sum(seg, axis=2, keepdims=True) < 2)
    use_clean = clean.sum(axis=1) > min_num_frames           # (C, S)
    out = np.where(use_clean[:, None, :], clean, seg)
    return np.transpose(out, (0, 2, 1)).copy()


def crop_chunks(wav: np.ndarray, num_chunks: int, chunk_duration: float, chunk_step: float) -> np.ndarray:
    """core/io.py:307-436 (mode="pad") as used by get_embeddings: start=floor(t0*sr), n=floor(dur*sr), right zero pad."""
    n = int(math.floor(chunk_duration * SR))
    out = np.zeros((num_chunks, n), dtype=np.float32)
    for c in range(num_chunks):
        s = int(math.floor((c * chunk_step) * SR))
        seg = wav[s:s + n]
        out[c, :seg.shape[0]] = seg
    return out


def filter_embeddings(embeddings: np.ndarray, binarized: np.ndarray, min_frames_ratio: float = 0.1):
    """clustering.py:111-157 (no subsampling: max_num_embeddings = inf for AHC)."""
    active = np.sum(binarized, axis=1) > 0
    valid = ~np.any(np.isnan(embeddings), axis=2)

    def by_frames(min_frames):
        single = (np.sum(binarized, axis=2, keepdims=True) == 1)
        return np.sum(binarized * single, axis=1) >= min_frames

    min_frames = round(min_frames_ratio * binarized.shape[1])
    ci, si = np.where(active * valid * by_frames(min_frames))
    if len(ci) < 2:
        ci, si = np.where(active * valid * by_frames(0))
    return embeddings[ci, si], ci, si


def set_num_clusters(num_embeddings, num_clusters=None, min_clusters=None, max_clusters=None):
    """clustering.py:88-109."""
    min_clusters = num_clusters or min_clusters or 1
    min_clusters = max(1, min(num_embeddings, min_clusters))
    max_clusters = num_clusters or max_clusters or num_embeddings
    max_clusters = max(1, min(num_embeddings, max_clusters))
    if min_clusters > max_clusters:
        raise ValueError("min_clusters must be smaller than (or equal to) max_clusters")
    if min_clusters == max_clusters:
        num_clusters = min_clusters
    return num_clusters, min_clusters, max_clusters


def ahc_cluster(embeddings: np.ndarray, threshold: float, min_cluster_size: int, min_clusters: int, max_clusters: int,
                num_clusters: Optional[int] = None, linkage_fn=None, fcluster_fn=None) -> np.ndarray:
    """clustering.py:363-513, method="centroid", metric="cosine".  `linkage_fn(unit_norm_embeddings) -> Z` and
    `fcluster_fn(Z, t) -> labels (1-based)` default to scipy; tests inject the device implementation here."""
    from scipy.cluster.hierarchy import fcluster, linkage
    from scipy.spatial.distance import cdist
    linkage_fn = linkage_fn or (lambda e: linkage(e, method="centroid", metric="euclidean"))
    fcluster_fn = fcluster_fn or (lambda Z, t: fcluster(Z, t, criterion="distance"))
    n = embeddings.shape[0]
    mcs = min(min_cluster_size, max(1, round(0.1 * n)))
    if n == 1:
        return np.zeros((1,), dtype=np.uint8)
    emb = embeddings.copy()
    with np.errstate(divide="ignore", invalid="ignore"):
        emb /= np.linalg.norm(emb, axis=-1, keepdims=True)
    Z = linkage_fn(emb)
    clusters = fcluster_fn(Z, threshold) - 1
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
        for it in np.argsort(np.abs(Z[:, 2] - threshold)):
            if _Z[it, 3] < mcs:
                continue
            clusters = fcluster_fn(_Z, it) - 1
            uniq, counts = np.unique(clusters, return_counts=True)
            large = uniq[counts >= mcs]
            nlarge = len(large)
            if abs(nlarge - num_clusters) < abs(best_n - num_clusters):
                best_it, best_n = it, nlarge
            if nlarge == num_clusters:
                break
        if best_n != num_clusters:
            clusters = fcluster_fn(_Z, best_it) - 1
            uniq, counts = np.unique(clusters, return_counts=True)
            large = uniq[counts >= mcs]
            nlarge = len(large)
    if nlarge == 0:
        clusters[:] = 0
        return clusters
    small = uniq[counts < mcs]
    if len(small) == 0:
        return clusters
    lc = np.vstack([np.mean(emb[clusters == k], axis=0) for k in large])
    scn = np.vstack([np.mean(emb[clusters == k], axis=0) for k in small])
    d = cdist(lc, scn, metric="cosine")
    for sk, lk in enumerate(np.argmin(d, axis=0)):
        clusters[clusters == small[sk]] = large[lk]
    _, clusters = np.unique(clusters, return_inverse=True)
    return clusters


def assign_embeddings(embeddings: np.ndarray, ci, si, train_clusters: np.ndarray, assign_fn=None):
    """clustering.py:175-245 with constrained (Hungarian) assignment :159-173."""
    from scipy.optimize import linear_sum_assignment
    from scipy.spatial.distance import cdist
    K = int(np.max(train_clusters)) + 1
    C, S, D = embeddings.shape
    train = embeddings[ci, si]
    centroids = np.vstack([np.mean(train[train_clusters == k], axis=0) for k in range(K)])
    soft = 2 - cdist(embeddings.reshape(C * S, D), centroids, metric="cosine").reshape(C, S, K)
    sc = np.nan_to_num(soft, nan=np.nanmin(soft))
    hard = -2 * np.ones((C, S), dtype=np.int8)
    if assign_fn is not None:
        return assign_fn(sc), soft, centroids
    for c in range(C):
        rows, cols = linear_sum_assignment(sc[c], maximize=True)
        for s, k in zip(rows, cols):
            hard[c, s] = k
    return hard, soft, centroids


def cluster_call(embeddings: np.ndarray, binarized: np.ndarray, threshold: float, min_cluster_size: int,
                 min_clusters: Optional[int], max_clusters: Optional[int], num_clusters: Optional[int] = None, **hooks):
    """clustering.py:247-322."""
    train, ci, si = filter_embeddings(embeddings, binarized)
    n = train.shape[0]
    num_clusters, min_c, max_c = set_num_clusters(n, num_clusters, min_clusters, max_clusters)
    if max_c < 2:
        C, S, _ = embeddings.shape
        return np.zeros((C, S), dtype=np.int8), np.ones((C, S, 1)), np.mean(train, axis=0, keepdims=True)
    tc = ahc_cluster(train, threshold, min_cluster_size, min_c, max_c, num_clusters,
                     hooks.get("linkage_fn"), hooks.get("fcluster_fn"))
    return assign_embeddings(embeddings, ci, si, tc, hooks.get("assign_fn"))


def reconstruct(segmentations: np.ndarray, hard_clusters: np.ndarray, count: np.ndarray, chunk_duration: float,
                chunk_step: float) -> np.ndarray:
    """speaker_diarization.py:377-425 + utils/diarization.py:193-239 -> (F, K) {0,1} float."""
    C, T, S = segmentations.shape
    K = int(np.max(hard_clusters)) + 1
    clustered = np.nan * np.zeros((C, T, K))
    for c in range(C):
        for k in np.unique(hard_clusters[c]):
            if k == -2:
                continue
            clustered[c, :, k] = np.max(segmentations[c][:, hard_clusters[c] == k], axis=1)
    act = aggregate(clustered, chunk_duration, chunk_step, skip_average=True)
    maxspf = int(np.max(count))
    if act.shape[1] < maxspf:
        act = np.pad(act, ((0, 0), (0, maxspf - act.shape[1])))
    F_ = min(act.shape[0], count.shape[0])
    act, cnt = act[:F_], count[:F_]
    # The reference calls np.argsort(-activations) with the default (unstable) kind: which of several speakers with
    # EQUAL summed activation is kept is platform dependent there (SIMD sort kernels).  The oracle - like the device
    # kernel - resolves ties towards the lower cluster index.
    order = np.argsort(-act, axis=-1, kind="stable")
    binary = np.zeros_like(act)
    for t in range(F_):
        for i in range(int(cnt[t, 0])):
            binary[t, order[t, i]] = 1.0
    return binary


def binarize(discrete: np.ndarray, onset: float = 0.5, offset: float = 0.5) -> List[Tuple[float, float, int]]:
    """utils/signal.py:254-317 with no padding / min durations -> [(start, end, label)] in creation order."""
    frames = SlidingWindow(0.0, FRAME_DURATION, FRAME_STEP)
    F_, K = discrete.shape
    ts = [frames.middle(i) for i in range(F_)]
    out = []
    for k in range(K):
        y = discrete[:, k]
        start, active = ts[0], y[0] > onset
        t = ts[0]
        for t, v in zip(ts[1:], y[1:]):
            if active:
                if v < offset:
                    out.append((start, t, k))
                    start, active = t, False
            elif v > onset:
                start, active = t, True
        if active:
            out.append((start, t, k))
    return out


def to_rttm(turns: List[Tuple[float, float, int]], uri: Optional[str]) -> str:
    """pyannote.core Annotation.to_rttm: tracks sorted by (start, end), then insertion."""
    u = uri if uri is not None else "<NA>"
    lines = []
    for s, e, k in sorted(turns, key=lambda x: (x[0], x[1], str(x[2]))):
        if not (e - s) > 1e-6:       # pyannote.core drops empty segments (Segment.__bool__, precision 1e-6)
            continue
        lines.append(f"SPEAKER {u} 1 {s:.3f} {e - s:.3f} <NA> <NA> {k} <NA> <NA>\n")
    return "".join(lines)


def run_pipeline(wav: np.ndarray, seg_fn, emb_fn, chunk_duration: float, step_ratio: float, threshold: float = 0.70,
                 min_cluster_size: int = 30, min_speakers: Optional[int] = 1, max_speakers: Optional[int] = 20,
                 apply_median_filtering: bool = True, **hooks) -> dict:
    """inference.py:121-192.  seg_fn(chunks (C,N)) -> multilabel (C,T,S) {0,1};  emb_fn(chunks (C,N), masks (C,S,T)) -> (C,S,256)."""
    from scipy.ndimage import median_filter
    window = int(math.floor(chunk_duration * SR))
    chunk_step = step_ratio * chunk_duration
    step = round(chunk_step * SR)
    chunks = slide_windows(wav, window, step)
    seg = np.asarray(seg_fn(chunks), dtype=np.float32)
    if apply_median_filtering:
        seg = median_filter(seg, size=(1, 11, 1), mode="reflect")
    count = speaker_count(seg, chunk_duration, chunk_step)
    C, T, S = seg.shape
    min_num_frames = math.ceil(T * 400 / (chunk_duration * SR))   # speaker_verification.py:677-691 -> 400 samples
    masks = embedding_masks(seg, min_num_frames)
    emb_chunks = crop_chunks(wav, C, chunk_duration, chunk_step)
    emb = np.asarray(emb_fn(emb_chunks, masks), dtype=np.float32)
    if hooks.get("cluster_fn") is not None:      # e.g. the VBx variant (oracle/vbx_oracle.py)
        hard, _, centroids = hooks["cluster_fn"](emb, seg)
    else:
        hard, _, centroids = cluster_call(emb, seg, threshold, min_cluster_size, min_speakers, max_speakers, **hooks)
    count = np.minimum(count, max_speakers).astype(np.int8)
    hard = hard.copy()
    hard[np.sum(seg, axis=1) == 0] = -2
    discrete = reconstruct(seg, hard, count, chunk_duration, chunk_step)
    turns = binarize(discrete)
    return {"segmentations": seg, "count": count, "embeddings": emb, "hard_clusters": hard, "discrete": discrete,
            "turns": turns}
