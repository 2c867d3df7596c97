"""The reference's own known-answer tests for this path, replayed against the oracle (CPU).
  pyannote-audio/tests/test_stats_pool.py:28-131   (5 KATs, values rounded to 4 decimals)
  pyannote-audio/tests/utils/test_powerset.py:29-76 (powerset <-> multilabel round trip, class order)
  pyannote-audio/tests/test_clustering.py:38-61     (centroid AHC must not over-merge 2 embeddings into 1 cluster)
"""
import itertools

import numpy as np
import torch

from oracle.emb_oracle import stats_pool
from oracle.pipeline_oracle import ahc_cluster
from oracle.seg_oracle import powerset_mapping, to_multilabel


def _sp(x, w):
    if w is None:
        return torch.cat([x.mean(dim=-1), x.std(dim=-1, correction=1)], dim=-1)
    if w.dim() == 2:
        return stats_pool(x, w[:, None]).squeeze(1)
    return stats_pool(x, w)


X = torch.Tensor([[[2.0, 4.0], [2.0, 4.0]], [[1.0, 1.0], [1.0, 1.0]]])


def test_stats_pool_weightless():
    assert torch.equal(torch.round(_sp(X, None), decimals=4), torch.Tensor([[3.0, 3.0, 1.4142, 1.4142], [1.0, 1.0, 0.0, 0.0]]))


def test_stats_pool_one_speaker():
    w = torch.Tensor([[0.5, 0.01], [0.2, 0.1]])
    assert torch.equal(torch.round(_sp(X, w), decimals=4), torch.Tensor([[2.0392, 2.0392, 1.4142, 1.4142], [1.0, 1.0, 0.0, 0.0]]))


def test_stats_pool_multi_speaker():
    w = torch.Tensor([[[0.1, 0.2], [0.2, 0.3]], [[0.001, 0.001], [0.2, 0.3]]])
    assert torch.equal(torch.round(_sp(X, w), decimals=4),
                       torch.Tensor([[[3.3333, 3.3333, 1.4142, 1.4142], [3.2, 3.2, 1.4142, 1.4142]],
                                     [[1.0, 1.0, 0.0, 0.0], [1.0, 1.0, 0.0, 0.0]]]))


def test_stats_pool_frame_mismatch():
    x = torch.Tensor([[[2.0, 2.0], [2.0, 2.0]], [[1.0, 1.0], [1.0, 1.0]]])
    w = torch.Tensor([[0.5], [0.2]])     # one weight frame for two feature frames -> nearest interpolation
    assert torch.equal(torch.round(_sp(x, w), decimals=4), torch.Tensor([[2.0, 2.0, 0.0, 0.0], [1.0, 1.0, 0.0, 0.0]]))


def test_stats_pool_all_zero_weights():
    w = torch.Tensor([[0.5, 0.01], [0.0, 0.0]])
    assert torch.equal(torch.round(_sp(X, w), decimals=4), torch.Tensor([[2.0392, 2.0392, 1.4142, 1.4142], [0.0, 0.0, 0.0, 0.0]]))


def test_powerset_roundtrip_and_order():
    """test_powerset.py: every multilabel vector with <= max_set_size active classes maps to exactly one powerset class
    and back; class order = by set size, then lexicographic combinations."""
    m = powerset_mapping(4, 2)
    assert m.shape == (11, 4)
    expected = [()] + [(i,) for i in range(4)] + list(itertools.combinations(range(4), 2))
    for row, comb in zip(m, expected):
        assert tuple(np.nonzero(row.numpy())[0]) == comb
    logp = torch.log_softmax(torch.eye(11) * 10, dim=-1)[None]      # one-hot-ish powerset scores
    assert torch.equal(to_multilabel(logp)[0], m)


def test_clustering_centroid_does_not_overmerge():
    """tests/test_clustering.py:38-61: 2 embeddings, threshold 0, min_cluster_size 0, 2 clusters requested -> [0, 1]."""
    emb = np.array([[1.0, 1.0, 1.0, 1.0], [1.0, 2.0, 1.0, 2.0]], dtype=np.float32)
    clusters = ahc_cluster(emb, threshold=0.0, min_cluster_size=0, min_clusters=2, max_clusters=2, num_clusters=2)
    assert np.array_equal(clusters, np.array([0, 1]))
