"""fcluster restatement vs scipy (numbering included) - CPU, scipy is the oracle (third-party)."""
import numpy as np
import pytest
from scipy.cluster.hierarchy import fcluster, linkage

from diarizen_b200.clustering import fcluster_distance


@pytest.mark.parametrize("n,seed", [(2, 0), (12, 1), (40, 2), (300, 3), (1000, 4)])
def test_fcluster_distance_matches_scipy(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, 16)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    Z = linkage(x, method="centroid", metric="euclidean")
    for t in (0.0, 0.3, 0.7, 1.0, 1.3, 5.0):
        assert np.array_equal(fcluster(Z, t, criterion="distance"), fcluster_distance(Z, t)), (n, t)
    _Z = Z.copy()
    _Z[:, 2] = np.arange(n - 1)
    for it in (0, (n - 1) // 2, n - 2):
        assert np.array_equal(fcluster(_Z, it, criterion="distance"), fcluster_distance(_Z, it))
