"""Parity at the BENCHMARKED configuration and precision (BASELINE.json configs[2]: wavlm_large_s80_md, 16 s windows):
the sizes `bench.py` runs, not reduced ones.  Tolerances are BASELINE.json's north star: log-probs within 1e-3 for the
fp32-class mode (bf16x3) and 1e-2 for the one-pass 16-bit mode (fp16); embeddings are compared relative to their scale
(fp32-class) and by angle (fp16: they are consumed through cosine / unit-norm Euclidean distances only); the linkage is
bit-identical to scipy at the recording-scale N, including constructed exact ties."""
import numpy as np
import pytest
import torch

from diarizen_b200.archs import get_arch, init_state_dict
from oracle.seg_oracle import seg_forward, to_multilabel

pytestmark = pytest.mark.gpu

N16 = 256000   # 16 s at 16 kHz -> T = 799 frames


@pytest.fixture(scope="module")
def large_ref():
    a = get_arch("wavlm_large_s80_md")
    sd = init_state_dict(a, 1)
    wav = 0.1 * torch.randn(4, N16, generator=torch.Generator().manual_seed(1234))
    return a, sd, wav, seg_forward(a, sd, wav)


@pytest.mark.parametrize("precision,attn,tol", [("fp16", "tc", 1e-2), ("bf16x3", "tc", 1e-3), ("bf16x3", "simt", 1e-3)])
def test_large_s80_16s(large_ref, precision, attn, tol):
    from diarizen_b200.segmentation import SegmentationModel
    a, sd, wav, ref = large_ref
    m = SegmentationModel(a, sd, precision=precision, attn_impl=attn)
    logp, ml = m.hard(wav.unsqueeze(1))
    torch.cuda.synchronize()
    logp, ml = logp.cpu(), ml.cpu()
    assert logp.shape == ref.shape == (4, 799, 11)
    err = (logp - ref).abs().max().item()
    assert err < tol, f"{precision}/{attn}: max |dlogp| = {err:.3e}"
    top2 = ref.topk(2, dim=-1).values
    safe = (top2[..., 0] - top2[..., 1]) > 2 * tol
    assert (ml[safe].float() == to_multilabel(ref)[safe]).all()


def test_large_s80_16s_engine_batch_independent(large_ref):
    """The pipeline calls the engine with 96 windows at a time; the result for a window must not depend on its batch."""
    from diarizen_b200.segmentation import SegmentationModel
    a, sd, wav, ref = large_ref
    m = SegmentationModel(a, sd, precision="fp16")
    big = wav.repeat(6, 1)[:22]
    l1, _ = m.hard(big.unsqueeze(1))
    l2, _ = m.hard(wav[:2].unsqueeze(1))
    assert torch.equal(l1[:2].cpu(), l2.cpu()) and torch.equal(l1[4:6].cpu(), l2.cpu())


@pytest.mark.parametrize("precision", ["bf16x3", "fp16"])
def test_embedding_16s(precision):
    from diarizen_b200.embedding import EmbeddingModel
    from oracle.emb_oracle import emb_forward, init_resnet_state_dict
    sd = init_resnet_state_dict(0)
    g = torch.Generator().manual_seed(7)
    B, S, T = 3, 4, 799
    wav = 0.1 * torch.randn(B, N16, generator=g)
    masks = (torch.rand(B, S, T, generator=g) > 0.4).float()
    masks[0, 3] = 0.0
    masks[1, 2, :] = 0.0
    masks[1, 2, 17:20] = 1.0          # a 3-frame speaker (just above min_num_frames)
    ref = emb_forward(sd, wav, masks)
    m = EmbeddingModel(sd, precision=precision)
    got = m.embed_windows(wav, masks).cpu()
    assert got.shape == (B, S, 256)
    if precision == "bf16x3":
        err = (got - ref).abs().max().item() / ref.abs().max().item()
        assert err < 1e-3, f"relative max err {err:.3e}"
    else:
        cos = torch.nn.functional.cosine_similarity(got.reshape(-1, 256), ref.reshape(-1, 256), dim=1)
        assert cos.min().item() > 0.9995, f"min cosine {cos.min().item():.6f}"
    assert torch.allclose(got[0, 3], sd["resnet.seg_1.bias"], atol=1e-6)


def _recording_scale_embeddings(n, seed, ties):
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((7, 256))
    x = (centers[rng.integers(0, 7, n)] + 0.6 * rng.standard_normal((n, 256))).astype(np.float32)
    if ties:
        # exact duplicates (distance 0 ties) and mirrored pairs (equal non-zero distances)
        x[1::97] = x[0::97][: len(x[1::97])]
        x[5::211] = x[3::211][: len(x[5::211])]
    x /= np.linalg.norm(x, axis=-1, keepdims=True)
    return x


def test_linkage_recording_scale_bitwise():
    """N = 2241 windows x 4 speakers = 8964 is the 60-min recording of the bench."""
    from scipy.cluster.hierarchy import linkage
    from diarizen_b200.clustering import device_linkage_centroid
    x = _recording_scale_embeddings(8964, 11, False)
    Zref = linkage(x, method="centroid", metric="euclidean")
    Z = device_linkage_centroid(x)
    assert np.array_equal(Z[:, [0, 1, 3]], Zref[:, [0, 1, 3]]), "merge order differs"
    assert np.array_equal(Z[:, 2], Zref[:, 2])


def test_linkage_with_exact_ties_same_partitions():
    """Duplicate embeddings give exactly tied distances.  scipy resolves such ties through the history of its binary heap;
    the device kernel always takes the tied pair with the lowest slot index.  Tied merges commute, so
    the merge heights (as a multiset) and every flat clustering are the same - only the order of the tied rows of Z, and
    with it the numbering of clusters created at a tie, can differ."""
    from scipy.cluster.hierarchy import fcluster, linkage
    from diarizen_b200.clustering import device_linkage_centroid
    x = _recording_scale_embeddings(4000, 11, True)
    Zref = linkage(x, method="centroid", metric="euclidean")
    Z = device_linkage_centroid(x)
    assert np.allclose(np.sort(Z[:, 2]), np.sort(Zref[:, 2]), rtol=0, atol=1e-12)
    for t in (1e-9, 0.3, 0.7, 1.0):
        a, b = fcluster(Z, t, criterion="distance"), fcluster(Zref, t, criterion="distance")
        assert a.max() == b.max()
        pairs = set(zip(a.tolist(), b.tolist()))
        assert len(pairs) == a.max(), f"partitions differ at t = {t}"
