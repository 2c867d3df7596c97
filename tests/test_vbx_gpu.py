"""GPU: the VBx clustering branch (dz_vbx_model / dz_vbx_resp + device AHC + device assignment) against the oracle
restatement of diarizen/clustering/VBx.py and VBxClustering.__call__ (clustering.py:601-700)."""
import os

import numpy as np
import pytest

import vbx_util

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "vbx.npz")


def test_device_vb_gmm_matches_reference_golden():
    from diarizen_b200.clustering import device_vb_gmm
    g = np.load(GOLD)
    gamma, pi, hist = device_vb_gmm(g["fea"], g["phi"], g["q0"], float(g["Fa"]), float(g["Fb"]), int(g["max_iters"]))
    # float64 throughout; only the summation order differs from numpy's
    np.testing.assert_allclose(gamma, g["gamma"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(pi, g["pi"], rtol=0, atol=1e-10)


@pytest.mark.parametrize("N,S,D,Fa,Fb", [(37, 1, 128, 0.07, 0.8), (500, 7, 128, 0.3, 10.0), (3000, 33, 64, 0.07, 0.8)])
def test_device_vb_gmm_matches_oracle(N, S, D, Fa, Fb):
    from diarizen_b200.clustering import device_vb_gmm
    from oracle.vbx_oracle import init_responsibilities, vb_gmm
    r = np.random.default_rng(N)
    spk = 2.0 * r.standard_normal((max(2, S // 2), D))
    lab_true = r.integers(0, len(spk), size=N)
    X = spk[lab_true] + r.standard_normal((N, D))
    Phi = np.linspace(30.0, 0.1, D)
    labels = r.integers(0, S, size=N)
    labels[:S] = np.arange(S)
    q0 = init_responsibilities(labels)
    g_ref, pi_ref, h_ref = vb_gmm(X, Phi, q0, Fa, Fb, 20)
    g, pi, h = device_vb_gmm(X, Phi, q0, Fa, Fb, 20)
    assert len(h) == len(h_ref)
    np.testing.assert_allclose(h, h_ref, rtol=1e-10)
    np.testing.assert_allclose(g, g_ref, rtol=0, atol=1e-8)
    np.testing.assert_allclose(pi, pi_ref, rtol=0, atol=1e-10)


@pytest.mark.parametrize("seed,crit,thr", [(0, "distance", 0.6), (1, "distance", 0.9), (2, "maxclust", 6), (3, "distance", 0.3)])
def test_vbx_clustering_call_matches_oracle(seed, crit, thr, tmp_path):
    from diarizen_b200.clustering import VBxClustering
    from oracle.vbx_oracle import vbx_cluster_call
    xt, plda = vbx_util.write_plda(str(tmp_path), seed)
    emb, seg = vbx_util.make_embeddings(seed, C=80, n_spk=4)
    clu = VBxClustering(plda_dir=str(tmp_path), device="cuda:0")
    clu.ahc_criterion, clu.ahc_threshold, clu.Fa, clu.Fb = crit, thr, 0.07, 0.8
    hard, soft, cent = clu(embeddings=emb, segmentations=seg)
    o_hard, o_soft, o_cent = vbx_cluster_call(emb, seg, xt, plda, thr, 0.07, 0.8, ahc_criterion=crit)
    assert cent.shape == o_cent.shape
    np.testing.assert_allclose(cent, o_cent, rtol=0, atol=1e-7)
    np.testing.assert_allclose(np.nan_to_num(soft, nan=-9), np.nan_to_num(o_soft, nan=-9), rtol=0, atol=1e-9)
    np.testing.assert_array_equal(hard, o_hard)


def test_vbx_too_few_embeddings(tmp_path):
    from diarizen_b200.clustering import VBxClustering
    vbx_util.write_plda(str(tmp_path), 0)
    emb = np.full((3, 4, 256), np.nan)
    emb[1, 0] = 1.0
    seg = np.zeros((3, 50, 4), dtype=np.float32)
    seg[1, :20, 0] = 1
    hard, soft, cent = VBxClustering(plda_dir=str(tmp_path), device="cuda:0")(embeddings=emb, segmentations=seg)
    assert hard.shape == (3, 4) and not hard.any() and soft.shape == (3, 4, 1) and cent.shape == (1, 256)


def test_pipeline_with_vbx(tmp_path):
    import torch
    from diarizen_b200.pipeline import DiariZenPipeline
    vbx_util.write_plda(str(tmp_path / "plda"), 0)
    pipe = DiariZenPipeline.from_random_init("tiny_large", seed=1, seg_duration=4.0, batch_size=4, classifier_gain=6.0,
                                             vbx={"plda_dir": str(tmp_path / "plda"), "ahc_threshold": 0.6})
    assert type(pipe.clustering).__name__ == "VBxClustering"
    g = torch.Generator().manual_seed(0)
    wav = 0.1 * torch.randn(1, 16000 * 20, generator=g)
    ann = pipe({"waveform": wav, "sample_rate": 16000}, sess_name="vbx")
    turns = list(ann.itertracks(yield_label=True))
    assert all(t[0].end > t[0].start for t in turns)
