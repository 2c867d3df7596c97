"""The numpy/scipy glue oracle (oracle/pipeline_oracle.py, oracle/vbx_oracle.py) against golden vectors PRODUCED BY THE
REFERENCE'S OWN CODE (scripts/make_glue_golden.py: the unmodified pyannote-audio / diarizen glue files executed behind
third-party stubs, oracle/ref_glue.py).  This is what pins the glue oracle; the CUDA path is compared with the same
goldens in tests/test_glue_golden_gpu.py.  CPU only."""
import os

import numpy as np
import pytest

from oracle import pipeline_oracle as po

G = os.path.join(os.path.dirname(__file__), "golden")


def _params(z, key):
    return eval(str(z[key]), {"__builtins__": {}}, {"dict": dict})


def _feed(z, variant=None):
    raw = z["raw_segmentations"].astype(np.float32)
    emb = z[f"{variant}__embeddings"] if variant and f"{variant}__embeddings" in z.files else z["embeddings"]
    return (lambda chunks: raw), (lambda chunks, masks: emb)


@pytest.mark.parametrize("arch", ["tiny_base", "tiny_large"])
def test_e2e_golden(arch):
    z = np.load(os.path.join(G, f"glue_e2e_{arch}.npz"))
    wav = z["wav_i16"].astype(np.float32) / 32768.0
    seg_fn, emb_fn = _feed(z)
    out = po.run_pipeline(wav, seg_fn, emb_fn, float(z["seg_duration"]), 0.1, threshold=0.70, min_cluster_size=int(z["min_cluster_size"]),
                          min_speakers=1, max_speakers=20)
    assert np.array_equal(out["segmentations"], z["segmentations"])
    assert np.array_equal(out["count"][:, 0], z["count"])
    assert np.array_equal(out["hard_clusters"], z["hard_clusters"])
    assert np.array_equal(out["discrete"], z["discrete"])
    assert po.to_rttm(out["turns"], "sess") == str(z["rttm"])


def test_e2e_large_s80_golden():
    """the 5-minute wavlm_large_s80_md golden: oracle glue on the stored window decisions / embeddings == the reference's output;
    also checks that the integer-synthetic recording is regenerated bit-exactly on this machine"""
    import hashlib
    from synth_audio import integer_meeting
    z = np.load(os.path.join(G, "glue_e2e_large_s80.npz"))
    w16 = integer_meeting(float(z["seconds"]), int(z["audio_seed"]))
    assert hashlib.sha1(w16.tobytes()).hexdigest() == str(z["audio_sha1"])
    seg_fn, emb_fn = _feed(z)
    out = po.run_pipeline(w16.astype(np.float32) / 32768.0, seg_fn, emb_fn, 16.0, 0.1, threshold=0.70, min_cluster_size=int(z["min_cluster_size"]),
                          min_speakers=1, max_speakers=20)
    assert np.array_equal(out["segmentations"], z["segmentations"]) and np.array_equal(out["hard_clusters"], z["hard_clusters"])
    assert np.array_equal(out["discrete"], z["discrete"]) and po.to_rttm(out["turns"], "sess") == str(z["rttm"])


def _synth_cases():
    out = []
    for name in ("5s", "16s"):
        z = np.load(os.path.join(G, f"glue_synth_{name}.npz"))
        out += [(name, str(v)) for v in z["variants"]]
    return out


@pytest.mark.parametrize("name,variant", _synth_cases())
def test_synth_golden(name, variant, tmp_path):
    z = np.load(os.path.join(G, f"glue_synth_{name}.npz"))
    kw = _params(z, f"{variant}__params")
    wav = np.zeros(int(z["num_samples"]), dtype=np.float32)
    seg_fn, emb_fn = _feed(z, variant)
    dur = float(z["seg_duration"])
    if "vbx" in kw:
        from vbx_util import make_plda
        from oracle import vbx_oracle as vo
        xt, plda = make_plda(int(z["plda_seed"]))
        v = kw["vbx"]
        hooks = {"cluster_fn": lambda emb, seg: vo.vbx_cluster_call(emb, seg, xt, plda, kw["ahc_threshold"], v["Fa"], v["Fb"], v["lda_dim"],
                                                                   v["max_iters"], v["ahc_criterion"])}
    else:
        hooks = {}
    out = po.run_pipeline(wav, seg_fn, emb_fn, dur, 0.1, threshold=kw["ahc_threshold"], min_cluster_size=kw.get("min_cluster_size", 30),
                          min_speakers=kw["min_speakers"], max_speakers=kw["max_speakers"],
                          apply_median_filtering=kw.get("apply_median_filtering", True), **hooks)
    seg_exp = z[f"{variant}__segmentations"] if f"{variant}__segmentations" in z.files else z["segmentations"]
    assert np.array_equal(out["segmentations"], seg_exp)
    assert np.array_equal(out["count"][:, 0], np.minimum(z[f"{variant}__count"], kw["max_speakers"]))
    assert np.array_equal(out["hard_clusters"], z[f"{variant}__hard_clusters"])
    assert np.array_equal(out["discrete"], z[f"{variant}__discrete"])
    assert po.to_rttm(out["turns"], "sess") == str(z[f"{variant}__rttm"])


def _clu_names():
    return [str(n) for n in np.load(os.path.join(G, "glue_clustering.npz"))["names"]]


@pytest.mark.parametrize("name", _clu_names())
def test_clustering_golden(name, tmp_path):
    z = np.load(os.path.join(G, "glue_clustering.npz"))
    prm = _params(z, f"{name}__params")
    emb, seg = z[f"{name}__embeddings"], z[f"{name}__segmentations"].astype(np.float32)
    if "vbx" in prm:
        from vbx_util import make_plda
        from oracle import vbx_oracle as vo
        xt, plda = make_plda(int(z["plda_seed"]))
        v = prm["vbx"]
        hard, soft, cent = vo.vbx_cluster_call(emb, seg, xt, plda, v["ahc_threshold"], v["Fa"], v["Fb"], v["lda_dim"], v["maxIters"],
                                               v["ahc_criterion"])
    else:
        hard, soft, cent = po.cluster_call(emb, seg, prm["threshold"], prm["mcs"], prm["min"], prm["max"], num_clusters=prm.get("num"))
    assert np.array_equal(np.asarray(hard).astype(np.int16), z[f"{name}__hard"])
    assert np.allclose(soft, z[f"{name}__soft"], atol=1e-9, equal_nan=True)
    assert np.allclose(cent, z[f"{name}__centroids"], atol=1e-9)
