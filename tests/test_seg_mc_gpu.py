"""(f4) the multi-channel segmentation model on the GPU vs golden vectors produced by the REFERENCE's own modules
(scripts/make_mc_golden.py: wav2vec2_model.extract_features_mc + CrossChannelAttention + ConformerEncoder assembled as
model_wavlm_conformer_mc.py:241-282).  Both the post-norm (base) and the pre-norm (large) wiring."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("arch", ["tiny_base", "tiny_large"])
@pytest.mark.parametrize("precision,tol", [("bf16x3", 1e-3), ("fp16", 1e-2)])
def test_multichannel_forward(arch, precision, tol):
    from diarizen_b200.archs import get_arch, init_state_dict
    from diarizen_b200.segmentation_mc import MCSegmentationModel
    z = np.load(os.path.join(G, f"seg_mc_{arch}.npz"))
    a = get_arch(arch)
    sd = init_state_dict(a, seed=int(z["seed"]))
    sd.update({k[len("fusion."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("fusion.")})
    m = MCSegmentationModel(a, sd, fusion_dim=int(z["cfg_fusion_dim"]), fusion_heads=int(z["cfg_fusion_heads"]),
                            fusion_layers=int(z["cfg_fusion_layers"]), precision=precision)
    logp, ml, att = m.hard(torch.from_numpy(z["wav"]))
    torch.cuda.synchronize()
    assert tuple(logp.shape) == z["logp"].shape and tuple(att.shape) == z["att"].shape
    e_att = float(np.abs(att.cpu().numpy() - z["att"]).max())
    e_logp = float(np.abs(logp.cpu().numpy() - z["logp"]).max())
    assert e_att < tol, f"channel attention max err {e_att:.3e}"
    assert e_logp < tol, f"max |dlogp| = {e_logp:.3e}"
    # a second call (plans cached) gives the same result; single-channel input degenerates to attention weights of 1
    logp2, _, _ = m.hard(torch.from_numpy(z["wav"]))
    assert torch.equal(logp, logp2)
    _, _, att1 = m.hard(torch.from_numpy(z["wav"][:, :1]))
    assert torch.allclose(att1, torch.ones_like(att1))


def test_multichannel_session_matches_reference_recipe():
    """`DiariZenPipeline.diarize_session` on a 3-channel recording == the reference recipe's `diarize_session`
    (recipes/diar_ssl_mc/infer_avg.py:47-118, run through oracle/ref_glue.py around the reference's own MC model modules):
    window decisions, channel weights, attention-weighted embeddings, clusters and RTTM text."""
    from diarizen_b200.archs import get_arch, init_state_dict
    from diarizen_b200.pipeline import DiariZenPipeline
    z = np.load(os.path.join(G, "glue_mc_session.npz"))
    a = get_arch("tiny_base_mc")
    sd = init_state_dict(a, int(z["weights_seed"]), float(z["classifier_gain"]))
    sd.update({k[len("fusion."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("fusion.")})
    pipe = DiariZenPipeline.from_random_init("tiny_base_mc", seed=int(z["weights_seed"]), seg_duration=float(z["seg_duration"]), batch_size=8,
                                             min_cluster_size=int(z["min_cluster_size"]), classifier_gain=float(z["classifier_gain"]),
                                             precision="bf16x3", seg_state_dict=sd,
                                             multichannel=dict(fusion_dim=int(z["cfg_fusion_dim"]), fusion_heads=int(z["cfg_fusion_heads"]),
                                                               fusion_layers=int(z["cfg_fusion_layers"])))
    wav = torch.from_numpy(z["wav_i16"].astype(np.float32) / 32768.0)
    ann = pipe.diarize_session({"waveform": wav, "sample_rate": 16000}, sess_name="sess")
    res = pipe.last
    flips = np.argwhere(pipe.last_raw.cpu().numpy() != z["raw_segmentations"])
    assert flips.size == 0, f"{len(flips)} window decisions differ, first {flips[:5].tolist()}"
    w_ref = z["attention3"].mean(axis=(1, 2))
    assert np.abs(pipe.last_channel_weights.cpu().numpy() - w_ref).max() < 1e-3
    assert np.array_equal(res["segmentations"].cpu().numpy(), z["segmentations"])
    scale = np.abs(z["embeddings"]).max()
    assert np.abs(res["embeddings"] - z["embeddings"]).max() / scale < 2e-3
    assert np.array_equal(res["hard_clusters"], z["hard_clusters"])
    assert np.array_equal(res["discrete"], z["discrete"])
    assert ann.to_rttm() == str(z["rttm"])


def test_multichannel_hub_directory(tmp_path):
    """config.toml with channel_fusion_* model args -> the multi-channel model is built by the hub-directory loader"""
    from diarizen_b200.archs import get_arch, init_resnet_state_dict, init_state_dict
    from diarizen_b200.checkpoints import write_hub_snapshot
    from diarizen_b200.pipeline import DiariZenPipeline
    z = np.load(os.path.join(G, "glue_mc_session.npz"))
    a = get_arch("tiny_base_mc")
    sd = init_state_dict(a, int(z["weights_seed"]), float(z["classifier_gain"]))
    sd.update({k[len("fusion."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("fusion.")})
    fusion = dict(fusion_dim=int(z["cfg_fusion_dim"]), fusion_heads=int(z["cfg_fusion_heads"]), fusion_layers=int(z["cfg_fusion_layers"]))
    write_hub_snapshot(tmp_path / "hub", a, sd, init_resnet_state_dict(int(z["weights_seed"])),
                       {"seg_duration": 5.0, "segmentation_step": 0.1, "batch_size": 8, "apply_median_filtering": True},
                       {"method": "AgglomerativeClustering", "min_speakers": 1, "max_speakers": 20, "ahc_criterion": "distance", "ahc_threshold": 0.70,
                        "min_cluster_size": int(z["min_cluster_size"])}, fusion=fusion)
    pipe = DiariZenPipeline.from_pretrained(str(tmp_path / "hub"), precision="bf16x3")
    wav = torch.from_numpy(z["wav_i16"].astype(np.float32) / 32768.0)
    assert pipe.diarize_session({"waveform": wav, "sample_rate": 16000}, sess_name="sess").to_rttm() == str(z["rttm"])
