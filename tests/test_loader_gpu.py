"""(f1 / f3) the loader paths of the reference on the GPU: hub-snapshot directory (config.toml + pytorch_model.bin + a WavLM
`{config, state_dict}` checkpoint + lightning-wrapped WeSpeaker checkpoint + plda/) and checkpoint-averaged inference
(recipes/diar_ssl/infer_avg.py:292-345) - each must give the RTTM of the same weights loaded directly."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _wav():
    z = np.load(os.path.join(G, "glue_e2e_tiny_base.npz"))
    return torch.from_numpy(z["wav_i16"].astype(np.float32) / 32768.0), str(z["rttm"])


INF = {"seg_duration": 5.0, "segmentation_step": 0.1, "batch_size": 16, "apply_median_filtering": True}
AHC = {"method": "AgglomerativeClustering", "min_speakers": 1, "max_speakers": 20, "ahc_criterion": "distance", "ahc_threshold": 0.70,
       "min_cluster_size": 3}


def test_hub_directory_equals_direct_weights(tmp_path):
    from diarizen_b200.archs import get_arch, init_resnet_state_dict, init_state_dict
    from diarizen_b200.checkpoints import write_hub_snapshot
    from diarizen_b200.pipeline import DiariZenPipeline
    wav, golden_rttm = _wav()
    a = get_arch("tiny_base")
    write_hub_snapshot(tmp_path / "hub", a, init_state_dict(a, 2, 40.0), init_resnet_state_dict(2), INF, AHC)
    pipe = DiariZenPipeline.from_pretrained(str(tmp_path / "hub"), rttm_out_dir=str(tmp_path / "rttm"), precision="bf16x3")
    ann = pipe(dict(waveform=wav[None], sample_rate=16000), sess_name="sess")
    assert ann.to_rttm() == golden_rttm            # == the reference pipeline on the same weights (tests/golden/glue_e2e_tiny_base.npz)
    assert (tmp_path / "rttm" / "sess.rttm").read_text() == golden_rttm
    # inherited surface
    assert pipe.model is pipe._segmentation.model and pipe._segmentation.model.specifications.powerset
    assert pipe.to(torch.device("cuda")) is pipe
    with pytest.raises(TypeError):
        pipe.to("cuda")


def test_hub_directory_with_vbx(tmp_path):
    from vbx_util import make_plda
    from diarizen_b200.archs import get_arch, init_resnet_state_dict, init_state_dict
    from diarizen_b200.checkpoints import write_hub_snapshot
    from diarizen_b200.pipeline import DiariZenPipeline
    wav, _ = _wav()
    a = get_arch("tiny_base")
    vbx = {"method": "VBxClustering", "min_speakers": 1, "max_speakers": 20, "ahc_criterion": "distance", "ahc_threshold": 0.6,
           "Fa": 0.07, "Fb": 0.8, "lda_dim": 128, "max_iters": 20}
    write_hub_snapshot(tmp_path / "hub", a, init_state_dict(a, 2, 40.0), init_resnet_state_dict(2), INF, vbx, plda=make_plda(3))
    pipe = DiariZenPipeline.from_pretrained(str(tmp_path / "hub"), precision="bf16x3")
    assert pipe.clustering.plda_dir == str(tmp_path / "hub" / "plda")
    ann = pipe(dict(waveform=wav[None], sample_rate=16000), sess_name="sess")
    ref = DiariZenPipeline.from_random_init("tiny_base", seed=2, seg_duration=5.0, batch_size=16, classifier_gain=40.0, precision="bf16x3",
                                            ahc_threshold=0.6, vbx={"plda_dir": str(tmp_path / "hub" / "plda")})
    assert ann.to_rttm() == ref(dict(waveform=wav[None], sample_rate=16000), sess_name="sess").to_rttm()


def test_checkpoint_averaged_inference(tmp_path):
    """infer_avg.py: segmentation=[ckpt paths] -> key-wise average at load time; same RTTM as loading the pre-averaged weights."""
    from diarizen_b200.archs import get_arch, init_resnet_state_dict, init_state_dict
    from diarizen_b200.checkpoints import average_states, write_hub_snapshot
    from diarizen_b200.pipeline import DiariZenPipeline
    wav, _ = _wav()
    a = get_arch("tiny_base")
    base = init_state_dict(a, 2, 40.0)
    g = torch.Generator().manual_seed(0)
    states = [{k: v + 0.01 * v.abs().mean() * torch.randn(v.shape, generator=g) for k, v in base.items()} for _ in range(3)]
    paths = []
    for i, sd in enumerate(states):
        p = tmp_path / f"epoch_{i:04d}"
        p.mkdir()
        torch.save(sd, p / "pytorch_model.bin")
        paths.append({"bin_path": p / "pytorch_model.bin"})
    write_hub_snapshot(tmp_path / "hub", a, base, init_resnet_state_dict(2), INF, AHC)
    avg_pipe = DiariZenPipeline(tmp_path / "hub", str(tmp_path / "hub" / "wespeaker" / "pytorch_model.bin"), segmentation=paths, precision="bf16x3")
    write_hub_snapshot(tmp_path / "hub2", a, average_states(states), init_resnet_state_dict(2), INF, AHC)
    direct = DiariZenPipeline.from_pretrained(str(tmp_path / "hub2"), precision="bf16x3")
    f = dict(waveform=wav[None], sample_rate=16000)
    r1, r2 = avg_pipe(f, sess_name="s").to_rttm(), direct(f, sess_name="s").to_rttm()
    assert r1 == r2 and len(r1) > 0
