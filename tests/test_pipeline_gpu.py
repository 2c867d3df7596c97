"""End-to-end pipeline on the GPU vs the numpy/scipy restatement of the reference glue.

(1) glue isolation: the oracle pipeline is fed the GPU networks' own outputs (multilabel windows, embeddings), so any
    difference can only come from windowing / median / count / masks / clustering / reconstruction / binarisation:
    those must agree EXACTLY (same RTTM text).
(2) full parity: oracle networks on the CPU (fp32 torch) vs the GPU pipeline in fp32-class precision, with a widened
    classifier margin (SURVEY.md section 7 "Hard parts"): identical RTTM turn boundaries and labels.
"""
import numpy as np
import pytest
import torch

from oracle import pipeline_oracle as po

pytestmark = pytest.mark.gpu


def _meeting(seconds, seed=0):
    """Synthetic 'meeting': 3 noise sources with different spectra switched on and off in turns."""
    g = torch.Generator().manual_seed(seed)
    n = int(seconds * 16000)
    t = torch.arange(n) / 16000.0
    wav = torch.zeros(n)
    for s, f0 in enumerate((180.0, 320.0, 520.0)):
        src = 0.05 * torch.randn(n, generator=g) + 0.1 * torch.sin(2 * np.pi * f0 * t) * (1 + 0.3 * torch.sin(2 * np.pi * (2 + s) * t))
        gate = torch.zeros(n)
        pos = int(torch.randint(0, 16000, (1,), generator=g))
        while pos < n:
            on = int(torch.randint(8000, 64000, (1,), generator=g))
            gate[pos:pos + on] = 1.0
            pos += on + int(torch.randint(8000, 80000, (1,), generator=g))
        wav += src * gate
    return wav.clamp_(-1, 1)


@pytest.mark.parametrize("arch,dur,secs", [("tiny_base", 5.0, 31.3), ("tiny_large", 16.0, 61.0)])
def test_glue_is_exact(arch, dur, secs):
    from diarizen_b200.pipeline import DiariZenPipeline
    pipe = DiariZenPipeline.from_random_init(arch, seed=2, seg_duration=dur, batch_size=16, min_cluster_size=3,
                                             classifier_gain=40.0, precision="bf16x3")
    wav = _meeting(secs, 1)
    res = pipe.diarize_waveform(wav)
    ann = pipe.to_annotation(res["discrete"], "sess")

    def seg_fn(chunks):
        _, ml = pipe._segmentation.hard(torch.from_numpy(chunks), want_logp=False)
        return ml.cpu().numpy()

    def emb_fn(chunks, masks):
        return pipe._embedding.embed_windows(torch.from_numpy(chunks), torch.from_numpy(masks)).cpu().numpy()

    ref = po.run_pipeline(wav.numpy(), seg_fn, emb_fn, dur, 0.1, threshold=0.70, min_cluster_size=3, min_speakers=1, max_speakers=20)
    assert np.array_equal(res["segmentations"].cpu().numpy().astype(np.float32), ref["segmentations"])
    assert np.array_equal(res["count"].cpu().numpy(), ref["count"][:, 0].astype(np.uint8))
    assert np.array_equal(res["hard_clusters"], ref["hard_clusters"])
    assert np.array_equal(res["discrete"].astype(np.float32), ref["discrete"])
    assert ann.to_rttm() == po.to_rttm(ref["turns"], "sess")
    assert len(ann) > 0


def test_full_parity_rttm():
    from diarizen_b200.archs import get_arch, init_resnet_state_dict, init_state_dict
    from diarizen_b200.pipeline import DiariZenPipeline
    from oracle.emb_oracle import emb_forward
    from oracle.seg_oracle import seg_forward, to_multilabel
    arch, dur = "tiny_base", 5.0
    pipe = DiariZenPipeline.from_random_init(arch, seed=2, seg_duration=dur, batch_size=16, min_cluster_size=3,
                                             classifier_gain=40.0, precision="bf16x3")
    a = get_arch(arch)
    sd = init_state_dict(a, 2, 40.0)
    esd = init_resnet_state_dict(2)
    wav = _meeting(31.3, 1)

    def seg_fn(chunks):
        return to_multilabel(seg_forward(a, sd, torch.from_numpy(chunks))).numpy()

    def emb_fn(chunks, masks):
        return emb_forward(esd, torch.from_numpy(chunks), torch.from_numpy(masks)).numpy()

    ref = po.run_pipeline(wav.numpy(), seg_fn, emb_fn, dur, 0.1, threshold=0.70, min_cluster_size=3, min_speakers=1, max_speakers=20)
    ann = pipe(dict(waveform=wav[None], sample_rate=16000), sess_name="sess")
    flips = np.argwhere(pipe.last["segmentations"].cpu().numpy().astype(np.float32) != ref["segmentations"])
    assert flips.size == 0, f"{len(flips)} frame decisions differ, first at {flips[:5].tolist()}"
    assert ann.to_rttm() == po.to_rttm(ref["turns"], "sess")


def test_example_wav_api(tmp_path):
    """from_random_init -> __call__(path) -> itertracks / RTTM file, on a real 16 kHz wav when it is available."""
    import os
    from diarizen_b200.pipeline import DiariZenPipeline
    path = os.path.join(os.path.dirname(__file__), "golden", "tone_10s.wav")
    pipe = DiariZenPipeline.from_random_init("tiny_base", seed=0, seg_duration=5.0, min_cluster_size=2, classifier_gain=40.0,
                                             rttm_out_dir=str(tmp_path))
    ann = pipe(path, sess_name="tone")
    assert ann.uri == "tone"
    for turn, _, speaker in ann.itertracks(yield_label=True):
        assert turn.end > turn.start and isinstance(speaker, (int, np.integer))
    assert (tmp_path / "tone.rttm").read_text() == ann.to_rttm()
