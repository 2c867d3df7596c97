"""Embedding forward on the GPU vs the fp32 torch oracle (oracle/emb_oracle.py, pinned against the reference
ResNet / StatsPool modules)."""
import pytest
import torch

from oracle.emb_oracle import compute_fbank, emb_forward, init_resnet_state_dict

pytestmark = pytest.mark.gpu


def _setup(B, N, S, T, seed=0):
    sd = init_resnet_state_dict(seed)
    g = torch.Generator().manual_seed(7)
    wav = 0.1 * torch.randn(B, N, generator=g)
    masks = (torch.rand(B, S, T, generator=g) > 0.4).float()
    masks[0, S - 1] = 0.0          # an inactive speaker: embedding must equal seg_1.bias exactly
    return sd, wav, masks


def test_fbank_matches_kaldi():
    from diarizen_b200.embedding import EmbeddingModel
    sd, wav, masks = _setup(2, 32000, 4, 99)
    m = EmbeddingModel(sd, precision="bf16x3")
    m.embed_windows(wav, masks)
    import torchaudio.compliance.kaldi as kaldi
    ref = torch.stack([kaldi.fbank(w[None] * 32768, num_mel_bins=80, frame_length=25, frame_shift=10, dither=0.0,
                                   sample_frequency=16000, window_type="hamming", use_energy=False) for w in wav])
    got = m.fbank().cpu()
    assert got.shape == ref.shape
    err = (got - ref).abs().max().item()
    assert err < 2e-3, f"log-mel max err {err:.3e}"      # fp32 FFT vs pocketfft, values ~ 10..25


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("N,T", [(32000, 99), (80000, 249)])
def test_embedding_fp32_class(impl, N, T):
    from diarizen_b200.embedding import EmbeddingModel
    sd, wav, masks = _setup(2, N, 4, T)
    ref = emb_forward(sd, wav, masks)
    m = EmbeddingModel(sd, precision="bf16x3", gemm_impl=impl)
    got = m.embed_windows(wav, masks).cpu()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item() / scale
    assert err < 1e-3, f"relative max err {err:.3e}"
    assert torch.allclose(got[0, 3], sd["resnet.seg_1.bias"], atol=1e-6), "all-zero mask must give seg_1.bias"


def test_embedding_fp16_cosine():
    """One-pass mode: embeddings are used through cosine distances, so the check is on the angle."""
    from diarizen_b200.embedding import EmbeddingModel
    sd, wav, masks = _setup(3, 80000, 4, 249)
    ref = emb_forward(sd, wav, masks)
    m = EmbeddingModel(sd, precision="fp16")
    got = m.embed_windows(wav, masks).cpu()
    cos = torch.nn.functional.cosine_similarity(got.reshape(-1, 256), ref.reshape(-1, 256), dim=1)
    assert cos.min().item() > 0.9995, f"min cosine {cos.min().item():.6f}"


def test_reference_call_convention():
    from diarizen_b200.embedding import EmbeddingModel
    sd, wav, masks = _setup(2, 32000, 1, 99)
    m = EmbeddingModel(sd, precision="bf16x3")
    out = m(wav[:, None, :], masks=masks[:, 0])
    assert out.shape == (2, 256) and out.dtype.name == "float32"
