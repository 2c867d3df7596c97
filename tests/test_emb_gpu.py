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


@pytest.mark.parametrize("C,B,H,W", [(128, 2, 20, 400), (128, 3, 7, 133), (128, 1, 1, 16), (64, 2, 40, 300), (32, 2, 9, 131)])
@pytest.mark.parametrize("fp16", [1, 0])
@pytest.mark.parametrize("with_res", [True, False])
def test_conv3x3_kernels_vs_torch(C, B, H, W, fp16, with_res):
    """The stride-1 3x3 convolution kernels of the ResNet trunk (resident-weight kernel for C = 32 / 64, streamed-weight
    two-row kernel for C = 128: odd H, a ragged last pixel tile and a single-row image included) against torch.conv2d in
    float64 on the same 16-bit-rounded operands."""
    import ctypes as C_
    from diarizen_b200 import _lib
    g = torch.Generator().manual_seed(C + H + W)
    dt = torch.float16 if fp16 else torch.bfloat16
    x = torch.randn(B, H, W, C, generator=g).to(dt)
    w = (torch.randn(C, C, 3, 3, generator=g) / (3 * C ** 0.5)).to(dt)      # [out][in][kh][kw]
    bias = torch.randn(C, generator=g)
    res = torch.randn(B, H, W, C, generator=g).to(dt)
    krun = (3 * C + 63) // 64 * 64
    ldw = 3 * krun
    wl = torch.zeros(C, 3, krun, dtype=dt)
    wl[:, :, :3 * C] = w.permute(0, 2, 3, 1).reshape(C, 3, 3 * C)           # (kh, kw, ci) at kh * krun + kw * C + ci
    pad = lambda t: torch.nn.functional.pad(t, (0, 0, 1, 1)).contiguous().cuda()   # zero border columns
    xin, rin = pad(x), pad(res)
    out = torch.full((B, H, W + 2, C), 7.0, dtype=dt, device="cuda")
    wd, bd = wl.reshape(C, ldw).contiguous().cuda(), bias.cuda()
    vp = C_.c_void_p
    _lib.check(_lib.lib().dz_conv3x3(vp(xin.data_ptr()), vp(out.data_ptr()), vp(rin.data_ptr()) if with_res else None, vp(wd.data_ptr()), ldw,
                                     vp(bd.data_ptr()), B, H, W, C, 1, fp16, None))
    ref = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w.double(), bias.double(), padding=1).permute(0, 2, 3, 1)
    if with_res:
        ref = ref + res.double()
    ref = torch.relu(ref)
    got = out[:, :, 1:W + 1].double().cpu()
    err = (got - ref).abs().max().item() / (ref.abs().max().item() + 1e-6)
    assert err < (1.5e-3 if fp16 else 1e-2), f"rel err {err:.3e}"           # output rounding of the 16-bit format dominates
    assert (out[:, :, 0] == 7.0).all() and (out[:, :, W + 1] == 7.0).all(), "border columns must not be written"
