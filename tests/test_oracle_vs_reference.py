"""Pins the oracle against the reference modules themselves (only where /root/reference is mounted)."""
import pytest
import torch

from diarizen_b200.archs import get_arch, init_state_dict, param_shapes
from oracle import ref_loader
from oracle.seg_oracle import seg_forward

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")


@pytest.mark.parametrize("name,N", [("tiny_base", 16000), ("tiny_large", 16000), ("wavlm_base_s80_md", 24000),
                                    ("wavlm_large_s80_md", 48000)])      # the benchmarked architecture
def test_seg_oracle_equals_reference(name, N):
    a = get_arch(name)
    m = ref_loader.RefSegModel(a).eval()
    ref_keys = {k: tuple(v.shape) for k, v in m.state_dict().items() if v.dtype.is_floating_point}
    assert ref_keys == param_shapes(a), "state_dict layout must match the reference Model"
    sd = init_state_dict(a, seed=4)
    m.load_state_dict(sd, strict=False)
    wav = 0.1 * torch.randn(2, N, generator=torch.Generator().manual_seed(9))
    with torch.inference_mode():
        ref = m(wav[:, None])
    assert (seg_forward(a, sd, wav) - ref).abs().max().item() < 1e-5


def test_reference_config_roundtrip():
    import sys
    sys.path.insert(0, ref_loader.REF)
    from diarizen.models.module.wavlm_config import get_config
    from diarizen_b200.archs import arch_from_reference_config
    for name in ("wavlm_base", "wavlm_large", "wavlm_base_s80_md", "wavlm_large_s80_md"):
        a = arch_from_reference_config(get_config(name), name)
        b = get_arch(name)
        assert (a.large, a.conv_channels, a.embed_dim, a.total_heads, a.heads, a.ffn) == \
               (b.large, b.conv_channels, b.embed_dim, b.total_heads, b.heads, b.ffn)
