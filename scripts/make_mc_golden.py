"""tests/golden/seg_mc_<arch>.npz from the REFERENCE multi-channel model parts (oracle/ref_loader.RefSegModelMC = the reference's
wav2vec2_model.extract_features_mc + CrossChannelAttention + ConformerEncoder assembled as model_wavlm_conformer_mc.py:241-282):
seeded weights (stored), a 3-channel input, the log-probabilities and the channel-attention maps."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diarizen_b200.archs import get_arch, init_state_dict  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
FUSION = dict(fusion_dim=64, fusion_heads=4, fusion_layers=2)


def main():
    assert ref_loader.available()
    for name in ("tiny_base", "tiny_large"):
        a = get_arch(name)
        torch.manual_seed(21)
        m = ref_loader.RefSegModelMC(a, **FUSION).eval()
        sd = init_state_dict(a, seed=11)
        m.load_state_dict(sd, strict=False)
        g = torch.Generator().manual_seed(5)
        fus = {}
        for k, v in m.state_dict().items():
            if k.startswith("channel_fusions."):
                if k.endswith("ln_norm.weight"):
                    t = 0.5 + 0.5 * torch.rand(v.shape, generator=g)      # the reference initialises it to 1e-2: too small to test anything
                elif k.endswith("bias"):
                    t = 0.1 * torch.randn(v.shape, generator=g)
                else:
                    t = torch.randn(v.shape, generator=g) / v.shape[-1] ** 0.5
                fus[k] = t
        m.load_state_dict(fus, strict=False)
        wav = 0.1 * torch.randn(2, 3, 16000, generator=g)
        with torch.inference_mode():
            logp, att = m(wav)
        np.savez_compressed(os.path.join(OUT, f"seg_mc_{name}.npz"), wav=wav.numpy(), logp=logp.numpy(), att=att.numpy(), seed=11,
                            **{"fusion." + k: v.numpy() for k, v in fus.items()}, **{f"cfg_{k}": v for k, v in FUSION.items()})
        print(name, tuple(logp.shape), tuple(att.shape), float(att.std()))


if __name__ == "__main__":
    main()
