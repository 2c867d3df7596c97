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
