"""Oracle vs the golden vectors produced FROM THE REFERENCE MODULES by scripts/make_golden.py (CPU; runs anywhere)."""
import os

import numpy as np
import pytest
import torch

from diarizen_b200.archs import get_arch, init_resnet_state_dict, init_state_dict
from oracle.emb_oracle import resnet_trunk, stats_pool
from oracle.seg_oracle import seg_forward

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["tiny_base", "tiny_large"])
def test_seg_oracle_matches_reference_output(name):
    z = np.load(os.path.join(G, f"seg_{name}.npz"))
    a = get_arch(name)
    y = seg_forward(a, init_state_dict(a, int(z["seed"])), torch.from_numpy(z["wav"]))
    assert np.abs(y.numpy() - z["logp"]).max() < 5e-6


def test_emb_oracle_matches_reference_output():
    z = np.load(os.path.join(G, "emb_resnet.npz"))
    sd = init_resnet_state_dict(int(z["seed"]))
    with torch.inference_mode():
        out = resnet_trunk(sd, torch.from_numpy(z["fbank"]))
        B, C, H, W = out.shape
        st = stats_pool(out.reshape(B, C * H, W), torch.from_numpy(z["masks"]))
        emb = torch.nn.functional.linear(st, sd["resnet.seg_1.weight"], sd["resnet.seg_1.bias"])
    assert np.abs(emb.numpy() - z["emb"]).max() < 1e-5


def test_stats_pool_matches_reference_module():
    z = np.load(os.path.join(G, "stats_pool.npz"))
    for k in ("one_speaker", "multi_speaker", "frame_mismatch", "all_zero"):
        x, w = torch.from_numpy(z[k + "_x"]), torch.from_numpy(z[k + "_w"])
        y = stats_pool(x, w[:, None] if w.dim() == 2 else w)
        y = y.squeeze(1) if w.dim() == 2 else y
        assert np.allclose(y.numpy(), z[k + "_y"], atol=1e-7), k
