"""TEST INFRASTRUCTURE ONLY - imports the *reference* modules when /root/reference is mounted.

The reference's `Model` wrapper cannot be imported (pyannote.core & co. are absent), so the
wrapper (model_wavlm_conformer.py:238-264) is re-assembled here from the importable reference parts:
`wav2vec2_model`, `ConformerEncoder`.  Used by scripts/make_golden.py and the oracle-pinning tests.
"""
from __future__ import annotations

import os
import sys

import torch
import torch.nn as nn

REF = os.environ.get("DIARIZEN_REF", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "diarizen"))


def _path():
    if REF not in sys.path:
        sys.path.insert(0, REF)


class RefSegModel(nn.Module):
    """Same submodule names as the reference Model => identical state_dict keys."""

    def __init__(self, arch):
        super().__init__()
        _path()
        from diarizen.models.module.conformer import ConformerEncoder
        from diarizen.models.module.wav2vec2.model import wav2vec2_model
        from diarizen_b200.archs import to_reference_config
        self.wavlm_model = wav2vec2_model(**to_reference_config(arch))
        self.weight_sum = nn.Linear(arch.num_layers + 1, 1, bias=False)
        self.proj = nn.Linear(arch.embed_dim, arch.head_dim_model)
        self.lnorm = nn.LayerNorm(arch.head_dim_model)
        self.conformer = ConformerEncoder(attention_in=arch.head_dim_model, ffn_hidden=arch.head_ffn,
                                          num_head=arch.head_heads, num_layer=arch.head_layers,
                                          kernel_size=arch.head_kernel, dropout=0.1, use_posi=False,
                                          output_activate_function=False)
        self.classifier = nn.Linear(arch.head_dim_model, arch.num_classes)
        self.activation = nn.LogSoftmax(dim=-1)

    def forward(self, waveforms):
        assert waveforms.dim() == 3
        waveforms = waveforms[:, 0, :]
        reps, _ = self.wavlm_model.extract_features(waveforms)
        x = torch.stack(reps, dim=-1)
        x = torch.squeeze(self.weight_sum(x), -1)
        x = self.lnorm(self.proj(x))
        x = self.conformer(x)
        return self.activation(self.classifier(x))
