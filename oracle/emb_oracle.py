"""TEST INFRASTRUCTURE ONLY - fp32 torch restatement of the WeSpeaker ResNet34 embedding forward.

Follows (pa = pyannote-audio/pyannote/audio):
  pa/models/embedding/wespeaker/__init__.py:80-103  compute_fbank (x 2^15, kaldi.fbank, CMN over frames)
  pa/models/embedding/wespeaker/__init__.py:190-204 forward
  pa/models/embedding/wespeaker/resnet.py:139-144   BasicBlock.forward
  pa/models/embedding/wespeaker/resnet.py:344-376   ResNet.forward (trunk, TSTP pooling, seg_1)
  pa/models/blocks/pooling.py:44-131                StatsPool (weighted mean / unbiased std, nearest interpolation)
Pinned against the reference modules loaded by file path (tests/test_oracle_vs_reference.py) and against the
StatsPool known-answer tests of pyannote-audio/tests/test_stats_pool.py (tests/test_reference_kats.py).
`kaldi.fbank` itself is third-party (torchaudio, pinned 2.1.1 by the reference; 2.11 installed here): the installed
implementation is the oracle for the CUDA fbank kernel - parity unpinned by any reference test.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from diarizen_b200.archs import (EMB_DIM, M_CHANNELS, NUM_MEL, RESNET34_BLOCKS, init_resnet_state_dict,  # noqa: E402,F401
                                 resnet_param_shapes)


def compute_fbank(wav: torch.Tensor) -> torch.Tensor:
    """(B, N) -> (B, frames, 80): wespeaker/__init__.py:80-103."""
    import torchaudio.compliance.kaldi as kaldi
    feats = [kaldi.fbank(w[None] * (1 << 15), num_mel_bins=NUM_MEL, frame_length=25, frame_shift=10, dither=0.0,
                         sample_frequency=16000, window_type="hamming", use_energy=False) for w in wav]
    f = torch.stack(feats)
    return f - f.mean(dim=1, keepdim=True)


def _bn(x, sd, name):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], training=False, eps=1e-5)


def resnet_trunk(sd: Dict[str, torch.Tensor], fbank: torch.Tensor, prefix: str = "resnet.") -> torch.Tensor:
    """(B, frames, 80) -> (B, 256, 10, frames/8): resnet.py:358-365."""
    x = fbank.permute(0, 2, 1).unsqueeze(1)
    out = F.relu(_bn(F.conv2d(x, sd[prefix + "conv1.weight"], padding=1), sd, prefix + "bn1"))
    cin = M_CHANNELS
    for li, nb in enumerate(RESNET34_BLOCKS):
        planes = M_CHANNELS * (2 ** li)
        for bi in range(nb):
            stride = (1 if li == 0 else 2) if bi == 0 else 1
            b = f"{prefix}layer{li + 1}.{bi}."
            y = F.relu(_bn(F.conv2d(out, sd[b + "conv1.weight"], stride=stride, padding=1), sd, b + "bn1"))
            y = _bn(F.conv2d(y, sd[b + "conv2.weight"], padding=1), sd, b + "bn2")
            if (b + "shortcut.0.weight") in sd:
                sc = _bn(F.conv2d(out, sd[b + "shortcut.0.weight"], stride=stride), sd, b + "shortcut.1")
            else:
                sc = out
            out = F.relu(y + sc)
            cin = planes
    return out


def stats_pool(seq: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
    """seq (B, feat, frames), weights (B, S, T) -> (B, S, 2*feat): pooling.py:44-131."""
    Fr = seq.shape[-1]
    if weights.shape[-1] != Fr:
        weights = F.interpolate(weights, size=Fr, mode="nearest")
    outs = []
    for s in range(weights.shape[1]):
        w = weights[:, s].unsqueeze(1)
        v1 = w.sum(dim=2) + 1e-8
        mean = torch.sum(seq * w, dim=2) / v1
        dx2 = torch.square(seq - mean.unsqueeze(2))
        v2 = torch.square(w).sum(dim=2)
        var = torch.sum(dx2 * w, dim=2) / (v1 - v2 / v1 + 1e-8)
        outs.append(torch.cat([mean, torch.sqrt(var)], dim=1))
    return torch.stack(outs, dim=1)


@torch.inference_mode()
def emb_forward(sd: Dict[str, torch.Tensor], wav: torch.Tensor, masks: torch.Tensor, prefix: str = "resnet.",
                taps: Optional[dict] = None) -> torch.Tensor:
    """wav (B, N), masks (B, S, T) -> (B, S, 256).  The trunk does not depend on the mask, so it runs once per
    window and is pooled S times - the same arithmetic the reference performs S times."""
    fb = compute_fbank(wav)
    if taps is not None:
        taps["fbank"] = fb
    out = resnet_trunk(sd, fb, prefix)
    if taps is not None:
        taps["trunk"] = out
    B, C, H, W = out.shape
    stats = stats_pool(out.reshape(B, C * H, W), masks.float())
    if taps is not None:
        taps["stats"] = stats
    return F.linear(stats, sd[prefix + "seg_1.weight"], sd[prefix + "seg_1.bias"])
