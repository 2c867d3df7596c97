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

RESNET34_BLOCKS = (3, 4, 6, 3)
M_CHANNELS = 32
NUM_MEL = 80
EMB_DIM = 256


def resnet_param_shapes(prefix: str = "resnet.") -> Dict[str, tuple]:
    """State-dict layout of WeSpeakerResNet34 (keys as in the pyannote checkpoint: `resnet.*`)."""
    P: Dict[str, tuple] = {}

    def bn(name, c):
        for leaf in ("weight", "bias", "running_mean", "running_var"):
            P[f"{name}.{leaf}"] = (c,)

    P[prefix + "conv1.weight"] = (M_CHANNELS, 1, 3, 3)
    bn(prefix + "bn1", M_CHANNELS)
    cin = M_CHANNELS
    for li, nb in enumerate(RESNET34_BLOCKS):
        planes = M_CHANNELS * (2 ** li)
        for bi in range(nb):
            stride = (1 if li == 0 else 2) if bi == 0 else 1
            b = f"{prefix}layer{li + 1}.{bi}."
            P[b + "conv1.weight"] = (planes, cin, 3, 3)
            bn(b + "bn1", planes)
            P[b + "conv2.weight"] = (planes, planes, 3, 3)
            bn(b + "bn2", planes)
            if stride != 1 or cin != planes:
                P[b + "shortcut.0.weight"] = (planes, cin, 1, 1)
                bn(b + "shortcut.1", planes)
            cin = planes
    P[prefix + "seg_1.weight"] = (EMB_DIM, (NUM_MEL // 8) * M_CHANNELS * 8 * 2)
    P[prefix + "seg_1.bias"] = (EMB_DIM,)
    return P


def init_resnet_state_dict(seed: int = 0, prefix: str = "resnet.") -> Dict[str, torch.Tensor]:
    """Seeded random-init weights with the shapes above (no checkpoint is reachable offline)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shp in resnet_param_shapes(prefix).items():
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "running_mean":
            t = 0.05 * torch.randn(shp, generator=g)
        elif leaf == "running_var":
            t = 0.8 + 0.4 * torch.rand(shp, generator=g)
        elif ("bn" in name or "shortcut.1" in name) and leaf == "weight":
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif leaf == "bias":
            t = 0.05 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            t = torch.randn(shp, generator=g) * (2.0 / fan_in) ** 0.5   # He init
            if name.endswith("conv2.weight"):
                t = t * 0.35                                             # weak residual branches keep the trunk O(10)
        sd[name] = t.float().contiguous()
    return sd


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
