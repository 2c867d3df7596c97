"""Architecture tables for the segmentation network (WavLM front end + Conformer head).

The numbers restate the *shapes* the reference instantiates
(reference: diarizen/models/module/wavlm_config.py:38-239 for the four WavLM
variants, diarizen/models/eend/model_wavlm_conformer.py:26-76 for the head).
They are data, not code: every kernel launch in the engine is sized from a
`SegArch`, and `param_shapes()` enumerates the reference state_dict layout
("Checkpoint / resume") so that a `pytorch_model.bin`
written by the reference loads unchanged.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

CONV_KERNELS = (10, 3, 3, 3, 3, 2, 2)
CONV_STRIDES = (5, 2, 2, 2, 2, 2, 2)
HEAD_DIM = 64
NUM_BUCKETS = 320
MAX_DISTANCE = 800
POS_CONV_KERNEL = 128
POS_CONV_GROUPS = 16
NUM_POWERSET = 11  # 4 speakers, <=2 simultaneous: 1 + 4 + 6
NUM_SPEAKERS = 4


@dataclass(frozen=True)
class SegArch:
    name: str
    large: bool                      # layer_norm extractor + pre-norm encoder + waveform norm
    conv_channels: Tuple[int, ...]   # 7 conv output widths
    embed_dim: int
    total_heads: int
    heads: Tuple[Tuple[int, ...], ...]   # remaining head indices per layer (empty = no attention)
    ffn: Tuple[int, ...]                 # FFN width per layer (0 = no FFN)
    # conformer head
    head_dim_model: int = 256
    head_ffn: int = 1024
    head_heads: int = 4
    head_layers: int = 4
    head_kernel: int = 31
    num_classes: int = NUM_POWERSET

    @property
    def num_layers(self) -> int:
        return len(self.ffn)

    def num_frames(self, num_samples: int) -> int:
        """reference: model_wavlm_conformer.py:98-124 (multi_conv_num_frames)."""
        n = num_samples
        for k, s in zip(CONV_KERNELS, CONV_STRIDES):
            n = (n - k) // s + 1
        return n

    def conv_frames(self, num_samples: int) -> List[int]:
        out, n = [], num_samples
        for k, s in zip(CONV_KERNELS, CONV_STRIDES):
            n = (n - k) // s + 1
            out.append(n)
        return out


def _full(n_layers: int, n_heads: int):
    return tuple(tuple(range(n_heads)) for _ in range(n_layers))


_BASE_S80_HEADS = ((1, 6), (5, 7, 8), (0, 3, 9), (0, 1, 4, 8, 11), (6, 8), (0,), (7, 8, 10, 11),
                   (0, 1, 4, 8), (), (), (4, 7), (5,))
_BASE_S80_FFN = (666, 660, 649, 1080, 237, 299, 437, 573, 53, 80, 211, 334)
_LARGE_S80_HEADS = ((1, 2, 4, 5, 6), (9, 10, 14), (0, 1, 2, 4, 5, 7), (1, 4, 7, 12, 13, 14),
                    (0, 2, 3, 4, 13), (1, 7, 13, 14, 15), (11, 13, 15), (2, 3, 4, 8, 15), (2, 5, 6, 15),
                    (), (0, 1), (1, 3, 5, 12), (), (4, 7, 11), (6, 9), (11,), (), (), (14,), (5, 15),
                    (0, 2, 8, 11, 13, 15), (0, 1, 3, 4, 5, 6, 7, 10, 13),
                    (0, 1, 3, 6, 7, 9, 10, 11, 12, 14), (1, 2, 3, 4, 7, 13, 14, 15))
_LARGE_S80_FFN = (1092, 925, 759, 646, 745, 615, 684, 958, 286, 294, 406, 377, 463, 542, 298, 236,
                  96, 104, 134, 211, 473, 1011, 1770, 1316)

ARCHS: Dict[str, SegArch] = {
    "wavlm_base": SegArch("wavlm_base", False, (512,) * 7, 768, 12, _full(12, 12), (3072,) * 12),
    "wavlm_large": SegArch("wavlm_large", True, (512,) * 7, 1024, 16, _full(24, 16), (4096,) * 24),
    "wavlm_base_s80_md": SegArch("wavlm_base_s80_md", False, (90, 161, 173, 181, 351, 155, 137), 768, 12,
                                 _BASE_S80_HEADS, _BASE_S80_FFN),
    "wavlm_large_s80_md": SegArch("wavlm_large_s80_md", True, (512, 153, 224, 255, 302, 368, 211), 1024, 16,
                                  _LARGE_S80_HEADS, _LARGE_S80_FFN),
    # tiny shapes for fixtures / CPU tests (not a reference variant)
    "tiny_base": SegArch("tiny_base", False, (24, 20, 28, 20, 36, 20, 28), 128, 2, ((0, 1), (), (1,)), (72, 40, 0),
                         head_dim_model=64, head_ffn=96, head_heads=1, head_layers=1),
    # five small post-norm layers: room for the four channel-fusion modules of the multi-channel recipe
    "tiny_base_mc": SegArch("tiny_base_mc", False, (24, 20, 28, 20, 36, 20, 28), 128, 2, ((0, 1), (1,), (0,), (0, 1), (1,)),
                            (72, 40, 56, 40, 48), head_dim_model=64, head_ffn=96, head_heads=1, head_layers=1),
    "tiny_large": SegArch("tiny_large", True, (32, 20, 28, 20, 36, 20, 28), 128, 2, ((1,), (), (0, 1)), (72, 0, 40),
                          head_dim_model=64, head_ffn=96, head_heads=1, head_layers=2),
}


def get_arch(name: str) -> SegArch:
    key = name.lower()
    if key not in ARCHS:
        raise ValueError(f"Unknown config name '{name}'. Available options: {', '.join(ARCHS)}.")
    return ARCHS[key]


def arch_from_reference_config(cfg: dict, name: str = "ckpt") -> SegArch:
    """Build a SegArch from a reference `{config: ...}` checkpoint dict
    (reference: model_wavlm_conformer.py:209-221, recipes/diar_ssl_pruning/apply_pruning.py:120-126)."""
    for k, v in cfg.items():
        if "prune" in k and v is not False:
            raise ValueError(f"Pruning must be disabled. Found: {k}={v}")
    convs = cfg["extractor_conv_layer_config"]
    if tuple(c[1] for c in convs) != CONV_KERNELS or tuple(c[2] for c in convs) != CONV_STRIDES:
        raise ValueError("unsupported conv geometry")
    if cfg.get("extractor_conv_bias", False):
        raise ValueError("conv bias not supported")
    large = cfg["extractor_mode"] == "layer_norm"
    if large != bool(cfg["encoder_layer_norm_first"]) or large != bool(cfg.get("normalize_waveform", False)):
        raise ValueError("unsupported norm combination")
    n = cfg["encoder_num_layers"]
    heads = tuple(tuple(cfg["encoder_remaining_heads"][i]) if cfg["encoder_use_attention"][i] else ()
                  for i in range(n))
    ffn = tuple(cfg["encoder_ff_interm_features"][i] if cfg["encoder_use_feed_forward"][i] else 0
                for i in range(n))
    return SegArch(name, large, tuple(c[0] for c in convs), cfg["encoder_embed_dim"],
                   cfg["encoder_total_num_heads"][0], heads, ffn)


def to_reference_config(a: SegArch) -> dict:
    """Inverse of arch_from_reference_config: kwargs for the reference's wav2vec2_model()."""
    n = a.num_layers
    return {
        "extractor_mode": "layer_norm" if a.large else "group_norm",
        "extractor_conv_layer_config": [(c, k, s) for c, k, s in zip(a.conv_channels, CONV_KERNELS, CONV_STRIDES)],
        "extractor_conv_bias": False,
        "encoder_embed_dim": a.embed_dim,
        "encoder_projection_dropout": 0.1,
        "encoder_pos_conv_kernel": POS_CONV_KERNEL,
        "encoder_pos_conv_groups": POS_CONV_GROUPS,
        "encoder_num_layers": n,
        "encoder_use_attention": [len(h) > 0 for h in a.heads],
        "encoder_use_feed_forward": [f > 0 for f in a.ffn],
        "encoder_total_num_heads": [a.total_heads] * n,
        "encoder_remaining_heads": [list(h) for h in a.heads],
        "encoder_num_buckets": NUM_BUCKETS,
        "encoder_max_distance": MAX_DISTANCE,
        "encoder_attention_dropout": 0.1,
        "encoder_ff_interm_features": [max(f, 1) for f in a.ffn],
        "encoder_ff_interm_dropout": 0.0,
        "encoder_dropout": 0.1,
        "encoder_layer_norm_first": a.large,
        "encoder_layer_drop": 0.05,
        "aux_num_out": None,
        "normalize_waveform": a.large,
        "extractor_prune_conv_channels": False,
        "encoder_prune_attention_heads": False,
        "encoder_prune_attention_layer": False,
        "encoder_prune_feed_forward_intermediate": False,
        "encoder_prune_feed_forward_layer": False,
    }


def param_shapes(a: SegArch) -> Dict[str, Tuple[int, ...]]:
    """Names and shapes of the reference `Model.state_dict()` (floating tensors only)."""
    P: Dict[str, Tuple[int, ...]] = {}
    fe = "wavlm_model.feature_extractor."
    cin = 1
    for i, (c, k) in enumerate(zip(a.conv_channels, CONV_KERNELS)):
        if a.large or i == 0:
            P[f"{fe}conv_layers.{i}.layer_norm.weight"] = (c,)
            P[f"{fe}conv_layers.{i}.layer_norm.bias"] = (c,)
        P[f"{fe}conv_layers.{i}.conv.weight"] = (c, cin, k)
        cin = c
    P[f"{fe}dummy_weight"] = (cin,)
    en = "wavlm_model.encoder."
    D = a.embed_dim
    P[f"{en}feature_projection.layer_norm.weight"] = (cin,)
    P[f"{en}feature_projection.layer_norm.bias"] = (cin,)
    P[f"{en}feature_projection.projection.weight"] = (D, cin)
    P[f"{en}feature_projection.projection.bias"] = (D,)
    tr = en + "transformer."
    P[f"{tr}pos_conv_embed.conv.bias"] = (D,)
    P[f"{tr}pos_conv_embed.conv.parametrizations.weight.original0"] = (1, 1, POS_CONV_KERNEL)
    P[f"{tr}pos_conv_embed.conv.parametrizations.weight.original1"] = (D, D // POS_CONV_GROUPS, POS_CONV_KERNEL)
    P[f"{tr}layer_norm.weight"] = (D,)
    P[f"{tr}layer_norm.bias"] = (D,)
    for l in range(a.num_layers):
        L = f"{tr}layers.{l}."
        h = len(a.heads[l])
        if h:
            at = L + "attention."
            P[at + "gru_rel_pos_const"] = (1, a.total_heads, 1, 1)
            for nm in ("k_proj", "v_proj", "q_proj"):
                P[at + nm + ".weight"] = (h * HEAD_DIM, D)
                P[at + nm + ".bias"] = (h * HEAD_DIM,)
            P[at + "out_proj.weight"] = (D, h * HEAD_DIM)
            P[at + "out_proj.bias"] = (D,)
            if l == 0:
                P[at + "rel_attn_embed.weight"] = (NUM_BUCKETS, a.total_heads)
            P[at + "gru_rel_pos_linear.weight"] = (8, D // a.total_heads)
            P[at + "gru_rel_pos_linear.bias"] = (8,)
        P[L + "layer_norm.weight"] = (D,)
        P[L + "layer_norm.bias"] = (D,)
        if a.ffn[l]:
            P[L + "feed_forward.intermediate_dense.weight"] = (a.ffn[l], D)
            P[L + "feed_forward.intermediate_dense.bias"] = (a.ffn[l],)
            P[L + "feed_forward.output_dense.weight"] = (D, a.ffn[l])
            P[L + "feed_forward.output_dense.bias"] = (D,)
        P[L + "final_layer_norm.weight"] = (D,)
        P[L + "final_layer_norm.bias"] = (D,)
    A, F = a.head_dim_model, a.head_ffn
    P["weight_sum.weight"] = (1, a.num_layers + 1)
    P["proj.weight"] = (A, D)
    P["proj.bias"] = (A,)
    P["lnorm.weight"] = (A,)
    P["lnorm.bias"] = (A,)
    for i in range(a.head_layers):
        C = f"conformer.conformer_layer.{i}."
        for ff in ("ffn1.", "ffn2."):
            P[C + ff + "ln_norm.weight"] = (A,)
            P[C + ff + "ln_norm.bias"] = (A,)
            P[C + ff + "w_1.weight"] = (F, A)
            P[C + ff + "w_1.bias"] = (F,)
            P[C + ff + "w_2.weight"] = (A, F)
            P[C + ff + "w_2.bias"] = (A,)
        P[C + "mha.ln_norm.weight"] = (A,)
        P[C + "mha.ln_norm.bias"] = (A,)
        for nm in ("linearQ", "linearK", "linearV", "linearO"):
            P[C + f"mha.mha.{nm}.weight"] = (A, A)
            P[C + f"mha.mha.{nm}.bias"] = (A,)
        P[C + "conv.ln_norm.weight"] = (A,)
        P[C + "conv.ln_norm.bias"] = (A,)
        P[C + "conv.pointwise_conv1.weight"] = (2 * A, A, 1)
        P[C + "conv.pointwise_conv1.bias"] = (2 * A,)
        P[C + "conv.depthwise_conv.weight"] = (A, 1, a.head_kernel)
        P[C + "conv.depthwise_conv.bias"] = (A,)
        for nm in ("weight", "bias", "running_mean", "running_var"):
            P[C + "conv.bn_norm." + nm] = (A,)
        P[C + "conv.pointwise_conv2.weight"] = (A, A, 1)
        P[C + "conv.pointwise_conv2.bias"] = (A,)
        P[C + "ln_norm.weight"] = (A,)
        P[C + "ln_norm.bias"] = (A,)
    P["classifier.weight"] = (a.num_classes, A)
    P["classifier.bias"] = (a.num_classes,)
    return P


def init_state_dict(a: SegArch, seed: int = 0, classifier_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights of the named architecture (no checkpoint is reachable offline).

    Matrices ~ U(-b, b) with b = 1/sqrt(fan_in) (the torch default for Linear/Conv), LayerNorm/BN affine
    perturbed around (1, 0) so that gamma/beta wiring errors are visible in parity tests, BN running
    stats around (0, 1).  `classifier_gain` widens the powerset margin for hard-decision parity runs
    ("Hard parts").
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shp in param_shapes(a).items():
        leaf = name.rsplit(".", 1)[-1]
        if name.endswith("dummy_weight"):
            t = 1.0 + 0.05 * torch.randn(shp, generator=g)
        elif name.endswith("gru_rel_pos_const"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif name.endswith("original0"):
            t = 0.5 + torch.rand(shp, generator=g)
        elif name.endswith("rel_attn_embed.weight"):
            t = 0.5 * torch.randn(shp, generator=g)
        elif name.endswith("weight_sum.weight"):
            t = (1.0 + 0.3 * torch.randn(shp, generator=g)) / shp[1]
        elif "norm" in name and leaf == "weight":
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif "norm" in name and leaf == "bias":
            t = 0.05 * torch.randn(shp, generator=g)
        elif leaf == "running_mean":
            t = 0.05 * torch.randn(shp, generator=g)
        elif leaf == "running_var":
            t = 0.8 + 0.4 * torch.rand(shp, generator=g)
        elif leaf == "bias":
            t = 0.02 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            b = 1.0 / (fan_in ** 0.5)
            t = (torch.rand(shp, generator=g) * 2 - 1) * b
            if name == "classifier.weight":
                t = t * classifier_gain
        sd[name] = t.to(torch.float32).contiguous()
    return sd


# ---------------------------------------------------------------------------------------------------------
# WeSpeaker ResNet34 embedding network (reference: pyannote-audio/pyannote/audio/models/embedding/wespeaker/resnet.py:213-260)
# ---------------------------------------------------------------------------------------------------------
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


