"""Multi-channel segmentation model: host-side mirror of
`diarizen.models.eend.model_wavlm_conformer_mc.Model` (reference model_wavlm_conformer_mc.py:26-288).

The reference runs the WavLM front end and its first `F - 1` transformer layers on every microphone channel, mixes the
channels with a `CrossChannelAttention` module (diarizen/models/module/utils_mc.py:13-64) after the pre-processing and after
each of those layers, averages over channels before layer `F` and continues as the single-channel model
(diarizen/models/module/wav2vec2/components.py:1026-1070); the layer mix uses the channel means of the first F states
(model_wavlm_conformer_mc.py:241-247).  It returns the log-probabilities and the (B, F, T, C, C) channel-attention weights
(mean over heads) that the recipe uses to weight per-channel speaker embeddings (recipes/diar_ssl_mc/infer_avg.py:33-45).

Here the same computation is driven through two instances of the single-channel engine - one planned for B*C windows (the
channels are just more windows), one for B - run step range by step range (`dz_seg_run_steps`), with the fusion modules
(`dz_fusion_*`: tcgen05 projections + a warp-per-frame attention over channels) applied to the first engine's residual
stream in between, and `dz_channel_mean` carrying residual stream and layer-mix accumulator across.  All arithmetic is in
libdiarizen_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from .archs import NUM_SPEAKERS, SegArch
from .segmentation import SegmentationModel

FUSION_KEYS = ("linearQ.weight", "linearQ.bias", "linearK.weight", "linearK.bias", "linearV.weight", "linearV.bias",
               "linearO.weight", "linearO.bias", "ln_norm.weight", "ln_norm.bias")


class MCSegmentationModel:
    """state_dict: the reference MC model's (same keys as the single-channel model + `channel_fusions.<i>.*`)."""

    def __init__(self, arch: SegArch, state_dict: Dict[str, torch.Tensor], fusion_dim: int = 256, fusion_heads: int = 8,
                 fusion_layers: int = 4, precision: str = "fp16", device: Optional[torch.device] = None):
        if fusion_layers < 1 or fusion_layers > arch.num_layers:
            raise ValueError("channel_fusion_layers must be in [1, number of transformer layers]")
        self.arch, self.precision = arch, precision
        self.device = torch.device(device if device is not None else "cuda")
        self.fusion_layers = fusion_layers
        self._L = _lib.lib()
        # the channels are more windows for the first engine; the second continues on the channel mean
        self._per_channel = SegmentationModel(arch, state_dict, precision=precision, device=self.device)
        self._merged = SegmentationModel(arch, state_dict, precision=precision, device=self.device)
        self._mix_w = state_dict["weight_sum.weight"].detach().to("cpu", torch.float32).reshape(-1).tolist()
        prec = {"bf16": 1, "fp16": 2, "bf16x3": 3}[precision]
        self._fusions = []
        with torch.cuda.device(self.device):
            for i in range(fusion_layers):
                h = self._L.dz_fusion_create(arch.embed_dim, fusion_dim, fusion_heads, prec)
                if not h:
                    raise _lib.DzError(self._L.dz_last_error().decode())
                self._fusions.append(h)
                for k in FUSION_KEYS:
                    t = state_dict[f"channel_fusions.{i}.{k}"].detach().to("cpu", torch.float32).contiguous()
                    _lib.check(self._L.dz_fusion_set_param(h, k.encode(), C.c_void_p(t.data_ptr()), t.numel()))
                _lib.check(self._L.dz_fusion_finalize(h))

    def __del__(self):
        for h in getattr(self, "_fusions", []):
            self._L.dz_fusion_destroy(h)
        self._fusions = []

    def num_frames(self, num_samples: int) -> int:
        return self.arch.num_frames(num_samples)

    def hard(self, waveforms: torch.Tensor, want_logp: bool = True) -> Tuple[Optional[torch.Tensor], torch.Tensor, torch.Tensor]:
        """waveforms (B, C, N) -> (log-probs (B,T,11) or None, multilabel (B,T,4) uint8, attention (B, F, T, C, C) fp32), on the device."""
        if waveforms.dim() != 3:
            raise ValueError(f"Expected (batch, channel, sample), got {tuple(waveforms.shape)}")
        B, Cn, N = waveforms.shape
        if Cn > 8:
            raise ValueError("at most 8 channels")
        dev, L, F = self.device, self._L, self.fusion_layers
        w = waveforms.to(dev, torch.float32).reshape(B * Cn, N).contiguous()
        T = self.num_frames(N)
        A, M = self._per_channel, self._merged
        A.plan(B * Cn, N)
        M.plan(B, N)
        xa, ma, ba = A.tap_info("rep0"), A.tap_info("mix"), A.tap_info("xbf")
        xm, mm, bm = M.tap_info("rep0"), M.tap_info("mix"), M.tap_info("xbf")
        D, ld = xa["cols"], xa["ld"]
        post_norm = not self.arch.large     # the 16-bit copy of the stream is the next layer's input, and states are mixed when produced
        fp16 = 1 if self.precision == "fp16" else 0
        planes = 2 if self.precision == "bf16x3" else 1
        att = torch.empty((B, F, T, Cn, Cn), device=dev, dtype=torch.float32)
        att_l = torch.empty((B * T, Cn, Cn), device=dev, dtype=torch.float32)
        logp = torch.empty((B, T, self.arch.num_classes), device=dev, dtype=torch.float32) if want_logp else None
        ml = torch.empty((B, T, NUM_SPEAKERS), device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            first = 0
            for i in range(F):
                last = A.tap_info(f"rep{i}")["step"]
                A.run_steps(B * Cn, N, first, last, wav=w)
                first = last
                _lib.check(L.dz_fusion_forward(self._fusions[i], C.c_void_p(xa["ptr"]), B, Cn, T, ld,
                                               C.c_void_p(ba["ptr"]) if post_norm else None, ba["plane"], ba["ld"],
                                               C.c_void_p(ma["ptr"]) if post_norm else None, float(self._mix_w[i]),
                                               C.c_void_p(att_l.data_ptr()), st))
                att[:, i] = att_l.view(B, T, Cn, Cn)
            # channel mean of the stream and of the layer-mix accumulator -> the single-channel continuation
            _lib.check(L.dz_channel_mean(C.c_void_p(xa["ptr"]), C.c_void_p(xm["ptr"]), B, Cn, T, D, ld, st))
            _lib.check(L.dz_channel_mean(C.c_void_p(ma["ptr"]), C.c_void_p(mm["ptr"]), B, Cn, T, D, ld, st))
            if post_norm:
                _lib.check(L.dz_rows_to_planes(C.c_void_p(xm["ptr"]), B * T, D, ld, C.c_void_p(bm["ptr"]), bm["plane"], bm["ld"], planes, fp16, st))
            M.run_steps(B, N, M.tap_info(f"rep{F - 1}")["step"], -1, wav=None, logp=logp, ml=ml)
        self._keep = w
        return logp, ml, att

    def __call__(self, waveforms: torch.Tensor):
        logp, _, att = self.hard(waveforms)
        return logp, att
