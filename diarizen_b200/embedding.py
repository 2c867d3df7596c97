"""Host-side mirror of the reference speaker-embedding callable.

`EmbeddingModel` stands where `PyannoteAudioPretrainedSpeakerEmbedding` stands in `get_embeddings`
(pyannote-audio/pyannote/audio/pipelines/speaker_diarization.py:348-350, speaker_verification.py:612-705): same
attributes (`sample_rate`, `dimension`, `metric`, `min_num_samples`) and the same call convention
`embedding(waveforms (B,1,N), masks=(B,T)) -> np.ndarray (B,256)`.  `embed_windows` is the B200-native form the
pipeline uses: one trunk pass per window, pooled with all S local-speaker masks.
All arithmetic happens in libdiarizen_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib

EMB_DIM = 256


class EmbeddingModel:
    sample_rate = 16000
    dimension = EMB_DIM
    metric = "cosine"
    min_num_samples = 400   # one 25 ms fbank frame (reference finds the same value by bisection, speaker_verification.py:677-691)

    def __init__(self, state_dict: Dict[str, torch.Tensor], precision: str = "fp16", gemm_impl: str = "tc",
                 device: Optional[torch.device] = None, prefix: str = "resnet."):
        if not torch.cuda.is_available():
            raise RuntimeError("diarizen_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda")
        self._L = _lib.lib()
        self._ctor = (state_dict, precision, gemm_impl, prefix)
        prec = {"bf16": 1, "fp16": 2, "bf16x3": 3}[precision]
        with torch.cuda.device(self.device):
            self._h = self._L.dz_emb_create(prec, {"tc": 0, "simt": 1}[gemm_impl])
            if not self._h:
                raise _lib.DzError(self._L.dz_last_error().decode())
            for name, t in state_dict.items():
                if not name.startswith(prefix) or not t.dtype.is_floating_point:
                    continue
                key = "resnet." + name[len(prefix):]
                t = t.detach().to("cpu", torch.float32).contiguous()
                _lib.check(self._L.dz_emb_set_param(self._h, key.encode(), C.c_void_p(t.data_ptr()), t.numel()))
            _lib.check(self._L.dz_emb_finalize(self._h))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.dz_emb_destroy(h)

    def to(self, device) -> "EmbeddingModel":
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("diarizen_b200 runs on CUDA devices only (no CPU fallback)")
        if device == self.device or (device.index is None and self.device.type == "cuda"):
            return self
        sd, precision, gi, prefix = self._ctor
        return EmbeddingModel(sd, precision=precision, gemm_impl=gi, device=device, prefix=prefix)

    def embed_windows(self, waveforms: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
        """waveforms (B, N) fp32, masks (B, S, T) -> (B, S, 256) fp32 on the device."""
        w = waveforms.to(self.device, torch.float32).contiguous()
        m = masks.to(self.device, torch.float32).contiguous()
        B, N = w.shape
        _, S, T = m.shape
        out = torch.empty((B, S, EMB_DIM), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(self._L.dz_emb_forward(self._h, C.c_void_p(w.data_ptr()), C.c_void_p(m.data_ptr()), B, N, S, T,
                                              C.c_void_p(out.data_ptr()), C.c_void_p(st)))
        self._keep = (w, m)
        return out

    def __call__(self, waveforms: torch.Tensor, masks: Optional[torch.Tensor] = None) -> np.ndarray:
        """Reference call convention: (B,1,N) waveforms, (B,T) masks -> np.ndarray (B,256)."""
        w = waveforms[:, 0, :] if waveforms.dim() == 3 else waveforms
        if masks is None:
            masks = torch.ones((w.shape[0], 1), dtype=torch.float32)
        return self.embed_windows(w, masks[:, None, :]).squeeze(1).cpu().numpy()

    def fbank(self) -> torch.Tensor:
        """Debug: log-mel features (B, F, 80) of the last call, before mean subtraction."""
        n = self._L.dz_emb_tap_fbank(self._h, None, 0)
        _lib.check(n)
        out = torch.empty(n, device=self.device, dtype=torch.float32)
        _lib.check(self._L.dz_emb_tap_fbank(self._h, C.c_void_p(out.data_ptr()), n))
        return out.view(-1, 80, self._L.dz_emb_num_fbank_frames(self._keep[0].shape[1])).transpose(1, 2)   # the engine keeps them mel-major

    def profile(self):
        n = self._L.dz_emb_num_steps(self._h)
        ms = (C.c_float * n)()
        fl = (C.c_double * n)()
        names = C.create_string_buffer(n * 64)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(self._L.dz_emb_profile(self._h, ms, fl, names, 64, n, C.c_void_p(st)))
        return [(names.raw[i * 64:(i + 1) * 64].split(b"\0")[0].decode(), float(ms[i]), float(fl[i])) for i in range(n)]

    @property
    def last_launches(self) -> int:
        return self._L.dz_emb_last_launches(self._h)
