"""Host-side mirror of the reference segmentation model object.

`SegmentationModel` stands where the reference's `Model` (diarizen/models/eend/model_wavlm_conformer.py:25)
stands in `Inference.infer` (pyannote-audio/pyannote/audio/core/inference.py:213-226): it is called with a
`(batch, channel, sample)` float tensor and returns `(batch, frame, classes)` log-probabilities; `.hard()`
additionally returns the powerset->multilabel decoding the reference applies right after the forward.
All arithmetic happens in libdiarizen_b200.so (sm_100a CUDA); this class only holds tensors and pointers.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from .archs import NUM_SPEAKERS, SegArch, get_arch, init_state_dict, param_shapes


def _arch_to_c(a: SegArch) -> _lib.SegArchC:
    c = _lib.SegArchC()
    c.large = int(a.large)
    for i, ch in enumerate(a.conv_channels):
        c.conv_channels[i] = ch
    c.embed_dim, c.total_heads, c.num_layers = a.embed_dim, a.total_heads, a.num_layers
    for l in range(a.num_layers):
        c.num_heads[l] = len(a.heads[l])
        for j, h in enumerate(a.heads[l]):
            c.head_index[l][j] = h
        c.ffn[l] = a.ffn[l]
    c.head_dim_model, c.head_ffn, c.head_heads = a.head_dim_model, a.head_ffn, a.head_heads
    c.head_layers, c.head_kernel, c.num_classes = a.head_layers, a.head_kernel, a.num_classes
    return c


class SegmentationModel:
    """precision: "fp16" (one tensor-core pass over IEEE-half operands, fp32 accumulation; log-probs within 1e-2),
    "bf16" (one pass over bfloat16 operands; ~4x the rounding error of fp16) or "bf16x3" (split precision,
    fp32-class, log-probs within 1e-3)."""

    def __init__(self, arch: SegArch, state_dict: Dict[str, torch.Tensor], precision: str = "fp16",
                 gemm_impl: str = "tc", attn_impl: str = "tc", device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("diarizen_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.arch = arch
        self.device = torch.device(device if device is not None else "cuda")
        self.precision = precision
        self._ctor = (state_dict, gemm_impl, attn_impl)
        self._L = _lib.lib()
        prec = {"bf16": 1, "fp16": 2, "bf16x3": 3}[precision]
        with torch.cuda.device(self.device):
            self._h = self._L.dz_seg_create(C.byref(_arch_to_c(arch)), prec, {"tc": 0, "simt": 1}[gemm_impl],
                                            {"tc": 0, "simt": 1}[attn_impl])
            if not self._h:
                raise _lib.DzError(self._L.dz_last_error().decode())
            shapes = param_shapes(arch)
            for name, shp in shapes.items():
                if name not in state_dict:
                    raise KeyError(f"state_dict is missing '{name}'")
                t = state_dict[name].detach().to("cpu", torch.float32).contiguous()
                if tuple(t.shape) != tuple(shp):
                    raise ValueError(f"'{name}' has shape {tuple(t.shape)}, expected {tuple(shp)}")
                _lib.check(self._L.dz_seg_set_param(self._h, name.encode(), C.c_void_p(t.data_ptr()), t.numel()))
            _lib.check(self._L.dz_seg_finalize(self._h))

    @classmethod
    def random_init(cls, name: str, seed: int = 0, **kw) -> "SegmentationModel":
        a = get_arch(name)
        return cls(a, init_state_dict(a, seed), **kw)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.dz_seg_destroy(h)

    def num_frames(self, num_samples: int) -> int:
        return self.arch.num_frames(num_samples)

    # --- the attributes of the reference `Model` / `Inference` objects that pipeline callers read (inference.py:93,140) ---
    @property
    def model(self) -> "SegmentationModel":
        """`pipeline._segmentation.model` in the reference is the nn.Module inside the Inference wrapper; here both are this object"""
        return self

    @property
    def specifications(self):
        from .annotation import Specifications
        return Specifications(duration=getattr(self, "duration", 16.0), classes=tuple(f"speaker#{i + 1}" for i in range(NUM_SPEAKERS)),
                              powerset_max_classes=2)

    @property
    def _receptive_field(self):
        """SlidingWindow of the conv stack (core/model.py:180-195 evaluated with the reference's receptive_field.py helpers for
        kernels [10,3,3,3,3,2,2] / strides [5,2,2,2,2,2,2]): 400-sample frames every 320 samples, the centre the helpers return
        for frame 0 is sample 79 -> start = (79 - 199.5) / 16000 s"""
        from .annotation import SlidingWindow
        return SlidingWindow(start=(79 - (400 - 1) / 2) / 16000, duration=400 / 16000, step=320 / 16000)

    def eval(self) -> "SegmentationModel":
        return self

    def to(self, device) -> "SegmentationModel":
        """-> this model when it already lives on `device`, else a new engine built there from the same weights"""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("diarizen_b200 runs on CUDA devices only (no CPU fallback)")
        if device == self.device or (device.index is None and self.device.type == "cuda"):
            return self
        sd, gi, ai = self._ctor
        return SegmentationModel(self.arch, sd, precision=self.precision, gemm_impl=gi, attn_impl=ai, device=device)

    def _prep(self, waveforms: torch.Tensor) -> torch.Tensor:
        if waveforms.dim() == 3:
            waveforms = waveforms[:, 0, :]
        if waveforms.dim() != 2:
            raise ValueError(f"Expected (batch, channel, sample) or (batch, sample), got {tuple(waveforms.shape)}")
        return waveforms.to(self.device, torch.float32).contiguous()

    def hard(self, waveforms: torch.Tensor, want_logp: bool = True,
             ml_out: Optional[torch.Tensor] = None) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
        """-> (log-probs (B,T,11) fp32 or None, multilabel (B,T,4) uint8), both on the device.
        `ml_out`: optional preallocated contiguous uint8 (B,T,4) slice to write the decoding into."""
        w = self._prep(waveforms)
        B, N = w.shape
        T = self.num_frames(N)
        logp = torch.empty((B, T, self.arch.num_classes), device=self.device, dtype=torch.float32) if want_logp else None
        ml = ml_out if ml_out is not None else torch.empty((B, T, NUM_SPEAKERS), device=self.device, dtype=torch.uint8)
        assert ml.is_contiguous() and ml.dtype == torch.uint8 and tuple(ml.shape) == (B, T, NUM_SPEAKERS)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(self._L.dz_seg_forward(self._h, C.c_void_p(w.data_ptr()), B, N,
                                              C.c_void_p(logp.data_ptr()) if want_logp else None,
                                              C.c_void_p(ml.data_ptr()), C.c_void_p(st)))
        self._keep = w  # the engine's debug taps replay from this buffer
        return logp, ml

    def __call__(self, waveforms: torch.Tensor) -> torch.Tensor:
        return self.hard(waveforms)[0]

    def forward_host(self, wav_host: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """End-to-end call on HOST tensors: H2D, forward, D2H all inside the library (pinned staging)."""
        if wav_host.dim() == 3:
            wav_host = wav_host[:, 0, :]
        w = wav_host.to("cpu", torch.float32).contiguous()
        B, N = w.shape
        T = self.num_frames(N)
        pin = w.is_pinned()
        logp = torch.empty((B, T, self.arch.num_classes), dtype=torch.float32, pin_memory=pin)
        ml = torch.empty((B, T, NUM_SPEAKERS), dtype=torch.uint8, pin_memory=pin)
        with torch.cuda.device(self.device):
            _lib.check(self._L.dz_seg_forward_host(self._h, C.c_void_p(w.data_ptr()), B, N,
                                                   C.c_void_p(logp.data_ptr()), C.c_void_p(ml.data_ptr())))
        return logp, ml

    # --- step-range execution (used by the multi-channel model to interleave its fusion modules with the layers) ---
    def plan(self, B: int, N: int) -> None:
        with torch.cuda.device(self.device):
            _lib.check(self._L.dz_seg_plan(self._h, B, N))

    def tap_info(self, name: str) -> dict:
        ptr, plane, rows = C.c_void_p(), C.c_int64(), C.c_int64()
        cols, ld, step, is16 = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _lib.check(self._L.dz_seg_tap_info(self._h, name.encode(), C.byref(ptr), C.byref(plane), C.byref(rows), C.byref(cols), C.byref(ld),
                                           C.byref(step), C.byref(is16)))
        return {"ptr": ptr.value, "plane": plane.value, "rows": rows.value, "cols": cols.value, "ld": ld.value, "step": step.value, "is16": bool(is16.value)}

    def run_steps(self, B: int, N: int, first: int, last: int, wav: Optional[torch.Tensor] = None, logp: Optional[torch.Tensor] = None,
                  ml: Optional[torch.Tensor] = None) -> None:
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(self._L.dz_seg_run_steps(self._h, C.c_void_p(wav.data_ptr()) if wav is not None else None, B, N, first, last,
                                                C.c_void_p(logp.data_ptr()) if logp is not None else None,
                                                C.c_void_p(ml.data_ptr()) if ml is not None else None, C.c_void_p(st)))

    @property
    def num_steps(self) -> int:
        return self._L.dz_seg_num_steps(self._h)

    def tap(self, name: str) -> torch.Tensor:
        """Debug: fp32 copy of a named intermediate of the last forward (rows, C)."""
        with torch.cuda.device(self.device):
            n = self._L.dz_seg_tap(self._h, name.encode(), None, 0)
            _lib.check(n)
            out = torch.empty(n, device=self.device, dtype=torch.float32)
            _lib.check(self._L.dz_seg_tap(self._h, name.encode(), C.c_void_p(out.data_ptr()), n))
        return out

    def profile(self, waveforms: torch.Tensor):
        """Per-launch device times of one forward: list of (name, ms, algorithmic_flops, algorithmic_bytes)."""
        import ctypes as C
        w = self._prep(waveforms)
        B, N = w.shape
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            ms = (C.c_float * 4096)()
            n = self._L.dz_seg_profile(self._h, C.c_void_p(w.data_ptr()), B, N, ms, 4096, C.c_void_p(st))
            _lib.check(n)
            out = []
            buf = C.create_string_buffer(128)
            fl, by = C.c_double(), C.c_double()
            for i in range(n):
                _lib.check(self._L.dz_seg_step_info(self._h, i, buf, 128, C.byref(fl), C.byref(by)))
                out.append((buf.value.decode(), float(ms[i]), float(fl.value), float(by.value)))
        self._keep = w
        return out

    @property
    def last_launches(self) -> int:
        return self._L.dz_seg_last_launches(self._h)
