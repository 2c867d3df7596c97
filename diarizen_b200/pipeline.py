"""`DiariZenPipeline`: drop-in for the reference class of the same name (diarizen/pipelines/inference.py:26-192).

Same constructor / `from_pretrained` / `__call__` signatures, same result protocol (`itertracks(yield_label=True)`,
`.uri`, `.to_rttm()`), same config.toml schema.  What differs is where the work happens: every stage between the
decoded waveform and the final (frames x speakers) decision matrix runs as sm_100a CUDA inside libdiarizen_b200.so -
sliding-window segmentation, median filter, speaker counting, embedding masks, ResNet34 embeddings, float64 distance
matrix + centroid linkage, constrained assignment, cluster-wise reconstruction and top-count selection.  The recording
stays on the device from the first window to the decision matrix; only the embeddings (C x 4 x 256) and a few
per-chunk counters visit the host, where the reference's own selection logic is applied to them.

Long recordings shard by window range across the ranks of a torch.distributed (NCCL) job: every rank runs the two
network forwards on its windows, one all-gather collects the binarised segmentations (uint8) and the embeddings, and
rank 0 clusters and reconstructs (`DiariZenPipeline.__call__` returns the Annotation on rank 0, None elsewhere).
"""
from __future__ import annotations

import ctypes as C
import io
import math
import os
from pathlib import Path
from typing import Any, Dict, Optional, Union

import numpy as np
import torch

from . import _lib
from .annotation import Annotation, Segment
from .archs import SegArch, arch_from_reference_config, get_arch, init_state_dict
from .clustering import AgglomerativeClustering, VBxClustering
from .embedding import EmbeddingModel
from .segmentation import SegmentationModel
from .sharding import gather_records, window_ranges

SR = 16000
FRAME_DURATION = 400 / SR    # receptive field of the conv stack (model_wavlm_conformer.py:126-176)
FRAME_STEP = 320 / SR
vp = C.c_void_p


def _closest_frame(t: float) -> int:
    """pyannote.core SlidingWindow.closest_frame with start=0."""
    return int(np.rint((t - 0.5 * FRAME_DURATION) / FRAME_STEP))


def _to_16k(w: torch.Tensor, sr: int) -> torch.Tensor:
    """Other sample rates go through the same resampler the reference uses (pyannote `Audio` -> torchaudio.functional.resample,
    pa/core/io.py:187-221); decode-time preprocessing on the host, before the recording is uploaded."""
    if sr == SR:
        return w
    try:
        import torchaudio.functional as AF
    except Exception as e:   # pragma: no cover
        raise ValueError(f"input is {sr} Hz and torchaudio is not available to resample it to {SR} Hz") from e
    return AF.resample(w[None], sr, SR)[0].contiguous()


def _read_wav(src):
    """RIFF/WAVE decode on the host (stands in for torchaudio.load, which needs a codec backend that is not in this image):
    PCM 8 / 16 / 24 / 32 bit, IEEE float 32 / 64, plain and WAVE_FORMAT_EXTENSIBLE headers.  -> ((frames, channels) float32 in
    [-1, 1], sample rate).  Compressed formats (flac, mp3, ...) are not decoded here: pass {"waveform", "sample_rate"}."""
    import struct
    raw = src.getvalue() if isinstance(src, io.BytesIO) else open(os.fspath(src), "rb").read()
    if len(raw) < 12 or raw[:4] != b"RIFF" or raw[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file (compressed audio must be decoded by the caller and passed as {'waveform', 'sample_rate'})")
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(raw):
        cid, size = raw[pos:pos + 4], struct.unpack("<I", raw[pos + 4:pos + 8])[0]
        body = raw[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = body
        elif cid == b"data":
            data = body
            break
        pos += 8 + size + (size & 1)
    if fmt is None or data is None:
        raise ValueError("WAVE file without fmt / data chunk")
    tag, nch, sr, _, align, bits = struct.unpack("<HHIIHH", fmt[:16])
    if tag == 0xFFFE and len(fmt) >= 26:          # WAVE_FORMAT_EXTENSIBLE: the real tag is the first word of the sub-format GUID
        tag = struct.unpack("<H", fmt[24:26])[0]
    width = bits // 8
    n = len(data) // (width * nch) * nch
    if tag == 1 and width == 2:
        x = np.frombuffer(data, dtype="<i2", count=n).astype(np.float32) / 32768.0
    elif tag == 1 and width == 4:
        x = np.frombuffer(data, dtype="<i4", count=n).astype(np.float32) / 2147483648.0
    elif tag == 1 and width == 3:
        b3 = np.frombuffer(data, dtype=np.uint8, count=3 * n).reshape(-1, 3).astype(np.int32)
        v = b3[:, 0] | (b3[:, 1] << 8) | (b3[:, 2] << 16)
        x = (v - ((v & 0x800000) << 1)).astype(np.float32) / 8388608.0
    elif tag == 1 and width == 1:
        x = (np.frombuffer(data, dtype=np.uint8, count=n).astype(np.float32) - 128.0) / 128.0
    elif tag == 3 and width in (4, 8):
        x = np.frombuffer(data, dtype="<f4" if width == 4 else "<f8", count=n).astype(np.float32)
    else:
        raise ValueError(f"unsupported WAVE encoding (format tag {tag}, {bits} bits)")
    return x.reshape(-1, nch), sr


def load_waveform(in_wav) -> torch.Tensor:
    """-> mono (N,) float32 in [-1, 1], 16 kHz.  Stands in for `torchaudio.load(in_wav)[0][0]` (inference.py:127-128)."""
    if isinstance(in_wav, dict):
        w = in_wav["waveform"] if "waveform" in in_wav else None
        if w is None:
            return load_waveform(in_wav["audio"])
        w = torch.as_tensor(w, dtype=torch.float32)
        w = w[0] if w.dim() == 2 else w
        return _to_16k(w, int(in_wav.get("sample_rate", SR)))
    if isinstance(in_wav, (str, os.PathLike, io.BytesIO)):
        x, sr = _read_wav(in_wav)
        return _to_16k(torch.from_numpy(x[:, 0].copy()), sr)   # force channel 0 (inference.py:128)
    raise TypeError(f"input must be either a str, BytesIO or a ProtocolFile; there was {type(in_wav)}")


def _load_toml(path: Path) -> dict:
    import tomllib
    with open(path, "rb") as f:
        return tomllib.load(f)


class DiariZenPipeline:
    def __init__(self, diarizen_hub, embedding_model, config_parse: Optional[Dict[str, Any]] = None,
                 rttm_out_dir: Optional[str] = None, *, precision: str = "fp16", device=None,
                 _seg: Optional[SegmentationModel] = None, _emb: Optional[EmbeddingModel] = None, _config: Optional[dict] = None,
                 segmentation: Optional[list] = None, _mc=None):
        if _config is None:
            diarizen_hub = Path(diarizen_hub)
            config = _load_toml(diarizen_hub / "config.toml")
        else:
            config = _config
        if config_parse is not None:
            print("Overriding with parsed config.")
            config["inference"]["args"] = config_parse["inference"]["args"]
            config["clustering"]["args"] = config_parse["clustering"]["args"]
        inf, clu = config["inference"]["args"], config["clustering"]["args"]
        self.config = config
        self.device = torch.device(device if device is not None else "cuda")
        self.seg_duration = float(inf["seg_duration"])
        self.segmentation_step = float(inf["segmentation_step"])
        self.segmentation_batch_size = int(inf["batch_size"])
        self.embedding_batch_size = int(inf["batch_size"])
        # Windows per engine call.  The reference's batch_size only trades memory for speed (results are identical for
        # any value); on a 180 GB part larger batches amortise wave tails of the persistent kernels.
        self.engine_windows = max(self.segmentation_batch_size, int(os.environ.get("DZ_ENGINE_WINDOWS", "96")))
        self.engine_emb_windows = max(1, self.embedding_batch_size // 4, int(os.environ.get("DZ_ENGINE_EMB_WINDOWS", "32")))
        self.embedding_exclude_overlap = True
        self.apply_median_filtering = bool(inf["apply_median_filtering"])
        self.min_speakers = clu["min_speakers"]
        self.max_speakers = clu["max_speakers"]
        if clu["method"] == "AgglomerativeClustering":
            self.clustering = AgglomerativeClustering(metric="cosine", device=self.device)
            self.clustering.method = "centroid"
            self.clustering.min_cluster_size = clu["min_cluster_size"]
            self.clustering.threshold = clu["ahc_threshold"]
        elif clu["method"] == "VBxClustering":
            # inference.py:72-83
            self.clustering = VBxClustering(metric="cosine", device=self.device)
            self.clustering.ahc_criterion = clu["ahc_criterion"]
            self.clustering.ahc_threshold = clu["ahc_threshold"]
            self.clustering.Fa = clu["Fa"]
            self.clustering.Fb = clu["Fb"]
            self.clustering.plda_dir = str(clu["plda_dir"]) if "plda_dir" in clu else str(Path(diarizen_hub) / "plda")
            self.clustering.lda_dim = clu["lda_dim"]
            self.clustering.maxIters = clu["max_iters"]
        else:
            raise ValueError(f"Unsupported clustering method: {clu['method']}")
        if _seg is None:
            margs = config["model"]["args"]
            src = margs.get("wavlm_src", "wavlm_base")
            if segmentation:
                # checkpoint-averaged inference (recipes/*/infer_avg.py:292-303, ckpt_utils.py:16-43): a list of
                # checkpoint paths / {'bin_path': ...} records is averaged key-wise at load time
                from .checkpoints import average_checkpoints
                sd = average_checkpoints(segmentation) if len(segmentation) > 1 else torch.load(
                    str(segmentation[0]["bin_path"] if isinstance(segmentation[0], dict) else segmentation[0]), map_location="cpu")
            else:
                sd = torch.load(str(diarizen_hub / "pytorch_model.bin"), map_location="cpu")
            if os.path.isfile(src):
                ck = torch.load(src, map_location="cpu")
                arch = arch_from_reference_config(ck["config"], name=Path(src).stem)
            else:
                arch = get_arch(src)
            # conformer head hyper-parameters of Model.__init__ (model_wavlm_conformer.py:26-48), when the config names them
            import dataclasses
            head = {ours: int(margs[theirs]) for theirs, ours in (("attention_in", "head_dim_model"), ("ffn_hidden", "head_ffn"), ("num_head", "head_heads"),
                                                                    ("num_layer", "head_layers"), ("kernel_size", "head_kernel")) if theirs in margs}
            if head:
                arch = dataclasses.replace(arch, **head)
            if "channel_fusion_layers" in margs or str(config["model"].get("path", "")).endswith("model_wavlm_conformer_mc.Model"):
                # multi-channel model (model_wavlm_conformer_mc.py:26-57): its single-channel continuation engine doubles as
                # `_segmentation` (frame geometry, specifications); every input goes through `diarize_session`
                from .segmentation_mc import MCSegmentationModel
                self._segmentation_mc = MCSegmentationModel(arch, sd, fusion_dim=int(margs.get("channel_fusion_dim", 768)),
                                                            fusion_heads=int(margs.get("channel_fusion_heads", 4)),
                                                            fusion_layers=int(margs.get("channel_fusion_layers", 4)),
                                                            precision=precision, device=self.device)
                _seg = self._segmentation_mc._merged
            else:
                _seg = SegmentationModel(arch, sd, precision=precision, device=self.device)
        if _emb is None:
            esd = torch.load(str(embedding_model), map_location="cpu")
            esd = esd.get("state_dict", esd)
            _emb = EmbeddingModel(esd, precision=precision, device=self.device)
        _seg.duration = self.seg_duration
        self._segmentation = _seg
        self._embedding = _emb
        assert self._segmentation.model.specifications.powerset is True   # inference.py:93
        if rttm_out_dir is not None:
            os.makedirs(rttm_out_dir, exist_ok=True)
        self.rttm_out_dir = rttm_out_dir
        self._L = _lib.lib()
        self.last = {}
        self._planned = {}
        if not hasattr(self, "_segmentation_mc"):
            self._segmentation_mc = _mc
        # window-sharded mode: fraction of an even window share that the clustering rank takes (None = even split); see
        # sharding.window_ranges.  DZ_ROOT_SHARE overrides.
        self.root_share = float(os.environ["DZ_ROOT_SHARE"]) if os.environ.get("DZ_ROOT_SHARE") else None
        self.collect_timing = os.environ.get("DZ_TIMING") is not None
        self._timing, self._t_last = {}, None

    # ------------------------------------------------------------------------------------------------
    def to(self, device) -> "DiariZenPipeline":
        """pyannote `Pipeline.to(device)` (pyannote-audio/pyannote/audio/core/pipeline.py:328-348): moves both models."""
        if not isinstance(device, torch.device):
            raise TypeError(f"`device` must be an instance of `torch.device`, got `{type(device).__name__}`")
        self._segmentation = self._segmentation.to(device)
        self._embedding = self._embedding.to(device)
        self.device = self._segmentation.device
        self.clustering.device = self.device
        self._planned = {}
        return self

    @property
    def model(self):
        """the segmentation model, as `SpeakerDiarization.model` (speaker_diarization.py:136)"""
        return self._segmentation

    @classmethod
    def from_pretrained(cls, repo_id: str, cache_dir: str = None, rttm_out_dir: str = None, **kw) -> "DiariZenPipeline":
        """`repo_id` may be a local directory laid out like the hub snapshot (config.toml, pytorch_model.bin and
        `wespeaker/pytorch_model.bin` or `embedding.bin`); otherwise the HF hub is queried as the reference does
        (inference.py:95-119)."""
        p = Path(repo_id)
        if p.is_dir():
            for cand in (p / "wespeaker" / "pytorch_model.bin", p / "embedding.bin"):
                if cand.exists():
                    return cls(diarizen_hub=p.expanduser().absolute(), embedding_model=str(cand), rttm_out_dir=rttm_out_dir, **kw)
            raise FileNotFoundError(f"no embedding checkpoint under {p} (expected wespeaker/pytorch_model.bin)")
        from huggingface_hub import hf_hub_download, snapshot_download
        hub = snapshot_download(repo_id=repo_id, cache_dir=cache_dir, local_files_only=cache_dir is not None)
        emb = hf_hub_download(repo_id="pyannote/wespeaker-voxceleb-resnet34-LM", filename="pytorch_model.bin",
                              cache_dir=cache_dir, local_files_only=cache_dir is not None)
        return cls(diarizen_hub=Path(hub).expanduser().absolute(), embedding_model=emb, rttm_out_dir=rttm_out_dir, **kw)

    @classmethod
    def from_random_init(cls, arch_name: str = "wavlm_large_s80_md", seed: int = 0, seg_duration: float = 16.0,
                         segmentation_step: float = 0.1, batch_size: int = 32, min_cluster_size: int = 30,
                         ahc_threshold: float = 0.70, min_speakers: int = 1, max_speakers: int = 20,
                         apply_median_filtering: bool = True, classifier_gain: float = 1.0, precision: str = "fp16",
                         rttm_out_dir: Optional[str] = None, device=None, emb_state_dict=None,
                         vbx: Optional[dict] = None, multichannel: Optional[dict] = None, seg_state_dict=None) -> "DiariZenPipeline":
        """Seeded random weights of the named architecture (no checkpoint is reachable offline: SURVEY.md 0.8)."""
        from .archs import init_resnet_state_dict
        dev = torch.device(device if device is not None else "cuda")
        arch = get_arch(arch_name)
        sd = seg_state_dict if seg_state_dict is not None else init_state_dict(arch, seed, classifier_gain)
        mc = None
        if multichannel is not None:
            # multichannel = {"fusion_dim": ..., "fusion_heads": ..., "fusion_layers": ...}; `seg_state_dict` must hold channel_fusions.*
            from .segmentation_mc import MCSegmentationModel
            mc = MCSegmentationModel(arch, sd, precision=precision, device=dev, **multichannel)
            seg = mc._merged
        else:
            seg = SegmentationModel(arch, sd, precision=precision, device=dev)
        emb = EmbeddingModel(emb_state_dict if emb_state_dict is not None else init_resnet_state_dict(seed), precision=precision, device=dev)
        config = {
            "model": {"args": {"wavlm_src": arch_name}},
            "inference": {"args": {"seg_duration": seg_duration, "segmentation_step": segmentation_step, "batch_size": batch_size,
                                   "apply_median_filtering": apply_median_filtering}},
            "clustering": {"args": {"method": "AgglomerativeClustering", "min_speakers": min_speakers, "max_speakers": max_speakers,
                                    "ahc_criterion": "distance", "ahc_threshold": ahc_threshold, "min_cluster_size": min_cluster_size}},
        }
        if vbx is not None:
            # vbx = {"plda_dir": ..., "Fa": ..., "Fb": ..., "lda_dim": ..., "max_iters": ...} switches to the VBx method
            config["clustering"]["args"].update({"method": "VBxClustering", "Fa": 0.07, "Fb": 0.8, "lda_dim": 128,
                                                 "max_iters": 20})
            config["clustering"]["args"].update(vbx)
        return cls(None, None, rttm_out_dir=rttm_out_dir, precision=precision, device=dev, _seg=seg, _emb=emb, _config=config, _mc=mc)

    # ------------------------------------------------------------------------------------------------
    def _windows(self, num_samples: int):
        window = int(math.floor(self.seg_duration * SR))                  # core/inference.py:265
        step = round(self.segmentation_step * self.seg_duration * SR)     # :266
        n_full = (num_samples - window) // step + 1 if num_samples >= window else 0
        has_last = (num_samples < window) or ((num_samples - window) % step > 0)
        return window, step, n_full + int(has_last)

    # ------------------------------------------------------------------------------------------------
    # stage 1 - every rank, on its own window range: the two networks and the per-window kernels between them
    # ------------------------------------------------------------------------------------------------
    def _front(self, wloc: torch.Tensor, n_loc: int, window: int, step: int, T: int, c0: int, c1: int, per: int):
        """wloc: the samples of windows c0..c1-1.  -> (raw, seg (per,T,S) uint8 median-filtered, stats (per,S,2) int32,
        emb (per,S,256) fp32); rows >= c1-c0 are padding."""
        dev, L, S = self.device, self._L, 4
        chunks = wloc.as_strided((max(n_loc, 0), window), (step, 1))
        raw = torch.zeros((per, T, S), device=dev, dtype=torch.uint8)
        bs = self._call_batch("seg", n_loc, self.engine_windows)
        # The engine plans (workspace + tensor maps) per batch shape: every call of a recording uses ONE batch size (the window
        # count split evenly over the fewest calls that fit `engine_windows`), the last call zero padded to it.
        for a in range(0, n_loc, bs):
            b = min(a + bs, n_loc)
            if b - a == bs:
                self._segmentation.hard(chunks[a:b].contiguous(), want_logp=False, ml_out=raw[a:b])
            else:
                wpad = torch.zeros((bs, window), device=dev, dtype=torch.float32)
                wpad[:b - a] = chunks[a:b]
                tail = torch.empty((bs, T, S), device=dev, dtype=torch.uint8)
                self._segmentation.hard(wpad, want_logp=False, ml_out=tail)
                raw[a:b] = tail[:b - a]
        self._mark("segmentation")
        return (raw,) + self._masks_and_embeddings(raw, wloc, chunks, window, step, T, c0, c1, per)

    def _masks_and_embeddings(self, raw, wloc, chunks, window, step, T, c0, c1, per):
        dev, L, S = self.device, self._L, 4
        st = vp(torch.cuda.current_stream(dev).cuda_stream)
        if self.apply_median_filtering:
            seg = torch.empty_like(raw)
            _lib.check(L.dz_median_filter(vp(raw.data_ptr()), vp(seg.data_ptr()), per, T, S, 11, st))
        else:
            seg = raw
        min_num_frames = math.ceil(T * self._embedding.min_num_samples / (self.seg_duration * SR))
        masks = torch.empty((per, S, T), device=dev, dtype=torch.float32)
        stats = torch.empty((per, S, 2), device=dev, dtype=torch.int32)
        _lib.check(L.dz_embedding_masks(vp(seg.data_ptr()), per, T, S, min_num_frames, vp(masks.data_ptr()), vp(stats.data_ptr()), st))
        # chunk crops as the reference computes them (io.py:359-364): start = floor(c * step_s * sr) in float64
        chunk_step_s = self.segmentation_step * self.seg_duration
        e_starts = [int(math.floor((c * chunk_step_s) * SR)) - c0 * step for c in range(c0, c1)]   # relative to wloc
        same = all(e == i * step for i, e in enumerate(e_starts))
        self._mark("count_masks")
        emb = torch.zeros((per, S, 256), device=dev, dtype=torch.float32)
        n_loc = c1 - c0
        ebs = self._call_batch("emb", n_loc, self.engine_emb_windows)
        for a in range(0, n_loc, ebs):
            b = min(a + ebs, n_loc)
            if same:
                wv = chunks[a:b].contiguous()
            else:
                wv = torch.stack([torch.nn.functional.pad(wloc[max(e_starts[i], 0):e_starts[i] + window], (0, max(0, e_starts[i] + window - wloc.shape[0])))[:window]
                                  for i in range(a, b)])
            mk = masks[a:b]
            if b - a < ebs:   # last call: zero padded to the recording's batch size
                wv = torch.cat([wv, torch.zeros((ebs - (b - a), window), device=dev, dtype=torch.float32)])
                mk = torch.cat([mk, torch.zeros((ebs - (b - a), S, T), device=dev, dtype=torch.float32)])
            emb[a:b] = self._embedding.embed_windows(wv, mk)[:b - a]
        self._mark("embedding")
        return seg, stats, emb

    # ------------------------------------------------------------------------------------------------
    # stage 2 - one rank: counting, clustering, reconstruction
    # ------------------------------------------------------------------------------------------------
    def _back(self, seg: torch.Tensor, stats: torch.Tensor, emb: torch.Tensor, Cn: int, T: int) -> Dict[str, Any]:
        dev, L, S = self.device, self._L, 4
        st = vp(torch.cuda.current_stream(dev).cuda_stream)
        chunk_step_s = self.segmentation_step * self.seg_duration
        starts = np.array([_closest_frame(c * chunk_step_s + 0.5 * FRAME_DURATION) for c in range(Cn)], dtype=np.int32)
        F = _closest_frame(self.seg_duration + (Cn - 1) * chunk_step_s + 0.5 * FRAME_DURATION) + 1
        dstart = torch.as_tensor(starts, device=dev)
        count = torch.empty(F, device=dev, dtype=torch.uint8)
        maxc = int(self.max_speakers) if self.max_speakers else 255
        _lib.check(L.dz_speaker_count(vp(seg.data_ptr()), vp(dstart.data_ptr()), Cn, T, S, F, maxc, vp(count.data_ptr()), st))
        cmax = count.max()                                   # stays on the device until the clustering is done
        emb_np = emb.cpu().numpy()
        stats_np = stats.cpu().numpy()
        self._mark("gather_to_host")
        hard, _, centroids = self.clustering(embeddings=emb_np, segmentations=None, min_clusters=self.min_speakers,
                                             max_clusters=self.max_speakers,
                                             frame_stats=(stats_np[..., 0], stats_np[..., 1], T))
        self._mark("clustering")
        if hard.size and int(hard.max()) > 127:
            raise ValueError(f"{int(hard.max()) + 1} clusters: cluster labels are int8 in the reference (at most 128 clusters)")
        hard = np.array(hard, dtype=np.int8, copy=True)
        hard[stats_np[..., 0] == 0] = -2                                   # inactive speakers (inference.py:166-170)
        K = max(int(hard.max()) + 1 if hard.size else 1, 1)
        Kout = max(K, int(cmax.item()))   # activations are zero padded up to the largest count (diarization.py:222-226)
        dh = torch.as_tensor(hard, device=dev)
        disc = torch.empty((F, Kout), device=dev, dtype=torch.uint8)
        _lib.check(L.dz_reconstruct(vp(seg.data_ptr()), vp(dh.data_ptr()), vp(dstart.data_ptr()), vp(count.data_ptr()), Cn, T, S,
                                    K, Kout, F, vp(disc.data_ptr()), None, st))
        discrete = disc.cpu().numpy()
        self._mark("reconstruct")
        out = {"segmentations": seg, "count": count, "embeddings": emb_np, "hard_clusters": hard, "discrete": discrete,
               "centroids": centroids, "num_chunks": Cn, "num_frames": T, "timing": dict(self._timing)}
        self.last = out
        return out

    def _call_batch(self, which: str, n: int, cap: int) -> int:
        """windows per engine call for a run of `n` windows: n split evenly over ceil(n / cap) calls (so that the zero padding
        of the last call is less than one window per call - with a fixed batch of `cap`, a rank holding 299 of the windows of a
        sharded recording would pad 37 of them).  The batch the engine is currently planned for is kept when it wastes < 3 %:
        a re-plan reallocates the workspace."""
        if n <= 0:
            return max(1, cap)
        calls = -(-n // cap)
        b = -(-n // calls)
        cur = self._planned.get(which)
        if cur is not None and cur <= cap and (-(-n // cur) * cur - n) <= 0.03 * n:
            return cur
        self._planned[which] = b
        return b

    def _mark(self, name: Optional[str]):
        """stage timer (only with `collect_timing`, which synchronises the device at every stage boundary)"""
        if not self.collect_timing:
            return
        import time
        torch.cuda.synchronize(self.device)
        now = time.perf_counter()
        if name is not None and self._t_last is not None:
            self._timing[name] = self._timing.get(name, 0.0) + 1e3 * (now - self._t_last)
        self._t_last = now

    def diarize_waveform(self, wav: torch.Tensor, shard: Optional[bool] = None, root: int = 0) -> Dict[str, Any]:
        """wav (N,) fp32 -> dict with every intermediate the parity tests compare.

        shard: window-shard this recording over the ranks of the initialised torch.distributed group: every rank runs both
        networks on its window range, ONE all-gather collects the packed per-window records (binarised segmentations,
        frame counters, embeddings) and rank `root` clusters and reconstructs (result there only, {} elsewhere).  Default:
        shard when a process group exists.  Policy for many recordings: `diarize_many`."""
        dist = torch.distributed if (torch.distributed.is_available() and torch.distributed.is_initialized()) else None
        if shard is False:
            dist = None
        rank = dist.get_rank() if dist else 0
        world = dist.get_world_size() if dist else 1
        dev = self.device
        self._timing, self._t_last = {}, None
        with torch.cuda.device(dev):
            self._mark(None)
            Nw = wav.shape[0]
            window, step, Cn = self._windows(Nw)
            T = self._segmentation.num_frames(window)
            ranges, per = window_ranges(Cn, world, root, self.root_share if world > 1 else None)
            c0, c1 = ranges[rank]
            # this rank's span of the recording: windows c0..c1-1 (zero padded past the end for the orphan chunk)
            s0, s1 = c0 * step, (max(c1, c0 + 1) - 1) * step + window
            if wav.device == dev and wav.dtype == torch.float32 and Nw >= s1:
                wloc = wav[s0:s1]
            else:
                wloc = torch.zeros(s1 - s0, device=dev, dtype=torch.float32)
                if Nw > s0:
                    wloc[:min(Nw, s1) - s0] = wav[s0:min(Nw, s1)].to(dev, torch.float32, non_blocking=True)
            self.last_h2d_bytes = 0 if wav.device == dev else 4 * max(0, min(Nw, s1) - s0)
            raw, seg, stats, emb = self._front(wloc, c1 - c0, window, step, T, c0, c1, per)
            self.last_raw = raw[:c1 - c0]      # this rank's window decisions before the median filter (parity tests)
            if world > 1:
                seg, stats, emb = gather_records(seg, stats, emb, Cn, world, ranges)
                self._mark("all_gather")
                if rank != root:
                    return {}
            else:
                seg, stats, emb = seg[:Cn], stats[:Cn], emb[:Cn]
            return self._back(seg, stats, emb, Cn, T)

    def tune_root_share(self, wav: torch.Tensor, root: int = 0) -> Optional[float]:
        """Window-sharded mode, recordings processed back to back: measure one sharded pass with stage timers and give the
        clustering rank the window share at which its networks + clustering take as long as the other ranks' networks
        (w_root = w_other - tail / t_window).  Collective: call on every rank.  -> the share (also stored in `root_share`)."""
        dist = torch.distributed if (torch.distributed.is_available() and torch.distributed.is_initialized()) else None
        if dist is None or dist.get_world_size() < 2:
            return None
        world, rank = dist.get_world_size(), dist.get_rank()
        keep, self.collect_timing, self.root_share = self.collect_timing, True, None
        import time
        res = self.diarize_waveform(wav, shard=True, root=root)
        t0 = time.perf_counter()
        if res:
            self.to_annotation(res["discrete"], None)
        share = torch.zeros(1, device=self.device, dtype=torch.float64)
        if rank == root:
            tm = res["timing"]
            Cn = res["num_chunks"]
            n_root = -(-Cn // world)
            t_window = (tm.get("segmentation", 0.0) + tm.get("count_masks", 0.0) + tm.get("embedding", 0.0)) / max(1, n_root)
            tail = tm.get("gather_to_host", 0.0) + tm.get("clustering", 0.0) + tm.get("reconstruct", 0.0) + 1e3 * (time.perf_counter() - t0)
            w_other = (Cn + tail / max(t_window, 1e-6)) / world
            share[0] = min(1.0, max(0.05, (w_other - tail / max(t_window, 1e-6)) / (Cn / world)))
        dist.all_reduce(share)
        self.collect_timing = keep
        self.root_share = float(share.item())
        return self.root_share

    # ------------------------------------------------------------------------------------------------
    # multi-channel recordings: recipes/diar_ssl_mc/infer_avg.py:47-118 `diarize_session`
    # ------------------------------------------------------------------------------------------------
    def diarize_multichannel(self, wav: torch.Tensor) -> Dict[str, Any]:
        """wav (channels, samples) fp32 -> the same dict as `diarize_waveform`.  Windows of all channels go through the
        multi-channel segmentation model; speaker embeddings are extracted per channel with the shared masks and averaged with
        the channel weights the model's 4th fusion module attends with (mean over frames and query channels of its attention
        map: infer_avg.py:33-45 `att_enhanced_emb`); everything after that is the single-channel path."""
        if self._segmentation_mc is None:
            raise RuntimeError("this pipeline was built with a single-channel segmentation model")
        dev, S = self.device, 4
        self._timing, self._t_last = {}, None
        with torch.cuda.device(dev):
            self._mark(None)
            Cch, Nw = wav.shape
            window, step, Cn = self._windows(Nw)
            T = self._segmentation.num_frames(window)
            pad_to = (Cn - 1) * step + window
            wdev = torch.zeros((Cch, max(pad_to, Nw)), device=dev, dtype=torch.float32)
            wdev[:, :Nw] = wav.to(dev, torch.float32)
            chunks = wdev.as_strided((Cn, Cch, window), (step, wdev.stride(0), 1))
            raw = torch.zeros((Cn, T, S), device=dev, dtype=torch.uint8)
            F = self._segmentation_mc.fusion_layers
            sel = 3 if F > 3 else F - 1                       # the recipe reads fusion module 3 (of 4)
            weights = torch.empty((Cn, Cch), device=dev, dtype=torch.float32)
            bs = max(1, self.segmentation_batch_size)
            for a in range(0, Cn, bs):
                b = min(a + bs, Cn)
                _, ml, att = self._segmentation_mc.hard(chunks[a:b].contiguous(), want_logp=False)
                raw[a:b] = ml
                weights[a:b] = att[:, sel].mean(dim=(1, 2))    # (windows, T, C, C) -> weight of channel j = mean over t, i
            self._mark("segmentation")
            # masks as in the single-channel path, then one embedding pass per channel on that channel's audio
            emb = torch.zeros((Cn, S, 256), device=dev, dtype=torch.float32)
            seg = stats = None
            for ch in range(Cch):
                seg, stats, e = self._masks_and_embeddings(raw, wdev[ch], wdev[ch].as_strided((Cn, window), (step, 1)), window, step, T, 0, Cn, Cn)
                emb += weights[:, ch, None, None] * e
            self.last_raw = raw
            self.last_channel_weights = weights
            return self._back(seg, stats, emb, Cn, T)

    def diarize_session(self, in_wav, sess_name=None):
        """multi-channel counterpart of `__call__`: every channel of the file is used (the single-channel `__call__` keeps
        channel 0 only, inference.py:128)."""
        if isinstance(in_wav, dict):
            w = torch.as_tensor(in_wav["waveform"], dtype=torch.float32)
            w = w[None] if w.dim() == 1 else w
            sr = int(in_wav.get("sample_rate", SR))
        else:
            x, sr = _read_wav(in_wav)
            w = torch.from_numpy(np.ascontiguousarray(x.T))
        if sr != SR:
            w = torch.stack([_to_16k(c, sr) for c in w])
        print("Extracting segmentations...")
        return self._finish(self.diarize_multichannel(w), sess_name)

    def diarize_segmentations(self, raw_segmentations, embeddings) -> Dict[str, Any]:
        """Stage 2 alone: (C,T,4) {0,1} window decisions as they leave the segmentation network and (C,4,256) embeddings ->
        the same dict as `diarize_waveform` (median filter, counting, clustering, reconstruction).  This is the seam the
        glue parity tests drive with reference-produced stage inputs."""
        dev, L, S = self.device, self._L, 4
        self._timing, self._t_last = {}, None
        with torch.cuda.device(dev):
            raw = torch.as_tensor(np.ascontiguousarray(raw_segmentations, dtype=np.uint8), device=dev)
            Cn, T, _ = raw.shape
            st = vp(torch.cuda.current_stream(dev).cuda_stream)
            seg = raw
            if self.apply_median_filtering:
                seg = torch.empty_like(raw)
                _lib.check(L.dz_median_filter(vp(raw.data_ptr()), vp(seg.data_ptr()), Cn, T, S, 11, st))
            masks = torch.empty((Cn, S, T), device=dev, dtype=torch.float32)
            stats = torch.empty((Cn, S, 2), device=dev, dtype=torch.int32)
            _lib.check(L.dz_embedding_masks(vp(seg.data_ptr()), Cn, T, S, 2, vp(masks.data_ptr()), vp(stats.data_ptr()), st))
            emb = torch.as_tensor(np.ascontiguousarray(embeddings, dtype=np.float32), device=dev)
            return self._back(seg, stats, emb, Cn, T)

    def diarize_many(self, waveforms, sess_names=None):
        """Several recordings over the ranks of the process group (BASELINE.json configs[4]).  Whole
        recordings go to ranks round-robin - no data-path collective, and the rank that owns a recording also clusters it,
        so the clustering of different recordings runs on different ranks at the same time.  The n mod world recordings left
        over are window-sharded over all ranks (one all-gather each) with the clustering rank rotating.
        -> list aligned with `waveforms`: the Annotation on the rank that finished the recording, None elsewhere."""
        dist = torch.distributed if (torch.distributed.is_available() and torch.distributed.is_initialized()) else None
        rank = dist.get_rank() if dist else 0
        world = dist.get_world_size() if dist else 1
        n = len(waveforms)
        names = list(sess_names) if sess_names is not None else [None] * n
        whole = (n // world) * world          # recordings handed out whole
        out = [None] * n
        for i in range(n):
            if i < whole:
                if i % world != rank:
                    continue
                res = self.diarize_waveform(load_waveform(waveforms[i]) if not torch.is_tensor(waveforms[i]) else waveforms[i], shard=False)
            else:
                res = self.diarize_waveform(load_waveform(waveforms[i]) if not torch.is_tensor(waveforms[i]) else waveforms[i], shard=True,
                                            root=i % world)
            if res:
                out[i] = self._finish(res, names[i])
        return out

    def _finish(self, res, sess_name):
        result = self.to_annotation(res["discrete"], sess_name)
        if self.rttm_out_dir is not None:
            assert sess_name is not None
            with open(os.path.join(self.rttm_out_dir, sess_name + ".rttm"), "w") as f:
                f.write(result.to_rttm())
        return result

    @staticmethod
    def to_annotation(discrete: np.ndarray, uri: Optional[str]) -> Annotation:
        """Binarize(onset=0.5, offset=0.5) on a {0,1} matrix (pyannote-audio/pyannote/audio/utils/signal.py:254-317):
        a region starts at the middle of the first active frame and ends at the middle of the first inactive one."""
        F, K = discrete.shape
        if F == 0 or K == 0:
            return Annotation(uri=uri)
        y = np.zeros((F + 2, K), dtype=np.int8)
        y[1:-1] = discrete
        d = np.diff(y, axis=0)                       # (F + 1, K): +1 at the first active frame of a run, -1 one past its last
        spk, first = np.nonzero(d.T == 1)            # row-major over (speaker, frame): the runs of one speaker in time order
        _, past = np.nonzero(d.T == -1)              # ... and the frame after each of those runs, in the same order
        last_excl = np.minimum(past, F - 1)          # a run reaching the end is closed at the last frame's middle
        s0 = first * FRAME_STEP
        e0 = last_excl * FRAME_STEP
        # middle of frame i = half the sum of its two ends: this order of operations decides the third decimal of the
        # RTTM times (pinned by tests/golden/glue_*.npz, produced by the reference's Binarize)
        return Annotation.from_arrays(0.5 * (s0 + (s0 + FRAME_DURATION)), 0.5 * (e0 + (e0 + FRAME_DURATION)), spk, uri=uri)

    def __call__(self, in_wav, sess_name=None, shard: Optional[bool] = None):
        wav = load_waveform(in_wav)
        print("Extracting segmentations.")
        res = self.diarize_waveform(wav, shard=shard)
        if not res:
            return None
        return self._finish(res, sess_name)
