"""TEST INFRASTRUCTURE ONLY - runs the REFERENCE's own glue code (unmodified files under /root/reference) so that the
glue goldens in tests/golden/glue_*.npz are produced by reference code, not by a restatement.

What is real (loaded by file path from /root/reference, executed as is):
  pyannote-audio/pyannote/audio/core/inference.py        Inference.infer / slide / aggregate / trim
  pyannote-audio/pyannote/audio/core/io.py               Audio (validate_file, crop with mode="pad")
  pyannote-audio/pyannote/audio/utils/{powerset,multi_task,reproducibility,signal}.py
  pyannote-audio/pyannote/audio/pipelines/utils/diarization.py   speaker_count / to_diarization
  pyannote-audio/pyannote/audio/pipelines/clustering.py  AgglomerativeClustering / VBxClustering (+ diarizen/clustering/VBx.py)
  pyannote-audio/pyannote/audio/pipelines/speaker_diarization.py get_segmentations / get_embeddings / reconstruct
  diarizen/pipelines/inference.py                        DiariZenPipeline.__call__
What is stubbed (third-party packages that are not in this image, none of them arithmetic on the path except
pyannote.core, see oracle/pyannote_core_stub.py): pyannote.core, pyannote.pipeline (attribute plumbing),
pyannote.metrics, pyannote.database, pytorch_lightning's is_oom_error, the `Model` base class / `Specifications`
dataclass (pyannote-audio/pyannote/audio/core/{model,task}.py need lightning + torchmetrics; their fields used by
the glue are mirrored below), the HF-hub loading constructors (the pipeline object is assembled with
object.__new__ and the attributes its __call__ reads).  numpy >= 2 removed `np.NaN` / `np.NAN`, which the reference
spells (core/inference.py:550, speaker_diarization.py:403): aliased here.

The two networks inside are the repo's pinned torch oracles (oracle/seg_oracle.py, oracle/emb_oracle.py - each
pinned to the reference modules separately); this file is about the glue around them.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from dataclasses import dataclass
from enum import Enum
from functools import cached_property
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("DIARIZEN_REF", "/root/reference")
PA = os.path.join(REF, "pyannote-audio", "pyannote", "audio")


def available() -> bool:
    return os.path.isfile(os.path.join(PA, "pipelines", "clustering.py"))


# ------------------------------------------------------------------------------------------------------------
# stubs
# ------------------------------------------------------------------------------------------------------------
class Problem(Enum):
    BINARY_CLASSIFICATION = 0
    MONO_LABEL_CLASSIFICATION = 1
    MULTI_LABEL_CLASSIFICATION = 2
    REPRESENTATION = 3
    REGRESSION = 4


class Resolution(Enum):
    FRAME = 1
    CHUNK = 2


@dataclass
class Specifications:
    """field-for-field mirror of pyannote-audio/pyannote/audio/core/task.py:79-136"""
    problem: Problem
    resolution: Resolution
    duration: float
    min_duration: Optional[float] = None
    warm_up: Optional[Tuple[float, float]] = (0.0, 0.0)
    classes: Optional[List[str]] = None
    powerset_max_classes: Optional[int] = None
    permutation_invariant: bool = False

    @cached_property
    def powerset(self) -> bool:
        return self.powerset_max_classes is not None

    def __len__(self):
        return 1

    def __iter__(self):
        yield self


class _Param:
    def __init__(self, *a, **k):
        pass


class _ParamDict(dict):
    def __init__(self, **k):
        super().__init__(**k)


class _PipelineBase:
    """pyannote.pipeline.Pipeline: hyper-parameters are plain attributes once `instantiate`d - here they are plain
    attributes from the start (the tests set them directly, as DiariZenPipeline.instantiate(PIPELINE_PARAMS) does)."""
    training = False

    def __init__(self):
        pass

    def instantiate(self, params):
        for k, v in params.items():
            if isinstance(v, dict):
                sub = getattr(self, k)
                if isinstance(sub, _PipelineBase):
                    sub.instantiate(v)
                else:
                    setattr(self, k, v)
            else:
                setattr(self, k, v)
        return self


def _mod(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    parent, _, leaf = name.rpartition(".")
    if parent in sys.modules:
        setattr(sys.modules[parent], leaf, m)
    return m


_NS = None


def load():
    """-> namespace with the reference classes (cached)."""
    global _NS
    if _NS is not None:
        return _NS
    assert available(), "needs /root/reference"
    if not hasattr(np, "NaN"):
        np.NaN = np.nan
    if not hasattr(np, "NAN"):
        np.NAN = np.nan
    from oracle import pyannote_core_stub as core
    if REF not in sys.path:
        sys.path.insert(0, REF)     # diarizen.clustering.VBx (pure numpy/scipy) imports for real
    _mod("pyannote")
    _mod("pyannote.core", Segment=core.Segment, SlidingWindow=core.SlidingWindow, SlidingWindowFeature=core.SlidingWindowFeature,
         Annotation=core.Annotation, Timeline=core.Timeline)
    _mod("pyannote.core.utils")
    _mod("pyannote.core.utils.types", Label=object)
    import itertools
    _mod("pyannote.core.utils.generators", pairwise=itertools.pairwise)
    _mod("pyannote.pipeline", Pipeline=_PipelineBase)
    _mod("pyannote.pipeline.parameter", Categorical=_Param, Integer=_Param, Uniform=_Param, ParamDict=_ParamDict)
    _mod("pyannote.metrics")
    _mod("pyannote.metrics.diarization", DiarizationErrorRate=object, GreedyDiarizationErrorRate=object)
    _mod("pyannote.database")
    _mod("pyannote.database.protocol")
    _mod("pyannote.database.protocol.protocol", ProtocolFile=type("ProtocolFile", (dict,), {}))
    _mod("pytorch_lightning")
    _mod("pytorch_lightning.utilities")
    _mod("pytorch_lightning.utilities.memory", is_oom_error=lambda e: "out of memory" in str(e))

    class Model(nn.Module):
        """the attributes of pyannote-audio/pyannote/audio/core/model.py:132-195 the glue reads"""

    _mod("pyannote.audio")
    _mod("pyannote.audio.core")
    _mod("pyannote.audio.utils")
    _mod("pyannote.audio.pipelines")
    io = _load("pyannote.audio.core.io", os.path.join(PA, "core", "io.py"))
    _mod("pyannote.audio.core.task", Problem=Problem, Resolution=Resolution, Specifications=Specifications)
    _mod("pyannote.audio.core.model", Model=Model, Specifications=Specifications)
    _load("pyannote.audio.utils.multi_task", os.path.join(PA, "utils", "multi_task.py"))
    ps = _load("pyannote.audio.utils.powerset", os.path.join(PA, "utils", "powerset.py"))
    _load("pyannote.audio.utils.reproducibility", os.path.join(PA, "utils", "reproducibility.py"))
    rf = _load("pyannote.audio.utils.receptive_field", os.path.join(PA, "utils", "receptive_field.py"))
    _mod("pyannote.audio.utils.permutation", permutate=None)
    inf = _load("pyannote.audio.core.inference", os.path.join(PA, "core", "inference.py"))
    sig = _load("pyannote.audio.utils.signal", os.path.join(PA, "utils", "signal.py"))
    dia = _load("pyannote.audio.pipelines.utils.diarization", os.path.join(PA, "pipelines", "utils", "diarization.py"))
    _mod("pyannote.audio.pipelines.utils", SpeakerDiarizationMixin=dia.SpeakerDiarizationMixin, oracle_segmentation=None,
         PipelineModel=object, get_model=None)
    clu = _load("pyannote.audio.pipelines.clustering", os.path.join(PA, "pipelines", "clustering.py"))

    class AudioPipeline(_PipelineBase):
        """pyannote.audio.core.pipeline.Pipeline (HF-hub loading, .to(device)): not on the __call__ path"""

    _mod("pyannote.audio", Audio=io.Audio, Inference=inf.Inference, Model=Model, Pipeline=AudioPipeline)
    _mod("pyannote.audio.pipelines.speaker_verification", PretrainedSpeakerEmbedding=None)
    sd = _load("pyannote.audio.pipelines.speaker_diarization", os.path.join(PA, "pipelines", "speaker_diarization.py"))
    _mod("pyannote.audio.pipelines", SpeakerDiarization=sd.SpeakerDiarization)
    _mod("diarizen")
    if "diarizen.pipelines" not in sys.modules:
        import diarizen.pipelines  # noqa: F401  (real package: __init__ is empty)
    dz = _load("diarizen.pipelines.inference", os.path.join(REF, "diarizen", "pipelines", "inference.py"))
    _mod("pyannote.metrics.segmentation", Annotation=core.Annotation, Segment=core.Segment)
    mc = _load("ref_recipe_diar_ssl_mc_infer_avg", os.path.join(REF, "recipes", "diar_ssl_mc", "infer_avg.py"))
    _NS = types.SimpleNamespace(mc_recipe=mc, core=core, io=io, inference=inf, signal=sig, diarization=dia, clustering=clu, speaker_diarization=sd,
                                dz=dz, powerset=ps, receptive_field=rf, Model=Model, Specifications=Specifications,
                                Problem=Problem, Resolution=Resolution)
    return _NS


# ------------------------------------------------------------------------------------------------------------
# a reference DiariZenPipeline object around the two pinned network oracles
# ------------------------------------------------------------------------------------------------------------
def build_reference_pipeline(arch, seg_state_dict, emb_state_dict, seg_duration: float, segmentation_step: float = 0.1,
                             batch_size: int = 8, method: str = "AgglomerativeClustering", ahc_threshold: float = 0.70,
                             min_cluster_size: int = 30, min_speakers=1, max_speakers=20, apply_median_filtering: bool = True,
                             vbx: Optional[dict] = None, seg_fn=None, emb_fn=None):
    """Assembles what DiariZenPipeline.__init__ (diarizen/pipelines/inference.py:27-93) and SpeakerDiarization.__init__
    (speaker_diarization.py:115-186) assemble, minus the checkpoint / hub loading.  `seg_fn(wave (B,N)) -> logp (B,T,11)`
    and `emb_fn(wave (B,N), masks (B,T)) -> (B,256)` default to the pinned oracles."""
    ns = load()
    from oracle.emb_oracle import emb_forward
    from oracle.seg_oracle import seg_forward
    core = ns.core

    class SegModel(ns.Model):
        def __init__(self):
            super().__init__()
            self.sample_rate = 16000
            self.audio = ns.io.Audio(sample_rate=16000, mono="downmix")
            self.specifications = ns.Specifications(problem=ns.Problem.MONO_LABEL_CLASSIFICATION, resolution=ns.Resolution.FRAME,
                                                    duration=seg_duration, warm_up=(0.0, 0.0),
                                                    classes=[f"speaker#{i + 1}" for i in range(4)], powerset_max_classes=2,
                                                    permutation_invariant=True)
            self._dummy = nn.Parameter(torch.zeros(1))
            # core/model.py:180-195 with the conv-stack geometry of model_wavlm_conformer.py:113-176, evaluated with the
            # reference's own receptive_field.py helpers
            ks, st, pd, dl = [10, 3, 3, 3, 3, 2, 2], [5, 2, 2, 2, 2, 2, 2], [0] * 7, [1] * 7
            size = ns.receptive_field.multi_conv_receptive_field_size(1, kernel_size=ks, stride=st, padding=pd, dilation=dl)
            step = ns.receptive_field.multi_conv_receptive_field_size(2, kernel_size=ks, stride=st, padding=pd, dilation=dl) - size
            center = ns.receptive_field.multi_conv_receptive_field_center(0, kernel_size=ks, stride=st, padding=pd, dilation=dl)
            self._receptive_field = core.SlidingWindow(start=(center - (size - 1) / 2) / 16000, duration=size / 16000, step=step / 16000)

        @property
        def device(self):
            return self._dummy.device

        def forward(self, waveforms):
            w = waveforms[:, 0, :]
            if seg_fn is not None:
                return torch.as_tensor(seg_fn(w))
            return seg_forward(arch, seg_state_dict, w)

    class Embedding:
        sample_rate, dimension, metric, min_num_samples = 16000, 256, "cosine", 400

        def __call__(self, waveforms, masks=None):
            w = waveforms[:, 0, :]
            if emb_fn is not None:
                return np.asarray(emb_fn(w, masks))
            return emb_forward(emb_state_dict, w, masks[:, None, :])[:, 0].numpy()

    P = ns.dz.DiariZenPipeline
    pipe = object.__new__(P)
    model = SegModel()
    pipe.model = model
    pipe.segmentation_step = segmentation_step
    pipe.embedding_batch_size = batch_size
    pipe.embedding_exclude_overlap = True
    pipe.klustering = method
    pipe._segmentation = ns.inference.Inference(model, duration=seg_duration, step=segmentation_step * seg_duration,
                                                skip_aggregation=True, batch_size=batch_size, device=torch.device("cpu"))
    pipe._embedding = Embedding()
    pipe._audio = ns.io.Audio(sample_rate=16000, mono="downmix")
    pipe.clustering = ns.clustering.Clustering[method].value(metric="cosine")
    pipe.apply_median_filtering = apply_median_filtering
    pipe.min_speakers, pipe.max_speakers = min_speakers, max_speakers
    if method == "AgglomerativeClustering":
        pipe.clustering.instantiate({"method": "centroid", "min_cluster_size": min_cluster_size, "threshold": ahc_threshold})
    else:
        pipe.clustering.instantiate({"ahc_criterion": vbx.get("ahc_criterion", "distance"), "ahc_threshold": ahc_threshold,
                                     "Fa": vbx["Fa"], "Fb": vbx["Fb"]})
        pipe.clustering.plda_dir = vbx["plda_dir"]
        pipe.clustering.lda_dim = vbx["lda_dim"]
        pipe.clustering.maxIters = vbx["max_iters"]
    pipe.rttm_out_dir = None
    return pipe


def run_reference_pipeline(pipe, wav: np.ndarray, sess_name: str = "sess"):
    """pipe(in_wav) with every intermediate the reference computes captured on the way -> dict of numpy arrays + RTTM text.
    `torchaudio.load` (broken in this image: needs torchcodec) is replaced by a closure returning `wav`."""
    ns = load()
    import torchaudio
    cap = {}
    P = type(pipe)
    orig = {k: getattr(P, k) for k in ("get_segmentations", "speaker_count", "get_embeddings", "reconstruct")}
    orig_clu = pipe.clustering.__class__.__call__

    def get_segmentations(self, file, hook=None, soft=False):
        out = orig["get_segmentations"](self, file, hook=hook, soft=soft)
        cap["raw_segmentations"] = out.data.copy()
        return out

    def speaker_count(binarized, frames, warm_up=(0.1, 0.1)):
        cap["segmentations"] = binarized.data.copy()
        out = orig["speaker_count"](binarized, frames, warm_up=warm_up)
        cap["count"] = out.data.copy()
        return out

    def get_embeddings(self, file, seg, exclude_overlap=False, hook=None):
        out = orig["get_embeddings"](self, file, seg, exclude_overlap=exclude_overlap, hook=hook)
        cap["embeddings"] = out.copy()
        return out

    def reconstruct(self, seg, hard, count):
        cap["hard_clusters"] = np.array(hard, copy=True)
        cap["count_capped"] = count.data.copy()
        out = orig["reconstruct"](self, seg, hard, count)
        cap["discrete"] = out[0].data.copy()
        cap["activations"] = out[1].data.copy()
        return out

    def clu_call(self, *a, **k):
        out = orig_clu(self, *a, **k)
        cap["clustering_hard"] = np.array(out[0], copy=True)
        cap["soft_clusters"] = np.array(out[1], copy=True)
        cap["centroids"] = np.array(out[2], copy=True)
        return out

    real_load = torchaudio.load
    w = torch.as_tensor(wav, dtype=torch.float32)
    torchaudio.load = lambda path: (w[None] if w.dim() == 1 else w, 16000)
    P.get_segmentations, P.get_embeddings, P.reconstruct = get_segmentations, get_embeddings, reconstruct
    P.speaker_count = staticmethod(speaker_count)
    pipe.clustering.__class__.__call__ = clu_call
    try:
        ann = pipe("in-memory.wav", sess_name=sess_name)
    finally:
        torchaudio.load = real_load
        P.get_segmentations, P.get_embeddings, P.reconstruct = orig["get_segmentations"], orig["get_embeddings"], orig["reconstruct"]
        P.speaker_count = staticmethod(orig["speaker_count"])
        pipe.clustering.__class__.__call__ = orig_clu
    cap["rttm"] = ann.to_rttm()
    cap["turns"] = [(s.start, s.end, l) for s, _, l in ann.itertracks(yield_label=True)]
    return cap


# ------------------------------------------------------------------------------------------------------------
# the multi-channel recipe: recipes/diar_ssl_mc/infer_avg.py `diarize_session` around the reference MC model parts
# ------------------------------------------------------------------------------------------------------------
def build_reference_mc_pipeline(arch, seg_state_dict, emb_state_dict, fusion: dict, seg_duration: float, segmentation_step: float = 0.1,
                                batch_size: int = 8, ahc_threshold: float = 0.70, min_cluster_size: int = 3):
    """-> a pyannote `SpeakerDiarization` object as recipes/diar_ssl_mc/infer_avg.py builds it (minus checkpoint loading): its
    segmentation model is oracle/ref_loader.RefSegModelMC (the reference's own wav2vec2 / fusion / conformer modules), the
    embedding model the pinned ResNet oracle."""
    ns = load()
    from oracle import ref_loader
    from oracle.emb_oracle import emb_forward
    core = ns.core
    inner = ref_loader.RefSegModelMC(arch, **fusion).eval()
    inner.load_state_dict(seg_state_dict, strict=False)

    class SegModel(ns.Model):
        def __init__(self):
            super().__init__()
            self.inner = inner
            self.sample_rate = 16000
            self.audio = ns.io.Audio(sample_rate=16000, mono=None)          # core/model.py:152-157: num_channels > 1 => no downmix
            self.specifications = ns.Specifications(problem=ns.Problem.MONO_LABEL_CLASSIFICATION, resolution=ns.Resolution.FRAME,
                                                    duration=seg_duration, warm_up=(0.0, 0.0), classes=[f"speaker#{i + 1}" for i in range(4)],
                                                    powerset_max_classes=2, permutation_invariant=True)
            self._receptive_field = core.SlidingWindow(start=(79 - 199.5) / 16000, duration=400 / 16000, step=320 / 16000)

        @property
        def device(self):
            return next(self.inner.parameters()).device

        def forward(self, waveforms):
            return self.inner(waveforms)

    class Embedding:
        sample_rate, dimension, metric, min_num_samples = 16000, 256, "cosine", 400

        def __call__(self, waveforms, masks=None):
            return emb_forward(emb_state_dict, waveforms[:, 0, :], masks[:, None, :])[:, 0].numpy()

    P = ns.speaker_diarization.SpeakerDiarization
    pipe = object.__new__(P)
    model = SegModel()
    pipe.model = model
    pipe.segmentation_step = segmentation_step
    pipe.embedding_batch_size = batch_size
    pipe.embedding_exclude_overlap = True
    pipe._segmentation = ns.inference.Inference(model, duration=seg_duration, step=segmentation_step * seg_duration, skip_aggregation=True,
                                                batch_size=batch_size, device=torch.device("cpu"))
    pipe._embedding = Embedding()
    pipe._audio = ns.io.Audio(sample_rate=16000, mono="downmix")
    pipe.clustering = ns.clustering.Clustering["AgglomerativeClustering"].value(metric="cosine")
    pipe.clustering.instantiate({"method": "centroid", "min_cluster_size": min_cluster_size, "threshold": ahc_threshold})
    return pipe


def run_reference_mc_session(pipe, wav_mc: np.ndarray, sess_name: str = "sess", min_speakers=1, max_speakers=20):
    """recipes/diar_ssl_mc/infer_avg.py:47-118 `diarize_session` on an in-memory (channels, samples) waveform."""
    ns = load()
    import torchaudio
    cap = {}
    P = type(pipe)
    orig_seg, orig_rec = P.get_segmentations, P.reconstruct
    orig_clu = pipe.clustering.__class__.__call__

    def get_segmentations(self, file, hook=None, soft=False):
        out = orig_seg(self, file, hook=hook, soft=soft)
        cap["raw_segmentations"] = out[0].data.copy()
        cap["attention"] = np.array(out[1], copy=True)
        return out

    def reconstruct(self, seg, hard, count):
        cap["segmentations"] = seg.data.copy()
        cap["hard_clusters"] = np.array(hard, copy=True)
        cap["count"] = count.data.copy()
        out = orig_rec(self, seg, hard, count)
        cap["discrete"] = out[0].data.copy()
        return out

    def clu_call(self, *a, **k):
        cap["embeddings"] = np.array(k["embeddings"], copy=True)
        return orig_clu(self, *a, **k)

    real_load = torchaudio.load
    w = torch.as_tensor(wav_mc, dtype=torch.float32)
    torchaudio.load = lambda path: (w, 16000)
    P.get_segmentations, P.reconstruct = get_segmentations, reconstruct
    pipe.clustering.__class__.__call__ = clu_call
    try:
        ann = ns.mc_recipe.diarize_session(sess_name, "in-memory.wav", pipe, min_speakers=min_speakers, max_speakers=max_speakers)
    finally:
        torchaudio.load = real_load
        P.get_segmentations, P.reconstruct = orig_seg, orig_rec
        pipe.clustering.__class__.__call__ = orig_clu
    cap["rttm"] = ann.to_rttm()
    return cap
