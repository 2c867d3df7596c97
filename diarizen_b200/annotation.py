"""Minimal stand-ins for the `pyannote.core` result types the reference returns (pyannote.core is a third-party
dependency that is not vendored in the reference tree; semantics restated from SURVEY.md Appendix B).

Only what `DiariZenPipeline.__call__` callers use is provided: `Annotation.itertracks(yield_label=True)`, `.uri`,
`.labels()`, `.to_rttm()` and `Segment(start, end)` (reference: README.md:37-38, diarizen/pipelines/inference.py:184-191).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, List, Optional, Tuple


@dataclass(frozen=True, order=True)
class Segment:
    start: float
    end: float

    @property
    def duration(self) -> float:
        return self.end - self.start

    @property
    def middle(self) -> float:
        return 0.5 * (self.start + self.end)

    def __str__(self) -> str:
        return f"[{self.start:.3f} --> {self.end:.3f}]"


@dataclass(frozen=True)
class SlidingWindow:
    """frame geometry (pyannote.core.SlidingWindow semantics): frame i covers [start + i * step, start + i * step + duration)"""
    start: float
    duration: float
    step: float

    def __getitem__(self, i: int) -> Segment:
        return Segment(self.start + i * self.step, self.start + i * self.step + self.duration)

    def closest_frame(self, t: float) -> int:
        import numpy as np
        return int(np.rint((t - self.start - 0.5 * self.duration) / self.step))


@dataclass(frozen=True)
class Specifications:
    """the fields of the reference's task specifications that callers of the pipeline read
    (pyannote-audio/pyannote/audio/core/task.py:79-136; diarizen/pipelines/inference.py:93)"""
    duration: float
    classes: tuple
    powerset_max_classes: int
    powerset: bool = True
    permutation_invariant: bool = True
    warm_up: tuple = (0.0, 0.0)


class Annotation:
    """Speaker turns.  Tracks added one by one (`ann[segment, track] = label`) and tracks handed over as arrays
    (`Annotation.from_arrays`, what the pipeline does: a recording can hold 10^5 turns) live side by side; Segment objects
    for the array part are only made when somebody iterates."""

    def __init__(self, uri: Optional[str] = None):
        self.uri = uri
        self._tracks: List[Tuple[Segment, int, object]] = []   # (segment, track, label) in insertion order
        self._arr = None                                         # (starts, ends, tracks) float64 / float64 / int64 arrays

    @classmethod
    def from_arrays(cls, starts, ends, tracks, uri: Optional[str] = None) -> "Annotation":
        """turns [starts[i], ends[i]) of integer track / label tracks[i]; empty ones (end - start <= 1e-6) are dropped"""
        import numpy as np
        ann = cls(uri)
        starts, ends, tracks = np.asarray(starts, dtype=np.float64), np.asarray(ends, dtype=np.float64), np.asarray(tracks, dtype=np.int64)
        keep = (ends - starts) > 1e-6
        ann._arr = (starts[keep], ends[keep], tracks[keep])
        return ann

    def __setitem__(self, key, label):
        segment, track = key
        if not (segment.end - segment.start) > 1e-6:    # empty segments are not kept (a turn made of the last frame only)
            return
        self._tracks.append((segment, track, label))

    def __len__(self) -> int:
        return len(self._tracks) + (len(self._arr[0]) if self._arr is not None else 0)

    def _all(self) -> List[Tuple[Segment, int, object]]:
        out = list(self._tracks)
        if self._arr is not None:
            s, e, t = self._arr
            out += [(Segment(float(a), float(b)), int(k), int(k)) for a, b, k in zip(s.tolist(), e.tolist(), t.tolist())]
        return out

    def labels(self) -> list:
        return sorted({l for _, _, l in self._all()}, key=str)

    def itertracks(self, yield_label: bool = False) -> Iterator:
        # segments in (start, end) order; tracks sharing a segment in the order of their names
        for seg, trk, lab in sorted(self._all(), key=lambda x: (x[0].start, x[0].end, str(x[1]), str(x[2]))):
            yield (seg, trk, lab) if yield_label else (seg, trk)

    def to_rttm(self) -> str:
        uri = self.uri if self.uri else "<NA>"
        return "".join(f"SPEAKER {uri} 1 {s.start:.3f} {s.duration:.3f} <NA> <NA> {lab} <NA> <NA>\n"
                       for s, _, lab in self.itertracks(yield_label=True))
