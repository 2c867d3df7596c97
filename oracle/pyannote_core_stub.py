"""TEST INFRASTRUCTURE ONLY - a minimal stand-in for the third-party package `pyannote.core` (pinned 5.0.0 by
pyannote-audio/requirements.txt:6), which is absent from /root/reference and from this image.

Only what the reference's hot-path glue calls is provided (call sites: pyannote-audio/pyannote/audio/core/inference.py
:377-379,577-666, pipelines/utils/diarization.py:143-238, pipelines/speaker_diarization.py:271-425,
utils/signal.py:254-317, diarizen/pipelines/inference.py:184-191).  With this stub registered as `pyannote.core`, the
reference's OWN glue code (Inference.slide/aggregate/trim, speaker_count, get_embeddings, reconstruct, to_diarization,
Binarize, the clustering classes, DiariZenPipeline.__call__) executes unmodified - see oracle/ref_glue.py - so the
golden vectors in tests/golden/glue_*.npz come from reference code, and only the arithmetic below is "restated":

  Segment        : (start, end); non-empty iff end - start > 1e-6; `&` = intersection; middle, duration
  SlidingWindow  : frame i = [start + i*step, start + i*step + duration); closest_frame(t) = rint((t - start -
                   duration/2) / step); range_to_segment; crop (loose / strict / center)
  SlidingWindowFeature : (data, sliding_window[, labels]); iteration yields (window[i], data[i]); numpy ufuncs and
                   reductions applied to it return a SlidingWindowFeature on the same window; extent; crop
  Annotation     : ann[segment, track] = label (empty segments are dropped); itertracks sorted by (start, end) then
                   track; to_rttm lines "SPEAKER {uri} 1 {start:.3f} {duration:.3f} <NA> <NA> {label} <NA> <NA>"
"""
from __future__ import annotations

import numbers
from typing import Iterator, Optional, Tuple

import numpy as np

SEGMENT_PRECISION = 1e-6


class Segment:
    __slots__ = ("start", "end")

    def __init__(self, start: float = 0.0, end: float = 0.0):
        object.__setattr__(self, "start", start)
        object.__setattr__(self, "end", end)

    def __setattr__(self, k, v):
        raise AttributeError("Segment is immutable")

    def __bool__(self):
        return bool((self.end - self.start) > SEGMENT_PRECISION)

    @property
    def duration(self) -> float:
        return self.end - self.start if self else 0.0

    @property
    def middle(self) -> float:
        return 0.5 * (self.start + self.end)

    def __iter__(self):
        yield self.start
        yield self.end

    def __and__(self, other: "Segment") -> "Segment":
        return Segment(max(self.start, other.start), min(self.end, other.end))

    def __or__(self, other: "Segment") -> "Segment":
        if not self:
            return other
        if not other:
            return self
        return Segment(min(self.start, other.start), max(self.end, other.end))

    def _key(self):
        return (self.start, self.end)

    def __eq__(self, other):
        return isinstance(other, Segment) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def __lt__(self, other):
        return self._key() < other._key()

    def __le__(self, other):
        return self._key() <= other._key()

    def __repr__(self):
        return f"<Segment({self.start:g}, {self.end:g})>"


class SlidingWindow:
    def __init__(self, duration: float = 0.030, step: float = 0.010, start: float = 0.000, end: Optional[float] = None):
        if duration <= 0:
            raise ValueError("'duration' must be a float > 0.")
        if step <= 0:
            raise ValueError("'step' must be a float > 0.")
        self.duration, self.step, self.start = duration, step, start
        self.end = np.inf if end is None else end

    def closest_frame(self, t: float) -> int:
        return int(np.rint((t - self.start - 0.5 * self.duration) / self.step))

    def samples(self, from_duration: float, mode: str = "strict") -> int:
        if mode == "strict":
            return int(np.floor((from_duration - self.duration) / self.step)) + 1
        if mode == "loose":
            return int(np.floor((from_duration + self.duration) / self.step))
        if mode == "center":
            return int(np.rint(from_duration / self.step))
        raise ValueError(mode)

    def __getitem__(self, i: int) -> Optional[Segment]:
        start = self.start + i * self.step
        if start >= self.end:
            return None
        return Segment(start, start + self.duration)

    def range_to_segment(self, i0: int, n: int) -> Segment:
        start = self.start + (i0 - 0.5) * self.step + 0.5 * self.duration
        end = start + n * self.step
        if i0 == 0:
            start = self.start
        return Segment(start, end)

    def crop(self, focus: Segment, mode: str = "loose", fixed: Optional[float] = None, return_ranges: bool = False):
        if mode == "loose":
            i = int(np.ceil((focus.start - self.duration - self.start) / self.step))
            if fixed is None:
                j = int(np.floor((focus.end - self.start) / self.step))
                rng = (i, j + 1)
            else:
                rng = (i, i + self.samples(fixed, mode="loose"))
        elif mode == "strict":
            i = int(np.ceil((focus.start - self.start) / self.step))
            if fixed is None:
                j = int(np.floor((focus.end - self.duration - self.start) / self.step))
                rng = (i, j + 1)
            else:
                rng = (i, i + self.samples(fixed, mode="strict"))
        elif mode == "center":
            i = self.closest_frame(focus.start)
            if fixed is None:
                rng = (i, self.closest_frame(focus.end) + 1)
            else:
                rng = (i, i + self.samples(fixed, mode="center"))
        else:
            raise ValueError("'mode' must be one of {'loose', 'strict', 'center'}.")
        if return_ranges:
            return [list(rng)]
        return np.array(range(*rng), dtype=np.int64)

    def __eq__(self, other):
        return (isinstance(other, SlidingWindow) and self.duration == other.duration and self.step == other.step
                and self.start == other.start and self.end == other.end)

    def __repr__(self):
        return f"<SlidingWindow(start={self.start:g}, duration={self.duration:g}, step={self.step:g})>"


class SlidingWindowFeature(np.lib.mixins.NDArrayOperatorsMixin):
    _HANDLED_TYPES = (np.ndarray, numbers.Number)

    def __init__(self, data: np.ndarray, sliding_window: SlidingWindow, labels=None):
        self.sliding_window = sliding_window
        self.data = data
        self.labels = labels
        self.__i = -1

    def __len__(self):
        return self.data.shape[0]

    @property
    def extent(self) -> Segment:
        return self.sliding_window.range_to_segment(0, len(self))

    @property
    def dimension(self):
        return self.data.shape[1]

    def __getitem__(self, i):
        return self.data[i]

    def __iter__(self):
        self.__i = -1
        return self

    def __next__(self) -> Tuple[Segment, np.ndarray]:
        self.__i += 1
        try:
            return self.sliding_window[self.__i], self.data[self.__i]
        except IndexError:
            raise StopIteration()

    def crop(self, focus: Segment, mode: str = "loose", fixed: Optional[float] = None, return_data: bool = True):
        ranges = self.sliding_window.crop(focus, mode=mode, fixed=fixed, return_ranges=True)
        n = self.data.shape[0]
        clipped, first, last = [], 0, 0
        for s, e in ranges:
            first += min(e, 0) - min(s, 0)
            last += max(e, n) - max(s, n)
            if e < 0 or s >= n:
                continue
            clipped.append([max(s, 0), min(e, n)])
        if clipped:
            data = np.vstack([self.data[s:e] for s, e in clipped])
        else:
            data = np.empty((0,) + self.data.shape[1:], dtype=self.data.dtype)
        if fixed is not None:
            data = np.vstack([np.tile(self.data[0], (first,) + (1,) * (self.data.ndim - 1)).reshape((first,) + self.data.shape[1:]),
                              data,
                              np.tile(self.data[n - 1], (last,) + (1,) * (self.data.ndim - 1)).reshape((last,) + self.data.shape[1:])])
        if return_data:
            return data
        sw = SlidingWindow(start=self.sliding_window[clipped[0][0]].start, duration=self.sliding_window.duration,
                           step=self.sliding_window.step)
        return SlidingWindowFeature(data, sw, labels=self.labels)

    def __array__(self, dtype=None, copy=None) -> np.ndarray:
        return self.data if dtype is None else self.data.astype(dtype)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        out = kwargs.get("out", ())
        for x in inputs + out:
            if not isinstance(x, self._HANDLED_TYPES + (SlidingWindowFeature,)):
                return NotImplemented
        inputs = tuple(x.data if isinstance(x, SlidingWindowFeature) else x for x in inputs)
        if out:
            kwargs["out"] = tuple(x.data if isinstance(x, SlidingWindowFeature) else x for x in out)
        data = getattr(ufunc, method)(*inputs, **kwargs)
        if type(data) is tuple:
            return tuple(type(self)(x, self.sliding_window) for x in data)
        if method == "at":
            return None
        return type(self)(data, self.sliding_window)


class Timeline:
    def __init__(self, segments=None, uri=None):
        self.uri = uri
        self.segments_ = sorted(s for s in (segments or []) if s)

    def __iter__(self) -> Iterator[Segment]:
        return iter(self.segments_)

    def __len__(self):
        return len(self.segments_)

    def add(self, segment: Segment):
        if segment and segment not in self.segments_:
            self.segments_.append(segment)
            self.segments_.sort()
        return self


class Annotation:
    def __init__(self, uri=None, modality=None):
        self.uri = uri
        self.modality = modality
        self._tracks = {}            # Segment -> {track: label}

    def __setitem__(self, key, label):
        segment, track = key
        if not segment:              # empty segments are silently dropped
            return
        self._tracks.setdefault(segment, {})[track] = label

    def __delitem__(self, key):
        segment, track = key
        del self._tracks[segment][track]
        if not self._tracks[segment]:
            del self._tracks[segment]

    def __len__(self):
        return len(self._tracks)

    def __bool__(self):
        return len(self._tracks) > 0

    def itertracks(self, yield_label: bool = False):
        for segment in sorted(self._tracks):
            for track, lbl in sorted(self._tracks[segment].items(), key=lambda tl: (str(tl[0]), str(tl[1]))):
                yield (segment, track, lbl) if yield_label else (segment, track)

    def labels(self):
        return sorted({l for t in self._tracks.values() for l in t.values()}, key=str)

    def rename_tracks(self, generator="string"):
        return self

    def support(self, collar: float = 0.0):
        raise NotImplementedError("Annotation.support is not on the DiariZen hot path (Binarize pads are 0)")

    def _iter_rttm(self):
        uri = self.uri if self.uri else "<NA>"
        if isinstance(uri, str) and " " in uri:
            raise ValueError(f'Space-separated RTTM file format does not allow file URIs containing spaces (got: "{uri}").')
        for segment, _, label in self.itertracks(yield_label=True):
            if isinstance(label, str) and " " in label:
                raise ValueError(f'Space-separated RTTM file format does not allow labels containing spaces (got: "{label}").')
            yield f"SPEAKER {uri} 1 {segment.start:.3f} {segment.duration:.3f} <NA> <NA> {label} <NA> <NA>\n"

    def to_rttm(self) -> str:
        return "".join(self._iter_rttm())
