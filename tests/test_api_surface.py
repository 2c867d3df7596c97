"""Public surface inherited from the reference's pipeline classes: CPU-checkable parts."""
import numpy as np
import pytest


def test_receptive_field_matches_reference_helpers():
    from oracle import ref_glue
    if not ref_glue.available():
        pytest.skip("needs /root/reference")
    from diarizen_b200.segmentation import SegmentationModel
    ns = ref_glue.load()
    ks, st, pd, dl = [10, 3, 3, 3, 3, 2, 2], [5, 2, 2, 2, 2, 2, 2], [0] * 7, [1] * 7
    rf = ns.receptive_field
    size = rf.multi_conv_receptive_field_size(1, kernel_size=ks, stride=st, padding=pd, dilation=dl)
    step = rf.multi_conv_receptive_field_size(2, kernel_size=ks, stride=st, padding=pd, dilation=dl) - size
    center = rf.multi_conv_receptive_field_center(0, kernel_size=ks, stride=st, padding=pd, dilation=dl)
    sw = SegmentationModel._receptive_field.fget(None)
    assert (sw.start, sw.duration, sw.step) == ((center - (size - 1) / 2) / 16000, size / 16000, step / 16000)


def test_annotation_drops_empty_segments_and_orders_tracks():
    from diarizen_b200.annotation import Annotation, Segment
    a = Annotation(uri="u")
    a[Segment(1.0, 1.0), 0] = 0           # a turn made of a single frame: empty, dropped (pyannote.core semantics)
    a[Segment(0.5, 2.0), 10] = 10
    a[Segment(0.5, 2.0), 2] = 2
    a[Segment(0.25, 0.75), 1] = 1
    got = [(s.start, s.end, l) for s, _, l in a.itertracks(yield_label=True)]
    assert got == [(0.25, 0.75, 1), (0.5, 2.0, 10), (0.5, 2.0, 2)]        # tracks of one segment in str order: "10" < "2"
    assert a.to_rttm().splitlines()[0] == "SPEAKER u 1 0.250 0.500 <NA> <NA> 1 <NA> <NA>"


def test_to_annotation_matches_reference_binarize():
    """pipeline.to_annotation on a {0,1} matrix == the reference's Binarize (run through oracle/ref_glue.py when mounted)."""
    from oracle import ref_glue
    if not ref_glue.available():
        pytest.skip("needs /root/reference")
    from diarizen_b200.pipeline import DiariZenPipeline
    ns = ref_glue.load()
    r = np.random.default_rng(0)
    disc = (r.random((4000, 3)) < 0.5).astype(np.float64)
    for k in range(3):                      # runs instead of salt and pepper
        disc[:, k] = np.repeat(r.random(400) < 0.4, 10)
    disc[-1, 0], disc[-2, 0] = 1.0, 0.0     # last frame only: zero-length turn
    swf = ns.core.SlidingWindowFeature(disc, ns.core.SlidingWindow(start=0.0, duration=400 / 16000, step=320 / 16000))
    ref = ns.signal.Binarize(onset=0.5, offset=0.5, min_duration_on=0.0, min_duration_off=0.0)(swf)
    ref.uri = "x"
    assert DiariZenPipeline.to_annotation(disc.astype(np.uint8), "x").to_rttm() == ref.to_rttm()
