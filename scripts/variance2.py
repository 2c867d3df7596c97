import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_meeting
from diarizen_b200.pipeline import DiariZenPipeline
pipe = DiariZenPipeline.from_random_init("wavlm_large_s80_md", seed=0, seg_duration=16.0, batch_size=32, classifier_gain=40.0)
wav = synth_meeting(20 * 60, 100).cuda()
window, step = 256000, 25600
Cn = (wav.shape[0] - window) // step + 1
chunks = wav.as_strided((Cn, window), (step, 1))
seg = torch.zeros((Cn, 799, 4), device="cuda", dtype=torch.uint8)
def loop(mode):
    evs = []
    cpu = []
    for a in range(0, Cn - 31, 32):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        if mode == "copy":
            w = chunks[a:a + 32].contiguous()
        else:
            w = fixed
        pipe._segmentation.hard(w, want_logp=False, ml_out=seg[a:a + 32])
        e1.record()
        cpu.append((time.perf_counter() - t0) * 1e3)
        evs.append((e0, e1))
    torch.cuda.synchronize()
    g = [e0.elapsed_time(e1) for e0, e1 in evs]
    return g, cpu
fixed = chunks[0:32].contiguous()
for mode in ("fixed", "copy", "copy", "fixed", "copy"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g, c = loop(mode)
    dt = (time.perf_counter() - t0) * 1e3
    print(mode, "total %.0f ms | gpu per batch: min %.1f max %.1f | cpu issue per batch: min %.2f max %.2f mean %.2f" % (dt, min(g), max(g), min(c), max(c), sum(c) / len(c)), flush=True)
    print("   gpu:", " ".join("%.0f" % x for x in g))
