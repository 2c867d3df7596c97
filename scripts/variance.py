import os, sys, time, subprocess, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DZ_TIMING"] = "1"
import torch
from bench import synth_meeting
from diarizen_b200.pipeline import DiariZenPipeline
pipe = DiariZenPipeline.from_random_init("wavlm_large_s80_md", seed=0, seg_duration=16.0, batch_size=32, classifier_gain=40.0)
wav = synth_meeting(20 * 60, 100).cuda()
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader", "-lms", "250"], stdout=subprocess.PIPE, text=True)
lines = []
threading.Thread(target=lambda: [lines.append((time.perf_counter(), l.strip())) for l in smi.stdout], daemon=True).start()
for i in range(8):
    t0 = time.perf_counter(); r = pipe.diarize_waveform(wav); torch.cuda.synchronize(); t1 = time.perf_counter()
    clk = [l for (t, l) in lines if t0 <= t <= t1]
    print("run", i, "%.3f s" % (t1 - t0), {k: round(v * 1e3) for k, v in r["timing"].items()}, "reserved GB %.1f" % (torch.cuda.memory_reserved() / 2**30), clk[:3], flush=True)
smi.terminate()
# isolate: segmentation engine only, same batch repeatedly
wb = wav[:256000].repeat(32, 1).contiguous()
for i in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        pipe._segmentation.hard(wb, want_logp=False)
    torch.cuda.synchronize(); print("seg x20 batches: %.1f ms/batch" % ((time.perf_counter() - t0) * 50), flush=True)
