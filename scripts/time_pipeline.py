import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DZ_TIMING"] = "1"
import torch
from bench import synth_meeting
from diarizen_b200.pipeline import DiariZenPipeline
mins = float(sys.argv[1]) if len(sys.argv) > 1 else 60
pipe = DiariZenPipeline.from_random_init("wavlm_large_s80_md", seed=0, seg_duration=16.0, batch_size=int(os.environ.get("DZ_BATCH", "32")), classifier_gain=40.0)
wav = synth_meeting(mins * 60, 100).cuda()
for i in range(3):
    t0 = time.perf_counter(); r = pipe.diarize_waveform(wav); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("run", i, "%.3f s" % dt, {k: round(v * 1e3, 1) for k, v in r["timing"].items()})
T = 799
wb = wav[:256000].repeat(pipe.engine_windows, 1).contiguous()
prof = pipe._segmentation.profile(wb); prof = pipe._segmentation.profile(wb)
tot = sum(p[1] for p in prof)
print("SEG per-batch total %.2f ms" % tot)
agg = {}
for n, ms, fl, by in prof:
    key = n.split("_", 1)[1] if n[0] in "LC" and "_" in n else n
    a = agg.setdefault(key, [0.0, 0.0, 0]); a[0] += ms; a[1] += fl; a[2] += 1
for k, (ms, fl, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:22]:
    print("  %-14s x%-3d %7.3f ms  %6.1f TF/s" % (k, c, ms, fl / ms / 1e9 if fl else 0))
pipe._embedding.embed_windows(wb[:pipe.engine_emb_windows], torch.ones(pipe.engine_emb_windows, 4, T, device="cuda"))
ep = pipe._embedding.profile(); ep = pipe._embedding.profile()
print("EMB per-batch(8) total %.2f ms" % sum(p[1] for p in ep))
for n, ms, fl in sorted(ep, key=lambda x: -x[1])[:24]:
    print("  %-16s %7.3f ms  %6.1f TF/s" % (n, ms, fl / ms / 1e9 if fl else 0))
