"""Generates tests/golden/glue_*.npz by running the REFERENCE's own glue code (oracle/ref_glue.py: the unmodified
files under /root/reference behind third-party stubs) - run in the container where /root/reference is mounted.

  glue_e2e_<arch>.npz  : DiariZenPipeline.__call__ of the reference on a synthetic meeting, with the pinned network oracles
                         inside: int16 waveform, raw / median-filtered segmentations, count, embeddings, hard clusters,
                         discrete diarization, RTTM text.
  glue_synth_<name>.npz: the same __call__ with scripted networks (a ground-truth turn script rendered to powerset
                         log-probabilities per window + prototype embeddings), so that several clusters, small clusters,
                         the num_clusters re-cut branches (min_speakers / max_speakers) and VBx are exercised; one set of
                         stage inputs, several hyper-parameter variants of the outputs.
  glue_clustering.npz  : direct calls of the reference AgglomerativeClustering / VBxClustering classes (incl. more than 32
                         clusters, NaN embeddings, every branch of `cluster`).
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from diarizen_b200.archs import get_arch, init_resnet_state_dict, init_state_dict  # noqa: E402
from oracle import ref_glue  # noqa: E402
from vbx_util import write_plda  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SR = 16000


def meeting(seconds, seed):
    g = torch.Generator().manual_seed(seed)
    n = int(seconds * SR)
    t = torch.arange(n) / SR
    wav = torch.zeros(n)
    for s, f0 in enumerate((180.0, 320.0, 520.0)):
        src = 0.05 * torch.randn(n, generator=g) + 0.1 * torch.sin(2 * np.pi * f0 * t) * (1 + 0.3 * torch.sin(2 * np.pi * (2 + s) * t))
        gate = torch.zeros(n)
        pos = int(torch.randint(0, SR, (1,), generator=g))
        while pos < n:
            on = int(torch.randint(8000, 64000, (1,), generator=g))
            gate[pos:pos + on] = 1.0
            pos += on + int(torch.randint(8000, 80000, (1,), generator=g))
        wav += src * gate
    w16 = (wav.clamp_(-1, 1) * 32767).round().to(torch.int16)
    return w16.numpy()


def pack(cap):
    return dict(raw_segmentations=cap["raw_segmentations"].astype(np.uint8), segmentations=cap["segmentations"].astype(np.uint8),
                count=cap["count"][:, 0].astype(np.uint8), embeddings=cap["embeddings"].astype(np.float32),
                hard_clusters=cap["hard_clusters"].astype(np.int8), discrete=cap["discrete"].astype(np.uint8),
                rttm=np.array(cap["rttm"]))


def e2e(arch_name, dur, seconds, seed, mcs):
    a = get_arch(arch_name)
    sd = init_state_dict(a, 2, 40.0)
    esd = init_resnet_state_dict(2)
    w16 = meeting(seconds, seed)
    pipe = ref_glue.build_reference_pipeline(a, sd, esd, seg_duration=dur, min_cluster_size=mcs)
    cap = ref_glue.run_reference_pipeline(pipe, w16.astype(np.float32) / 32768.0)
    np.savez_compressed(os.path.join(OUT, f"glue_e2e_{arch_name}.npz"), wav_i16=w16, seg_duration=dur, weights_seed=2,
                        classifier_gain=40.0, min_cluster_size=mcs, **pack(cap))
    print(arch_name, cap["segmentations"].shape, "K =", cap["discrete"].shape[1], len(cap["turns"]), "turns")


def e2e_large(seconds=300.0, seed=7):
    """The BENCHMARKED architecture end to end through the reference's __call__: wavlm_large_s80_md, 16 s windows, a 5-min
    integer-synthetic recording (tests/synth_audio.py - regenerated bit-exactly by the test, not stored), embedding bias
    centred so that several clusters form.  Stores the oracle top-2 log-prob margin per window frame so that the GPU test
    can explain any flipped decision without re-running the CPU oracle."""
    import hashlib
    from synth_audio import integer_meeting
    from oracle.emb_oracle import emb_forward
    from oracle.seg_oracle import seg_forward
    a = get_arch("wavlm_large_s80_md")
    sd = init_state_dict(a, 2, 40.0)
    esd = init_resnet_state_dict(2)
    w16 = integer_meeting(seconds, seed)
    wav = w16.astype(np.float32) / 32768.0
    # centre the embeddings on 16 calibration windows (all-ones masks)
    N = 256000
    cal = torch.from_numpy(np.stack([wav[s:s + N] for s in np.linspace(0, len(wav) - N - 1, 16).astype(int)]))
    mean = emb_forward(esd, cal, torch.ones(16, 1, 799))[:, 0].mean(0)
    esd = dict(esd)
    esd["resnet.seg_1.bias"] = esd["resnet.seg_1.bias"] - mean
    margins = []

    def seg_fn(w):
        logp = seg_forward(a, sd, torch.as_tensor(w))
        t2 = logp.topk(2, dim=-1).values
        margins.append((t2[..., 0] - t2[..., 1]).numpy())
        return logp
    pipe = ref_glue.build_reference_pipeline(a, sd, esd, seg_duration=16.0, min_cluster_size=10, seg_fn=seg_fn, batch_size=8)
    cap = ref_glue.run_reference_pipeline(pipe, wav)
    np.savez_compressed(os.path.join(OUT, "glue_e2e_large_s80.npz"), seconds=seconds, audio_seed=seed, audio_sha1=np.array(hashlib.sha1(w16.tobytes()).hexdigest()),
                        seg_duration=16.0, weights_seed=2, classifier_gain=40.0, min_cluster_size=10,
                        emb_bias=esd["resnet.seg_1.bias"].numpy(), margin=np.concatenate(margins).astype(np.float32), **pack(cap))
    print("large_s80", cap["segmentations"].shape, "K =", cap["discrete"].shape[1], "clusters", int(cap["hard_clusters"].max()) + 1, len(cap["turns"]), "turns")


# ------------------------------------------------------------------------------------------------------------
POWERSET = [(), (0,), (1,), (2,), (3,), (0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]   # pa/utils/powerset.py:68-90 order


def script_activity(num_frames, sizes, seed):
    """ground truth (frames x speakers) {0,1}: speaker g talks in turns; `sizes[g]` ~ share of the recording it is present"""
    r = np.random.default_rng(seed)
    G = len(sizes)
    act = np.zeros((num_frames, G), dtype=np.uint8)
    for g, share in enumerate(sizes):
        lo = int(r.integers(0, max(1, int(num_frames * (1 - share)))))
        hi = lo + int(num_frames * share)
        t = lo
        while t < hi:
            on = int(r.integers(40, 400))
            act[t:min(t + on, hi), g] = 1
            t += on + int(r.integers(30, 500))
    return act


def scripted_networks(act, T, frames_per_step, protos, noise, seed):
    """-> seg_fn, emb_fn (stateful: the reference calls them in chunk order) + the tables they serve"""
    r = np.random.default_rng(seed)
    state = {"c": 0, "e": 0}
    slots = {}
    D = protos.shape[1]
    const = r.standard_normal(D).astype(np.float32) * 0.01

    def chunk_logp(c):
        a = act[c * frames_per_step:c * frames_per_step + T]
        if a.shape[0] < T:
            a = np.pad(a, ((0, T - a.shape[0]), (0, 0)))
        who = [g for g in np.argsort(-a.sum(0), kind="stable") if a[:, g].sum() > 0][:4]
        who = list(r.permutation(who))                       # arbitrary local slot order, like a PIT-trained model
        slots[c] = who
        logp = np.full((T, 11), -20.0, dtype=np.float32)
        for t in range(T):
            on = tuple(s for s, g in enumerate(who) if a[t, g])[:2]
            if r.random() < 0.02:                            # salt noise for the median filter
                on = tuple(sorted(r.choice(4, size=int(r.integers(0, 3)), replace=False)))
            logp[t, POWERSET.index(on)] = -1e-4
        return logp

    def seg_fn(w):
        out = np.stack([chunk_logp(state["c"] + i) for i in range(w.shape[0])])
        state["c"] += w.shape[0]
        return out

    def emb_fn(w, masks):
        out = np.empty((w.shape[0], D), dtype=np.float32)
        for i in range(w.shape[0]):
            c, s = divmod(state["e"] + i, 4)
            who = slots[c]
            if s < len(who) and float(masks[i].sum()) > 0:
                out[i] = protos[who[s]] + noise * r.standard_normal(D)
            else:
                out[i] = const
        state["e"] += w.shape[0]
        return out
    return seg_fn, emb_fn


def synth(name, dur, seconds, sizes, variants, seed, noise=0.35):
    T = {5.0: 249, 16.0: 799}[dur]
    n = int(seconds * SR)
    a = get_arch("tiny_base")
    frames_per_step = int(round(0.1 * dur / 0.02))
    act = script_activity(int(seconds / 0.02) + T + 2, sizes, seed)
    protos = np.random.default_rng(seed + 1).standard_normal((len(sizes), 256)).astype(np.float32)
    w16 = (np.random.default_rng(seed + 2).standard_normal(n) * 1000).astype(np.int16)
    out = {}
    plda_dir = tempfile.mkdtemp()
    write_plda(plda_dir, seed=3)
    for vname, kw in variants.items():
        seg_fn, emb_fn = scripted_networks(act, T, frames_per_step, protos, noise, seed + 3)
        pipe = ref_glue.build_reference_pipeline(a, None, None, seg_duration=dur, seg_fn=seg_fn, emb_fn=emb_fn, batch_size=8,
                                                 **({**kw, "vbx": {**kw["vbx"], "plda_dir": plda_dir}} if "vbx" in kw else kw))
        cap = ref_glue.run_reference_pipeline(pipe, w16.astype(np.float32) / 32768.0)
        p = pack(cap)
        if not out:
            out.update(num_samples=n, seg_duration=dur, raw_segmentations=p["raw_segmentations"], segmentations=p["segmentations"],
                       embeddings=p["embeddings"], variants=np.array(list(variants)))
        elif not (np.array_equal(out["embeddings"], p["embeddings"]) and np.array_equal(out["segmentations"], p["segmentations"])):
            out[f"{vname}__segmentations"], out[f"{vname}__embeddings"] = p["segmentations"], p["embeddings"]
        for k in ("count", "hard_clusters", "discrete", "rttm"):
            out[f"{vname}__{k}"] = p[k]
        out[f"{vname}__params"] = np.array(repr(kw))
        print(name, vname, "C =", p["segmentations"].shape[0], "K =", p["discrete"].shape[1], "clusters", int(p["hard_clusters"].max()) + 1)
    np.savez_compressed(os.path.join(OUT, f"glue_synth_{name}.npz"), plda_seed=3, **out)


# ------------------------------------------------------------------------------------------------------------
def clustering_cases():
    ns = ref_glue.load()
    core = ns.core
    out = {}
    plda_dir = tempfile.mkdtemp()
    write_plda(plda_dir, seed=5)
    cases = {
        # name: (C, T, n_spk, noise, seed, params)
        "plain":        (100, 60, 4, 0.35, 0, dict(threshold=0.7, mcs=8, min=1, max=20)),
        "small":        (110, 60, 9, 0.35, 1, dict(threshold=0.7, mcs=25, min=1, max=20)),
        "recut_down":   (100, 60, 6, 0.35, 2, dict(threshold=0.7, mcs=6, min=1, max=3)),
        "recut_up":     (100, 60, 3, 0.30, 3, dict(threshold=0.9, mcs=6, min=5, max=20)),
        "recut_exact":  (100, 60, 5, 0.35, 4, dict(threshold=0.7, mcs=6, min=1, max=20, num=4)),
        "tiny":         (12, 60, 6, 0.35, 5, dict(threshold=0.2, mcs=30, min=1, max=20)),
        "many":         (220, 60, 45, 0.20, 6, dict(threshold=0.7, mcs=1, min=1, max=None)),
        "loose":        (90, 60, 3, 0.80, 7, dict(threshold=1.1, mcs=4, min=2, max=20)),
        "vbx":          (120, 60, 4, 0.35, 8, dict(vbx=dict(ahc_threshold=0.6, ahc_criterion="distance", Fa=0.07, Fb=0.8, lda_dim=128, maxIters=20))),
        "vbx_maxclust": (120, 60, 5, 0.35, 9, dict(vbx=dict(ahc_threshold=7, ahc_criterion="maxclust", Fa=0.1, Fb=1.0, lda_dim=64, maxIters=10))),
    }
    for name, (C, T, n_spk, noise, seed, prm) in cases.items():
        r = np.random.default_rng(seed)
        spk = r.standard_normal((n_spk, 256))
        S = 4
        emb = np.full((C, S, 256), np.nan, dtype=np.float32)
        seg = np.zeros((C, T, S), dtype=np.float32)
        pop = r.dirichlet(np.ones(n_spk) * (0.6 if name in ("small", "recut_up") else 3.0))
        for c in range(C):
            k = int(r.integers(1, min(S, n_spk) + 1))
            who = r.choice(n_spk, size=k, replace=False, p=pop)
            for s, w in enumerate(who):
                emb[c, s] = spk[w] + noise * r.standard_normal(256)
                a0 = int(r.integers(0, T // 2))
                seg[c, a0:a0 + int(r.integers(3, T // 2)), s] = 1.0
        if name == "plain":
            emb[3, 0] = np.nan          # an active speaker whose embedding extraction failed
        swf = core.SlidingWindowFeature(seg, core.SlidingWindow(start=0.0, duration=5.0, step=0.5))
        if "vbx" in prm:
            v = prm["vbx"]
            cl = ns.clustering.VBxClustering(metric="cosine")
            cl.instantiate({"ahc_criterion": v["ahc_criterion"], "ahc_threshold": v["ahc_threshold"], "Fa": v["Fa"], "Fb": v["Fb"]})
            cl.plda_dir, cl.lda_dim, cl.maxIters = plda_dir, v["lda_dim"], v["maxIters"]
            hard, soft, cent = cl(embeddings=emb.copy(), segmentations=swf, min_clusters=1, max_clusters=20)
        else:
            cl = ns.clustering.AgglomerativeClustering(metric="cosine")
            cl.instantiate({"method": "centroid", "min_cluster_size": prm["mcs"], "threshold": prm["threshold"]})
            hard, soft, cent = cl(embeddings=emb.copy(), segmentations=swf, num_clusters=prm.get("num"), min_clusters=prm["min"],
                                  max_clusters=prm["max"])
        out[f"{name}__embeddings"] = emb
        out[f"{name}__segmentations"] = seg.astype(np.uint8)
        out[f"{name}__hard"] = np.asarray(hard).astype(np.int16)
        out[f"{name}__soft"] = np.asarray(soft, dtype=np.float64)
        out[f"{name}__centroids"] = np.asarray(cent, dtype=np.float64)
        out[f"{name}__params"] = np.array(repr(prm))
        print("clustering", name, "K =", np.asarray(soft).shape[-1], "labels", np.unique(hard)[:8])
    np.savez_compressed(os.path.join(OUT, "glue_clustering.npz"), names=np.array(list(cases)), plda_seed=5, **out)


def main():
    assert ref_glue.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["e2e", "synth", "clustering"]
    if "clustering" in which:
        clustering_cases()
    if "synth" in which:
        V = dict(ahc_threshold=0.7, min_cluster_size=10, min_speakers=1, max_speakers=20)
        synth("5s", 5.0, 123.37, [0.9, 0.7, 0.5, 0.35, 0.06, 0.04], {
            "default": dict(V),
            "max3": dict(V, max_speakers=3),
            "min8": dict(V, min_speakers=8),
            "nomedian": dict(V, apply_median_filtering=False),
            "vbx": dict(method="VBxClustering", ahc_threshold=0.6, min_speakers=1, max_speakers=20,
                        vbx=dict(ahc_criterion="distance", Fa=0.07, Fb=0.8, lda_dim=128, max_iters=20)),
        }, seed=10)
        synth("16s", 16.0, 187.9, [0.8, 0.6, 0.5, 0.08], {"default": dict(V, min_cluster_size=4)}, seed=20)
    if "large" in which:
        e2e_large()
    if "e2e" in which:
        e2e("tiny_base", 5.0, 31.3, 1, 3)
        e2e("tiny_large", 16.0, 61.0, 4, 2)


if __name__ == "__main__":
    main()
