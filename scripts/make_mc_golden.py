"""tests/golden/seg_mc_<arch>.npz from the REFERENCE multi-channel model parts (oracle/ref_loader.RefSegModelMC = the reference's
wav2vec2_model.extract_features_mc + CrossChannelAttention + ConformerEncoder assembled as model_wavlm_conformer_mc.py:241-282):
seeded weights (stored), a 3-channel input, the log-probabilities and the channel-attention maps."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diarizen_b200.archs import get_arch, init_state_dict  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
FUSION = dict(fusion_dim=64, fusion_heads=4, fusion_layers=2)


def main():
    assert ref_loader.available()
    for name in ("tiny_base", "tiny_large"):
        a = get_arch(name)
        torch.manual_seed(21)
        m = ref_loader.RefSegModelMC(a, **FUSION).eval()
        sd = init_state_dict(a, seed=11)
        m.load_state_dict(sd, strict=False)
        g = torch.Generator().manual_seed(5)
        fus = {}
        for k, v in m.state_dict().items():
            if k.startswith("channel_fusions."):
                if k.endswith("ln_norm.weight"):
                    t = 0.5 + 0.5 * torch.rand(v.shape, generator=g)      # the reference initialises it to 1e-2: too small to test anything
                elif k.endswith("bias"):
                    t = 0.1 * torch.randn(v.shape, generator=g)
                else:
                    t = torch.randn(v.shape, generator=g) / v.shape[-1] ** 0.5
                fus[k] = t
        m.load_state_dict(fus, strict=False)
        wav = 0.1 * torch.randn(2, 3, 16000, generator=g)
        with torch.inference_mode():
            logp, att = m(wav)
        np.savez_compressed(os.path.join(OUT, f"seg_mc_{name}.npz"), wav=wav.numpy(), logp=logp.numpy(), att=att.numpy(), seed=11,
                            **{"fusion." + k: v.numpy() for k, v in fus.items()}, **{f"cfg_{k}": v for k, v in FUSION.items()})
        print(name, tuple(logp.shape), tuple(att.shape), float(att.std()))


if __name__ == "__main__" and "session" not in sys.argv:
    main()


def session():
    """tests/golden/glue_mc_session.npz: the reference recipe's `diarize_session` (recipes/diar_ssl_mc/infer_avg.py:47-118) on a
    3-channel synthetic recording, 5 s windows, the `tiny_base_mc` architecture with FOUR fusion modules (the recipe reads the
    attention map of module 3)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from synth_audio import integer_meeting
    from diarizen_b200.archs import init_resnet_state_dict
    from oracle import ref_glue
    a = get_arch("tiny_base_mc")
    fusion = dict(fusion_dim=64, fusion_heads=4, fusion_layers=4)
    torch.manual_seed(3)
    sd = init_state_dict(a, seed=2, classifier_gain=40.0)
    probe = ref_loader.RefSegModelMC(a, **fusion)
    g = torch.Generator().manual_seed(9)
    for k, v in probe.state_dict().items():
        if k.startswith("channel_fusions."):
            sd[k] = (0.5 + 0.5 * torch.rand(v.shape, generator=g)) if k.endswith("ln_norm.weight") else \
                    (0.1 * torch.randn(v.shape, generator=g) if k.endswith("bias") else torch.randn(v.shape, generator=g) / v.shape[-1] ** 0.5)
    esd = init_resnet_state_dict(2)
    seconds = 21.7
    chans = np.stack([integer_meeting(seconds, 40 + c, speakers=3) for c in range(3)])
    chans[1] = (chans[0].astype(np.int32) * 3 // 4 + chans[1].astype(np.int32) // 4).astype(np.int16)     # correlated microphones
    chans[2] = (chans[0].astype(np.int32) // 2 + chans[2].astype(np.int32) // 2).astype(np.int16)
    pipe = ref_glue.build_reference_mc_pipeline(a, sd, esd, fusion, seg_duration=5.0, min_cluster_size=3)
    cap = ref_glue.run_reference_mc_session(pipe, chans.astype(np.float32) / 32768.0)
    np.savez_compressed(os.path.join(OUT, "glue_mc_session.npz"), wav_i16=chans, seg_duration=5.0, weights_seed=2, classifier_gain=40.0, min_cluster_size=3,
                        raw_segmentations=cap["raw_segmentations"].astype(np.uint8), segmentations=cap["segmentations"].astype(np.uint8),
                        attention3=cap["attention"][:, 3].astype(np.float32), embeddings=cap["embeddings"].astype(np.float32),
                        hard_clusters=cap["hard_clusters"].astype(np.int8), discrete=cap["discrete"].astype(np.uint8), rttm=np.array(cap["rttm"]),
                        **{"fusion." + k: v.numpy() for k, v in sd.items() if k.startswith("channel_fusions.")},
                        **{f"cfg_{k}": v for k, v in fusion.items()})
    print("session", cap["segmentations"].shape, cap["attention"].shape, cap["discrete"].shape, len(cap["rttm"].splitlines()), "turns")


if __name__ == "__main__" and "session" in sys.argv:
    session()
